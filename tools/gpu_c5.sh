timeout 600 python -m pytest tests/test_cuda_parity.py tests/test_cuda_full_size.py -x -q -k "pairs or left_overlaps or golden or c5 or C5" 2>&1 | tail -2
for v in default default; do
  echo -n "$v: "; timeout 600 python tools/bench_c5_mgpu.py 2>/dev/null | tail -1 | cut -c1-120
done
