timeout 600 python -m pytest tests/test_cuda_parity.py -x -q -k "pairs or left_overlaps or golden" 2>&1 | tail -2
timeout 600 python tools/bench_c5_mgpu.py 2>/dev/null | tail -1
