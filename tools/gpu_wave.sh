set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_wavelet.py tests/test_cuda_parity.py tests/test_cuda_pipe_variants.py tests/test_cuda_full_size.py -x -q > gpurun_out/wave_test.log 2>&1; echo "rc=$?" >> gpurun_out/wave_test.log; tail -5 gpurun_out/wave_test.log
timeout 600 python tools/bench_configs.py c3ops > gpurun_out/c3ops_wave.jsonl 2> gpurun_out/c3ops_wave.err; cat gpurun_out/c3ops_wave.jsonl; tail -3 gpurun_out/c3ops_wave.err
GRB_PLANAR_ROWS=0 timeout 600 python tools/bench_configs.py c3 > gpurun_out/c3_noplanar.jsonl 2>/dev/null; cat gpurun_out/c3_noplanar.jsonl
