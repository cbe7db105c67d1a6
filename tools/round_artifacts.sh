# Round-2 evidence run (one B200): everything below lands in gpurun_out/ and is summarised into profiles/r2/ afterwards.
# Every ncu capture comes after the same command has exited 0 without ncu; numbers printed under ncu are never bench values.
set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r2.log; tail -3 gpurun_out/pytest_r2.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r2_ref.json 2> gpurun_out/bench_r2_ref.err; tail -c 400 gpurun_out/bench_r2_ref.json
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2_n1.json 2> gpurun_out/bench_r2_n1.err; tail -c 300 gpurun_out/bench_r2_n1.json
python tools/bench_configs.py c3ops > gpurun_out/c3ops_r2.jsonl 2> gpurun_out/c3ops_r2.err; cat gpurun_out/c3ops_r2.jsonl
# launch list of the timed map passes and the index builds: our kernels only (-k regex:k_), so the list is not cut off inside torch's generator
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/launches_r2.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/ncu_l_r2.log 2>&1
# full captures: the headline kernel, and the wavelet runs of k_sweep3 (median only, 30M x 30M)
ncu --set full --clock-control none --import-source on -k regex:k_map_pipe -s 3 -c 1 -o gpurun_out/prof_pipe_r2 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e --no-configs > gpurun_out/ncu_p_r2.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep3 -s 1 -c 1 -o gpurun_out/prof_wave_r2 -f python tools/bench_configs.py c3ops --scale 0.3 > gpurun_out/ncu_w_r2.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none -k regex:"k_map_pipe|k_coarse|k_fine|k_unbucket" -s 8 -c 24 --csv --log-file gpurun_out/variants_kernels_r2.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e > gpurun_out/ncu_v_r2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tile|k_sweep|k_median|k_minmax|k_map_pipe" -c 60 --csv --log-file gpurun_out/launches_c3_r2.csv python tools/bench_configs.py c3ops > gpurun_out/ncu_lc3_r2.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:k_ --csv --log-file gpurun_out/build_launches_r2.csv python tools/time_build.py 100000000 1 > gpurun_out/ncu_b_r2.log 2>&1
python tools/time_build.py > gpurun_out/build_times_r2.txt 2>&1; tail -6 gpurun_out/build_times_r2.txt
python tools/time_bucket.py > gpurun_out/unsorted_times_r2.txt 2>&1; tail -8 gpurun_out/unsorted_times_r2.txt
ls -la gpurun_out/*.ncu-rep | tail -3
