set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1_final.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_r1_final.log; tail -3 gpurun_out/pytest_r1_final.log
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_r1s3_ref.json 2> gpurun_out/bench_r1s3_ref.err; tail -c 600 gpurun_out/bench_r1s3_ref.json
python bench.py > gpurun_out/bench_r1s3_n1.json 2> gpurun_out/bench_r1s3_n1.err; python tools/benchline.py n1 < gpurun_out/bench_r1s3_n1.json
python tools/bench_configs.py > gpurun_out/configs_r1s3.jsonl 2> gpurun_out/configs_r1s3.err; cat gpurun_out/configs_r1s3.jsonl; tail -3 gpurun_out/configs_r1s3.err
python tools/bench_configs.py c3ops > gpurun_out/c3ops_r1s3.jsonl 2>/dev/null; cat gpurun_out/c3ops_r1s3.jsonl
python tools/bench_configs.py c4ops > gpurun_out/c4ops_r1s3.jsonl 2>/dev/null; cat gpurun_out/c4ops_r1s3.jsonl
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches_r1s3.csv python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_l_r1s3.log 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"k_tile|k_sweep|k_median|k_minmax|k_overlap|k_weighted|k_map_pipe" -c 400 --csv --log-file gpurun_out/launches_c3_r1s3.csv python tools/bench_configs.py c3ops > gpurun_out/ncu_lc3_r1s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_map_pipe -c 1 -o gpurun_out/prof_pipe_r1s3 -f python bench.py --steps 1 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_p_r1s3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:k_sweep3 -c 1 -o gpurun_out/prof_sweep3_r1s3 -f python tools/bench_configs.py c3 > gpurun_out/ncu_s3_r1s3.log 2>&1
bash tools/bench_cli.sh 20000000 > gpurun_out/cli_r1s3.txt 2>&1; tail -25 gpurun_out/cli_r1s3.txt
python tools/e2e_phases.py > gpurun_out/e2e_phases_r1s3.txt 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
