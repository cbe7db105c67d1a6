mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:"k_coarse|k_fine|k_unbucket|k_map_pipe" -c 10 --csv --log-file gpurun_out/bucket_kernels_b.csv python tools/time_bucket.py 100000000 > gpurun_out/ncu_bucket_b.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/bucket_kernels_b.csv')) if len(r)>10]
h=rows[0]
for r in rows[1:]:
    d=dict(zip(h,r)); print(d['Kernel Name'][:40], d['Metric Name'], d['Metric Value'])
PY
