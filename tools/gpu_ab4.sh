for v in default prev default prev; do
  if [ "$v" = default ]; then unset GRB_LIB_PATH; else export GRB_LIB_PATH=$PWD/build_ab/lib_$v.so; fi
  echo -n "$v: "
  timeout 600 python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
print(round(d['ms_per_step'],4), {k:round(v['ms'],3) for k,v in d['configs'].items()}, {k:round(v['ms'],3) for k,v in d['variants'].items()})"
done
