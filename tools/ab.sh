#!/bin/bash
# A/B runs of alternative builds of the library (build_ab/lib_<name>.so, same ABI) through bench.py: ms per step and roofline fraction.
# usage: tools/ab.sh "<bench args>" name1 name2 ...   ("default" = the in-tree build)
args=$1; shift
for v in "$@"; do
  if [ "$v" = default ]; then unset GRB_LIB_PATH; else export GRB_LIB_PATH=$PWD/build_ab/lib_$v.so; fi
  echo -n "$v: "
  timeout 300 python bench.py --no-cpu --no-e2e $args 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['roofline']['frac'],4), d['step_ms_rank0'])" 2>&1 | tail -1
done
