#!/bin/bash
# usage: tools/sweep_pipe.sh "L W LUT" ...   (rebuilds the library per configuration and prints the bench line)
for cfg in "$@"; do
  set -- $cfg
  sed -i "s/^constexpr int PIPE_LSTAGES = [0-9]*;/constexpr int PIPE_LSTAGES = $1;/; s/^constexpr int PIPE_WSTAGES = [0-9]*;/constexpr int PIPE_WSTAGES = $2;/; s/^constexpr u32 PIPE_LUT = [0-9]*; /constexpr u32 PIPE_LUT = $3; /" granges_b200/csrc/map_pipe.cuh
  make -C granges_b200/csrc -j8 2>&1 | grep -iE "error" -A3
  python bench.py --no-cpu --no-e2e 2>&1 | python tools/benchline.py "L=$1,W=$2,LUT=$3"
done
