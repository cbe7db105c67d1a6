#!/bin/bash
# Build build_ab/lib_<name>.so: the in-tree library with csrc/<file>.cu recompiled with extra flags (A/B runs via GRB_LIB_PATH / tools/ab.sh).
# usage: tools/build_variant.sh <name> <file-without-.cu> "<extra nvcc flags>"
set -e
name=$1; file=$2; extra=$3
cd "$(dirname "$0")/../granges_b200/csrc"
mkdir -p ../../build_ab
NV="/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC,-Wall,-Wno-unused-function,-pthread -I../../include --cudart static"
$NV $extra -c $file.cu -o ../../build_ab/$file.$name.o 2>/dev/null
objs=""
for f in ctx scan sort index query pipe bucket reduce windows shard merge whole; do
  if [ $f = $file ]; then objs="$objs ../../build_ab/$file.$name.o"; else objs="$objs $f.o"; fi
done
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -shared --cudart static -Xcompiler -pthread -o ../../build_ab/lib_$name.so $objs -lpthread
echo built build_ab/lib_$name.so
