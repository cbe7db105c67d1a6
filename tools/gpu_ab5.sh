A="--steps 20 --warmup 5 --no-configs"
bash tools/ab.sh "$A" default pair default pair
GRB_LIB_PATH=$PWD/build_ab/lib_pair.so timeout 900 python -m pytest tests/test_cuda_pipe_variants.py tests/test_cuda_parity.py tests/test_cuda_full_size.py -x -q 2>&1 | tail -3
