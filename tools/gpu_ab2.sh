A="--steps 20 --warmup 5 --no-configs"
for m in 1 2 3; do echo "== GRB_PIPE_MODE=$m"; GRB_PIPE_MODE=$m bash tools/ab.sh "$A" default; done
