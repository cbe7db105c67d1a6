A="--steps 20 --warmup 5 --no-configs"
echo "== round-robin (default)"; bash tools/ab.sh "$A" default
export GRB_PIPE_ASSIGN=1
echo "== runs of 4"; bash tools/ab.sh "$A" default
echo "== runs of 2 / 8 / 16"; bash tools/ab.sh "$A" run2 run8 run16
unset GRB_PIPE_ASSIGN
echo "== round-robin again"; bash tools/ab.sh "$A" default
