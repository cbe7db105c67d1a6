A="--steps 20 --warmup 5 --no-configs"
bash tools/ab.sh "$A" default diag3 diag4 diag5 diag2 default
