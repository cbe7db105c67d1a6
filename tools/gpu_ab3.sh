A="--steps 20 --warmup 5 --no-configs"
bash tools/ab.sh "$A" default cons1 l12w3 l10w3 l4w5 default cons1
