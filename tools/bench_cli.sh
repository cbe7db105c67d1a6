#!/bin/bash
# End-to-end CLI timing (text in, text out): tools/bench_cli.sh [rows]
set -e
N=${1:-20000000}
D=$(mktemp -d)
python - > $D/hg38.tsv <<'PY'
import sys; sys.path.insert(0, '.')
import bench
for i, l in enumerate(bench.HG38): print(f"chr{i+1}\t{l}")
PY
G=granges_b200/granges
tm() { local label="$1"; shift; local t0=$(date +%s%N); "$@"; local t1=$(date +%s%N); echo "$label: $(( (t1 - t0) / 1000000 )) ms"; }
tm "random-bed left" $G random-bed $D/hg38.tsv -n $N -s --seed 13 -o $D/left.bed
tm "random-bed right" $G random-bed $D/hg38.tsv -n $N -s -c --seed 14 -o $D/right.bed
ls -la $D | tail -3
export GRANGES_TIMING=1
for t in "" 1; do
  export GRANGES_THREADS=$t
  tm "granges map --func mean ($N x $N rows, GRANGES_THREADS='$t')" $G map -g $D/hg38.tsv -l $D/left.bed -r $D/right.bed -f mean -o $D/out.tsv
done
unset GRANGES_THREADS
tm "granges filter" $G filter -g $D/hg38.tsv -l $D/left.bed -r $D/right.bed -o $D/filt.tsv
tm "granges windows 1kb/500bp" $G windows -g $D/hg38.tsv -w 1000 -s 500 -o $D/win.bed
tm "granges merge (BED5, --func mean, $N sorted rows)" $G merge -b $D/right.bed -f mean -o $D/merged.bed
awk 'BEGIN{OFS="\t"} NR<=2000000{print $1,$2,$3,"f" (NR%5)}' $D/right.bed > $D/feat.bed
tm "granges feature-density (2M BED4 rows, 5 features, 100 kb windows)" $G feature-density -g $D/hg38.tsv -b $D/feat.bed -w 100000 -l -o $D/fd.tsv
wc -l $D/out.tsv $D/filt.tsv $D/win.bed $D/merged.bed $D/fd.tsv; head -2 $D/out.tsv; head -2 $D/merged.bed; head -2 $D/fd.tsv
rm -rf $D
