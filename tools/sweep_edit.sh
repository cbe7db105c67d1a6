#!/bin/bash
# usage: tools/sweep_edit.sh "label|file|sed-expr" ...   applies each edit on top of a pristine copy, rebuilds, benches, restores
for spec in "$@"; do
  IFS='|' read -r label file expr <<< "$spec"
  cp "$file" /tmp/sweep_backup.$$ 
  sed -i "$expr" "$file"
  make -C granges_b200/csrc -j8 2>&1 | grep -iE " error" -A3
  python bench.py --no-cpu --no-e2e 2>&1 | python tools/benchline.py "$label"
  cp /tmp/sweep_backup.$$ "$file"
done
make -C granges_b200/csrc -j8 > /dev/null 2>&1
