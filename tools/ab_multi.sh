#!/bin/bash
# usage: ab2.sh "ENV1" "ENV2" ... ; each a space-separated env assignment list ("" = default); 2 reps interleaved
for r in 1 2; do
  for kv in "$@"; do
    out=$(env $kv timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-arm --no-optimizer 2>/dev/null | tail -1)
    echo "[$kv] $(echo $out | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"])')"
  done
done
