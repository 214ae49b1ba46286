#!/bin/bash
# A/B of one environment switch on the SAME box (boxes differ by +-1-2 %): bash tools/ab_env.sh VAR=VALUE [reps]
# prints device-resident ms/step for the default build and with the variable set, interleaved.
KV=$1; REPS=${2:-3}
for r in $(seq 1 $REPS); do
  for mode in default "$KV"; do
    if [ "$mode" = "default" ]; then
      out=$(timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-arm --no-optimizer 2>/dev/null | tail -1)
    else
      out=$(env $KV timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-torch-arm --no-optimizer 2>/dev/null | tail -1)
    fi
    echo "$mode $(echo $out | python -c 'import sys,json; d=json.loads(sys.stdin.read()); print(d["ms_per_step"], d["value"], d["e2e"]["value"])')"
  done
done
