#!/bin/bash
# Quick GPU iteration pass: selected tests (-k expr), then a short bench line.  bash tools/gpu_quick.sh <tag> "<pytest -k expr>" [bench-env...]
TAG=${1:-q}
OUT=gpurun_out/$TAG
mkdir -p $OUT
K="$2"
if [ -n "$K" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q -k "$K" > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
  tail -15 $OUT/pytest.log
fi
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-torch-arm --no-optimizer --breakdown --shapes > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cp gpurun_out/kernel_breakdown.tsv $OUT/ 2>/dev/null
python - <<PY
import json
try:
    d = json.loads(open("$OUT/bench.json").read().strip().splitlines()[-1])
    print("value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], "launches", d["gpu_launches"])
except Exception as e:
    print("bench parse failed", e)
    print(open("$OUT/bench.err").read()[-2000:])
PY
