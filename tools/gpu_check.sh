#!/bin/bash
# One GPU-box pass: parity tests, bench (+ per-kernel breakdown), ncu launch list and one full capture of the top kernels.
# Usage (from the repo root, under gpurun): bash tools/gpu_check.sh [tag]
TAG=${1:-run}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -3 $OUT/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --shapes > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cp gpurun_out/kernel_breakdown.tsv $OUT/ 2>/dev/null
tail -1 $OUT/bench.json
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $OUT/launches.csv python tools/prof_step.py 16 1 > $OUT/ncu_launch.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 400 -c 6 -o $OUT/gemm_full -f python tools/prof_step.py 16 1 > $OUT/ncu_full.log 2>&1
fi
ls -la $OUT
