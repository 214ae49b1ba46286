#!/bin/bash
# Round-end evidence pass (1 GPU): all GPU tests, smoke, bench line (+ optional ncu passes: `bash tools/gpu_final.sh tag ncu`).
TAG=${1:-final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log
tail -2 $OUT/pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; grep "\[smoke\]" $OUT/smoke.log
timeout 600 python bench.py --steps 10 --warmup 3 --breakdown --shapes > $OUT/bench.json 2> $OUT/bench.err; echo "bench exit $?"
cp gpurun_out/kernel_breakdown.tsv $OUT/ 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm exit $?"
tail -c 400 $OUT/bench_reference.json
if [ "$2" == "ncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python tools/prof_step.py 16 1 > $OUT/ncu_launch.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"ln_fwd_wide" --launch-skip 20 -c 2 -o $OUT/lnfwd_full -f python tools/prof_step.py 16 1 > $OUT/ncu_ln.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"adam_ema|grad_sqsum" --launch-skip 2 -c 2 -o $OUT/optim_full -f python tools/prof_optim.py > $OUT/ncu_opt.log 2>&1
fi
ls $OUT
