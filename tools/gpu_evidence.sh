#!/bin/bash
# Round-2 evidence pass (1 GPU): tests, smoke, bench (+ reference arm), ncu launch list, ncu --set full of the GEMM shapes (+ traffic
# json), compute-sanitizer on the op tests, other configs.   bash tools/gpu_evidence.sh <tag>
TAG=${1:-r02_final}
OUT=gpurun_out/$TAG
mkdir -p $OUT
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $OUT/smi.txt 2>&1
timeout 1500 python -m pytest tests -m gpu -q > $OUT/pytest.log 2>&1; echo "pytest exit $?" >> $OUT/pytest.log; tail -3 $OUT/pytest.log
timeout 300 python __graft_entry__.py smoke > $OUT/smoke.log 2>&1; echo "smoke exit $?"; grep "\[smoke\]" $OUT/smoke.log
timeout 900 python bench.py --breakdown > $OUT/bench_n1.json 2> $OUT/bench_n1.err; echo "bench exit $?"
cp gpurun_out/kernel_breakdown.tsv $OUT/ 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $OUT/bench_reference.json 2> $OUT/bench_reference.err; echo "reference arm exit $?"
DCPT_OPERAND=fp16 timeout 600 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-torch-arm --no-optimizer > $OUT/bench_n1_fp16.json 2> $OUT/bench_n1_fp16.err; echo "fp16 bench exit $?"
timeout 600 python tools/bench_configs.py > $OUT/other_configs.json 2> $OUT/other_configs.err; echo "configs exit $?"
timeout 300 python tools/gemm_bench.py > $OUT/gemm_bench.txt 2>&1
if [ "$2" != "noncu" ]; then
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/ncu_launches.csv python tools/prof_step.py 16 1 > $OUT/ncu_launch.log 2>&1; echo "ncu launches exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -o $OUT/gemm_full -f python tools/gemm_bench.py --ncu > $OUT/ncu_full.log 2>&1; echo "ncu full exit $?"
cp gpurun_out/gemm_bench_manifest.json $OUT/ 2>/dev/null
fi
if [ "$3" != "nosan" ]; then
K='(fused_layernorm and 130) or (fused_layernorm_backward and 130) or (gate32 and 130) or (gemm_store and 130-72-40 and tcgen05)'
for tool in memcheck racecheck synccheck; do
  timeout 420 compute-sanitizer --tool $tool --error-exitcode 9 python -m pytest tests/test_gpu_kernels.py -m gpu -q -x -k "$K" > $OUT/sanitizer_$tool.log 2>&1; echo "sanitizer $tool exit $?"
  tail -4 $OUT/sanitizer_$tool.log
done
fi
ls -la $OUT
