mkdir -p gpurun_out/r02b
for pf in 0 1 2; do
  echo "=== DCPT_GEMM_PF=$pf" >> gpurun_out/r02b/trace.txt
  DCPT_GEMM_PF=$pf DCPT_LIB=dcpt_b200/libdcpt_sm100_trace.so timeout 120 python tools/gemm_trace.py 2>&1 | head -9 >> gpurun_out/r02b/trace.txt
  DCPT_GEMM_PF=$pf timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-torch-arm --no-optimizer 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('PF=$pf value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])" >> gpurun_out/r02b/trace.txt
done
cat gpurun_out/r02b/trace.txt
