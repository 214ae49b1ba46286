mkdir -p gpurun_out/r02k
torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/dp_overlap_check.py 2>&1 | tail -3
for ov in 1 0; do
DCPT_DP_OVERLAP=$ov torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 3 --no-optimizer > gpurun_out/r02k/bench_n2_ov$ov.json 2> gpurun_out/r02k/bench_n2_ov$ov.err
python - <<PY
import json
try:
    d=json.loads(open("gpurun_out/r02k/bench_n2_ov$ov.json").read().strip().splitlines()[-1]); print("overlap=$ov value", d["value"], "ms", d["ms_per_step"], "e2e", d["e2e"]["value"], d["e2e"]["ms_per_step"])
except Exception as e:
    print("fail", e); print(open("gpurun_out/r02k/bench_n2_ov$ov.err").read()[-1500:])
PY
done
python bench.py --steps 20 --warmup 3 --no-optimizer --no-cpu-baseline --no-torch-arm 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e']['value'])"
