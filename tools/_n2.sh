T="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
$T bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2_flat.json 2> gpurun_out/bench_n2_flat.err
tail -c 400 gpurun_out/bench_n2_flat.err
for f in gpurun_out/bench_n2_flat.json; do python -c "import sys,json; d=json.loads(open('$f').read().strip().splitlines()[-1]); print('$f', d['value'], d['ms_per_step'], d['e2e'])"; done
