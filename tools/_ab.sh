timeout 600 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_nafnet.py -x -q 2>&1 | tail -2
run() {
  python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-optimizer --breakdown --shapes 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$1', d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e']['ms_per_step'])"
  grep "^ln_fwd\|^scale_rows\|^sca_" gpurun_out/kernel_breakdown.tsv | head -9
}
run new
run new_again
