mkdir -p gpurun_out/r01t
timeout 600 python tools/bench_configs.py --iters 10 --out gpurun_out/r01t/configs.json > gpurun_out/r01t/configs.log 2>&1; tail -40 gpurun_out/r01t/configs.log | grep -v INFO
timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 430 -c 8 -o gpurun_out/r01t/gemm_full -f python tools/prof_step.py 16 1 > gpurun_out/r01t/ncu_gemm.log 2>&1
ls -la gpurun_out/r01t
