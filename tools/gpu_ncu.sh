#!/bin/bash
# ncu evidence pass (1 GPU): launch list of one fwd+bwd step and full captures of the heaviest kernels.
TAG=${1:-ncu}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches.csv python tools/prof_step.py 16 1 > $OUT/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel --launch-skip 420 -c 8 -o $OUT/gemm_full -f python tools/prof_step.py 16 1 > $OUT/ncu_gemm.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"ln_bwd_wide|dwgate_bwd_a|ln_fwd_kernel|dwconv_bwd_data|dwgate_fwd" --launch-skip 60 -c 6 -o $OUT/elt_full -f python tools/prof_step.py 16 1 > $OUT/ncu_elt.log 2>&1
ls -la $OUT
