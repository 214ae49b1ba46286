#!/bin/bash
# One `ncu --set full` capture of the depthwise-stencil kernels of a training step at the L4 shape (16384 x 512) plus the
# per-kernel summary lines.  bash tools/ncu_stencil.sh <tag>
TAG=${1:-stencil}
OUT=gpurun_out/$TAG
mkdir -p $OUT
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"dwgate_fwd|dwgate_bwd_a|dwconv_bwd_data" \
  --launch-skip 30 --launch-count 18 -f -o $OUT/stencil python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-torch-arm \
  --no-optimizer > $OUT/ncu.log 2>&1
echo "ncu exit $?"
ncu -i $OUT/stencil.ncu-rep --page raw --csv > $OUT/stencil_raw.csv 2>/dev/null
python - <<PY
import csv
rows = list(csv.reader(open("$OUT/stencil_raw.csv")))
hdr = rows[0]
want = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__registers_per_thread", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "sm__inst_executed.avg.per_cycle_elapsed", "smsp__issue_active.avg.pct", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "sm__throughput.avg.pct_of_peak_sustained_elapsed", "launch__occupancy_limit_registers",
        "launch__occupancy_limit_shared_mem"]
idx = [hdr.index(w) if w in hdr else -1 for w in want]
for r in rows[2:]:
    print(" | ".join((r[i][:60] if i >= 0 else "-") for i in idx))
PY
