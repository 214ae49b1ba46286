"""Compact per-launch summary of an `ncu --set full` report (read here, no GPU):
    python tools/ncu_summary.py gpurun_out/x.ncu-rep [out.tsv]
"""
import csv
import io
import subprocess
import sys

KEYS = [
    ("dur_us", "gpu__time_duration.sum"),
    ("grid", "launch__grid_size"),
    ("regs", "launch__registers_per_thread"),
    ("dram_rd_MB", "dram__bytes_read.sum"),
    ("dram_wr_MB", "dram__bytes_write.sum"),
    ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
    ("l2_pct", "lts__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
    ("tensor_pct_rt", "TPC.TriageCompute.sm__pipe_tensor_cycles_active_realtime.avg.pct_of_peak_sustained_elapsed"),
    ("sm_pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed"),
    ("warps_act_pct", "sm__warps_active.avg.pct_of_peak_sustained_active"),
    ("sm_mhz", "sm__cycles_elapsed.avg.per_second"),
]


def main():
    rep = sys.argv[1]
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    out = ["\t".join(["id", "kernel"] + [k for k, _ in KEYS])]
    for r in data:
        name = r[col["Kernel Name"]][:70]
        vals = []
        for k, m in KEYS:
            if m not in col:
                vals.append("-")
                continue
            v, u = r[col[m]].replace(",", ""), units[col[m]]
            try:
                f = float(v)
                if k == "dur_us":
                    f = f / 1e3 if u in ("ns", "nsecond") else (f * 1e3 if u.startswith("ms") else f)
                if k.endswith("_MB"):
                    f = {"byte": 1e-6, "Kbyte": 1e-3, "Mbyte": 1.0, "Gbyte": 1e3}.get(u, 1e-6) * f
                if k == "sm_mhz":
                    f = {"hz": 1e-6, "Khz": 1e-3, "Mhz": 1.0, "Ghz": 1e3}.get(u, 1.0) * f
                vals.append(f"{f:.2f}")
            except ValueError:
                vals.append(v)
        out.append("\t".join([r[col["ID"]], name] + vals))
    s = "\n".join(out)
    print(s)
    if len(sys.argv) > 2:
        open(sys.argv[2], "w").write(s + "\n")


if __name__ == "__main__":
    main()
