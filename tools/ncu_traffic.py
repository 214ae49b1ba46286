"""profiles/ncu_traffic.json from an `ncu --set full` capture of `tools/gemm_bench.py --ncu` (the number in the
bench line's `roofline.traffic` must be reproducible from the newest capture, by script):

    ncu --set full --clock-control none --import-source on -k regex:gemm_tc_kernel -o gpurun_out/<tag>/gemm_full python tools/gemm_bench.py --ncu
    python tools/ncu_traffic.py gpurun_out/<tag>/gemm_full.ncu-rep gpurun_out/gemm_bench_manifest.json [profiles/ncu_traffic.json]

gemm_bench.py --ncu launches each shape ONCE (cold L2) in manifest order; this script pairs the
capture's gemm_tc_kernel launches with the manifest and writes, per "<bench tag> <MxNxK>", dram__bytes_read.sum +
dram__bytes_write.sum of the launch (plus duration, tensor-pipe and DRAM utilisation for the record)."""
import csv
import io
import json
import os
import subprocess
import sys


def main():
    rep, man_path = sys.argv[1], sys.argv[2]
    out_path = sys.argv[3] if len(sys.argv) > 3 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles",
                                                                  "ncu_traffic.json")
    man = json.load(open(man_path))
    txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}
    data = [r for r in data if "gemm_tc_kernel" in r[col["Kernel Name"]]]

    def val(r, m, scale):
        v, u = float(r[col[m]].replace(",", "")), units[col[m]]
        return v * scale.get(u, 1.0)
    B = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    T = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3}
    per = len(data) // len(man)               # launches per manifest entry (bench() runs fn() once in --ncu mode)
    assert per >= 1 and per * len(man) == len(data), f"{len(data)} gemm launches in the capture vs {len(man)} manifest entries"
    res = {"_how": f"dram__bytes_read.sum + dram__bytes_write.sum per launch, `ncu --set full --clock-control none` of tools/gemm_bench.py --ncu "
                   f"({os.path.basename(rep)}; cold L2, one launch per shape), written by tools/ncu_traffic.py. Keys: '<bench.py kernel tag> <MxNxK>'."}
    for i, m in enumerate(man):
        r = data[i * per + per - 1]
        res[f"{m['tag']} {m['shape']}"] = {
            "bytes": round(val(r, "dram__bytes_read.sum", B) + val(r, "dram__bytes_write.sum", B)),
            "dram_read": round(val(r, "dram__bytes_read.sum", B)), "dram_write": round(val(r, "dram__bytes_write.sum", B)),
            "dur_us": round(val(r, "gpu__time_duration.sum", T), 2),
            "tensor_pct": float(r[col["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]]) if
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed" in col else None,
            "dram_pct": float(r[col["gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"]]),
            "note": m["note"]}
    json.dump(res, open(out_path, "w"), indent=1)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
