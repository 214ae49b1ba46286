"""Per-kernel breakdown (CUDA events around every library launch, dcpt_prof_*) of ONE eager Restormer forward on 128 x 128 tiles:
python tools/restormer_breakdown.py [batch]   -> gpurun_out/restormer_breakdown_b<batch>.tsv"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as BM
from basicsr.archs import build_network
from dcpt_b200 import lib as L

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
lib = L.load_library()
net = build_network(dict(type="Restormer", window_size=8)).cuda().eval()
net.engine().use_graphs = False
x = torch.rand(B, 3, 128, 128, device="cuda")
with torch.no_grad():
    for _ in range(3):
        net(x)
    torch.cuda.synchronize()
    lib.dcpt_prof_enable(2)
    net(x)
    torch.cuda.synchronize()
    rows = BM.prof_table(lib)
    lib.dcpt_prof_enable(0)
rows.sort(key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in rows)
os.makedirs("gpurun_out", exist_ok=True)
with open(f"gpurun_out/restormer_breakdown_b{B}.tsv", "w") as f:
    f.write("tag\tlaunches\tms\tshare\tTFLOP/s\tGB/s\n")
    for r in rows:
        f.write(f"{r['tag']}\t{r['launches']}\t{r['ms']:.3f}\t{r['ms'] / tot:.4f}\t{r['flops'] / (r['ms'] * 1e-3 + 1e-12) / 1e12:.1f}\t"
                f"{r['bytes'] / (r['ms'] * 1e-3 + 1e-12) / 1e9:.0f}\n")
print(f"b{B}: {sum(r['launches'] for r in rows)} profiled launches, kernel time {tot:.2f} ms")
for r in rows[:22]:
    print(f"{r['tag'][:58]:58s} {r['launches']:4d} {r['ms']:7.3f} ms {r['ms'] / tot:6.1%}")
