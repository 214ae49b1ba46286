"""Per-kernel breakdown of ONE eager forward + backward of the DC head at the C4 size (f = [64..512], 8 x 256 x 256 features)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as BM
from basicsr.archs import build_network
from dcpt_b200 import lib as L

lib = L.load_library()
dims = [64, 128, 256, 512]
head = build_network(dict(type="PromptIR_NoImg_DC", feature_dims=dims, num_res_blocks=2, num_classes=5)).cuda()
head.engine().use_graphs = False
feats = [torch.randn(8, c, 256 >> i, 256 >> i, device="cuda").requires_grad_(True) for i, c in enumerate(dims)]
lq = torch.rand(8, 3, 256, 256, device="cuda")
idx = torch.randint(0, 5, (8,), device="cuda")


def step():
    head.zero_grad(set_to_none=True)
    torch.nn.functional.cross_entropy(head(lq, feats), idx).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
lib.dcpt_prof_enable(2)
step()
torch.cuda.synchronize()
rows = BM.prof_table(lib)
lib.dcpt_prof_enable(0)
rows.sort(key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in rows)
print(f"{sum(r['launches'] for r in rows)} profiled launches, kernel time {tot:.2f} ms")
for r in rows[:16]:
    print(f"{r['tag'][:60]:60s} {r['launches']:4d} {r['ms']:7.3f} ms {r['ms'] / tot:6.1%} {r['flops'] / (r['ms'] * 1e-3 + 1e-12) / 1e12:7.1f} TF/s {r['bytes'] / (r['ms'] * 1e-3 + 1e-12) / 1e9:6.0f} GB/s")
