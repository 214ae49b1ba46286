"""Per-kernel breakdown of ONE eager Restormer training step (fwd + L1 + bwd, batch 4 x 128 x 128) -> top kernels."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench as BM
from basicsr.archs import build_network
from dcpt_b200 import lib as L

lib = L.load_library()
net = build_network(dict(type="Restormer", window_size=8)).cuda().train()
eng = net.engine()
eng.use_graphs = eng.use_train_graphs = False
x, t = torch.rand(4, 3, 128, 128, device="cuda"), torch.rand(4, 3, 128, 128, device="cuda")


def step():
    net.zero_grad(set_to_none=True)
    torch.nn.functional.l1_loss(net(x), t).backward()


for _ in range(2):
    step()
torch.cuda.synchronize()
lib.dcpt_prof_enable(1)
step()
torch.cuda.synchronize()
rows = BM.prof_table(lib)
lib.dcpt_prof_enable(0)
rows.sort(key=lambda r: -r["ms"])
tot = sum(r["ms"] for r in rows)
print(f"{sum(r['launches'] for r in rows)} profiled launches, kernel time {tot:.2f} ms")
for r in rows[:14]:
    print(f"{r['tag'][:50]:50s} {r['launches']:4d} {r['ms']:7.3f} ms {r['ms'] / tot:6.1%}")
