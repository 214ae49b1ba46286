"""DC head (PromptIR_NoImg_DC, f = [64,128,256,512], 8 x 256 x 256 decoder features) forward + backward alone: device ms, wall ms, and
the library's launch count - is the classifier part of the C4 step bound by the GPU or by host launches?"""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from basicsr.archs import build_network
from dcpt_b200 import lib as L

lib = L.load_library()
dims = [64, 128, 256, 512]
head = build_network(dict(type="PromptIR_NoImg_DC", feature_dims=dims, num_res_blocks=2, num_classes=5)).cuda()
B = int(os.getenv("BATCH", "8"))
feats = [torch.randn(B, c, 256 >> i, 256 >> i, device="cuda").requires_grad_(True) for i, c in enumerate(dims)]
lq = torch.rand(B, 3, 256, 256, device="cuda")
idx = torch.randint(0, 5, (B,), device="cuda")


def step():
    for f in feats:
        f.grad = None
    head.zero_grad(set_to_none=True)
    loss = torch.nn.functional.cross_entropy(head(lq, feats), idx)
    loss.backward()
    return loss


for _ in range(3):
    step()
torch.cuda.synchronize()
n0 = lib.dcpt_launch_count()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
t0 = time.perf_counter()
a.record()
for _ in range(10):
    step()
b.record()
t_enq = (time.perf_counter() - t0) * 100
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 100
print(f"DC head fwd+bwd b{B}: device {a.elapsed_time(b) / 10:.2f} ms, host enqueue {t_enq:.2f} ms, wall {wall:.2f} ms, "
      f"library launches per step {(lib.dcpt_launch_count() - n0) // 10}")
