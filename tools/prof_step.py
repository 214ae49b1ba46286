"""One NAFNet-w64 forward+backward (batch B) through the C-ABI engine — the command profiled under ncu.
    python tools/prof_step.py [B] [steps]
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from bench import CFG, H, W  # noqa: E402
from dcpt_b200.nafnet import NAFNetEngine  # noqa: E402
from oracle import nafnet_oracle as O  # noqa: E402  (synthetic weight generator only)

B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 1
dev = torch.device("cuda", 0)
sd = O.random_nafnet_state_dict(seed=0, **CFG)
params = [v.to(dev).contiguous() for v in sd.values()]
eng = NAFNetEngine(3, CFG["width"], CFG["middle_blk_num"], CFG["enc_blk_nums"], CFG["dec_blk_nums"])
g = torch.Generator(device=dev).manual_seed(1)
inp = torch.rand(B, 3, H, W, device=dev, generator=g)
gt = torch.rand(B, 3, H, W, device=dev, generator=g)
flat, grads = eng.alloc_flat_grads(params)
for _ in range(steps):
    flat.zero_()
    out, _, saved = eng.forward(params, inp)
    dout = torch.sign(out - gt) / inp.numel()
    eng.backward(params, inp, saved, dout, grads=grads)
torch.cuda.synchronize()
print("done", float(flat.abs().sum()))
