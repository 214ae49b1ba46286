"""Two fused parameter updates (clip + AdamW + EMA) over the NAFNet-w64 parameter set, for ncu captures."""
import os, sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from basicsr.archs import build_network
from dcpt_b200.optim import FusedAdamW
import bench as BM

net = build_network(dict(type="NAFNetBaseline", window_size=16, **BM.CFG)).cuda()
params = list(net.parameters())
ema = [p.detach().clone() for p in params]
for p in params:
    p.grad = torch.randn_like(p) * 1e-3
opt = FusedAdamW(params, lr=1e-4, betas=(0.9, 0.9), weight_decay=1e-4)
for _ in range(2):
    opt.step(grad_clip=0.01, ema_params=ema, ema_decay=0.999)
torch.cuda.synchronize()
