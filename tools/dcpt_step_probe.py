"""Host-side timeline of one DCPTModel.optimize_parameters iteration at the C4 size (N = 1): wall time of each phase as the
host sees it (enqueue only - no synchronisation inside the step) and the device time of the whole step.  Where the host takes
longer than the device needs, the GPU is waiting for Python."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from collections import OrderedDict
from contextlib import ExitStack
from basicsr.models import build_model

adamw = {"type": "AdamW", "lr": 3e-4, "weight_decay": 1e-4, "betas": [0.9, 0.9]}
opt = {"name": "c4", "model_type": "DCPTModel", "scale": 1, "num_gpu": 1, "dist": False, "is_train": True, "rank": 0, "world_size": 1,
       "hook_names": "decoder", "path": {"pretrain_network_g": None},
       "network_g": dict(type="NAFNetBaseline", width=64, enc_blk_nums=[1, 1, 1, 28], middle_blk_num=1, dec_blk_nums=[1, 1, 1, 1], window_size=16),
       "network_dc": dict(type="PromptIR_NoImg_DC", feature_dims=[64, 128, 256, 512], num_res_blocks=2, num_classes=5),
       "train": {"optim_g": dict(adamw), "optim_dc": dict(adamw), "scheduler": {"type": "MultiStepLR", "milestones": [10 ** 6], "gamma": 0.5},
                 "pixel_opt": {"type": "L1Loss", "loss_weight": 1.0, "reduction": "mean"}, "classify_opt": {"type": "CrossEntropyLoss", "loss_weight": 1.0}}}
torch.manual_seed(0)
m = build_model(opt)
g = torch.Generator().manual_seed(0)
gt, lq = torch.rand(8, 3, 256, 256, generator=g).pin_memory(), torch.rand(8, 3, 256, 256, generator=g).pin_memory()
idx = torch.randint(0, 5, (8,), generator=g).pin_memory()
T = OrderedDict()


def mark(name, t0):
    T[name] = T.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
    return time.perf_counter()


def step(prof):
    t = time.perf_counter()
    m.feed_data({"lq": lq, "gt": gt, "dataset_idx": idx}); t = mark("feed_data", t) if prof else t
    m.net_g.train(); m.net_dc.eval(); m.optimizer_g.zero_grad(); t = mark("zero_grad_g", t) if prof else t
    pix = m.net_g(m.gt, hook=False); m.hook_outputs = []; t = mark("fwd_pixel", t) if prof else t
    l_pix = m.cri_pixel(pix, m.gt); t = mark("loss_pix", t) if prof else t
    m.net_dc.train(); m.optimizer_dc.zero_grad(); t = mark("zero_grad_dc", t) if prof else t
    m.net_g(m.lq, hook=True); t = mark("fwd_hooked", t) if prof else t
    cls = m.net_dc(m.lq, m.hook_outputs[::-1]); t = mark("fwd_head", t) if prof else t
    l_cls = m.cri_classify(cls, m.dataset_idx); t = mark("loss_cls", t) if prof else t
    (l_pix + l_cls).backward(); t = mark("backward", t) if prof else t
    m.optimizer_g.step(); t = mark("opt_g", t) if prof else t
    m.optimizer_dc.step(); t = mark("opt_dc", t) if prof else t
    m.prepack(m.net_g, m.net_dc); t = mark("prepack", t) if prof else t
    m.hook_outputs = []
    log = m.reduce_loss_dict(OrderedDict(l_pix=l_pix, l_classify=l_cls)); t = mark("log_sync", t) if prof else t
    return log


for _ in range(4):
    step(False)
torch.cuda.synchronize()
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
N = 10
for _ in range(N):
    step(True)
b.record()
torch.cuda.synchronize()
print("device ms per step %.2f" % (a.elapsed_time(b) / N))
print("host ms per phase: " + ", ".join(f"{k} {v / N:.2f}" for k, v in T.items()) + f" | sum {sum(T.values()) / N:.2f}")
