"""Golden vectors for the fused parameter update, produced by the REAL third-party implementation the reference calls
(torch.optim.Adam / AdamW, torch.nn.utils.clip_grad_norm_) plus the reference's own ``BaseModel.model_ema`` loop
(basicsr/models/base_model.py:86-95, imported from /root/reference when present, restated otherwise) on CPU.

    python tests/golden/make_golden_optim.py        ->  tests/golden/optim_step.npz
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

SHAPES = [(64, 3, 3, 3), (64,), (1, 64, 1, 1), (128, 64, 1, 1), (128,), (128, 1, 3, 3), (37,), (8193,), (3, 5, 7), (1,)]
CASES = {  # name -> (optimizer class name, kwargs, grad_clip, ema_decay, steps)
    "adamw_clip_ema": ("AdamW", dict(lr=1e-3, betas=(0.9, 0.9), weight_decay=1e-4), 0.01, 0.999, 3),
    "adam_l2": ("Adam", dict(lr=2e-4, betas=(0.9, 0.99), weight_decay=1e-3), None, 0.0, 2),
    "adamw_default_noclip": ("AdamW", dict(lr=3e-4), 10.0, 0.99, 2),   # max_norm above the norm: coefficient clamps to 1
}


def tensors(seed):
    g = torch.Generator().manual_seed(seed)
    params = [torch.randn(s, generator=g) * 0.1 for s in SHAPES]
    grads = [[torch.randn(s, generator=g) * (0.02 if i % 2 == 0 else 1e-4) for i, s in enumerate(SHAPES)] for _ in range(3)]
    return params, grads


def reference_model_ema():
    try:
        from oracle._ref_import import reference_scope
        with reference_scope():      # the function object outlives the scope; this process keeps its own `basicsr` afterwards
            from basicsr.models.base_model import BaseModel

        class _Bare:
            get_bare_model = staticmethod(lambda net: net)
        return lambda net, ema, decay: BaseModel.model_ema(type("M", (), {"net_g": net, "net_g_ema": ema,
                                                                         "get_bare_model": staticmethod(lambda n: n)})(), decay)
    except Exception:
        return None


def run_case(name):
    cls, kw, clip, decay, steps = CASES[name]
    params, grads = tensors(sum(map(ord, name)))
    net = torch.nn.ParameterList([torch.nn.Parameter(p.clone()) for p in params])
    ema = torch.nn.ParameterList([torch.nn.Parameter(p.clone()) for p in params])
    opt = getattr(torch.optim, cls)(net.parameters(), **kw)
    ref_ema = reference_model_ema()
    norms = []
    for t in range(steps):
        for p, g in zip(net, grads[t]):
            p.grad = g.clone()
        if clip:
            norms.append(float(torch.nn.utils.clip_grad_norm_(net.parameters(), clip)))   # sr_model.py:166-167
        opt.step()                                                                         # :169
        if decay > 0:                                                                      # :173-174
            if ref_ema is not None:
                ref_ema(net, ema, decay)
            else:
                for e, p in zip(ema.parameters(), net.parameters()):
                    e.data.mul_(decay).add_(p.data, alpha=1 - decay)
    out = {}
    for i, p in enumerate(net):
        st = opt.state[p]
        out[f"{name}|p{i}"] = p.detach().numpy()
        out[f"{name}|m{i}"] = st["exp_avg"].numpy()
        out[f"{name}|v{i}"] = st["exp_avg_sq"].numpy()
        out[f"{name}|e{i}"] = ema[i].detach().numpy()
    out[f"{name}|norms"] = np.asarray(norms, dtype=np.float64)
    return out


if __name__ == "__main__":
    torch.set_num_threads(1)
    arrays = {}
    for name in CASES:
        arrays.update(run_case(name))
    path = os.path.join(HERE, "optim_step.npz")
    np.savez_compressed(path, **arrays)
    print("wrote", path, os.path.getsize(path) // 1024, "KiB; reference model_ema used:", reference_model_ema() is not None)
