"""Golden vectors for the Restormer path, produced by running the REAL reference (``/root/reference``) on CPU.

Run in the build container only:  ``python tests/golden/make_golden_restormer.py``
The reference modules are built with their own constructors, then loaded (strict) with seeded weights from
``oracle.restormer_oracle.random_restormer_state_dict`` (which also proves key / shape equality), seeds recorded.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle._ref_import import import_reference  # noqa: E402
from oracle import restormer_oracle as RO  # noqa: E402


def npz(path, **arrays):
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v)) for k, v in arrays.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    import_reference()
    from basicsr.archs.restormer_arch import Restormer, TransformerBlock

    torch.set_num_threads(4)
    # ---- one TransformerBlock (MDTA + GDFN), fwd + autograd bwd -----------------------------------------
    for dim, heads, hw, lnt in ((48, 1, (16, 24), "BiasFree"), (96, 2, (8, 16), "BiasFree"), (32, 4, (8, 8), "WithBias")):
        torch.manual_seed(100 + dim)
        blk = TransformerBlock(dim, heads, 2.66, False, lnt)
        shapes = {k: tuple(p.shape) for k, p in blk.named_parameters()}
        g = torch.Generator().manual_seed(200 + dim)
        sd = {}
        for k, shp in shapes.items():
            if k.endswith("temperature"):
                sd[k] = 0.5 + torch.rand(shp, generator=g)
            elif "norm" in k and k.endswith("weight"):
                sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
            elif k.endswith("bias"):
                sd[k] = 0.1 * torch.randn(shp, generator=g)
            else:
                sd[k] = torch.randn(shp, generator=g) / (shp[1] * shp[2] * shp[3]) ** 0.5
        blk.load_state_dict(sd, strict=True)
        x = torch.randn(2, dim, *hw).requires_grad_(True)
        y = blk(x)
        dy = torch.randn_like(y)
        y.backward(dy)
        arrays = {"x": x, "y": y, "dy": dy, "dx": x.grad, "heads": heads, "seed": 100 + dim, "ln_bias": int(lnt != "BiasFree")}
        for k, p in blk.named_parameters():
            arrays["p." + k] = p
            arrays["g." + k] = p.grad
        npz(os.path.join(HERE, f"restormer_block_d{dim}.npz"), **arrays)

    # ---- tiny Restormer: output, hook=True decoder features, L1-loss gradients ----------------------------
    cfg = dict(dim=16, num_blocks=[1, 2, 1, 1], num_refinement_blocks=1, heads=[1, 2, 4, 8])
    net = Restormer(**cfg)
    sd = RO.random_restormer_state_dict(seed=7, **cfg)
    net.load_state_dict(sd, strict=True)      # proves the oracle's key / shape table against the reference
    assert list(sd.keys()) == [k for k, _ in net.named_parameters()]
    torch.manual_seed(8)
    inp, gt = torch.rand(2, 3, 32, 48), torch.rand(2, 3, 32, 48)
    out = net(inp)
    loss = (out - gt).abs().mean()
    loss.backward()
    arrays = {"inp": inp, "gt": gt, "out": out, "loss": loss, "seed": 7, "cfg_dim": 16, "cfg_blocks": [1, 2, 1, 1], "cfg_refine": 1,
              "cfg_heads": [1, 2, 4, 8]}
    for k, p in net.named_parameters():
        arrays["g." + k] = p.grad
    feats, hooks = [], []
    for name, m in net.named_modules():       # degradation_classification_pretrain_model.py:65-68 ('decoder', one dot)
        if "decoder" in name and name.count(".") == 1:
            hooks.append(m.register_forward_hook(lambda mod, i, o: feats.append(o)))
    with torch.no_grad():
        r = net(inp, hook=True)
    assert r is None and len(feats) == 3
    for i, f in enumerate(feats):
        arrays[f"feat{i}"] = f
    npz(os.path.join(HERE, "restormer_tiny.npz"), **arrays)


if __name__ == "__main__":
    main()
