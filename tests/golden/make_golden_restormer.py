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


def dcpt_case():
    """The backbone side of DCPTModel.optimize_parameters with a Restormer (degradation_classification_pretrain_model.py:
    133-169): the hooked pass on lq (hook=True -> None, restormer_arch.py:403) whose decoder features receive the classifier's
    gradient.  The classifier is replaced by fixed functionals of the hooked features so that the golden gradients test
    exactly what the backbone must do with an injected gradient (the head has its own golden):
      gs.*  l = 0.5 * sum_i mean(feat_i^2)            (smooth gradient, every reachable parameter)
      gw.*  l = sum_i <feat_i, white noise / sqrt(n)>  (decoder_level1 parameters only: they see feat_2's gradient alone)."""
    import_reference()
    from basicsr.archs.restormer_arch import Restormer

    torch.set_num_threads(4)
    cfg = dict(dim=16, num_blocks=[1, 2, 1, 1], num_refinement_blocks=1, heads=[1, 2, 4, 8])
    net = Restormer(**cfg)
    net.load_state_dict(RO.random_restormer_state_dict(seed=7, **cfg), strict=True)
    g = torch.Generator().manual_seed(21)
    gt, lq = torch.rand(2, 3, 32, 48, generator=g), torch.rand(2, 3, 32, 48, generator=g)
    hook_outputs = []
    for name, m in net.named_modules():       # :65-68 with hook_names = 'decoder'
        if "decoder" in name and name.count(".") == 1:
            m.register_forward_hook(lambda mod, i, o: hook_outputs.append(o))
    assert net(lq, hook=True) is None         # :154
    assert len(hook_outputs) == 3
    (0.5 * sum((f ** 2).mean() for f in hook_outputs)).backward()
    arrays = {"seed": 7, "cfg_dim": 16, "cfg_blocks": [1, 2, 1, 1], "cfg_refine": 1, "cfg_heads": [1, 2, 4, 8]}
    for i, f in enumerate(hook_outputs):
        arrays[f"feat{i}"] = f                # gt, lq and the white noise are regenerated by the test from the same generator
    for k, p in net.named_parameters():       # refinement / output are never reached: no gradient at all
        assert (p.grad is None) == (k.startswith("refinement.") or k.startswith("output.")), k
        if p.grad is not None:
            arrays["gs." + k] = p.grad
    dfe = [torch.randn(f.shape, generator=g) / f.numel() ** 0.5 for f in hook_outputs]
    net.zero_grad(set_to_none=True)
    hook_outputs.clear()
    net(lq, hook=True)
    sum((f * d).sum() for f, d in zip(hook_outputs, dfe)).backward()
    for k, p in net.named_parameters():
        if k.startswith("decoder_level1."):
            arrays["gw." + k] = p.grad
    npz(os.path.join(HERE, "restormer_dcpt_tiny.npz"), **arrays)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "dcpt":
        dcpt_case()
    else:
        main()
        dcpt_case()
