"""Generate golden vectors by running the REAL reference (``/root/reference``) on CPU.

Run in the build container only:  ``python tests/golden/make_golden.py``
Writes small fp32 ``.npz`` fixtures next to this file.  They pin ``oracle/`` to
the reference (tests/test_oracle_cpu.py) and travel to the GPU box, where
``/root/reference`` does not exist.

Weights: the reference modules' own default init, then beta/gamma ~ N(0,0.3)
and LN affine perturbed (zero-init beta/gamma makes every NAFBlock an identity,
nafnet_arch.py:162-163), seeds recorded in each file.
"""
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from oracle._ref_import import import_reference  # noqa: E402


def perturb(module, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for k, p in module.named_parameters():
            if k.endswith("beta") or k.endswith("gamma"):
                p.copy_(torch.randn(p.shape, generator=g) * 0.3)
            elif "norm" in k and k.endswith("weight"):
                p.copy_(1.0 + 0.1 * torch.randn(p.shape, generator=g))
            elif "norm" in k and k.endswith("bias"):
                p.copy_(0.1 * torch.randn(p.shape, generator=g))
            elif k.endswith("temperature"):
                p.copy_(0.5 + torch.rand(p.shape, generator=g))


def npz(path, **arrays):
    np.savez_compressed(path, **{k: (v.detach().numpy() if torch.is_tensor(v) else np.asarray(v))
                                 for k, v in arrays.items()})
    print("wrote", path, os.path.getsize(path) // 1024, "KiB")


def main():
    import_reference()
    from basicsr.archs.nafnet_arch import LayerNorm2d, NAFBlock, NAFNetBaseline

    torch.manual_seed(0)
    torch.set_num_threads(4)

    # ---- LayerNorm2d fwd + the reference's custom backward -----------------
    ln = LayerNorm2d(24)
    perturb(ln, 1)
    with torch.no_grad():
        ln.weight.copy_(1.0 + 0.2 * torch.randn(24))
        ln.bias.copy_(0.2 * torch.randn(24))
    x = (torch.randn(2, 24, 5, 7) * 1.5 + 0.3).requires_grad_(True)
    y = ln(x)
    dy = torch.randn_like(y)
    y.backward(dy)
    npz(os.path.join(HERE, "layernorm2d.npz"), x=x, weight=ln.weight, bias=ln.bias, y=y, dy=dy,
        dx=x.grad, dweight=ln.weight.grad, dbias=ln.bias.grad, seed=0)

    # ---- NAFBlock fwd + bwd -------------------------------------------------
    for c, hw in ((16, (12, 10)), (64, (8, 8))):
        torch.manual_seed(10 + c)
        blk = NAFBlock(c)
        perturb(blk, 2 + c)
        x = torch.randn(2, c, *hw).requires_grad_(True)
        y = blk(x)
        dy = torch.randn_like(y)
        y.backward(dy)
        arrays = {"x": x, "y": y, "dy": dy, "dx": x.grad, "seed": 10 + c}
        for k, p in blk.named_parameters():
            arrays["p." + k] = p
            arrays["g." + k] = p.grad
        npz(os.path.join(HERE, f"nafblock_c{c}.npz"), **arrays)

    # ---- NAFNetBaseline (tiny) fwd + L1-loss bwd, and hook=True features ----
    cfg = dict(width=8, enc_blk_nums=[1, 1, 2], middle_blk_num=1, dec_blk_nums=[1, 1, 1])
    torch.manual_seed(3)
    net = NAFNetBaseline(**cfg)
    perturb(net, 4)
    inp = torch.rand(2, 3, 32, 48)
    gt = torch.rand(2, 3, 32, 48)
    out = net(inp)
    loss = (out - gt).abs().mean()          # L1Loss, mean reduction (losses/basic_loss.py:57-86)
    loss.backward()
    arrays = {"inp": inp, "gt": gt, "out": out, "loss": loss, "seed": 3,
              "cfg_width": 8, "cfg_enc": [1, 1, 2], "cfg_mid": 1, "cfg_dec": [1, 1, 1]}
    for k, p in net.named_parameters():
        arrays["p." + k] = p
        arrays["g." + k] = p.grad
    feats = []
    hooks = []
    for name, m in net.named_modules():       # degradation_classification_pretrain_model.py:65-68
        if "decoder" in name and name.count(".") == 1:
            hooks.append(m.register_forward_hook(lambda mod, i, o: feats.append(o)))
    with torch.no_grad():
        r = net(inp, hook=True)
    assert r is None and len(feats) == 3
    for i, f in enumerate(feats):
        arrays[f"feat{i}"] = f
    npz(os.path.join(HERE, "nafnet_w8.npz"), **arrays)

    # ---- degradation-classifier head (PromptIR_NoImg_DC): logits + CE-loss gradients -------------------
    from basicsr.archs.degrad_classify_arch import PromptIR_NoImg_DC
    torch.manual_seed(5)
    dims = [8, 16, 32]
    head = PromptIR_NoImg_DC(feature_dims=dims, num_res_blocks=2, num_classes=5)
    perturb(head, 6)
    with torch.no_grad():
        head.mixing_weights.copy_(1.0 + 0.3 * torch.randn(3))
    feats = [torch.randn(2, c, 32 >> i, 48 >> i).requires_grad_(True) for i, c in enumerate(dims)]
    logits = head(None, list(feats))
    labels = torch.tensor([1, 3])
    loss = torch.nn.functional.cross_entropy(logits, labels)   # CrossEntropyLoss (losses/basic_loss.py:39-55)
    loss.backward()
    arrays = {"logits": logits, "labels": labels, "loss": loss, "dims": dims, "seed": 5}
    for i, f in enumerate(feats):
        arrays[f"feat{i}"] = f
        arrays[f"dfeat{i}"] = f.grad
    for k, p in head.named_parameters():
        arrays["p." + k] = p
        arrays["g." + k] = p.grad
    npz(os.path.join(HERE, "dchead.npz"), **arrays)


if __name__ == "__main__":
    main()
