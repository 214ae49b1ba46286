"""Pin the oracle (oracle/nafnet_oracle.py) to the reference.

1. against the committed golden vectors (outputs of the real reference, made by
   tests/golden/make_golden.py) — always runs;
2. against the live reference when /root/reference is present (build container).
Tolerances: fp32 CPU, same ATen kernels, different op grouping -> 2e-5 relative.
"""
import os

import numpy as np
import pytest
import torch

from oracle import nafnet_oracle as O

RTOL = 2e-5


def rel(a, b):
    a = torch.as_tensor(a).double()
    b = torch.as_tensor(b).double()
    return float((a - b).norm() / (b.norm() + 1e-30))


def load(golden_dir, name):
    z = np.load(os.path.join(golden_dir, name))
    return {k: torch.from_numpy(np.asarray(z[k])) for k in z.files}


def test_layernorm_golden(golden_dir):
    z = load(golden_dir, "layernorm2d.npz")
    y, y_hat, var = O.layernorm2d_fwd(z["x"], z["weight"], z["bias"])
    assert rel(y, z["y"]) < RTOL
    dx, dw, db = O.layernorm2d_bwd(z["dy"], y_hat, var, z["weight"])
    assert rel(dx, z["dx"]) < RTOL
    assert rel(dw, z["dweight"]) < RTOL
    assert rel(db, z["dbias"]) < RTOL
    # autograd through the plain-op forward agrees with the hand-written backward
    x = z["x"].clone().requires_grad_(True)
    O.layernorm2d_fwd(x, z["weight"], z["bias"])[0].backward(z["dy"])
    assert rel(x.grad, z["dx"]) < RTOL


@pytest.mark.parametrize("c", [16, 64])
def test_nafblock_golden(golden_dir, c):
    z = load(golden_dir, f"nafblock_c{c}.npz")
    sd = {k[2:]: v.clone().requires_grad_(True) for k, v in z.items() if k.startswith("p.")}
    assert {k: tuple(v.shape) for k, v in sd.items()} == O.NAFBLOCK_PARAM_SHAPES(c)
    x = z["x"].clone().requires_grad_(True)
    y = O.nafblock_fwd(x, sd)
    assert rel(y, z["y"]) < RTOL
    y.backward(z["dy"])
    assert rel(x.grad, z["dx"]) < RTOL
    for k, v in sd.items():
        assert rel(v.grad, z["g." + k]) < 5e-5, k


def test_nafnet_golden(golden_dir):
    z = load(golden_dir, "nafnet_w8.npz")
    enc, mid, dec = z["cfg_enc"].tolist(), int(z["cfg_mid"]), z["cfg_dec"].tolist()
    sd = {k[2:]: v for k, v in z.items() if k.startswith("p.")}
    keys = O.nafnet_state_dict_keys(int(z["cfg_width"]), enc, mid, dec)
    assert [k for k, _ in keys] == [k[2:] for k in z if k.startswith("p.")]
    assert all(tuple(sd[k].shape) == s for k, s in keys)
    out, loss, grads = O.nafnet_fwd_bwd(z["inp"], z["gt"], sd, enc, mid, dec)
    assert rel(out, z["out"]) < RTOL
    assert abs(float(loss) - float(z["loss"])) < 1e-6
    for k in sd:
        assert rel(grads[k], z["g." + k]) < 2e-4, k
    feats = []
    with torch.no_grad():
        r = O.nafnet_fwd(z["inp"], sd, enc, mid, dec, hook=True, decoder_feats=feats)
    assert r is None and len(feats) == 3
    for i, f in enumerate(feats):
        assert rel(f, z[f"feat{i}"]) < RTOL


def test_rounding_hook_is_small(golden_dir):
    """bf16 rounding points (the CUDA path's storage precision) move the block
    output by O(bf16 eps), not more — guards the hook placement."""
    z = load(golden_dir, "nafblock_c64.npz")
    sd = {k[2:]: v for k, v in z.items() if k.startswith("p.")}
    y = O.nafblock_fwd(z["x"], sd)
    yq = O.nafblock_fwd(z["x"], sd, q=lambda t: t.bfloat16().float())
    assert 1e-5 < rel(yq, y) < 1e-2


@pytest.mark.skipif(not os.path.isdir("/root/reference/basicsr"), reason="reference not mounted")
def test_oracle_vs_live_reference():
    import subprocess
    import sys
    code = r'''
import sys, torch
sys.path.insert(0, %r)
from oracle._ref_import import import_reference
import_reference()
from basicsr.archs.nafnet_arch import NAFNetBaseline
from oracle import nafnet_oracle as O
torch.manual_seed(7)
cfg = dict(width=8, enc_blk_nums=[2, 1], middle_blk_num=2, dec_blk_nums=[1, 2])
net = NAFNetBaseline(**cfg)
sd0 = O.random_nafnet_state_dict(8, [2, 1], 2, [1, 2], seed=5)
net.load_state_dict(sd0, strict=True)
inp, gt = torch.rand(1, 3, 16, 24), torch.rand(1, 3, 16, 24)
out = net(inp); loss = (out - gt).abs().mean(); loss.backward()
o2, l2, g2 = O.nafnet_fwd_bwd(inp, gt, sd0, [2, 1], 2, [1, 2])
rel = lambda a, b: float((a.double() - b.double()).norm() / (b.double().norm() + 1e-30))
assert rel(o2, out) < 2e-5, rel(o2, out)
for k, p in net.named_parameters():
    assert rel(g2[k], p.grad) < 2e-4, (k, rel(g2[k], p.grad))
print("OK")
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "OK" in r.stdout, r.stdout + r.stderr


def test_dchead_golden(golden_dir):
    from oracle import dchead_oracle as D
    z = load(golden_dir, "dchead.npz")
    dims = z["dims"].tolist()
    sd = {k[2:]: v.clone().requires_grad_(True) for k, v in z.items() if k.startswith("p.")}
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == D.dchead_param_shapes(dims, 2, 5)
    feats = [z[f"feat{i}"].clone().requires_grad_(True) for i in range(len(dims))]
    logits = D.dchead_fwd(feats, sd)
    assert rel(logits, z["logits"]) < RTOL
    loss = torch.nn.functional.cross_entropy(logits, z["labels"])
    assert abs(float(loss) - float(z["loss"])) < 1e-6
    loss.backward()
    for i, f in enumerate(feats):
        assert rel(f.grad, z[f"dfeat{i}"]) < 1e-4, i
    for k, v in sd.items():
        assert rel(v.grad, z["g." + k]) < 2e-4, k


def test_nafnet_tlc_golden(golden_dir):
    """oracle `tlc_avgpool` / `tlc_kernels` (NAFNet = Local_Base + NAFNetBaseline) against the real reference's output."""
    import sys
    sys.path.insert(0, golden_dir)
    from make_golden_tlc import tlc_state_dict
    z = load(golden_dir, "nafnet_tlc_w16.npz")
    cfg = dict(width=16, enc_blk_nums=[1, 1], middle_blk_num=1, dec_blk_nums=[1, 1])
    sd = tlc_state_dict(cfg)
    ts = tuple(int(v) for v in z["train_size"])
    with torch.no_grad():
        out = O.nafnet_fwd(z["inp"], sd, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"], tlc_train_size=ts)
        small = O.nafnet_fwd(z["small"], sd, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"], tlc_train_size=ts)
        base = O.nafnet_fwd(z["inp"], sd, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"])
    assert rel(out, z["out"]) < RTOL and rel(small, z["out_small"]) < RTOL
    assert rel(base, z["out"]) > 5e-3          # the fixture separates local from global pooling
    assert O.tlc_kernels((1, 3, 128, 128), 5) == [(192, 192), (96, 96), (48, 48), (24, 24), (12, 12)]


def test_dchead_img_golden(golden_dir):
    """PromptIR_DC (conv_embed 7x7 stride 2 + LayerNorm on lq, degrad_classify_arch.py:480-556): the oracle against logits /
    gradients of the reference's own module (tests/golden/make_golden_dchead_img.py)."""
    from oracle import dchead_oracle as D
    z = load(golden_dir, "dchead_img.npz")
    dims = z["dims"].tolist()
    sd = {k[2:]: v.clone().requires_grad_(True) for k, v in z.items() if k.startswith("p.")}
    assert [(k, tuple(v.shape)) for k, v in sd.items()] == D.dchead_param_shapes(dims, 2, 5, img_embed=True)
    feats = [z[f"feat{i}"].clone().requires_grad_(True) for i in range(len(dims))]
    logits = D.dchead_fwd(feats, sd, lq=z["lq"])
    assert rel(logits, z["logits"]) < RTOL
    torch.nn.functional.cross_entropy(logits, z["labels"]).backward()
    for i, f in enumerate(feats):
        assert rel(f.grad, z[f"dfeat{i}"]) < 1e-4
    for k, v in sd.items():
        assert rel(v.grad, z["g." + k]) < 2e-4, k
