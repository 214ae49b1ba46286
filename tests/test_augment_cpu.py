"""dcpt_b200.augment (batched crop + flip / transpose by index gather) and dcpt_b200.dist.enlarged_indices against the reference's
own per-sample functions (basicsr/data/transforms.py:48-196, basicsr/data/data_sampler.py:8-48) on the same random draws."""
import os
import subprocess
import sys

import pytest
import torch

from oracle._ref_import import reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_crop_augment_properties():
    from dcpt_b200.augment import crop_augment_batch, draw_params
    g = torch.Generator().manual_seed(0)
    gt, lq = torch.rand(3, 3, 40, 48, generator=g), torch.rand(3, 3, 20, 24, generator=g)
    # no flip (code 2), no transpose (code 0): a plain window, GT window = LQ window * scale
    a, b, _ = crop_augment_batch(gt, lq, 16, scale=2, params=[(1, 2, 2, 0), (0, 0, 2, 0), (12, 16, 2, 0)])
    assert torch.equal(a[0], gt[0, :, 2:18, 4:20]) and torch.equal(b[0], lq[0, :, 1:9, 2:10])
    assert torch.equal(a[2], gt[2, :, 24:40, 32:48]) and torch.equal(b[2], lq[2, :, 12:20, 16:24])
    # horizontal flip, vertical flip, transpose
    a, b, _ = crop_augment_batch(gt, lq, 16, scale=2, params=[(1, 2, 0, 0), (1, 2, 1, 0), (1, 2, 2, 3)])
    assert torch.equal(a[0], gt[0, :, 2:18, 4:20].flip(-1)) and torch.equal(a[1], gt[1, :, 2:18, 4:20].flip(-2))
    assert torch.equal(a[2], gt[2, :, 2:18, 4:20].transpose(-1, -2)) and torch.equal(b[2], lq[2, :, 1:9, 2:10].transpose(-1, -2))
    import random
    ps = draw_params(1000, 20, 24, 8, random.Random(1))
    assert all(0 <= t <= 12 and 0 <= l <= 16 and f in (0, 1, 2) and r in (0, 1, 2, 3) for t, l, f, r in ps)
    with pytest.raises(ValueError, match="Scale mismatches"):
        crop_augment_batch(gt, lq[:, :, :19], 16, scale=2)
    with pytest.raises(ValueError, match="smaller than patch"):
        crop_augment_batch(gt, lq, 64, scale=2)


@pytest.mark.skipif(not reference_available(), reason="/root/reference is not present")
def test_against_the_reference_transforms_and_sampler():
    code = r'''
import json, random, sys
import numpy as np, torch
sys.path.insert(0, %r)
from dcpt_b200.augment import crop_augment_batch
from dcpt_b200.dist import enlarged_indices
from oracle._ref_import import import_reference
import_reference()
from basicsr.data import transforms as T
from basicsr.data.data_sampler import EnlargedSampler
rng = random.Random(5)
worst = 0.0
for scale, P, (H, W) in ((1, 32, (50, 61)), (2, 24, (40, 52)), (4, 32, (64, 48))):
    B = 6
    g = torch.Generator().manual_seed(scale)
    gt = torch.rand(B, 3, H * scale, W * scale, generator=g)
    lq = torch.rand(B, 3, H, W, generator=g)
    params = [(rng.randint(0, H - P // scale), rng.randint(0, W - P // scale), rng.randint(0, 2), rng.randint(0, 3)) for _ in range(B)]
    a, b, _ = crop_augment_batch(gt, lq, P, scale=scale, params=params)
    for i, (top, left, flip, rot) in enumerate(params):
        draws = iter([top, left, flip, rot])
        T.random.randint = lambda lo, hi: next(draws)            # the reference's own code, our draws
        g_np = np.ascontiguousarray(gt[i].permute(1, 2, 0).numpy())
        l_np = np.ascontiguousarray(lq[i].permute(1, 2, 0).numpy())
        g_c, l_c = T.paired_random_crop(g_np, l_np, P, scale)
        g_a, l_a = T.augment([np.ascontiguousarray(g_c), np.ascontiguousarray(l_c)])
        worst = max(worst, float(np.abs(a[i].permute(1, 2, 0).numpy() - g_a).max()), float(np.abs(b[i].permute(1, 2, 0).numpy() - l_a).max()))
same = True
for n, rep, ratio, ep in ((10, 2, 1, 0), (37, 4, 3, 7), (5, 8, 100, 2)):
    for r in range(rep):
        s = EnlargedSampler(list(range(n)), rep, r, ratio)
        s.set_epoch(ep)
        same = same and list(iter(s)) == enlarged_indices(n, rep, r, ratio, ep)
print("RESULT " + json.dumps({"worst": worst, "sampler": same}))
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    import json
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert res["worst"] == 0.0 and res["sampler"], res
