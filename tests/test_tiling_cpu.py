"""Host logic of the full-resolution inference path (dcpt_b200/tiling.py) against the reference's own SRModel methods
(basicsr/models/sr_model.py:244-361) - geometry only, on CPU, with a small torch conv net standing in for the network."""
import json
import os
import subprocess
import sys

import pytest
import torch

from dcpt_b200 import tiling as T
from oracle._ref_import import reference_available

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _net():
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.GELU(), torch.nn.Conv2d(8, 3, 3, padding=1))
    return net.eval()


def _loop(net, lq, infer_size, tile_pad, scale=1):
    """Own-words restatement of the reference's per-tile loop (sr_model.py:273-361): one forward per tile, in order."""
    b, c, h, w = lq.shape
    out = lq.new_zeros(b, c, h * scale, w * scale)
    for (yp0, yp1, xp0, xp1), (y0, y1, x0, x1) in T.tile_plan(h, w, infer_size, tile_pad):
        with torch.no_grad():
            o = net(lq[:, :, yp0:yp1, xp0:xp1])
        oy, ox = (y0 - yp0) * scale, (x0 - xp0) * scale
        out[:, :, y0 * scale:y1 * scale, x0 * scale:x1 * scale] = o[:, :, oy:oy + (y1 - y0) * scale, ox:ox + (x1 - x0) * scale]
    return out


@pytest.mark.parametrize("h,w,infer,pad,batch", [(97, 130, 32, 8, 1), (64, 64, 32, 4, 2), (50, 45, 64, 16, 1), (128, 96, 32, 0, 3), (33, 200, 48, 12, 1)])
def test_batched_tiles_equal_the_per_tile_loop(h, w, infer, pad, batch):
    net = _net()
    lq = torch.rand(batch, 3, h, w, generator=torch.Generator().manual_seed(h + w))
    a = T.tile_forward(net, lq, infer, pad, 1, max_batch=8)
    b = _loop(net, lq, infer, pad)
    assert torch.allclose(a, b, atol=1e-6), float((a - b).abs().max())
    # at most 4 x 4 crop shapes, every output pixel written exactly once
    shapes = {(t[0][1] - t[0][0], t[0][3] - t[0][2]) for t in T.tile_plan(h, w, infer, pad)}
    assert len(shapes) <= 16    # per axis: first, interior, next-to-last (clipped padding), last (partial) tile
    cover = torch.zeros(h, w)
    for _, (y0, y1, x0, x1) in T.tile_plan(h, w, infer, pad):
        cover[y0:y1, x0:x1] += 1
    assert bool((cover == 1).all())


def test_pre_post_test_roundtrip():
    lq = torch.rand(1, 3, 100, 120)
    padded, pads = T.pre_test(lq, 16)
    assert tuple(padded.shape) == (1, 3, 112, 128) and pads == (12, 8)
    assert torch.equal(padded[:, :, :100, :120], lq) and torch.equal(padded[:, :, 100:, :120], lq[:, :, 87:99, :].flip(2))   # reflect
    assert torch.equal(T.post_test(padded, pads), lq)
    same, pads0 = T.pre_test(torch.rand(1, 3, 64, 32), [8, 16])
    assert tuple(same.shape) == (1, 3, 64, 32) and pads0 == (0, 0)


@pytest.mark.skipif(not reference_available(), reason="/root/reference is not present")
def test_against_the_reference_sr_model_methods():
    """The reference's SRModel.pre_test / test_tile / post_test executed on a bare object, vs dcpt_b200.tiling (child process:
    importing the reference replaces the `basicsr` package)."""
    code = r'''
import json, sys, types, torch
sys.path.insert(0, %r)
from dcpt_b200 import tiling as T
from oracle._ref_import import import_reference
import_reference()
from basicsr.models.sr_model import SRModel
torch.manual_seed(0)
net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.GELU(), torch.nn.Conv2d(8, 3, 3, padding=1)).eval()
res = {}
for (h, w, infer, pad, ws) in [(97, 130, 32, 8, 16), (50, 45, 64, 16, 8), (128, 96, 32, 0, 16)]:
    lq = torch.rand(1, 3, h, w)
    m = types.SimpleNamespace(lq=lq.clone(), net_g=net, scale=1, opt={"network_g": {"window_size": ws}, "scale": 1, "tile": {"infer_size": infer, "tile_pad": pad}})
    m.check_window_size = lambda wss: SRModel.check_window_size(m, wss)
    SRModel.pre_test(m)
    padded, pads = T.pre_test(lq, ws)
    ok_pad = torch.equal(m.lq, padded) and (m.mod_pad_h, m.mod_pad_w) == pads
    SRModel.test_tile(m)
    mine = T.tile_forward(net, padded, infer, pad, 1, max_batch=6)
    ok_tile = torch.allclose(m.output, mine, atol=1e-6)
    SRModel.post_test(m)
    ok_post = torch.allclose(m.output, T.post_test(mine, pads), atol=1e-6) and tuple(m.output.shape) == (1, 3, h, w)
    res["%%dx%%d" %% (h, w)] = [ok_pad, ok_tile, ok_post]
print("RESULT " + json.dumps(res))
''' % ROOT
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert r.returncode == 0, r.stderr[-3000:]
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert all(all(v) for v in res.values()), res
