"""Full-resolution inference path around the networks: the reference's ``SRModel.pre_test`` /
``post_test`` / ``test_tile`` (basicsr/models/sr_model.py:244-361) with the tile loop BATCHED.

The reference runs ``net_g`` once per tile at batch 1 (a 1280 x 720 GoPro frame at ``infer_size`` 256 is 15 launch-bound
forwards).  A tile's padded crop has one of at most 4 x 4 shapes (per axis: first, interior, next-to-last with clipped padding,
last partial tile), so the tiles are grouped by crop shape and each group goes through the network in batches: identical
arithmetic per tile, hence identical output, a handful of forwards per image (and the engines' per-shape CUDA-graph cache sees
a bounded set of shapes).

    from dcpt_b200.tiling import patch_sr_model
    patch_sr_model(SRModel)          # SRModel.test_tile now batches; pre_test / post_test are unchanged reference code

or call ``tile_forward(net, lq, infer_size, tile_pad, scale)`` directly.  ``net`` is any callable NCHW -> NCHW module (the
sm_100a networks in production; any torch module in the CPU tests of the geometry).
"""
import math

import torch
import torch.nn.functional as F


def window_size_of(window_size):
    """``SRModel.check_window_size`` (sr_model.py:233-242): a list / tuple collapses to its maximum."""
    while isinstance(window_size, (tuple, list)):
        window_size = max(window_size)
    return window_size


def pre_test(lq, window_size):
    """Reflect-pad H, W up to multiples of ``window_size`` (sr_model.py:244-260).  Returns (padded, (mod_pad_h, mod_pad_w))."""
    ws = window_size_of(window_size)
    _, _, h, w = lq.shape
    ph = (ws - h % ws) % ws
    pw = (ws - w % ws) % ws
    return F.pad(lq, (0, pw, 0, ph), "reflect"), (ph, pw)


def post_test(output, mod_pad, scale=1):
    """Crop the padding off again (sr_model.py:262-271)."""
    ph, pw = mod_pad
    _, _, h, w = output.shape
    return output[:, :, 0:h - ph * scale, 0:w - pw * scale]


def tile_plan(height, width, infer_size, tile_pad):
    """The reference's tile geometry (sr_model.py:285-320), one entry per tile in its loop order:
    (input box incl. padding (y0, y1, x0, x1), input box without padding (y0, y1, x0, x1))."""
    tiles = []
    for y in range(math.ceil(height / infer_size)):
        for x in range(math.ceil(width / infer_size)):
            x0, y0 = x * infer_size, y * infer_size
            x1, y1 = min(x0 + infer_size, width), min(y0 + infer_size, height)
            xp0, xp1 = max(x0 - tile_pad, 0), min(x1 + tile_pad, width)
            yp0, yp1 = max(y0 - tile_pad, 0), min(y1 + tile_pad, height)
            tiles.append(((yp0, yp1, xp0, xp1), (y0, y1, x0, x1)))
    return tiles


@torch.no_grad()
def tile_forward(net, lq, infer_size, tile_pad, scale=1, max_batch=16):
    """``SRModel.test_tile`` (sr_model.py:273-361) with the per-tile forwards batched by crop shape.  lq: [B, C, H, W]."""
    b, c, height, width = lq.shape
    out = lq.new_zeros((b, c, height * scale, width * scale))
    groups = {}
    for t in tile_plan(height, width, infer_size, tile_pad):
        (yp0, yp1, xp0, xp1), _ = t
        groups.setdefault((yp1 - yp0, xp1 - xp0), []).append(t)
    per = max(1, max_batch // b)                       # tiles per forward (each tile carries the image batch)
    from .params import frozen_weights
    with frozen_weights():                             # nothing can change the weights inside this loop: one fingerprint check
        for tiles in groups.values():
            for i in range(0, len(tiles), per):
                chunk = tiles[i:i + per]
                inp = torch.cat([lq[:, :, yp0:yp1, xp0:xp1] for (yp0, yp1, xp0, xp1), _ in chunk], dim=0)
                res = net(inp)
                for k, ((yp0, yp1, xp0, xp1), (y0, y1, x0, x1)) in enumerate(chunk):
                    o = res[k * b:(k + 1) * b]
                    oy, ox = (y0 - yp0) * scale, (x0 - xp0) * scale
                    out[:, :, y0 * scale:y1 * scale, x0 * scale:x1 * scale] = o[:, :, oy:oy + (y1 - y0) * scale, ox:ox + (x1 - x0) * scale]
    return out


def patch_sr_model(sr_model_cls, max_batch=16):
    """Replace ``SRModel.test_tile`` by the batched loop (same attributes read and written: ``self.lq``, ``self.opt['tile']``,
    ``self.opt['scale']``, ``self.net_g`` / ``self.net_g_ema``, ``self.output``)."""
    def test_tile(self):
        ema = hasattr(self, "net_g_ema")
        net = self.net_g_ema if ema else self.net_g
        net.eval()
        self.output = tile_forward(net, self.lq, self.opt["tile"]["infer_size"], self.opt["tile"]["tile_pad"], self.opt["scale"], max_batch)
        if not ema:
            self.net_g.train()
    sr_model_cls.test_tile = test_tile
    return sr_model_cls
