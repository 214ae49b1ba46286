"""CPU oracle for the Restormer hot path — TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Own-words functional restatement of ``/root/reference/basicsr/archs/restormer_arch.py`` (MDTA / GDFN
transformer block and the 4-level U-Net around it) on plain ``torch`` CPU ops.  Parameters travel as a
``state_dict``-style mapping with the reference's key names, so weights are shared with the reference
module (``load_state_dict(strict=True)``) and with the CUDA path.  The reference trains this network through
autograd (no hand-written backward), so gradients of the oracle are taken with ``torch.autograd`` too.

Pinned against the reference itself: ``tests/golden/restormer_*.npz`` are outputs of the real reference
modules (``tests/golden/make_golden.py``); ``tests/test_oracle_cpu.py`` checks this file against them and
against the live reference when ``/root/reference`` is present.  The reference ships no golden vectors for
this path.
"""
from __future__ import annotations

from typing import Dict, List, Optional, Sequence

import torch
import torch.nn.functional as F

Tensor = torch.Tensor

# Rounding hook (as in nafnet_oracle.py): the CUDA path stores branch tensors as bf16.  Tests that want to separate
# "operand rounding" from "wrong math" pass q = lambda t: t.bfloat16().float(), which rounds at the same points the
# kernels do (LN output, qkv / dw-conv outputs, the folded attention-projection matrix, GDFN intermediates, conv inputs
# and packed weights).  q = None (default) is the exact fp32 reference math.
_Q = None


class rounding:
    """Context manager: ``with rounding(q): ...`` runs the oracle with the rounding hook q."""

    def __init__(self, q):
        self.q = q

    def __enter__(self):
        global _Q
        self.prev, _Q = _Q, self.q

    def __exit__(self, *a):
        global _Q
        _Q = self.prev


def _q(t: Tensor) -> Tensor:
    return t if _Q is None else _Q(t)


# --------------------------------------------------------------------------
# LayerNorm over channels — restormer_arch.py:26-72
# --------------------------------------------------------------------------
def layernorm_chan(x: Tensor, weight: Tensor, bias: Optional[Tensor], eps: float = 1e-6) -> Tensor:
    """Per-pixel LayerNorm over dim 1 of NCHW (the reference goes through to_3d/to_4d, :18-23, :70-72).

    ``bias is None`` = BiasFree_LayerNorm (:38-40): the variance is taken about the mean (biased) but the mean
    is NOT subtracted in the numerator.  Otherwise WithBias_LayerNorm (:56-59)."""
    C = x.shape[1]
    var = x.var(dim=1, keepdim=True, unbiased=False)
    w = weight.view(1, C, 1, 1)
    if bias is None:
        return x / torch.sqrt(var + eps) * w
    mu = x.mean(dim=1, keepdim=True)
    return (x - mu) / torch.sqrt(var + eps) * w + bias.view(1, C, 1, 1)


def _ln(x: Tensor, P: Dict[str, Tensor], prefix: str) -> Tensor:
    return _q(layernorm_chan(x, P[prefix + ".body.weight"], P.get(prefix + ".body.bias")))


# --------------------------------------------------------------------------
# MDTA — restormer_arch.py:103-145
# --------------------------------------------------------------------------
def mdta(x: Tensor, P: Dict[str, Tensor], prefix: str, heads: int, softmax: bool = False) -> Tensor:
    """Transposed (channel) attention with ReLU instead of softmax (:135-136).  ``softmax=True``: the PromptIR transformer
    blocks' variant (promptir_arch.py:108-149: identical but for ``attn.softmax(dim=-1)`` at :140)."""
    b, c, h, w = x.shape
    qkv = _q(F.conv2d(x, _q(P[prefix + ".qkv.weight"]), P.get(prefix + ".qkv.bias")))                 # :124 1x1, d -> 3d
    qkv = _q(F.conv2d(qkv, P[prefix + ".qkv_dwconv.weight"], P.get(prefix + ".qkv_dwconv.bias"), padding=1,
                      groups=3 * c))                                                                   # :124 dw 3x3
    q, k, v = qkv.chunk(3, dim=1)                                                                      # :125
    ch = c // heads
    q = q.reshape(b, heads, ch, h * w)                                                                 # :127-129
    k = k.reshape(b, heads, ch, h * w)
    v = v.reshape(b, heads, ch, h * w)
    q = F.normalize(q, dim=-1)                                                                         # :131-132 (eps 1e-12)
    k = F.normalize(k, dim=-1)
    attn = (q @ k.transpose(-2, -1)) * P[prefix + ".temperature"].view(1, heads, 1, 1)                 # :134
    attn = attn.softmax(dim=-1) if softmax else F.relu(attn)                                           # :136 / promptir_arch.py:140
    if _Q is not None:  # the kernels fold attn into the projection: W_eff = W_out * blockdiag(attn), rounded to bf16
        wo = P[prefix + ".project_out.weight"].reshape(c, heads, ch)                                   # [o, head, i]
        weff = _q(torch.einsum("ohi,bhij->bohj", wo, attn).reshape(b, c, c))
        return torch.einsum("boj,bjp->bop", weff, v.reshape(b, c, h * w)).reshape(b, c, h, w)
    out = (attn @ v).reshape(b, c, h, w)                                                               # :138-142
    return F.conv2d(out, P[prefix + ".project_out.weight"], P.get(prefix + ".project_out.bias"))       # :144


# --------------------------------------------------------------------------
# GDFN — restormer_arch.py:75-100
# --------------------------------------------------------------------------
def gdfn(x: Tensor, P: Dict[str, Tensor], prefix: str) -> Tensor:
    y = _q(F.conv2d(x, _q(P[prefix + ".project_in.weight"]), P.get(prefix + ".project_in.bias")))      # :96
    hid2 = y.shape[1]
    y = F.conv2d(y, P[prefix + ".dwconv.weight"], P.get(prefix + ".dwconv.bias"), padding=1, groups=hid2)  # :97
    x1, x2 = y.chunk(2, dim=1)
    y = _q(F.gelu(x1) * x2)                                                                            # :98 (exact erf GELU)
    return F.conv2d(y, _q(P[prefix + ".project_out.weight"]), P.get(prefix + ".project_out.bias"))     # :99


def transformer_block(x: Tensor, P: Dict[str, Tensor], prefix: str, heads: int, softmax: bool = False) -> Tensor:
    """restormer_arch.py:156-159 (``softmax=True``: promptir_arch.py:171-186, the same block around the softmax attention)."""
    x = x + mdta(_ln(x, P, prefix + ".norm1"), P, prefix + ".attn", heads, softmax)
    x = x + gdfn(_ln(x, P, prefix + ".norm2"), P, prefix + ".ffn")
    return x


def _stage(x: Tensor, P: Dict[str, Tensor], name: str, nblk: int, heads: int) -> Tensor:
    for j in range(nblk):                                                                              # SequentialTransformerBlock :217-231
        x = transformer_block(x, P, f"{name}.body.{j}", heads)
    return x


def _down(x, P, name):   # Downsample :175-188: 3x3 conv C -> C/2 (no bias) + PixelUnshuffle(2)
    return F.pixel_unshuffle(F.conv2d(_q(x), _q(P[name + ".body.0.weight"]), None, padding=1), 2)


def _up(x, P, name):     # Upsample :191-202: 3x3 conv C -> 2C (no bias) + PixelShuffle(2)
    return F.pixel_shuffle(F.conv2d(_q(x), _q(P[name + ".body.0.weight"]), None, padding=1), 2)


def restormer_fwd(inp: Tensor, P: Dict[str, Tensor], num_blocks: Sequence[int] = (4, 6, 6, 8), num_refinement_blocks: int = 4,
                  heads: Sequence[int] = (1, 2, 4, 8), hook: bool = False, return_feats: bool = False):
    """Restormer.forward (restormer_arch.py:376-422), scale == 1, dual_pixel_task == False.

    ``hook`` truthy -> stops after decoder_level1 and returns None (:403); ``return_feats`` additionally returns the
    three decoder-level outputs [decoder_level3, decoder_level2, decoder_level1] (what DCPT's forward hooks on
    ``decoder_level{k}.body`` capture, degradation_classification_pretrain_model.py:60-68)."""
    x1 = F.conv2d(inp, P["patch_embed.proj.weight"], P.get("patch_embed.proj.bias"), padding=1)        # :377
    e1 = _stage(x1, P, "encoder_level1", num_blocks[0], heads[0])
    e2 = _stage(_down(e1, P, "down1_2"), P, "encoder_level2", num_blocks[1], heads[1])
    e3 = _stage(_down(e2, P, "down2_3"), P, "encoder_level3", num_blocks[2], heads[2])
    lat = _stage(_down(e3, P, "down3_4"), P, "latent", num_blocks[3], heads[3])
    d3 = torch.cat([_up(lat, P, "up4_3"), e3], 1)                                                      # :389-390
    d3 = F.conv2d(_q(d3), _q(P["reduce_chan_level3.weight"]), P.get("reduce_chan_level3.bias"))
    d3 = _stage(d3, P, "decoder_level3", num_blocks[2], heads[2])
    d2 = torch.cat([_up(d3, P, "up3_2"), e2], 1)
    d2 = F.conv2d(_q(d2), _q(P["reduce_chan_level2.weight"]), P.get("reduce_chan_level2.bias"))
    d2 = _stage(d2, P, "decoder_level2", num_blocks[1], heads[1])
    d1 = torch.cat([_up(d2, P, "up2_1"), e1], 1)                                                       # :398-399 (no reduce)
    d1 = _stage(d1, P, "decoder_level1", num_blocks[0], heads[0])
    feats = [d3, d2, d1]
    if hook:
        return (None, feats) if return_feats else None
    r = _stage(d1, P, "refinement", num_refinement_blocks, heads[0])
    out = F.conv2d(r, P["output.weight"], P.get("output.bias"), padding=1) + inp                       # :413
    return (out, feats) if return_feats else out


# --------------------------------------------------------------------------
# synthetic parameters with the reference's keys / shapes (restormer_arch.py:236-368)
# --------------------------------------------------------------------------
def restormer_param_shapes(inp_channels=3, out_channels=3, dim=48, num_blocks=(4, 6, 6, 8), num_refinement_blocks=4,
                           heads=(1, 2, 4, 8), ffn_expansion_factor=2.66, bias=False, LayerNorm_type="BiasFree"):
    """Ordered {key: shape} in the reference's ``named_parameters()`` order."""
    shapes: Dict[str, tuple] = {}

    def conv(name, co, ci, k, b=bias, groups=1):
        shapes[name + ".weight"] = (co, ci // groups, k, k)
        if b:
            shapes[name + ".bias"] = (co,)

    def block(prefix, d, h):
        shapes[prefix + ".norm1.body.weight"] = (d,)
        if LayerNorm_type != "BiasFree":
            shapes[prefix + ".norm1.body.bias"] = (d,)
        shapes[prefix + ".attn.temperature"] = (h, 1, 1)
        # the reference's Attention / FeedForward ignore their `bias` argument: every conv is bias=False (:109-119, :81-93)
        conv(prefix + ".attn.qkv", 3 * d, d, 1, b=False)
        conv(prefix + ".attn.qkv_dwconv", 3 * d, 3 * d, 3, b=False, groups=3 * d)
        conv(prefix + ".attn.project_out", d, d, 1, b=False)
        shapes[prefix + ".norm2.body.weight"] = (d,)
        if LayerNorm_type != "BiasFree":
            shapes[prefix + ".norm2.body.bias"] = (d,)
        hid = int(d * ffn_expansion_factor)
        conv(prefix + ".ffn.project_in", 2 * hid, d, 1, b=False)
        conv(prefix + ".ffn.dwconv", 2 * hid, 2 * hid, 3, b=False, groups=2 * hid)
        conv(prefix + ".ffn.project_out", d, hid, 1, b=False)

    def stage(name, d, h, n):
        for j in range(n):
            block(f"{name}.body.{j}", d, h)

    conv("patch_embed.proj", dim, inp_channels, 3, b=False)
    stage("encoder_level1", dim, heads[0], num_blocks[0])
    conv("down1_2.body.0", dim // 2, dim, 3, b=False)
    stage("encoder_level2", dim * 2, heads[1], num_blocks[1])
    conv("down2_3.body.0", dim, dim * 2, 3, b=False)
    stage("encoder_level3", dim * 4, heads[2], num_blocks[2])
    conv("down3_4.body.0", dim * 2, dim * 4, 3, b=False)
    stage("latent", dim * 8, heads[3], num_blocks[3])
    conv("up4_3.body.0", dim * 16, dim * 8, 3, b=False)
    conv("reduce_chan_level3", dim * 4, dim * 8, 1)
    stage("decoder_level3", dim * 4, heads[2], num_blocks[2])
    conv("up3_2.body.0", dim * 8, dim * 4, 3, b=False)
    conv("reduce_chan_level2", dim * 2, dim * 4, 1)
    stage("decoder_level2", dim * 2, heads[1], num_blocks[1])
    conv("up2_1.body.0", dim * 4, dim * 2, 3, b=False)
    stage("decoder_level1", dim * 2, heads[0], num_blocks[0])
    stage("refinement", dim * 2, heads[0], num_refinement_blocks)
    conv("output", out_channels, dim * 2, 3)
    return shapes


def random_restormer_state_dict(seed: int = 0, gain: float = 1.0, **cfg) -> Dict[str, Tensor]:
    """Seeded weights that exercise every branch: conv weights ~ N(0, gain^2 / fan_in) (the reference's own
    trunc_normal_(std=0.02) init, :370-374, leaves both residual branches ~1e-3 of the stream and would hide
    errors), LN weights ~ 1 + 0.1 N, temperature ~ U(0.5, 1.5)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for k, shp in restormer_param_shapes(**cfg).items():
        if k.endswith("temperature"):
            sd[k] = 0.5 + torch.rand(shp, generator=g)
        elif ".norm" in k and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            sd[k] = torch.randn(shp, generator=g) * (gain / fan_in ** 0.5)
    return sd
