"""CPU oracle for the PromptIR network - TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).

Own-words functional restatement of ``/root/reference/basicsr/archs/promptir_arch.py`` (PromptGenBlock :238-263, PromptIR
:267-518) on plain torch CPU ops, built on the transformer block of ``oracle/restormer_oracle.py`` with the softmax attention
(:140).  Parameters travel as a mapping with the reference's ``state_dict`` keys.  Pinned against the reference module itself
by ``tests/test_oracle_promptir_cpu.py`` (live, when ``/root/reference`` is present) and by ``tests/golden/promptir_net.npz``
(made by ``tests/golden/make_golden_promptir.py`` from the real ``PromptIR`` class).  The reference ships no golden vectors
for this path."""
from __future__ import annotations

from typing import Dict, Sequence

import torch
import torch.nn.functional as F

from .restormer_oracle import _q, transformer_block

Tensor = torch.Tensor

# (prompt_dim, prompt_len, prompt_size, lin_dim) of prompt1, prompt2, prompt3 - literals in the reference (:290-299)
PROMPTS = ((64, 5, 64, 96), (128, 5, 32, 192), (320, 5, 16, 384))


def prompt_gen(x: Tensor, P: Dict[str, Tensor], name: str) -> Tensor:
    """PromptGenBlock.forward (:252-263)."""
    B, _, H, W = x.shape
    emb = x.mean(dim=(-2, -1))                                                                   # :254
    wts = F.softmax(F.linear(emb, P[name + ".linear_layer.weight"], P[name + ".linear_layer.bias"]), dim=1)   # :255
    prompt = (wts[:, :, None, None, None] * P[name + ".prompt_param"]).sum(dim=1)                # :256-259 ([B,L,1,1,1] * [1,L,D,S,S])
    prompt = F.interpolate(prompt, (H, W), mode="bilinear")                                      # :260
    return F.conv2d(_q(prompt), _q(P[name + ".conv3x3.weight"]), None, padding=1)                # :261


def _stage(x, P, name, n, heads):
    for j in range(n):                                                                           # plain nn.Sequential stages (:301-311 ...)
        x = transformer_block(x, P, f"{name}.{j}", heads, softmax=True)
    return x


def _conv3(x, P, name):
    return F.conv2d(_q(x), _q(P[name + ".body.0.weight"]), None, padding=1)


def _conv1(x, P, name):
    return F.conv2d(_q(x), _q(P[name + ".weight"]), P.get(name + ".bias"))


def promptir_fwd(inp: Tensor, P: Dict[str, Tensor], num_blocks: Sequence[int] = (4, 6, 6, 8), num_refinement_blocks: int = 4,
                 heads: Sequence[int] = (1, 2, 4, 8)) -> Tensor:
    """PromptIR.forward(inp_img, hook=False) with ``decoder=True`` (:465-518)."""
    x1 = F.conv2d(inp, P["patch_embed.proj.weight"], P.get("patch_embed.proj.bias"), padding=1)
    e1 = _stage(x1, P, "encoder_level1", num_blocks[0], heads[0])
    e2 = _stage(F.pixel_unshuffle(_conv3(e1, P, "down1_2"), 2), P, "encoder_level2", num_blocks[1], heads[1])
    e3 = _stage(F.pixel_unshuffle(_conv3(e2, P, "down2_3"), 2), P, "encoder_level3", num_blocks[2], heads[2])
    lat = _stage(F.pixel_unshuffle(_conv3(e3, P, "down3_4"), 2), P, "latent", num_blocks[3], heads[3])
    lat = torch.cat([lat, prompt_gen(lat, P, "prompt3")], 1)                                     # :480-481
    lat = transformer_block(lat, P, "noise_level3", heads[2], softmax=True)
    lat = _conv1(lat, P, "reduce_noise_level3")
    d3 = torch.cat([F.pixel_shuffle(_conv3(lat, P, "up4_3"), 2), e3], 1)                         # :485-487
    d3 = _stage(_conv1(d3, P, "reduce_chan_level3"), P, "decoder_level3", num_blocks[2], heads[2])
    d3 = torch.cat([d3, prompt_gen(d3, P, "prompt2")], 1)                                        # :491-494
    d3 = _conv1(transformer_block(d3, P, "noise_level2", heads[2], softmax=True), P, "reduce_noise_level2")
    d2 = torch.cat([F.pixel_shuffle(_conv3(d3, P, "up3_2"), 2), e2], 1)
    d2 = _stage(_conv1(d2, P, "reduce_chan_level2"), P, "decoder_level2", num_blocks[1], heads[1])
    d2 = torch.cat([d2, prompt_gen(d2, P, "prompt1")], 1)                                        # :502-505
    d2 = _conv1(transformer_block(d2, P, "noise_level1", heads[2], softmax=True), P, "reduce_noise_level1")
    d1 = torch.cat([F.pixel_shuffle(_conv3(d2, P, "up2_1"), 2), e1], 1)                          # :508-509
    d1 = _stage(d1, P, "decoder_level1", num_blocks[0], heads[0])
    r = _stage(d1, P, "refinement", num_refinement_blocks, heads[0])
    return F.conv2d(r, P["output.weight"], P.get("output.bias"), padding=1) + inp                # :514


def promptir_param_shapes(dim=48, num_blocks=(4, 6, 6, 8), num_refinement_blocks=4, heads=(1, 2, 4, 8), ffn_expansion_factor=2.66,
                          bias=False, LayerNorm_type="WithBias"):
    """Ordered {key: shape} in ``PromptIR.named_parameters()`` order (registration order of :284-462).  ``bias=True`` would
    also put biases on the convs INSIDE the blocks (promptir_arch.py:82-98, 114-127 - unlike the fork's Restormer); only the
    shipped ``bias=False`` is restated."""
    assert not bias, "PromptIR(bias=True) is not restated"
    shapes: Dict[str, tuple] = {}

    def block(prefix, d, h):
        def norm(n):
            shapes[f"{prefix}.{n}.body.weight"] = (d,)
            if LayerNorm_type != "BiasFree":
                shapes[f"{prefix}.{n}.body.bias"] = (d,)
        norm("norm1")
        shapes[prefix + ".attn.temperature"] = (h, 1, 1)
        shapes[prefix + ".attn.qkv.weight"] = (3 * d, d, 1, 1)
        shapes[prefix + ".attn.qkv_dwconv.weight"] = (3 * d, 1, 3, 3)
        shapes[prefix + ".attn.project_out.weight"] = (d, d, 1, 1)
        norm("norm2")
        hid = int(d * ffn_expansion_factor)
        shapes[prefix + ".ffn.project_in.weight"] = (2 * hid, d, 1, 1)
        shapes[prefix + ".ffn.dwconv.weight"] = (2 * hid, 1, 3, 3)
        shapes[prefix + ".ffn.project_out.weight"] = (d, hid, 1, 1)

    def stage(name, d, h, n):
        for j in range(n):
            block(f"{name}.{j}", d, h)

    def conv1(name, co, ci):
        shapes[name + ".weight"] = (co, ci, 1, 1)
        if bias:
            shapes[name + ".bias"] = (co,)

    shapes["patch_embed.proj.weight"] = (dim, 3, 3, 3)
    for i, (D, L, S, lin) in enumerate(PROMPTS):
        shapes[f"prompt{i + 1}.prompt_param"] = (1, L, D, S, S)
        shapes[f"prompt{i + 1}.linear_layer.weight"] = (L, lin)
        shapes[f"prompt{i + 1}.linear_layer.bias"] = (L,)
        shapes[f"prompt{i + 1}.conv3x3.weight"] = (D, D, 3, 3)
    stage("encoder_level1", dim, heads[0], num_blocks[0])
    shapes["down1_2.body.0.weight"] = (dim // 2, dim, 3, 3)
    stage("encoder_level2", dim * 2, heads[1], num_blocks[1])
    shapes["down2_3.body.0.weight"] = (dim, dim * 2, 3, 3)
    stage("encoder_level3", dim * 4, heads[2], num_blocks[2])
    shapes["down3_4.body.0.weight"] = (dim * 2, dim * 4, 3, 3)
    stage("latent", dim * 8, heads[3], num_blocks[3])
    shapes["up4_3.body.0.weight"] = (dim * 8, dim * 4, 3, 3)
    conv1("reduce_chan_level3", dim * 4, dim * 2 + 192)
    block("noise_level3", dim * 4 + 512, heads[2])
    conv1("reduce_noise_level3", dim * 4, dim * 4 + 512)
    stage("decoder_level3", dim * 4, heads[2], num_blocks[2])
    shapes["up3_2.body.0.weight"] = (dim * 8, dim * 4, 3, 3)
    conv1("reduce_chan_level2", dim * 2, dim * 4)
    block("noise_level2", dim * 2 + 224, heads[2])
    conv1("reduce_noise_level2", dim * 4, dim * 2 + 224)
    stage("decoder_level2", dim * 2, heads[1], num_blocks[1])
    shapes["up2_1.body.0.weight"] = (dim * 4, dim * 2, 3, 3)
    block("noise_level1", dim * 2 + 64, heads[2])
    conv1("reduce_noise_level1", dim * 2, dim * 2 + 64)
    stage("decoder_level1", dim * 2, heads[0], num_blocks[0])
    stage("refinement", dim * 2, heads[0], num_refinement_blocks)
    shapes["output.weight"] = (3, dim * 2, 3, 3)
    if bias:
        shapes["output.bias"] = (3,)
    return shapes


def random_promptir_state_dict(seed: int = 0, gain: float = 1.0, **cfg) -> Dict[str, Tensor]:
    """Seeded weights that exercise every branch (as random_restormer_state_dict): conv / linear weights ~ N(0, gain^2 / fan_in),
    LN weights ~ 1 + 0.1 N, biases ~ 0.1 N, temperature ~ U(0.5, 1.5), prompt components ~ U(0, 1) (the reference's own init)."""
    g = torch.Generator().manual_seed(seed)
    sd: Dict[str, Tensor] = {}
    for k, shp in promptir_param_shapes(**cfg).items():
        if k.endswith("temperature"):
            sd[k] = 0.5 + torch.rand(shp, generator=g)
        elif k.endswith("prompt_param"):
            sd[k] = torch.rand(shp, generator=g)
        elif ".norm" in k and k.endswith("weight"):
            sd[k] = 1.0 + 0.1 * torch.randn(shp, generator=g)
        elif k.endswith(".bias"):
            sd[k] = 0.1 * torch.randn(shp, generator=g)
        elif k.endswith("linear_layer.weight"):
            sd[k] = torch.randn(shp, generator=g) * (2.0 / shp[1] ** 0.5)
        else:
            fan_in = shp[1] * shp[2] * shp[3]
            sd[k] = torch.randn(shp, generator=g) * (gain / fan_in ** 0.5)
    return sd
