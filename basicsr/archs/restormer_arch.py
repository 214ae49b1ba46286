"""Restormer on the B200 hot path — drop-in for the reference's ``basicsr/archs/restormer_arch.py``.

Same class names, ctor kwargs, sub-module / parameter names and ``state_dict`` shapes as the reference
(restormer_arch.py:26-72 LayerNorm, :75-100 FeedForward, :103-145 Attention, :148-159 TransformerBlock,
:162-231 patch embed / resampling / SequentialTransformerBlock, :234-422 Restormer), so the reference's
``options/all_in_one/test/test_Restormer_5d.yml`` ``network_g`` section and checkpoints load with
``strict=True``.  The modules are parameter containers; ``Restormer.forward`` hands the whole network to the
sm_100a kernels through the C ABI (``dcpt_restormer_fwd`` for inference, ``dcpt_restormer_fwd_train`` /
``dcpt_restormer_bwd`` under autograd).
"""
import numbers

import torch
import torch.nn as nn

from basicsr.utils.registry import ARCH_REGISTRY
from dcpt_b200.lib import DcptError
from dcpt_b200.restormer import RestormerEngine, restormer_apply


def _trunc_normal_(w, std=0.02):  # arch_util.py:259-282 (the reference's own helper), a = -2, b = 2
    return nn.init.trunc_normal_(w, mean=0.0, std=std, a=-2.0, b=2.0)


class BiasFree_LayerNorm(nn.Module):
    def __init__(self, normalized_shape):
        super().__init__()
        if isinstance(normalized_shape, numbers.Integral):
            normalized_shape = (normalized_shape,)
        assert len(normalized_shape) == 1
        self.weight = nn.Parameter(torch.ones(torch.Size(normalized_shape)))
        self.normalized_shape = torch.Size(normalized_shape)


class WithBias_LayerNorm(nn.Module):
    def __init__(self, normalized_shape):
        super().__init__()
        if isinstance(normalized_shape, numbers.Integral):
            normalized_shape = (normalized_shape,)
        assert len(normalized_shape) == 1
        self.weight = nn.Parameter(torch.ones(torch.Size(normalized_shape)))
        self.bias = nn.Parameter(torch.zeros(torch.Size(normalized_shape)))
        self.normalized_shape = torch.Size(normalized_shape)


class LayerNorm(nn.Module):
    def __init__(self, dim, LayerNorm_type):
        super().__init__()
        self.body = BiasFree_LayerNorm(dim) if LayerNorm_type == "BiasFree" else WithBias_LayerNorm(dim)


class FeedForward(nn.Module):
    """GDFN parameters (restormer_arch.py:75-93); every conv is bias-free in the reference regardless of ``bias``."""

    def __init__(self, dim, ffn_expansion_factor, bias):
        super().__init__()
        hidden = int(dim * ffn_expansion_factor)
        self.project_in = nn.Conv2d(dim, hidden * 2, kernel_size=1, bias=False)
        self.dwconv = nn.Conv2d(hidden * 2, hidden * 2, kernel_size=3, stride=1, padding=1, groups=hidden * 2, bias=False)
        self.project_out = nn.Conv2d(hidden, dim, kernel_size=1, bias=False)


class Attention(nn.Module):
    """MDTA parameters (restormer_arch.py:103-119)."""

    def __init__(self, dim, num_heads, bias):
        super().__init__()
        self.num_heads = num_heads
        self.temperature = nn.Parameter(torch.ones(num_heads, 1, 1))
        self.qkv = nn.Conv2d(dim, dim * 3, kernel_size=1, bias=False)
        self.qkv_dwconv = nn.Conv2d(dim * 3, dim * 3, kernel_size=3, stride=1, padding=1, groups=dim * 3, bias=False)
        self.project_out = nn.Conv2d(dim, dim, kernel_size=1, bias=False)


class TransformerBlock(nn.Module):
    def __init__(self, dim, num_heads, ffn_expansion_factor, bias, LayerNorm_type):
        super().__init__()
        self.norm1 = LayerNorm(dim, LayerNorm_type)
        self.attn = Attention(dim, num_heads, bias)
        self.norm2 = LayerNorm(dim, LayerNorm_type)
        self.ffn = FeedForward(dim, ffn_expansion_factor, bias)


class OverlapPatchEmbed(nn.Module):
    def __init__(self, in_c=3, embed_dim=48, bias=False):
        super().__init__()
        self.proj = nn.Conv2d(in_c, embed_dim, kernel_size=3, stride=1, padding=1, bias=False)


class Downsample(nn.Module):
    def __init__(self, n_feat):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(n_feat, n_feat // 2, kernel_size=3, stride=1, padding=1, bias=False), nn.PixelUnshuffle(2))


class Upsample(nn.Module):
    def __init__(self, n_feat):
        super().__init__()
        self.body = nn.Sequential(nn.Conv2d(n_feat, n_feat * 2, kernel_size=3, stride=1, padding=1, bias=False), nn.PixelShuffle(2))


class SequentialTransformerBlock(nn.Module):
    def __init__(self, dim, head, num_block, ffn_expansion_factor=2.66, bias=False, LayerNorm_type="BiasFree"):
        super().__init__()
        self.body = nn.Sequential(*[TransformerBlock(dim, head, ffn_expansion_factor, bias, LayerNorm_type) for _ in range(num_block)])


@ARCH_REGISTRY.register()
class Restormer(nn.Module):
    """restormer_arch.py:234-422.  ``window_size`` is accepted and ignored by the arch, as in the reference."""

    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4, heads=[1, 2, 4, 8],
                 ffn_expansion_factor=2.66, bias=False, LayerNorm_type="BiasFree", dual_pixel_task=False, scale=1, window_size=8,
                 _plain_stages=False):
        super().__init__()
        if dual_pixel_task or scale != 1:
            raise DcptError("dual_pixel_task / scale > 1 are not on the hot path (no shipped config uses them)")
        a = (ffn_expansion_factor, bias, LayerNorm_type)
        # a stage is `SequentialTransformerBlock` (blocks under `.body`, restormer_arch.py:205-231) for Restormer and a plain
        # nn.Sequential of blocks for Restormer_origin (:444-470): same parameters in the same order, different state_dict keys
        if _plain_stages:
            stage = lambda d, h, n: nn.Sequential(*[TransformerBlock(d, h, *a) for _ in range(n)])   # noqa: E731
        else:
            stage = lambda d, h, n: SequentialTransformerBlock(d, h, n, *a)                          # noqa: E731
        self.patch_embed = OverlapPatchEmbed(inp_channels, dim)
        self.encoder_level1 = stage(dim, heads[0], num_blocks[0])
        self.down1_2 = Downsample(dim)
        self.encoder_level2 = stage(dim * 2, heads[1], num_blocks[1])
        self.down2_3 = Downsample(dim * 2)
        self.encoder_level3 = stage(dim * 4, heads[2], num_blocks[2])
        self.down3_4 = Downsample(dim * 4)
        self.latent = stage(dim * 8, heads[3], num_blocks[3])
        self.up4_3 = Upsample(dim * 8)
        self.reduce_chan_level3 = nn.Conv2d(dim * 8, dim * 4, kernel_size=1, bias=bias)
        self.decoder_level3 = stage(dim * 4, heads[2], num_blocks[2])
        self.up3_2 = Upsample(dim * 4)
        self.reduce_chan_level2 = nn.Conv2d(dim * 4, dim * 2, kernel_size=1, bias=bias)
        self.decoder_level2 = stage(dim * 2, heads[1], num_blocks[1])
        self.up2_1 = Upsample(dim * 2)
        self.decoder_level1 = stage(dim * 2, heads[0], num_blocks[0])
        self.refinement = stage(dim * 2, heads[0], num_refinement_blocks)
        self.dual_pixel_task = dual_pixel_task
        self.scale = scale
        self.output = nn.Conv2d(dim * 2, out_channels, kernel_size=3, stride=1, padding=1, bias=bias)
        if not _plain_stages:
            self.apply(self._init_weights)   # (Restormer_origin keeps torch's default init, as in the reference)
        self._cfg = dict(inp_channels=inp_channels, out_channels=out_channels, dim=dim, num_blocks=tuple(num_blocks),
                         num_refinement_blocks=num_refinement_blocks, heads=tuple(heads), ffn_expansion_factor=ffn_expansion_factor,
                         bias=bias, ln_with_bias=LayerNorm_type != "BiasFree")
        self._engine = None

    def _init_weights(self, m):  # restormer_arch.py:370-374
        if isinstance(m, nn.Conv2d):
            _trunc_normal_(m.weight, std=0.02)
            if m.bias is not None:
                nn.init.constant_(m.bias, 0)

    def train(self, mode=True):
        """(as NAFNetBaseline.train: no module-tree walk when the mode does not change)"""
        if self.training == bool(mode):
            return self
        return super().train(mode)

    def prepack(self):
        """(as NAFNetBaseline.prepack: re-pack the weights' operand images right after the optimizer step)"""
        params = list(self.parameters())
        if params and params[0].is_cuda:
            self.engine().packed_for([p.detach() for p in params])

    def engine(self):
        if self._engine is None:
            self._engine = RestormerEngine(**self._cfg)
        return self._engine

    def _hook_targets(self):
        """DCPT registers forward hooks on ``decoder_level{k}.body`` (one-dot rule) or, under DDP, on
        ``module.decoder_level{k}`` (degradation_classification_pretrain_model.py:60-68): serve both."""
        out = []
        for k in (3, 2, 1):
            st = getattr(self, f"decoder_level{k}")
            mods = (st, st.body) if hasattr(st, "body") else (st,)
            out.append([m for m in mods if len(m._forward_hooks) > 0])
        return out

    def forward(self, inp_img, hook=None):
        params = list(self.parameters())
        targets = self._hook_targets()
        want = any(len(t) > 0 for t in targets)
        if torch.is_grad_enabled() and any(p.requires_grad for p in params):
            # training step (SRModel.optimize_parameters, sr_model.py:132-174; DCPTModel.optimize_parameters,
            # degradation_classification_pretrain_model.py:133-169): autograd Function over the C-ABI fwd / bwd; the hooked
            # decoder features are differentiable outputs, so the classifier's gradient flows back into the backbone
            if not inp_img.is_cuda:
                raise DcptError("dcpt_b200 has no CPU path: input is on %s" % inp_img.device)
            dead = [i for i, (k, _) in enumerate(self.named_parameters()) if k.startswith(("refinement.", "output."))] if hook else ()
            out, feats = restormer_apply(self.engine(), inp_img, params, hook=bool(hook), want_feats=want, dead=dead)
            self._fire_hooks(targets, feats)
            return None if hook else out
        out, feats = self.engine().forward([p.detach() for p in params], inp_img, hook=bool(hook), want_feats=want)
        if want:
            self._fire_hooks(targets, [f.permute(0, 3, 1, 2) for f in feats])
        return None if hook else out

    @staticmethod
    def _fire_hooks(targets, feats):
        if not feats:
            return
        for mods, f in zip(targets, feats):
            for m in mods:
                for fn in list(m._forward_hooks.values()):
                    fn(m, (None,), f)


@ARCH_REGISTRY.register()
class Restormer_origin(Restormer):
    """restormer_arch.py:425-518 - the original Restormer: plain ``nn.Sequential`` stages (state_dict keys
    ``encoder_level1.0.attn...`` instead of ``encoder_level1.body.0.attn...``), ``WithBias`` LayerNorm by default, no ``scale`` /
    ``hook`` arguments, torch's default initialisation.  Same parameters in the same order as ``Restormer``, hence the same
    engine (dcpt_restormer_*) underneath."""

    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4, heads=[1, 2, 4, 8],
                 ffn_expansion_factor=2.66, bias=False, LayerNorm_type="WithBias", dual_pixel_task=False, window_size=8):
        super().__init__(inp_channels=inp_channels, out_channels=out_channels, dim=dim, num_blocks=num_blocks,
                         num_refinement_blocks=num_refinement_blocks, heads=heads, ffn_expansion_factor=ffn_expansion_factor, bias=bias,
                         LayerNorm_type=LayerNorm_type, dual_pixel_task=dual_pixel_task, scale=1, window_size=window_size,
                         _plain_stages=True)

    def forward(self, inp_img):   # :481 - no hook argument
        return super().forward(inp_img, hook=None)
