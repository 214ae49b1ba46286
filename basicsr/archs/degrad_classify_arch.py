"""Degradation-classifier head on the B200 hot path — drop-in for the reference's
``basicsr/archs/degrad_classify_arch.py`` (``PromptIR_NoImg_DC`` :558-641, the "ResNet18 decoder" of the DCPT paper).

Same registry name, ctor kwargs, sub-module / parameter names and shapes (``bottleneck_layers.{l}.{j}.conv{1,2,3}.weight``,
``….conv{k}.norm.{weight,bias}``, ``downsample_layers.{l}.0.weight``, ``last_stage.{j}.…``, ``mixing_weights``, ``fc.*``), so
checkpoints load with ``strict=True``.  Modules are parameter containers; ``forward`` runs the whole head through the
sm_100a kernels (``dcpt_b200/dchead.py``).  No CPU path.
"""
import torch
import torch.nn as nn

from basicsr.utils.registry import ARCH_REGISTRY
from dcpt_b200.dchead import DCHeadEngine, dchead_apply
from dcpt_b200.lib import DcptError


class LayerNorm(nn.Module):
    """channels_first LayerNorm parameters (:17-44): weight/bias (C,), eps 1e-6."""

    def __init__(self, normalized_shape, eps=1e-6, data_format="channels_first"):
        super().__init__()
        if data_format != "channels_first":
            raise DcptError("only the channels_first LayerNorm is on the hot path")
        self.weight = nn.Parameter(torch.ones(normalized_shape))
        self.bias = nn.Parameter(torch.zeros(normalized_shape))
        self.eps = eps


class Conv2d(nn.Conv2d):
    """Conv + norm container (:69-103): ``self.norm`` holds the LayerNorm so the keys read ``convK.norm.weight``."""

    def __init__(self, *args, norm=None, activation=None, **kwargs):
        super().__init__(*args, **kwargs)
        self.norm = norm
        self.activation = activation


class BottleneckBlock(nn.Module):
    """Parameter layout of the reference BottleneckBlock with norm="LN" and in == out channels (:132-225)."""

    def __init__(self, in_channels, out_channels, *, bottleneck_channels, norm="LN"):
        super().__init__()
        if in_channels != out_channels or norm != "LN":
            raise DcptError("the DC head only builds in == out channel bottlenecks with LN (PromptIR_NoImg_DC :577-613)")
        self.in_channels, self.out_channels, self.stride = in_channels, out_channels, 1
        self.shortcut = None
        self.conv1 = Conv2d(in_channels, bottleneck_channels, kernel_size=1, bias=False, norm=LayerNorm(bottleneck_channels))
        self.conv2 = Conv2d(bottleneck_channels, bottleneck_channels, kernel_size=3, padding=1, bias=False,
                            norm=LayerNorm(bottleneck_channels))
        self.conv3 = Conv2d(bottleneck_channels, out_channels, kernel_size=1, bias=False, norm=LayerNorm(out_channels))
        for layer in (self.conv1, self.conv2, self.conv3):      # c2_msra_fill (:211-213)
            nn.init.kaiming_normal_(layer.weight, mode="fan_out", nonlinearity="relu")


def _make_stage(num_blocks, channels):
    return [BottleneckBlock(channels, channels, bottleneck_channels=int(channels * 2), norm="LN") for _ in range(num_blocks)]


@ARCH_REGISTRY.register()
class PromptIR_NoImg_DC(nn.Module):
    def __init__(self, feature_dims, num_res_blocks=2, num_classes=3, downsample=False):
        super().__init__()
        if downsample:
            raise DcptError("downsample=True (token features from SwinIR) is not on the hot path")
        self.feature_dims = list(feature_dims)
        self.downsample = downsample
        self.bottleneck_layers = nn.ModuleList()
        self.downsample_layers = nn.ModuleList()
        for l, f in enumerate(self.feature_dims):
            self.bottleneck_layers.append(nn.Sequential(*_make_stage(num_res_blocks, f)))
            nxt = self.feature_dims[l + 1] if l < len(self.feature_dims) - 1 else f
            self.downsample_layers.append(nn.Sequential(nn.Conv2d(f, nxt, 1, bias=False), nn.MaxPool2d(2, 2), nn.ReLU()))
        self.last_stage = nn.Sequential(*_make_stage(num_res_blocks, self.feature_dims[-1]))
        self.mixing_weights = nn.Parameter(torch.ones(len(self.bottleneck_layers)), requires_grad=True)
        self.fc = nn.Linear(self.feature_dims[-1], num_classes)
        self._engine = DCHeadEngine(self.feature_dims, num_res_blocks, num_classes)
        assert [k for k, _ in self.named_parameters()] == self._engine.names, "parameter order differs from the C-side plan"

    def engine(self):
        return self._engine

    def train(self, mode=True):
        if self.training == bool(mode):          # (no module-tree walk when the mode does not change; nothing here depends on it)
            return self
        return super().train(mode)

    def prepack(self):
        """Re-pack the 16-bit operand images of the weights now (see NAFNetBaseline.prepack)."""
        params = list(self.parameters())
        if params and params[0].is_cuda:
            self._engine._pack([p.detach().contiguous() for p in params])

    def forward(self, lq, features):
        """``lq`` is unused, as in the reference (:622-641).  features: fine -> coarse, logical NCHW CUDA tensors."""
        return dchead_apply(self._engine, list(features), list(self.parameters()))


@ARCH_REGISTRY.register()
class PromptIR_DC(PromptIR_NoImg_DC):
    """degrad_classify_arch.py:480-556: the classifier variant that embeds the degraded image itself,
    ``lq_feats = conv_embed(lq)`` with ``conv_embed = Conv2d(3, f0, 7, stride 2, pad 3) + LayerNorm(f0)`` (:497-500), before the
    same feature-mixing trunk.  ``features[0]`` therefore lives at half the image resolution (the reference pairs this head
    with token backbones whose features are resampled; NAFNet / Restormer decoder features at full resolution go with
    ``PromptIR_NoImg_DC``)."""

    def __init__(self, feature_dims, num_res_blocks=2, num_classes=3):
        nn.Module.__init__(self)
        self.feature_dims = list(feature_dims)
        self.conv_embed = nn.Sequential(nn.Conv2d(3, self.feature_dims[0], 7, 2, 3), LayerNorm(self.feature_dims[0]))
        self.bottleneck_layers = nn.ModuleList()
        self.downsample_layers = nn.ModuleList()
        for l, f in enumerate(self.feature_dims):
            self.bottleneck_layers.append(nn.Sequential(*_make_stage(num_res_blocks, f)))
            nxt = self.feature_dims[l + 1] if l < len(self.feature_dims) - 1 else f
            self.downsample_layers.append(nn.Sequential(nn.Conv2d(f, nxt, 1, bias=False), nn.MaxPool2d(2, 2), nn.ReLU()))
        self.last_stage = nn.Sequential(*_make_stage(num_res_blocks, self.feature_dims[-1]))
        self.mixing_weights = nn.Parameter(torch.ones(len(self.bottleneck_layers)), requires_grad=True)
        self.fc = nn.Linear(self.feature_dims[-1], num_classes)
        self._engine = DCHeadEngine(self.feature_dims, num_res_blocks, num_classes, img_embed=True)
        assert [k for k, _ in self.named_parameters()] == self._engine.names, "parameter order differs from the C-side plan"

    def forward(self, lq, features):
        return dchead_apply(self._engine, list(features), list(self.parameters()), lq=lq)
