"""Losses the DCPT / fine-tune steps use (reference: basicsr/losses/__init__.py:25-38 ``build_loss``; basic_loss.py:40-88,
236-303).  Each is ``loss_weight * F.<loss>(pred, target, reduction=...)`` on a tensor the size of the network OUTPUT (3 channels)
or of the logits: a few hundred KB per step next to gigabytes of activations, so they stay torch ops on the caller's stream - the
networks' backward (the hot path) starts from the gradient they hand back."""
from copy import deepcopy

import torch
from torch import nn
from torch.nn import functional as F

from basicsr.utils import get_root_logger
from basicsr.utils.registry import LOSS_REGISTRY

__all__ = ["build_loss", "L1Loss", "MSELoss", "CharbonnierLoss", "CrossEntropyLoss"]

_REDUCTIONS = ("none", "mean", "sum")


class _Weighted(nn.Module):
    def __init__(self, loss_weight=1.0, reduction="mean"):
        super().__init__()
        if reduction not in _REDUCTIONS:
            raise ValueError(f"Unsupported reduction mode: {reduction}. Supported ones are: {list(_REDUCTIONS)}")
        self.loss_weight, self.reduction = loss_weight, reduction

    def _reduce(self, elementwise):
        if self.reduction == "mean":
            return elementwise.mean()
        return elementwise.sum() if self.reduction == "sum" else elementwise


@LOSS_REGISTRY.register()
class L1Loss(_Weighted):
    def forward(self, pred, target, weight=None, **kwargs):
        diff = (pred - target).abs()
        return self.loss_weight * self._reduce(diff if weight is None else diff * weight)


@LOSS_REGISTRY.register()
class MSELoss(_Weighted):
    def forward(self, pred, target, weight=None, **kwargs):
        diff = (pred - target) ** 2
        return self.loss_weight * self._reduce(diff if weight is None else diff * weight)


@LOSS_REGISTRY.register()
class CharbonnierLoss(_Weighted):
    """sqrt((pred - target)^2 + eps), eps = 1e-12 by default (basic_loss.py:268-303)."""

    def __init__(self, loss_weight=1.0, reduction="mean", eps=1e-12):
        super().__init__(loss_weight, reduction)
        self.eps = eps

    def forward(self, pred, target, weight=None, **kwargs):
        diff = torch.sqrt((pred - target) ** 2 + self.eps)
        return self.loss_weight * self._reduce(diff if weight is None else diff * weight)


@LOSS_REGISTRY.register()
class CrossEntropyLoss(_Weighted):
    def forward(self, pred, target):
        return self.loss_weight * F.cross_entropy(pred, target, reduction=self.reduction)


def build_loss(opt):
    opt = deepcopy(opt)
    loss = LOSS_REGISTRY.get(opt.pop("type"))(**opt)
    get_root_logger().info(f"Loss [{loss.__class__.__name__}] is created.")
    return loss
