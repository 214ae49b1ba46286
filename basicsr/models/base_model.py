"""What the two step classes share (reference: basicsr/models/base_model.py).  Only the pieces a training / test step touches
are mirrored - device placement, the data-parallel wrapper, the optimizer factory, EMA, learning-rate schedules, checkpoint
load / save in the reference's ``{"params": ..., "params_ema": ...}`` format, loss logging.  Validation loops, image saving,
TensorBoard and resume-state files are the reference's control plane and stay with it (a maintainer keeps their own
``base_model.py``; INTEGRATION.md lists the three methods to swap there).

B200-native differences, all behind the reference's method names:
  * ``model_to_device``: ``opt["dist"]`` wraps in ``FlatGradDataParallel`` (ONE mean all-reduce of the engine's flat gradient
    buffer) where the reference wraps in ``DistributedDataParallel`` (base_model.py:100-118);
  * ``get_optimizer``: Adam / AdamW on CUDA parameters are the fused multi-tensor kernel (``dcpt_b200.optim``), which also
    does ``clip_grad_norm_`` and the EMA update in the same launches (base_model.py:86-95, 120-139; sr_model.py:164-174);
  * ``reduce_loss_dict``: one device->host read for all logged losses (base_model.py:432-457 reads one ``.item()`` per key)."""
import math
import os
from collections import OrderedDict

import torch

from basicsr.utils import get_root_logger


class BaseModel:
    def __init__(self, opt):
        self.opt = opt
        self.device = torch.device("cuda" if opt["num_gpu"] != 0 else "cpu")
        self.is_train = opt["is_train"]
        self.optimizers, self.schedulers = [], []
        self.log_dict = OrderedDict()

    # ------------------------------------------------------------------ placement
    def model_to_device(self, net, dist=True, find_unused_parameters=None):
        net = net.to(self.device)
        if self.opt.get("dist") and dist:
            from dcpt_b200.dist import FlatGradDataParallel
            net = FlatGradDataParallel(net)
        return net

    @staticmethod
    def get_bare_model(net):
        return net.module if hasattr(net, "module") and hasattr(net, "no_sync") else net

    # ------------------------------------------------------------------ optimizers / schedules
    def fused_step(self):
        """True when the optimizers are the fused sm_100a ones (CUDA parameters; ``train.fused_optimizer: false`` opts out)."""
        return self.device.type == "cuda" and self.opt.get("train", {}).get("fused_optimizer", True)

    def get_optimizer(self, optim_type, params, lr, **kwargs):
        if self.fused_step():
            from dcpt_b200.optim import get_optimizer
            return get_optimizer(optim_type, params, lr, **kwargs)
        table = {"Adam": torch.optim.Adam, "AdamW": torch.optim.AdamW, "Adamax": torch.optim.Adamax, "SGD": torch.optim.SGD,
                 "ASGD": torch.optim.ASGD, "RMSprop": torch.optim.RMSprop, "Rprop": torch.optim.Rprop}
        if optim_type not in table:
            raise NotImplementedError(f"optimizer {optim_type} is not supported yet.")
        return table[optim_type](params, lr, **kwargs)

    def _trainable(self, net):
        keep = []
        for name, p in net.named_parameters():
            if p.requires_grad:
                keep.append(p)
            else:
                get_root_logger().warning(f"Params {name} will not be optimized.")
        return keep

    def setup_schedulers(self):
        """``train.scheduler``: MultiStepLR / MultiStepRestartLR and CosineAnnealingRestartLR (base_model.py:141-160;
        lr_scheduler.py:8-131), as closed-form functions of the iteration applied by ``update_learning_rate``."""
        sched = dict(self.opt["train"].get("scheduler") or {})
        kind = sched.pop("type", None)
        if kind in (None, "none"):
            self._lr_factor = None
        elif kind in ("MultiStepLR", "MultiStepRestartLR"):
            milestones = list(sched["milestones"])
            gamma = sched.get("gamma", 0.1)
            restarts = list(sched.get("restarts", [0]))
            weights = list(sched.get("restart_weights", [1]))

            def factor(epoch):
                # RECURSIVE, like the reference (lr_scheduler.py:33-47): a restart resets to initial * weight, a milestone scales
                # the group's CURRENT lr, anything else leaves it alone (so a warm-up value sticks until the next milestone)
                if epoch in restarts:
                    return ("reset", weights[restarts.index(epoch)])
                return ("scale", gamma ** milestones.count(epoch) if epoch in milestones else 1.0)
            self._lr_factor = factor
        elif kind == "CosineAnnealingRestartLR":
            periods = list(sched["periods"])
            weights = list(sched.get("restart_weights", [1] * len(periods)))
            eta_min = sched.get("eta_min", 0)
            ends = [sum(periods[:i + 1]) for i in range(len(periods))]

            def factor(it):
                idx = next((i for i, e in enumerate(ends) if it <= e), len(ends) - 1)
                begin = 0 if idx == 0 else ends[idx - 1]
                cos = 0.5 * (1 + math.cos(math.pi * (it - begin) / periods[idx]))
                # a per-period list, as the reference indexes it (lr_scheduler.py:109-117); a scalar applies to every period
                return ("cosine", weights[idx], cos, eta_min[idx] if isinstance(eta_min, (list, tuple)) else eta_min)
            self._lr_factor = factor
        else:
            raise NotImplementedError(f"Scheduler {kind} is not implemented yet.")
        self._init_lrs = [[g["lr"] for g in o.param_groups] for o in self.optimizers]

    def update_learning_rate(self, current_iter, warmup_iter=-1):
        if getattr(self, "_lr_factor", None) is not None and current_iter > 1:
            f = self._lr_factor(current_iter - 1)       # the reference steps its scheduler from iteration 2 on: epoch = iter - 1
            for opt, inits in zip(self.optimizers, self._init_lrs):
                for group, lr0 in zip(opt.param_groups, inits):
                    if f[0] == "cosine":          # eta_min + w * 0.5 * (lr0 - eta_min) * (1 + cos)
                        _, w, cos, eta_min = f
                        group["lr"] = eta_min + w * (lr0 - eta_min) * cos
                    elif f[0] == "reset":
                        group["lr"] = lr0 * f[1]
                    else:
                        group["lr"] = group["lr"] * f[1]
        if current_iter < warmup_iter:            # linear warm-up from zero (base_model.py:173-186)
            for opt, inits in zip(self.optimizers, self._init_lrs):
                for group, lr0 in zip(opt.param_groups, inits):
                    group["lr"] = lr0 / warmup_iter * current_iter

    def get_current_learning_rate(self):
        return [g["lr"] for g in self.optimizers[0].param_groups]

    def prepack(self, *nets):
        """After optimizer.step(): let the engines refresh their packed weights while the GPU is still busy with the update
        (the arch mirrors' ``prepack``; a network without one is left alone)."""
        if self.device.type != "cuda":
            return
        for net in nets:
            fn = getattr(self.get_bare_model(net), "prepack", None)
            if fn is not None:
                fn()

    # ------------------------------------------------------------------ EMA
    def model_ema(self, decay=0.999):
        """net_g_ema = decay * net_g_ema + (1 - decay) * net_g over the parameters (base_model.py:86-95), one multi-tensor
        launch.  (``optimize_parameters`` normally folds this into the fused optimizer step and never calls it.)"""
        src = list(self.get_bare_model(self.net_g).parameters())
        dst = list(self.net_g_ema.parameters())
        with torch.no_grad():
            if decay == 0:
                for d, s in zip(dst, src):
                    d.copy_(s)
            else:
                torch._foreach_mul_(dst, decay)
                torch._foreach_add_(dst, src, alpha=1 - decay)

    # ------------------------------------------------------------------ checkpoints
    def load_network(self, net, load_path, strict=True, param_key="params", remove_norm=False):
        net = self.get_bare_model(net)
        blob = torch.load(load_path, map_location="cpu")
        if param_key is not None:
            if param_key not in blob and "params" in blob:
                param_key = "params"
                get_root_logger().info("Loading: params_ema does not exist, use params.")
            blob = blob[param_key]
        get_root_logger().info(f"Loading {net.__class__.__name__} model from {load_path}, with param key: [{param_key}].")
        state = OrderedDict((k[7:] if k.startswith("module.") else k, v) for k, v in blob.items())
        if remove_norm:
            state = OrderedDict((k, v) for k, v in state.items() if "norm" not in k)
        if not strict:     # same-named tensors of another size are skipped, as the reference's key report does
            own = net.state_dict()
            state = OrderedDict((k, v) for k, v in state.items() if k not in own or own[k].shape == v.shape)
        net.load_state_dict(state, strict=strict)

    def save_network(self, net, net_label, current_iter, param_key="params"):
        nets = net if isinstance(net, list) else [net]
        keys = param_key if isinstance(param_key, list) else [param_key]
        assert len(nets) == len(keys), "The lengths of net and param_key should be the same."
        name = f"{net_label}_{'latest' if current_iter == -1 else current_iter}.pth"
        blob = {k: OrderedDict((n[7:] if n.startswith("module.") else n, t.cpu())
                               for n, t in self.get_bare_model(m).state_dict().items()) for m, k in zip(nets, keys)}
        path = os.path.join(self.opt["path"]["models"], name)
        if self.opt.get("rank", 0) == 0:
            os.makedirs(os.path.dirname(path), exist_ok=True)
            torch.save(blob, path)
        return path

    # ------------------------------------------------------------------ logging
    def reduce_loss_dict(self, loss_dict):
        from dcpt_b200.dist import reduce_loss_dict
        with torch.no_grad():
            return OrderedDict(reduce_loss_dict({k: v.mean() for k, v in loss_dict.items()}))

    def get_current_log(self):
        return self.log_dict
