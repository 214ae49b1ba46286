"""Fused parameter update for the training step — host side of ``dcpt_optim_*``.

Replaces, in the reference's ``SRModel.optimize_parameters`` (basicsr/models/sr_model.py:164-174) and
``DCPTModel.optimize_parameters`` (degradation_classification_pretrain_model.py:163-165):

    torch.nn.utils.clip_grad_norm_(net_g.parameters(), grad_clip)     # :166-167 (optional)
    optimizer_g.step()                                                # :169, torch.optim.Adam / AdamW (base_model.py:120-139)
    self.model_ema(decay)                                             # :173-174, a Python loop over 664 tensors (base_model.py:86-95)

by ``FusedAdam.step(grad_clip=..., ema_params=..., ema_decay=...)``: two multi-tensor sm_100a kernels through the C ABI.
``FusedAdam`` is a ``torch.optim.Optimizer`` whose ``param_groups`` / ``state`` (``step``, ``exp_avg``, ``exp_avg_sq``) are
those of ``torch.optim.Adam`` / ``AdamW``, so the reference's ``.state`` resume files (base_model.py:413-430) load, and the
reference's LR schedulers (which edit ``param_groups[i]['lr']``) keep working.  No CPU path: parameters must be CUDA fp32.
"""
import ctypes as C

import torch

from . import lib as _l
from .ops import _stream


class _Plan:
    """C plan + device workspace for one list of tensors (a param group's parameters that currently have gradients)."""

    def __init__(self, lib, numels, device):
        self.lib = lib
        arr = (C.c_longlong * len(numels))(*numels)
        h = lib.dcpt_optim_create(arr, len(numels))
        if not h:
            raise _l.DcptError("dcpt_optim_create: " + lib.dcpt_last_error().decode())
        self.h = C.c_void_p(h)
        self.work = torch.empty(lib.dcpt_optim_workspace_bytes(self.h), dtype=torch.uint8, device=device)
        self.bound = None

    def bind(self, ptrs):
        """ptrs: tuple of 5 tuples of device addresses (params, grads, exp_avg, exp_avg_sq, ema or None)."""
        if ptrs == self.bound:
            return
        arrs = [None if p is None else _l.ptr_array(list(p)) for p in ptrs]
        _l.check(self.lib.dcpt_optim_bind(self.h, C.c_void_p(self.work.data_ptr()), *arrs, _stream()), "optim_bind")
        self.bound = ptrs

    def __del__(self):
        try:
            if self.h:
                self.lib.dcpt_optim_destroy(self.h)
                self.h = None
        except Exception:
            pass


class FusedAdam(torch.optim.Optimizer):
    """torch.optim.Adam (``decoupled_weight_decay=False``) / AdamW (``True``) on the sm_100a fused kernel.

    ``step(grad_clip=None, ema_params=None, ema_decay=0.0)``: ``grad_clip`` = ``max_norm`` of ``clip_grad_norm_`` over ALL
    parameters of the optimizer (the reference clips ``net_g.parameters()``, which is what ``optimizer_g`` holds);
    ``ema_params`` = the EMA network's parameters in the same order as the optimizer's (``net_g_ema.parameters()``).
    Returns the total gradient norm (a 0-dim CUDA tensor, no host sync) when clipping, else None."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled_weight_decay=False,
                 amsgrad=False, maximize=False):
        if amsgrad or maximize:
            raise _l.DcptError("FusedAdam: amsgrad / maximize are not built (the reference never sets them)")
        if lr < 0 or eps < 0 or not 0 <= betas[0] < 1 or not 0 <= betas[1] < 1 or weight_decay < 0:
            raise ValueError("invalid Adam hyper-parameters")   # torch.optim.Adam raises ValueError for these
        super().__init__(params, dict(lr=lr, betas=tuple(betas), eps=eps, weight_decay=weight_decay,
                                      decoupled_weight_decay=decoupled_weight_decay, amsgrad=False, maximize=False))
        self._lib = _l.load_library()
        self._plans = {}
        self._fast = {}

    def zero_grad(self, set_to_none=True):
        """torch.optim.Optimizer.zero_grad with set_to_none=True, minus its per-call bookkeeping (hooks, per-device foreach
        grouping): 1.7 -> 0.2 ms for 664 tensors, host time during which the GPU idles at the start of every iteration (the
        step classes call optimizer.zero_grad() first, sr_model.py:134, degradation_classification_pretrain_model.py:139)."""
        if not set_to_none:
            return super().zero_grad(set_to_none=False)
        for group in self.param_groups:
            for p in group["params"]:
                p.grad = None

    def _plan(self, key, numels, device):
        if key not in self._plans:
            self._plans[key] = _Plan(self._lib, numels, device)
        return self._plans[key]

    def load_state_dict(self, state_dict):
        super().load_state_dict(state_dict)
        self._fast = {}                                             # moments / step counts were replaced

    def _prepare(self, gi, group, sig, ema_list):
        """Slow path, taken when the set of (parameter, gradient, ema) addresses of a group changed: validate, create the
        optimizer state like torch.optim.adam._init_group, split by step count, (re)bind the C plans."""
        jobs, by_step = [], {}
        for pi, p in enumerate(group["params"]):
            if p.grad is None:
                continue                                            # torch skips parameters without a gradient
            if not p.is_cuda:
                raise _l.DcptError("dcpt_b200 has no CPU path: parameter is on %s" % p.device)
            if p.device != group["params"][0].device or p.grad.device != p.device:
                raise _l.DcptError("FusedAdam: all parameters and gradients of a param group must live on one device")
            if p.dtype != torch.float32 or p.grad.dtype != torch.float32 or not p.is_contiguous() or p.grad.is_sparse \
                    or not p.grad.is_contiguous():
                raise _l.DcptError("FusedAdam: parameters and gradients must be dense contiguous fp32")
            st = self.state[p]
            if len(st) == 0:
                st["step"] = torch.tensor(0.0, dtype=torch.float32)
                st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
            by_step.setdefault(int(st["step"]), []).append(pi)
        for step0, idx in by_step.items():
            ps = [group["params"][pi] for pi in idx]
            plan = self._plan((gi, tuple(idx)), [p.numel() for p in ps], ps[0].device)
            emas = None
            if ema_list is not None:
                emas = []
                for pi, p in zip(idx, ps):
                    e = ema_list[pi]
                    if e.shape != p.shape or e.dtype != torch.float32 or not e.is_cuda or not e.is_contiguous():
                        raise _l.DcptError("ema parameter does not match its parameter (shape / fp32 / CUDA / contiguous)")
                    emas.append(e.data_ptr())
                emas = tuple(emas)
            plan.bind((tuple(p.data_ptr() for p in ps), tuple(p.grad.data_ptr() for p in ps),
                       tuple(self.state[p]["exp_avg"].data_ptr() for p in ps),
                       tuple(self.state[p]["exp_avg_sq"].data_ptr() for p in ps), emas))
            jobs.append([plan, step0, [self.state[p]["step"] for p in ps]])
        return {"sig": sig, "jobs": jobs}

    @torch.no_grad()
    def step(self, closure=None, grad_clip=None, ema_params=None, ema_decay=0.0):
        if closure is not None:
            raise _l.DcptError("FusedAdam.step: closures are not supported")
        ema_all = None
        if ema_params is not None and ema_decay > 0:
            ema_all = list(ema_params)
            if len(ema_all) != sum(len(g["params"]) for g in self.param_groups):
                raise _l.DcptError(f"ema_params has {len(ema_all)} tensors, the optimizer "
                                   f"{sum(len(g['params']) for g in self.param_groups)}")
        prepared, off = [], 0
        for gi, group in enumerate(self.param_groups):
            ps = group["params"]
            ema_list = ema_all[off:off + len(ps)] if ema_all is not None else None
            off += len(ps)
            # per-step host work is one pass collecting addresses; everything else is cached until an address changes
            sig = (tuple(p.data_ptr() for p in ps), tuple(0 if p.grad is None else p.grad.data_ptr() for p in ps),
                   None if ema_list is None else tuple(e.data_ptr() for e in ema_list))
            fast = self._fast.get(gi)
            if fast is None or fast["sig"] != sig:
                fast = self._fast[gi] = self._prepare(gi, group, sig, ema_list)
            prepared.extend((group, job) for job in fast["jobs"])
        if not prepared:
            return None
        total_norm = None
        if grad_clip is not None and grad_clip > 0:
            # clip_grad_norm_ is over every parameter of the optimizer: one job in the normal case (one group, one step count);
            # with several jobs (param groups, or parameters that got their first gradient later) the per-job norms are
            # combined on the device, sqrt(sum n_i^2), and handed back to every job before its update
            dev = prepared[0][1][0].work.device
            norms = torch.empty(len(prepared), dtype=torch.float32, device=dev)
            for i, (_, job) in enumerate(prepared):
                plan = job[0]
                _l.check(self._lib.dcpt_optim_grad_norm(plan.h, C.c_void_p(plan.work.data_ptr()),
                                                        C.c_void_p(norms.data_ptr() + 4 * i), _stream()), "optim_grad_norm")
            if len(prepared) == 1:
                total_norm = norms[0]
            else:
                total_norm = torch.linalg.vector_norm(norms.double()).float()
                for _, job in prepared:
                    plan = job[0]
                    _l.check(self._lib.dcpt_optim_set_norm(plan.h, C.c_void_p(plan.work.data_ptr()),
                                                           C.c_void_p(total_norm.data_ptr()), _stream()), "optim_set_norm")
        for group, job in prepared:
            plan, step0, step_tensors = job
            b1, b2 = group["betas"]
            if plan.work.device.index != torch.cuda.current_device():
                raise _l.DcptError(f"FusedAdam: parameters live on {plan.work.device} but the current CUDA device is "
                                   f"{torch.cuda.current_device()} (the library launches on the current device: use torch.cuda.device)")
            _l.check(self._lib.dcpt_optim_step(plan.h, C.c_void_p(plan.work.data_ptr()), int(bool(group["decoupled_weight_decay"])),
                                               float(group["lr"]), float(b1), float(b2), float(group["eps"]),
                                               float(group["weight_decay"]), step0 + 1, float(grad_clip or 0.0),
                                               float(ema_decay if ema_all is not None else 0.0), _stream()), "optim_step")
            torch._foreach_add_(step_tensors, 1.0)                  # the per-parameter `step` tensors of torch's state layout
            job[1] = step0 + 1
        # The kernels wrote through raw device pointers: tell torch (and the engines' packed-operand caches, which key on the
        # version counters) that the parameters and the EMA copies changed, as an in-place torch op would have.
        touched = [p for group, _ in prepared for p in group["params"] if p.grad is not None]
        if ema_all is not None:
            touched += ema_all
        torch.autograd.graph.increment_version(touched)
        return total_norm


class FusedAdamW(FusedAdam):
    """torch.optim.AdamW: decoupled weight decay, default 1e-2."""

    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=1e-2, amsgrad=False, maximize=False):
        super().__init__(params, lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, decoupled_weight_decay=True,
                         amsgrad=amsgrad, maximize=maximize)


def get_optimizer(optim_type, params, lr, **kwargs):
    """Drop-in for BaseModel.get_optimizer (base_model.py:120-139) for the two types the DCPT configs use."""
    if optim_type == "Adam":
        return FusedAdam(params, lr, **kwargs)
    if optim_type == "AdamW":
        return FusedAdamW(params, lr, **kwargs)
    raise NotImplementedError(f"optimizer {optim_type} is not on the B200 hot path (Adam / AdamW are)")
