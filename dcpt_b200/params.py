"""Packed-operand cache validity (shared by the NAFNet / Restormer / DC-head engines).

The engines keep bf16 GEMM-operand copies of the fp32 parameters.  They must be rebuilt whenever a parameter changes:

* writes that go through torch (optimizers, ``load_state_dict``, ``copy_``) bump the tensor's version counter, which the cache
  key contains;
* ``FusedAdam`` / ``FusedAdamW`` update through raw device pointers and bump the counters themselves
  (``torch.autograd.graph.increment_version``, optim.py);
* writes through ``p.data`` (the reference's ``BaseModel.model_ema``: ``ema.data.mul_(decay).add_(p.data, ...)``,
  basicsr/models/base_model.py:86-95) bump NOTHING.  The network they touch (``net_g_ema``) is only ever run under
  ``torch.no_grad()`` (validation, sr_model.py:176-185), so every no-grad forward also compares a 64-bit device fingerprint of
  the parameter bits (one multi-tensor kernel over the weights, ~45 us for NAFNet-w64, plus an 8-byte read-back).
  ``DCPT_PARAM_FINGERPRINT=0`` turns that check off.
"""
import ctypes as C
import os
import threading

import torch

from . import lib as _l
from .ops import _stream


class ParamFingerprint:
    """64-bit order-independent hash of a fixed list of CUDA fp32 tensors (dcpt_optim_param_hash)."""

    def __init__(self, tensors):
        self.lib = _l.load_library()
        self.ptrs = tuple(t.data_ptr() for t in tensors)
        dev = tensors[0].device
        numels = [t.numel() for t in tensors]
        arr = (C.c_longlong * len(numels))(*numels)
        h = self.lib.dcpt_optim_create(arr, len(numels))
        if not h:
            raise _l.DcptError("dcpt_optim_create: " + self.lib.dcpt_last_error().decode())
        self.h = C.c_void_p(h)
        self.work = torch.empty(self.lib.dcpt_optim_workspace_bytes(self.h), dtype=torch.uint8, device=dev)
        self.out = torch.zeros(1, dtype=torch.int64, device=dev)
        pp = _l.ptr_array(list(self.ptrs))
        _l.check(self.lib.dcpt_optim_bind(self.h, C.c_void_p(self.work.data_ptr()), pp, pp, pp, pp, None, _stream()), "optim_bind")

    def value(self):
        _l.check(self.lib.dcpt_optim_param_hash(self.h, C.c_void_p(self.work.data_ptr()), C.c_void_p(self.out.data_ptr()), _stream()),
                 "optim_param_hash")
        return int(self.out.item())

    def __del__(self):
        try:
            if self.h:
                self.lib.dcpt_optim_destroy(self.h)
                self.h = None
        except Exception:
            pass


_tls = threading.local()


class training_pass:
    """``with training_pass():`` around the forward / backward of a gradient-needing call.  Inside ``autograd.Function.forward``
    and ``.backward`` grad mode is always off, so "no-grad" cannot be read from ``torch.is_grad_enabled()`` there; without this
    marker every training forward and backward would run the weight fingerprint and its 8-byte read-back - a host sync that
    keeps the CPU from queueing the next pass (the DCPT step has three forwards and three backwards per iteration)."""

    def __enter__(self):
        _tls.depth = getattr(_tls, "depth", 0) + 1
        return self

    def __exit__(self, *exc):
        _tls.depth -= 1
        return False


def in_training_pass():
    return getattr(_tls, "depth", 0) > 0


class frozen_weights:
    """``with frozen_weights():`` around a loop of no-grad forwards that THIS package drives and inside which no user code can
    touch the parameters (the tile loop of dcpt_b200/tiling.py): the weight fingerprint - a host sync per forward - is compared
    by the first forward of each engine inside the block and skipped by the following ones."""

    def __enter__(self):
        _tls.frozen = getattr(_tls, "frozen", 0) + 1
        if _tls.frozen == 1:
            _tls.frozen_seen = set()
        return self

    def __exit__(self, *exc):
        _tls.frozen -= 1
        if _tls.frozen == 0:
            _tls.frozen_seen = set()
        return False


def _fingerprint_already_checked(cache_key_obj):
    """Inside frozen_weights(): True from the second call on for this PackedCacheKey."""
    if getattr(_tls, "frozen", 0) <= 0:
        return False
    seen = _tls.frozen_seen
    if id(cache_key_obj) in seen:
        return True
    seen.add(id(cache_key_obj))
    return False


class PackedCacheKey:
    """Decides when an engine's packed operand cache must be rebuilt (see the module docstring)."""

    def __init__(self):
        self.key = None
        self.hash = None
        self.fp = None
        self.use_fp = os.getenv("DCPT_PARAM_FINGERPRINT", "1") != "0"

    def invalidate(self):
        self.key = None

    def stale(self, params):
        """True when the cache must be rebuilt for `params` (and records the new state: call the pack right after)."""
        key = tuple((p.data_ptr(), p._version) for p in params)
        h = None
        if self.use_fp and not torch.is_grad_enabled() and not in_training_pass() and params[0].is_cuda and \
                not torch.cuda.is_current_stream_capturing() and not (key == self.key and _fingerprint_already_checked(self)):
            ptrs = tuple(k[0] for k in key)
            if self.fp is None or self.fp.ptrs != ptrs:
                self.fp = ParamFingerprint(params)
            h = self.fp.value()
        changed = key != self.key or (h is not None and h != self.hash)
        if changed:
            self.key, self.hash = key, h
        return changed


class LRUCache:
    """Small LRU map for the engines' per-shape resources (CUDA-graph slots with their saved-activation arenas, workspaces):
    validation over images of many sizes, or parameters re-allocated by ``.to()``, must not pin one arena per shape forever
    (ADVICE r1).  At most ``cap`` keys (``DCPT_CACHE_SHAPES``, default 4); an entry for which ``can_evict`` is False (a slot
    whose backward has not run yet) is skipped."""

    def __init__(self, cap=None, can_evict=None):
        from collections import OrderedDict
        self.cap = int(os.getenv("DCPT_CACHE_SHAPES", "4")) if cap is None else cap
        self.d = OrderedDict()
        self.can_evict = can_evict or (lambda v: True)

    def get(self, key):
        v = self.d.get(key)
        if v is not None:
            self.d.move_to_end(key)
        return v

    def put(self, key, value):
        self.d[key] = value
        self.d.move_to_end(key)
        while len(self.d) > self.cap:
            victim = next((k for k, v in self.d.items() if k != key and self.can_evict(v)), None)
            if victim is None:
                break
            del self.d[victim]
        return value

    def setdefault(self, key, factory):
        v = self.get(key)
        return v if v is not None else self.put(key, factory())

    def __contains__(self, key):
        return key in self.d

    def __len__(self):
        return len(self.d)

    def clear(self):
        self.d.clear()


def fp16_grad_scale(tensors, target=256.0):
    """IEEE-half operand build (DCPT_OPERAND=fp16) only: the gradients entering a backward pass are tiny (a mean L1 loss over
    16x3x256x256 pixels hands out +-3e-7, below fp16's normal range), so the pass runs on gradients multiplied by a power of two
    that brings the largest incoming magnitude to ~`target`, and its results are multiplied back (the backward is linear in the
    incoming gradient; powers of two are exact).  Computed on the device, no host sync.  Returns (scale, 1/scale) 0-dim
    tensors, or None for the bf16 build (8 exponent bits: no scaling needed)."""
    ts = [t for t in tensors if t is not None]
    if not ts or _l.operand_dtype() != torch.float16:
        return None
    amax = torch.stack([t.detach().abs().max().float() for t in ts]).max().clamp_min(1e-30)
    k = torch.floor(torch.log2(target / amax))
    return torch.exp2(k), torch.exp2(-k)
