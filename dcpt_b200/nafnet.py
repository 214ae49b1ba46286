"""Host-side engine for NAFNetBaseline (reference: basicsr/archs/nafnet_arch.py:189-274).

``NAFNetEngine`` owns the C plan, the packed bf16 weight cache and the per-call arenas; the
``basicsr`` mirror's ``NAFNetBaseline`` module calls it through ``nafnet_apply`` (an
``autograd.Function``), so ``loss.backward()``, DDP and optimizers work as with the reference.
"""
import ctypes as C

import torch

from . import lib as _l
from .ops import _p, _stream


class NAFNetEngine:
    def __init__(self, img_channel, width, middle_blk_num, enc_blk_nums, dec_blk_nums):
        self.lib = _l.load_library()
        enc = (C.c_int * len(enc_blk_nums))(*enc_blk_nums)
        dec = (C.c_int * len(dec_blk_nums))(*dec_blk_nums)
        self.plan = self.lib.dcpt_nafnet_create(img_channel, width, middle_blk_num, enc, len(enc_blk_nums), dec,
                                                len(dec_blk_nums))
        if not self.plan:
            raise _l.DcptError("dcpt_nafnet_create: " + self.lib.dcpt_last_error().decode())
        self.plan = C.c_void_p(self.plan)
        self.n_dec = len(dec_blk_nums)
        self.n_enc = len(enc_blk_nums)
        self.width = width
        self.num_params = self.lib.dcpt_nafnet_num_params(self.plan)
        self.shapes = []
        dims = (C.c_int * 4)()
        for i in range(self.num_params):
            self.lib.dcpt_nafnet_param_shape(self.plan, i, dims)
            self.shapes.append(tuple(dims))
        self._packed = None
        self._packed_key = None
        self._scratch = {}

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.dcpt_nafnet_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    # ---- parameters -------------------------------------------------------------------------
    def _check_params(self, params):
        if len(params) != self.num_params:
            raise _l.DcptError(f"expected {self.num_params} parameters, got {len(params)}")
        for p in params:
            if not p.is_cuda:
                raise _l.DcptError("dcpt_b200 has no CPU path: move the network to a CUDA device (B200)")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _l.DcptError("parameters must be contiguous fp32 (master weights)")

    def packed_for(self, params):
        """bf16 operand cache; refreshed whenever a parameter was modified in place or replaced."""
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._packed is None or self._packed.device != params[0].device:
            self._packed = torch.empty(self.lib.dcpt_nafnet_packed_bytes(self.plan), dtype=torch.uint8,
                                       device=params[0].device)
            self._packed_key = None
        if key != self._packed_key:
            pp = _l.ptr_array([p.data_ptr() for p in params])
            _l.check(self.lib.dcpt_nafnet_pack(self.plan, pp, _p(self._packed), _stream()), "nafnet_pack")
            self._packed_key = key
        return self._packed

    @staticmethod
    def alloc_flat_grads(params, align=64):
        """One zeroed fp32 buffer holding every parameter gradient (each view starts on a 256-byte
        boundary: the kernels use 16-byte vector loads / reductions), returned as (flat, views)."""
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + align - 1) // align * align
        flat = torch.zeros(off, dtype=torch.float32, device=params[0].device)
        return flat, [flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, params)]

    # ---- forward / backward -------------------------------------------------------------------
    def forward(self, params, inp, hook=False, want_feats=False, keep_for_backward=True):
        """inp fp32 NCHW [N,3,H,W].  Returns (out or None, feats list (NHWC fp32) or None, saved arena)."""
        self._check_params(params)
        if not inp.is_cuda:
            raise _l.DcptError("dcpt_b200 has no CPU path: input is on %s" % inp.device)
        inp = inp.contiguous().float()
        N, _, H, W = inp.shape
        dev = inp.device
        packed = self.packed_for(params)
        nbytes = self.lib.dcpt_nafnet_saved_bytes(self.plan, N, H, W)
        if keep_for_backward:
            saved = torch.empty(nbytes, dtype=torch.uint8, device=dev)
        else:  # inference: one reusable arena per shape
            k = ("saved", N, H, W, dev)
            if k not in self._scratch:
                self._scratch[k] = torch.empty(nbytes, dtype=torch.uint8, device=dev)
            saved = self._scratch[k]
        out = None if hook else torch.empty_like(inp)
        feats = fp = None
        if want_feats:
            feats, c, h, w = [], self.width << self.n_enc, H >> self.n_enc, W >> self.n_enc
            for _ in range(self.n_dec):
                c, h, w = c // 2, h * 2, w * 2
                feats.append(torch.empty(N, h, w, c, dtype=torch.float32, device=dev))
            fp = _l.ptr_array([f.data_ptr() for f in feats])
        pp = _l.ptr_array([p.data_ptr() for p in params])
        _l.check(self.lib.dcpt_nafnet_fwd(self.plan, pp, _p(packed), _p(inp), _p(out), _p(saved), fp, int(bool(hook)), N, H,
                                          W, _stream()), "nafnet_fwd")
        return out, feats, saved

    def backward(self, params, inp, saved, dout, dfeats=None, grads=None):
        """Accumulates parameter gradients into ``grads`` (list of fp32 tensors; allocated zeroed if None)."""
        N, _, H, W = inp.shape
        dev = inp.device
        if grads is None:
            grads = self.alloc_flat_grads(params)[1]
        k = ("work", N, H, W, dev)
        if k not in self._scratch:
            self._scratch[k] = torch.empty(self.lib.dcpt_nafnet_workspace_bytes(self.plan, N, H, W), dtype=torch.uint8,
                                           device=dev)
        work = self._scratch[k]
        pp = _l.ptr_array([p.data_ptr() for p in params])
        gp = _l.ptr_array([g.data_ptr() for g in grads])
        dfp = None
        if dfeats is not None and any(d is not None for d in dfeats):
            dfeats = [None if d is None else d.contiguous().float() for d in dfeats]
            dfp = _l.ptr_array([0 if d is None else d.data_ptr() for d in dfeats])
        if dout is not None:
            dout = dout.contiguous().float()
        packed = self.packed_for(params)
        _l.check(self.lib.dcpt_nafnet_bwd(self.plan, pp, _p(packed), _p(saved), _p(inp), _p(dout), dfp, gp, _p(work), N, H,
                                          W, _stream()), "nafnet_bwd")
        return grads


class _NAFNetFunction(torch.autograd.Function):
    """out, feat_0..feat_{n-1} = NAFNet(inp; params).  feats are NHWC storage viewed as logical NCHW."""

    @staticmethod
    def forward(ctx, engine, inp, hook, want_feats, need_grad, *params):
        # need_grad is decided by the caller: grad mode is always OFF inside Function.forward, and a forward whose
        # activations are needed by a later backward must own its saved-activation arena (DCPT runs two forwards
        # before one backward).
        dparams = [p.detach() for p in params]
        inp_c = inp.detach().contiguous().float()
        out, feats, saved = engine.forward(dparams, inp_c, hook=hook, want_feats=want_feats, keep_for_backward=need_grad)
        ctx.engine, ctx.hook, ctx.n_feats = engine, hook, len(feats) if feats else 0
        ctx.inp, ctx.saved, ctx.params = inp_c, saved, dparams
        outs = []
        if out is None:
            out = inp_c.new_zeros(())  # placeholder (hook=True returns None to the caller)
            ctx.mark_non_differentiable(out)
        outs.append(out)
        if feats:
            outs.extend(f.permute(0, 3, 1, 2) for f in feats)  # logical NCHW, channels_last memory
        return tuple(outs)

    @staticmethod
    def backward(ctx, dout, *dfeats):
        eng = ctx.engine
        dfe = None
        if ctx.n_feats:
            dfe = [None if d is None else d.permute(0, 2, 3, 1).contiguous() for d in dfeats]
        grads = eng.backward(ctx.params, ctx.inp, ctx.saved, None if ctx.hook else dout, dfe)
        ctx.saved = None
        return (None, None, None, None, None) + tuple(grads)


def nafnet_apply(engine, inp, params, hook=False, want_feats=False):
    need_grad = torch.is_grad_enabled() and (inp.requires_grad or any(p.requires_grad for p in params))
    res = _NAFNetFunction.apply(engine, inp, hook, want_feats, need_grad, *params)
    out = None if hook else res[0]
    return out, list(res[1:])
