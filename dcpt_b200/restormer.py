"""Host-side engine for Restormer (reference: basicsr/archs/restormer_arch.py:234-422).

``RestormerEngine`` owns the C plan, the packed bf16 operand cache and per-shape workspaces; the ``basicsr`` mirror's
``Restormer`` module calls ``forward`` (inference, CUDA-graph replayed) or ``restormer_apply`` (an ``autograd.Function``
over ``dcpt_restormer_fwd_train`` / ``dcpt_restormer_bwd``).  No PyTorch implementation of the math lives here and there is
no CPU path.
"""
import ctypes as C
import os
import warnings
import weakref

import torch

from . import lib as _l
from .ops import _p, _stream
from .params import LRUCache, PackedCacheKey, fp16_grad_scale, training_pass


class RestormerEngine:
    _ABI = "dcpt_restormer"     # prefix of the create / packed_bytes / workspace_bytes / pack entry points (PromptIREngine: dcpt_promptir)

    def _fn(self, name):
        return getattr(self.lib, f"{self._ABI}_{name}")

    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=(4, 6, 6, 8), num_refinement_blocks=4, heads=(1, 2, 4, 8),
                 ffn_expansion_factor=2.66, bias=False, ln_with_bias=False, attn_softmax=False):
        """attn_softmax: the transposed attention map goes through softmax(dim=-1) (PromptIR's transformer blocks,
        promptir_arch.py:140) instead of the fork's ReLU (restormer_arch.py:136)."""
        self.lib = _l.load_library()
        nb = (C.c_int * 4)(*num_blocks)
        hd = (C.c_int * 4)(*heads)
        plan = self._fn("create")(inp_channels, out_channels, dim, nb, num_refinement_blocks, hd,
                                              float(ffn_expansion_factor), int(bool(bias)), int(bool(ln_with_bias)))
        if not plan:
            raise _l.DcptError(f"{self._ABI}_create: " + self.lib.dcpt_last_error().decode())
        self.plan = C.c_void_p(plan)
        self.dim = dim
        if attn_softmax:
            _l.check(self.lib.dcpt_restormer_set_attention(self.plan, 1), "restormer_set_attention")
        self.num_params = self.lib.dcpt_restormer_num_params(self.plan)
        dims = (C.c_int * 4)()
        self.numels = [self.lib.dcpt_restormer_param_shape(self.plan, i, dims) for i in range(self.num_params)]
        self._packed = None
        self._packed_key = PackedCacheKey()
        self._work = LRUCache()       # eager-path workspaces per shape (graph entries / train slots own theirs)
        # whole-forward CUDA-graph replay (~700 launches per 128x128 tile become one); DCPT_CUDA_GRAPH=0 launches eagerly
        self.use_graphs = os.getenv("DCPT_CUDA_GRAPH", "1") != "0"
        self._graphs = LRUCache()
        self._seen = LRUCache(cap=64)
        self.grad_sync = None      # set by dcpt_b200.dist.FlatGradDataParallel: callable(flat fp32 gradient buffer)
        # training path: CUDA-graph replay of forward-with-save and of backward (~1900 launches per step otherwise);
        # DCPT_RESTORMER_TRAIN_GRAPH=0 keeps eager launches
        self.use_train_graphs = self.use_graphs and os.getenv("DCPT_RESTORMER_TRAIN_GRAPH", "1") != "0"
        self._tslots = LRUCache(can_evict=lambda slots: not any(s_.busy for s_ in slots))

    def __del__(self):
        try:
            if getattr(self, "plan", None):
                self.lib.dcpt_restormer_destroy(self.plan)
                self.plan = None
        except Exception:
            pass

    def _check_params(self, params):
        if len(params) != self.num_params or [p.numel() for p in params] != self.numels:
            raise _l.DcptError(f"parameter layout mismatch between module and C plan ({len(params)} vs {self.num_params} tensors)")
        for p in params:
            if not p.is_cuda:
                raise _l.DcptError("dcpt_b200 has no CPU path: move the network to a CUDA device (B200)")
            if p.dtype != torch.float32 or not p.is_contiguous():
                raise _l.DcptError("parameters must be contiguous fp32")

    def packed_for(self, params):
        if self._packed is None or self._packed.device != params[0].device:
            self._packed = torch.empty(self._fn("packed_bytes")(self.plan), dtype=torch.uint8, device=params[0].device)
            self._packed_key.invalidate()
        if self._packed_key.stale(params):     # version counters + no-grad weight fingerprint, dcpt_b200/params.py
            pp = _l.ptr_array([p.data_ptr() for p in params])
            _l.check(self._fn("pack")(self.plan, pp, _p(self._packed), _stream()), f"{self._ABI}_pack")
        return self._packed

    def invalidate_packed(self):
        self._packed_key.invalidate()

    def forward(self, params, inp, hook=False, want_feats=False):
        """inp fp32 NCHW [N,3,H,W] -> (out or None, feats [decoder_level3, 2, 1] as NHWC fp32 or None)."""
        self._check_params(params)
        if not inp.is_cuda:
            raise _l.DcptError("dcpt_b200 has no CPU path: input is on %s" % inp.device)
        inp = inp.contiguous().float()
        N, _, H, W = inp.shape
        dev = inp.device
        packed = self.packed_for(params)
        if self.use_graphs and not torch.cuda.is_current_stream_capturing():
            # a shape earns a graph entry (static buffers + workspace + capture) when it comes back; first sighting is eager
            gkey = (N, H, W, dev, bool(hook), bool(want_feats), tuple(p.data_ptr() for p in params))
            if gkey in self._graphs or self._seen.get(gkey) is not None:
                return self._graph_forward(params, packed, inp, bool(hook), bool(want_feats), gkey)
            self._seen.put(gkey, True)
        work = self._work.setdefault((N, H, W, dev), lambda: torch.empty(
            self.lib.dcpt_restormer_workspace_bytes(self.plan, N, H, W), dtype=torch.uint8, device=dev))
        out = None if hook else torch.empty_like(inp)
        feats = fp = None
        if want_feats:
            d = self.dim
            feats = [torch.empty(N, H // 4, W // 4, 4 * d, dtype=torch.float32, device=dev),
                     torch.empty(N, H // 2, W // 2, 2 * d, dtype=torch.float32, device=dev),
                     torch.empty(N, H, W, 2 * d, dtype=torch.float32, device=dev)]
            fp = _l.ptr_array([f.data_ptr() for f in feats])
        pp = _l.ptr_array([p.data_ptr() for p in params])
        _l.check(self.lib.dcpt_restormer_fwd(self.plan, pp, _p(packed), _p(inp), _p(out), _p(work), fp, int(bool(hook)),
                                             N, H, W, _stream()), "restormer_fwd")
        return out, feats

    def _graph_forward(self, params, packed, inp, hook, want_feats, key):
        N, _, H, W = inp.shape
        dev = inp.device
        ent = self._graphs.get(key)
        if ent is None:
            d = self.dim
            ent = {"inp": torch.empty_like(inp), "out": None if hook else torch.empty_like(inp), "feats": None, "graph": None,
                   "work": torch.empty(self.lib.dcpt_restormer_workspace_bytes(self.plan, N, H, W), dtype=torch.uint8, device=dev)}
            if want_feats:
                ent["feats"] = [torch.empty(N, H // 4, W // 4, 4 * d, dtype=torch.float32, device=dev),
                                torch.empty(N, H // 2, W // 2, 2 * d, dtype=torch.float32, device=dev),
                                torch.empty(N, H, W, 2 * d, dtype=torch.float32, device=dev)]
            self._graphs.put(key, ent)
        ent["inp"].copy_(inp)
        work = ent["work"]
        pp = _l.ptr_array([p.data_ptr() for p in params])
        fp = _l.ptr_array([f.data_ptr() for f in ent["feats"]]) if ent["feats"] else None

        def run():
            _l.check(self.lib.dcpt_restormer_fwd(self.plan, pp, _p(packed), _p(ent["inp"]), _p(ent["out"]), _p(work), fp, int(hook),
                                                 N, H, W, _stream()), "restormer_fwd")
        if ent["graph"] is None:
            run()                       # eager once: one-time initialisation inside the library must not be captured
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                run()
            ent["graph"] = g
        else:
            ent["graph"].replay()
        out = None if hook else ent["out"].clone()
        feats = [f.clone() for f in ent["feats"]] if ent["feats"] else None
        return out, feats

    # ---- training -----------------------------------------------------------------------------------------------
    def _feat_buffers(self, N, H, W, dev):
        d = self.dim                                     # decoder_level3, 2, 1 outputs, NHWC fp32
        return [torch.empty(N, H // 4, W // 4, 4 * d, dtype=torch.float32, device=dev),
                torch.empty(N, H // 2, W // 2, 2 * d, dtype=torch.float32, device=dev),
                torch.empty(N, H, W, 2 * d, dtype=torch.float32, device=dev)]

    def forward_train(self, params, inp, hook=False, want_feats=False):
        """Returns (out or None, feats or None, saved arena).  inp fp32 NCHW on CUDA."""
        self._check_params(params)
        inp = inp.contiguous().float()
        N, _, H, W = inp.shape
        dev = inp.device
        packed = self.packed_for(params)
        work = self._work.setdefault((N, H, W, dev), lambda: torch.empty(
            self.lib.dcpt_restormer_workspace_bytes(self.plan, N, H, W), dtype=torch.uint8, device=dev))
        saved = torch.empty(self.lib.dcpt_restormer_saved_bytes(self.plan, N, H, W), dtype=torch.uint8, device=dev)
        out = None if hook else torch.empty_like(inp)
        feats = self._feat_buffers(N, H, W, dev) if want_feats else None
        fp = _l.ptr_array([f.data_ptr() for f in feats]) if feats else None
        pp = _l.ptr_array([p.data_ptr() for p in params])
        _l.check(self.lib.dcpt_restormer_fwd_train(self.plan, pp, _p(packed), _p(inp), _p(out), _p(saved), _p(work), fp,
                                                   int(bool(hook)), N, H, W, _stream()), "restormer_fwd_train")
        return out, feats, saved

    def backward(self, params, inp, saved, dout, dfeats=None):
        """Parameter gradients (one flat fp32 buffer, views per parameter) of a forward_train call.  dout: fp32 NCHW or None
        (hook pass); dfeats: gradients of the decoder-level features (NHWC fp32, entries may be None) or None."""
        N, _, H, W = inp.shape
        dev = inp.device
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + 63) // 64 * 64
        flat = torch.zeros(off, dtype=torch.float32, device=dev)
        grads = [flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, params)]
        bwork = self._work.setdefault(("bwd", N, H, W, dev), lambda: torch.empty(
            self.lib.dcpt_restormer_bwd_workspace_bytes(self.plan, N, H, W), dtype=torch.uint8, device=dev))
        pp = _l.ptr_array([p.data_ptr() for p in params])
        gp = _l.ptr_array([g.data_ptr() for g in grads])
        dout = None if dout is None else dout.contiguous().float()
        dfe = dfp = None
        if dfeats and any(d is not None for d in dfeats):
            dfe = [None if d is None else d.contiguous().float() for d in dfeats]
            dfp = _l.ptr_array([0 if d is None else d.data_ptr() for d in dfe])
        if dout is None and dfp is None:
            return grads                                  # nothing reached this forward: all-zero gradients
        sc = fp16_grad_scale([dout] + (dfe or []))        # IEEE-half operand build only (params.py)
        if sc is not None:
            dout = None if dout is None else dout * sc[0]
            if dfp is not None:
                dfe = [None if d is None else d * sc[0] for d in dfe]
                dfp = _l.ptr_array([0 if d is None else d.data_ptr() for d in dfe])
        _l.check(self.lib.dcpt_restormer_bwd(self.plan, pp, _p(self.packed_for(params)), _p(saved), _p(inp), _p(dout), dfp,
                                             gp, _p(bwork), N, H, W, _stream()), "restormer_bwd")
        if sc is not None:
            flat.mul_(sc[1])
        if self.grad_sync is not None:
            self.grad_sync(flat)
        return grads


class _TrainSlot:
    """Static buffers and captured graphs of one (shape, hook, feats, parameter storage) training signature.  The slot owns
    the saved-activation arena: it is busy from a forward until its backward ran (DCPT: two forwards before one backward
    take two slots)."""

    MAX_PER_KEY = 3

    def __init__(self, eng, N, H, W, dev, hook, want_feats):
        self.N, self.H, self.W, self.hook = N, H, W, hook
        self.inp = torch.empty(N, 3, H, W, dtype=torch.float32, device=dev)
        self.out = None if hook else torch.empty_like(self.inp)
        self.saved = torch.empty(eng.lib.dcpt_restormer_saved_bytes(eng.plan, N, H, W), dtype=torch.uint8, device=dev)
        self.feats = eng._feat_buffers(N, H, W, dev) if want_feats else None
        # the slot owns the workspaces its captured graphs point to: evicting the slot (LRU) frees everything together
        self.work = torch.empty(eng.lib.dcpt_restormer_workspace_bytes(eng.plan, N, H, W), dtype=torch.uint8, device=dev)
        self.bwork = None
        self.dout = self.dfeats = self.flat = self.shapes = self.gp = None
        self.graphs = {}            # "fwd" / backward mask -> captured graph
        self.busy = False
        self.owner = None

    def acquire(self):
        """Marks the slot busy for one forward; the returned token releases it only while that forward still owns it (the
        finalizer of an OLD autograd node may fire after a newer forward has taken the slot)."""
        self.busy, self.owner = True, object()
        return self.owner

    def release(self, token=None):
        if token is None or token is self.owner:
            self.busy, self.owner = False, None


def _run_or_replay(eng, slot, key, run):
    """First use: launch eagerly (one-time initialisation inside the library must not be captured; the results of this run
    are the ones used), then capture for the following steps.  If the runtime refuses the capture, keep launching eagerly."""
    g = slot.graphs.get(key)
    if g is not None:
        g.replay()
        return
    run()
    if not eng.use_train_graphs:
        return
    try:
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, capture_error_mode="thread_local"):
            run()
        slot.graphs[key] = g
    except RuntimeError as e:
        warnings.warn(f"dcpt_b200: CUDA-graph capture of the Restormer training path failed ({e}); launching eagerly")
        eng.use_train_graphs = False


def _train_graph_forward(eng, params, inp, hook, want_feats):
    """(out, feats, slot), or None when no slot is free / capture is unavailable (caller launches eagerly)."""
    eng._check_params(params)
    N, _, H, W = inp.shape
    dev = inp.device
    key = (N, H, W, dev, hook, want_feats, tuple(p.data_ptr() for p in params))
    slots = eng._tslots.setdefault(key, list)
    slot = next((s_ for s_ in slots if not s_.busy), None)
    if slot is None:
        if len(slots) >= _TrainSlot.MAX_PER_KEY:
            return None
        slot = _TrainSlot(eng, N, H, W, dev, hook, want_feats)
        slots.append(slot)
    packed = eng.packed_for(params)                       # re-packed eagerly when a parameter changed; static address
    work = slot.work
    slot.inp.copy_(inp)
    pp = _l.ptr_array([p.data_ptr() for p in params])
    fp = _l.ptr_array([f.data_ptr() for f in slot.feats]) if slot.feats else None

    def run():
        _l.check(eng.lib.dcpt_restormer_fwd_train(eng.plan, pp, _p(packed), _p(slot.inp), _p(slot.out), _p(slot.saved), _p(work), fp,
                                                  int(hook), N, H, W, _stream()), "restormer_fwd_train")
    token = slot.acquire()
    _run_or_replay(eng, slot, "fwd", run)
    out = None if hook else slot.out.clone()
    feats = [f.clone() for f in slot.feats] if slot.feats else None
    return out, feats, slot, token


def _train_graph_backward(eng, params, slot, dout, dfeats):
    N, H, W = slot.N, slot.H, slot.W
    dev = slot.inp.device
    if slot.flat is None:
        offs, off = [], 0
        for p in params:
            offs.append(off)
            off += (p.numel() + 63) // 64 * 64
        slot.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        slot.shapes = [(o, p.numel(), p.shape) for o, p in zip(offs, params)]
        slot.gp = _l.ptr_array([slot.flat[o:o + n].data_ptr() for o, n, _ in slot.shapes])
    has_df = bool(dfeats) and any(d is not None for d in dfeats)
    mask = (dout is not None, tuple(d is not None for d in dfeats) if has_df else None)
    sc = fp16_grad_scale([dout] + (list(dfeats) if has_df else []))   # IEEE-half operand build only (params.py)
    if mask[0]:
        if slot.dout is None:
            slot.dout = torch.empty_like(slot.inp)
        slot.dout.copy_(dout)
        if sc is not None:
            slot.dout.mul_(sc[0])
    if has_df:
        if slot.dfeats is None:
            slot.dfeats = [torch.empty_like(f) for f in slot.feats]
        for s_, d in zip(slot.dfeats, dfeats):
            if d is not None:
                s_.copy_(d)
                if sc is not None:
                    s_.mul_(sc[0])
    if not mask[0] and not has_df:
        slot.release()
        flat = torch.zeros_like(slot.flat)                 # nothing reached this forward: all-zero gradients
        return [flat[o:o + n].view(shp) for o, n, shp in slot.shapes]
    if slot.bwork is None:
        slot.bwork = torch.empty(eng.lib.dcpt_restormer_bwd_workspace_bytes(eng.plan, N, H, W), dtype=torch.uint8, device=dev)
    work = slot.bwork
    packed = eng.packed_for(params)
    pp = _l.ptr_array([p.data_ptr() for p in params])
    dfp = _l.ptr_array([s_.data_ptr() if m else 0 for s_, m in zip(slot.dfeats, mask[1])]) if has_df else None
    dptr = _p(slot.dout) if mask[0] else None

    def run():
        slot.flat.zero_()
        _l.check(eng.lib.dcpt_restormer_bwd(eng.plan, pp, _p(packed), _p(slot.saved), _p(slot.inp), dptr, dfp, slot.gp, _p(work),
                                            N, H, W, _stream()), "restormer_bwd")
    _run_or_replay(eng, slot, mask, run)
    slot.release()
    if sc is not None:
        slot.flat.mul_(sc[1])
    if eng.grad_sync is not None:
        eng.grad_sync(slot.flat)                           # data-parallel wrapper: ONE mean all-reduce of the flat buffer
    flat = slot.flat.clone()                               # autograd owns the returned gradients; the slot's buffer is reused
    return [flat[o:o + n].view(shp) for o, n, shp in slot.shapes]


class _RestormerFunction(torch.autograd.Function):
    """out, feat_0..2 = Restormer(inp; params).  feats (decoder_level3, 2, 1) are NHWC storage viewed as logical NCHW."""

    @staticmethod
    def forward(ctx, engine, inp, hook, want_feats, dead, *params):
        ctx.set_materialize_grads(False)                  # unused outputs (e.g. the pixel pass's features) arrive as None
        dparams = [p.detach() for p in params]
        inp_c = inp.detach().contiguous().float()
        res = None
        with training_pass():       # restormer_apply is only reached when gradients are needed (no fingerprint sync, params.py)
            if engine.use_train_graphs and inp_c.is_cuda and not torch.cuda.is_current_stream_capturing():
                res = _train_graph_forward(engine, dparams, inp_c, hook, want_feats)
            if res is not None:
                out, feats, saved, token = res
                try:
                    weakref.finalize(ctx, saved.release, token)   # a forward whose backward never runs must not pin the slot
                except TypeError:
                    pass
            else:
                out, feats, saved = engine.forward_train(dparams, inp_c, hook=hook, want_feats=want_feats)
        ctx.engine, ctx.inp, ctx.saved, ctx.params, ctx.hook, ctx.n_feats = engine, inp_c, saved, dparams, hook, len(feats) if feats else 0
        ctx.dead = frozenset(dead) if hook else frozenset()
        outs = []
        if out is None:
            out = inp_c.new_zeros(())                     # placeholder (hook=True returns None to the caller)
            ctx.mark_non_differentiable(out)
        outs.append(out)
        if feats:
            outs.extend(f.permute(0, 3, 1, 2) for f in feats)
        return tuple(outs)

    @staticmethod
    def backward(ctx, dout, *dfeats):
        dfe = [None if d is None else d.permute(0, 2, 3, 1) for d in dfeats] if ctx.n_feats else None
        with training_pass():
            if isinstance(ctx.saved, _TrainSlot):
                grads = _train_graph_backward(ctx.engine, ctx.params, ctx.saved, None if ctx.hook else dout, dfe)
            else:
                grads = ctx.engine.backward(ctx.params, ctx.inp, ctx.saved, None if ctx.hook else dout, dfe)
        ctx.saved = None
        # a hook pass stops after decoder_level1 (restormer_arch.py:403): refinement / output get NO gradient (None), as in
        # the reference; the input image receives none either (as for NAFNet: it is data)
        return (None, None, None, None, None) + tuple(None if i in ctx.dead else g for i, g in enumerate(grads))


def restormer_apply(engine, inp, params, hook=False, want_feats=False, dead=()):
    """dead: indices (named_parameters() order) of the parameters a hook=True pass never reaches."""
    res = _RestormerFunction.apply(engine, inp, bool(hook), bool(want_feats), tuple(dead), *params)
    return (None if hook else res[0]), list(res[1:])
