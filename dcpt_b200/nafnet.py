"""Host-side engine for NAFNetBaseline (reference: basicsr/archs/nafnet_arch.py:189-274).

``NAFNetEngine`` owns the C plan, the packed bf16 weight cache and the per-call arenas; the
``basicsr`` mirror's ``NAFNetBaseline`` module calls it through ``nafnet_apply`` (an
``autograd.Function``), so ``loss.backward()``, DDP and optimizers work as with the reference.
"""
import ctypes as C
import operator
import os
import weakref

import torch

from . import lib as _l
from .ops import _p, _stream
from .params import LRUCache, PackedCacheKey, fp16_grad_scale, training_pass


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
        self.blk_nums = (tuple(enc_blk_nums), int(middle_blk_num), tuple(dec_blk_nums))
        self.split_event = None     # data-parallel overlap (enable_grad_overlap)
        self.num_params = self.lib.dcpt_nafnet_num_params(self.plan)
        self.shapes = []
        dims = (C.c_int * 4)()
        for i in range(self.num_params):
            self.lib.dcpt_nafnet_param_shape(self.plan, i, dims)
            self.shapes.append(tuple(dims))
        self._packed = None
        self._packed_key = PackedCacheKey()
        # per-shape resources are LRU-bounded (params.py): workspaces / inference arenas of the eager path, and the CUDA-graph
        # slots (each owns a saved-activation arena, ~12 KB per pixel at width 64) of the autograd path
        self._scratch = LRUCache()
        # CUDA-graph replay of the autograd path (nafnet_apply): DCPT_CUDA_GRAPH=0 launches every kernel from the host
        self.use_graphs = os.getenv("DCPT_CUDA_GRAPH", "1") != "0"
        self._gslots = LRUCache(can_evict=lambda slots: not any(s.busy for s in slots))
        self._seen_nograd = LRUCache(cap=64)
        self.grad_sync = None      # set by dcpt_b200.dist.FlatGradDataParallel: callable(flat fp32 gradient buffer)
        self.tlc = False

    def set_hook_blocks(self, idx):
        """idx[i]: block of decoder level i whose output is the hooked feature (-1 = the level's output)."""
        idx = tuple(int(v) for v in idx)
        if idx == getattr(self, "_hook_blocks", None):
            return
        arr = (C.c_int * max(len(idx), 1))(*idx)
        _l.check(self.lib.dcpt_nafnet_set_hook_blocks(self.plan, arr, len(idx)), "nafnet_set_hook_blocks")
        self._hook_blocks = idx
        self._gslots.clear()            # captured graphs baked the old feature taps in

    # ---- data-parallel overlap ----------------------------------------------------------------
    def enable_grad_overlap(self, on=True, external=True):
        """The backward records ``self.split_event`` once the gradients of the deepest encoder level, the middle blocks, the up
        convs, the decoders and the ending conv are final (dcpt_nafnet_set_bwd_split_event); dcpt_b200.dist starts the
        all-reduce of that slice (``early_grad_range``) on a second stream while the shallower levels are still differentiated."""
        """external: inside a stream capture the record becomes an external-event node (the all-reduce is issued OUTSIDE the
        captured graph, as FlatGradDataParallel does); external=False when the all-reduce is captured into the same graph."""
        if on:
            if self.split_event is None:
                ev = torch.cuda.Event()
                ev.record()                                    # instantiates the cudaEvent_t handle
                self.split_event = ev
            _l.check(self.lib.dcpt_nafnet_set_bwd_split_event(self.plan, C.c_void_p(self.split_event.cuda_event), int(external)),
                     "set_bwd_split_event")
            self._gslots.clear()                               # captured backward graphs do not contain the record node yet
        elif self.split_event is not None:
            _l.check(self.lib.dcpt_nafnet_set_bwd_split_event(self.plan, None, 1), "set_bwd_split_event")
            self.split_event = None
            self._gslots.clear()

    def early_grad_range(self, align=64):
        """[a, b) element range, in the flat gradient buffer of ``alloc_flat_grads`` / the graph slots, of the parameters whose
        gradients are final when ``split_event`` fires: encoders.{n_enc-1}.*, middle_blks.*, ups.* - contiguous in
        named_parameters() order (nafnet_arch.py:202-248) and ~90 % of NAFNet-w64's 67.9 M parameters."""
        enc, mid, dec = self.blk_nums
        first = 4 + 18 * sum(enc[:-1]) if enc else 4
        last = 4 + 18 * (sum(enc) + mid) + len(dec)          # exclusive: first downs.* parameter
        offs, off = [], 0
        for shp in self.shapes:
            offs.append(off)
            n = 1
            for d in shp:
                n *= d
            off += (n + align - 1) // align * align
        offs.append(off)
        return offs[first], offs[last]

    def set_tlc(self, kernels):
        """kernels: [(kh, kw)] per resolution level (0 = full resolution .. n_enc) - NAFNet's test-time local converter."""
        n = len(kernels)
        kh = (C.c_int * max(n, 1))(*[k[0] for k in kernels])
        kw = (C.c_int * max(n, 1))(*[k[1] for k in kernels])
        _l.check(self.lib.dcpt_nafnet_set_tlc(self.plan, kh, kw, n), "nafnet_set_tlc")
        self.tlc = n > 0
        self._gslots.clear()

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
        """bf16 operand cache; refreshed whenever a parameter was modified or replaced (dcpt_b200/params.py: version
        counters, plus a device fingerprint of the weights on no-grad forwards for writes through ``p.data``)."""
        if self._packed is None or self._packed.device != params[0].device:
            self._packed = torch.empty(self.lib.dcpt_nafnet_packed_bytes(self.plan), dtype=torch.uint8,
                                       device=params[0].device)
            self._packed_key.invalidate()
        if self._packed_key.stale(params):
            pp = _l.ptr_array([p.data_ptr() for p in params])
            _l.check(self.lib.dcpt_nafnet_pack(self.plan, pp, _p(self._packed), _stream()), "nafnet_pack")
        return self._packed

    def invalidate_packed(self):
        """Force a re-pack on the next call (for writers that bypass both torch's version counters and no-grad forwards)."""
        self._packed_key.invalidate()

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
            saved = self._scratch.setdefault(("saved", N, H, W, dev), lambda: torch.empty(nbytes, dtype=torch.uint8, device=dev))
        out = None if hook else torch.empty_like(inp)
        feats = fp = None
        if want_feats:
            feats, c, h, w = [], self.width << self.n_enc, H >> self.n_enc, W >> self.n_enc
            for _ in range(self.n_dec):
                c, h, w = c // 2, h * 2, w * 2
                feats.append(torch.empty(N, h, w, c, dtype=torch.float32, device=dev))
            fp = _l.ptr_array([f.data_ptr() for f in feats])
        pp = _l.ptr_array([p.data_ptr() for p in params])
        # inference passes skip the stores only a backward would read (dcpt_nafnet_set_keep_activations)
        _l.check(self.lib.dcpt_nafnet_set_keep_activations(self.plan, _keep(keep_for_backward)), "nafnet_set_keep_activations")
        _l.check(self.lib.dcpt_nafnet_fwd(self.plan, pp, _p(packed), _p(inp), _p(out), _p(saved), fp, int(bool(hook)), N, H,
                                          W, _stream()), "nafnet_fwd")
        return out, feats, saved

    def backward(self, params, inp, saved, dout, dfeats=None, grads=None):
        """Accumulates parameter gradients into ``grads`` (list of fp32 tensors; allocated zeroed if None)."""
        N, _, H, W = inp.shape
        dev = inp.device
        flat = None
        if grads is None:
            flat, grads = self.alloc_flat_grads(params)
        work = self._scratch.setdefault(("work", N, H, W, dev), lambda: torch.empty(
            self.lib.dcpt_nafnet_workspace_bytes(self.plan, N, H, W), dtype=torch.uint8, device=dev))
        pp = _l.ptr_array([p.data_ptr() for p in params])
        gp = _l.ptr_array([g.data_ptr() for g in grads])
        dfp = None
        if dfeats is not None and any(d is not None for d in dfeats):
            dfeats = [None if d is None else d.contiguous().float() for d in dfeats]
            dfp = _l.ptr_array([0 if d is None else d.data_ptr() for d in dfeats])
        if dout is not None:
            dout = dout.contiguous().float()
        sc = fp16_grad_scale([dout] + (dfeats or []))            # IEEE-half operand build only (params.py)
        acc_into = None
        if sc is not None:
            if dout is not None:
                dout = dout * sc[0]
            if dfp is not None:
                dfeats = [None if d is None else d * sc[0] for d in dfeats]
                dfp = _l.ptr_array([0 if d is None else d.data_ptr() for d in dfeats])
            if flat is None:                                     # caller's buffers accumulate: scale back a private copy
                acc_into = grads
                flat, grads = self.alloc_flat_grads(params)
                gp = _l.ptr_array([g.data_ptr() for g in grads])
        packed = self.packed_for(params)
        _l.check(self.lib.dcpt_nafnet_bwd(self.plan, pp, _p(packed), _p(saved), _p(inp), _p(dout), dfp, gp, _p(work), N, H,
                                          W, _stream()), "nafnet_bwd")
        if sc is not None:
            flat.mul_(sc[1])
            if acc_into is not None:
                torch._foreach_add_(acc_into, grads)
                return acc_into
        if flat is not None and self.grad_sync is not None:
            self.grad_sync(flat)            # data-parallel wrapper (dcpt_b200.dist.FlatGradDataParallel): one all-reduce
        return grads


class _ParamView:
    """Per-engine cache of everything derived from the parameter list that is costly to rebuild every step for 664
    tensors (detached views, pointer table, storage key): ~2.5 ms of host time per forward otherwise, during which the
    GPU idles before the graph launch.  Valid while the same Parameter objects keep their storage (in-place optimizer
    updates and load_state_dict do; .to() / .cuda() re-allocate and are detected through the probe pointers)."""

    def __init__(self, eng, params):
        self.params = list(params)
        self.n = len(params)
        self.dparams = [p.detach() for p in params]
        eng._check_params(self.dparams)
        self.ptrs = tuple(p.data_ptr() for p in self.dparams)
        self.probe = (self.ptrs[0], self.ptrs[self.n // 2], self.ptrs[-1])
        self.pp = _l.ptr_array(list(self.ptrs))

    def matches(self, params):
        return (len(params) == self.n and all(map(operator.is_, params, self.params))
                and (params[0].data_ptr(), params[self.n // 2].data_ptr(), params[-1].data_ptr()) == self.probe)


def _param_view(eng, params):
    pv = getattr(eng, "_pview", None)
    if pv is None or not pv.matches(params):
        pv = eng._pview = _ParamView(eng, list(params))
    return pv


class _GraphSlot:
    """Static buffers and captured CUDA graphs of one (shape, hook, feats, parameter storage) signature.

    A whole forward (~140 launches) or backward (~360 launches) of the w64 network is one graph launch: the host cost of
    encoding the TMA descriptors and launching each kernel (several ms per step, during which the GPU idles) is paid once at
    capture.  The slot owns the saved-activation arena, so it is ``busy`` from a forward that needs gradients until its
    backward ran (DCPT runs two forwards before one backward: it gets two slots)."""

    MAX_PER_KEY = 3

    def __init__(self, eng, params, N, H, W, dev, hook, want_feats):
        self.N, self.H, self.W, self.hook = N, H, W, hook
        self.inp = torch.empty(N, 3, H, W, dtype=torch.float32, device=dev)
        self.out = None if hook else torch.empty_like(self.inp)
        self.saved = torch.empty(eng.lib.dcpt_nafnet_saved_bytes(eng.plan, N, H, W), dtype=torch.uint8, device=dev)
        self.feats = None
        if want_feats:
            self.feats, c, h, w = [], eng.width << eng.n_enc, H >> eng.n_enc, W >> eng.n_enc
            for _ in range(eng.n_dec):
                c, h, w = c // 2, h * 2, w * 2
                self.feats.append(torch.empty(N, h, w, c, dtype=torch.float32, device=dev))
        self.dout = None
        self.dfeats = None
        self.flat = self.grads = None
        self.work = None     # backward workspace: owned by the slot, so that evicting the slot frees everything its graphs point to
        self.fgraph = None
        self.bgraphs = {}
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


def _graph_forward(eng, pv, inp, hook, want_feats, need_grad):
    """Returns (out, feats, slot) or None when no slot is free (caller falls back to eager launches)."""
    params = pv.dparams
    inp = inp.contiguous().float()
    N, _, H, W = inp.shape
    # (a no-grad slot's captured forward does not store what only a backward reads: never shared with a gradient-needing forward)
    key = (N, H, W, inp.device, bool(hook), bool(want_feats), pv.ptrs, bool(need_grad))
    if not need_grad and key not in eng._gslots:
        # inference over images of many sizes (validation sets): a shape earns a graph slot (static buffers + a saved-activation
        # arena + a capture) only when it comes back; the first sighting launches eagerly through the shared scratch arena
        if eng._seen_nograd.get(key) is None:
            eng._seen_nograd.put(key, True)
            return None
    slots = eng._gslots.setdefault(key, list)
    slot = next((s for s in slots if not s.busy), None)
    if slot is None:
        if len(slots) >= _GraphSlot.MAX_PER_KEY:
            return None
        slot = _GraphSlot(eng, params, N, H, W, inp.device, bool(hook), bool(want_feats))
        slots.append(slot)
    packed = eng.packed_for(params)   # re-packed eagerly when a parameter changed; its address is static
    slot.inp.copy_(inp)
    pp = pv.pp
    fp = _l.ptr_array([f.data_ptr() for f in slot.feats]) if slot.feats else None

    def run():
        _l.check(eng.lib.dcpt_nafnet_set_keep_activations(eng.plan, _keep(need_grad)), "nafnet_set_keep_activations")
        _l.check(eng.lib.dcpt_nafnet_fwd(eng.plan, pp, _p(packed), _p(slot.inp), _p(slot.out), _p(slot.saved), fp, int(slot.hook),
                                         N, H, W, _stream()), "nafnet_fwd")
    if slot.fgraph is None:
        run()                           # eager once: lazy one-time initialisation inside the library must not be captured
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=_capture_stream(slot.inp.device), capture_error_mode="thread_local"):
            run()
        slot.fgraph = g
    else:
        slot.fgraph.replay()
    slot.token = slot.acquire() if need_grad else None
    out = None if hook else slot.out.clone()
    feats = [f.clone() for f in slot.feats] if slot.feats else None
    return out, feats, slot


def _graph_backward(eng, pv, slot, dout, dfeats):
    params = pv.dparams
    N, H, W = slot.N, slot.H, slot.W
    dev = slot.inp.device
    if slot.flat is None:
        slot.offs, off = [], 0
        for p in params:
            slot.offs.append(off)
            off += (p.numel() + 63) // 64 * 64
        slot.flat = torch.zeros(off, dtype=torch.float32, device=dev)
        slot.grads = [slot.flat[o:o + p.numel()].view(p.shape) for o, p in zip(slot.offs, params)]
        slot.gp = _l.ptr_array([g.data_ptr() for g in slot.grads])
        slot.shapes = [(o, p.numel(), p.shape) for o, p in zip(slot.offs, params)]
    mask = (dout is not None, tuple(d is not None for d in dfeats) if dfeats else None)
    sc = fp16_grad_scale([dout] + list(dfeats or []))            # IEEE-half operand build only (params.py)
    if dout is not None:
        if slot.dout is None:
            slot.dout = torch.empty_like(slot.inp)
        slot.dout.copy_(dout)
        if sc is not None:
            slot.dout.mul_(sc[0])
    if dfeats and any(d is not None for d in dfeats):
        if slot.dfeats is None:
            slot.dfeats = [torch.empty_like(f) for f in slot.feats]
        for s_, d in zip(slot.dfeats, dfeats):
            if d is not None:
                s_.copy_(d)
                if sc is not None:
                    s_.mul_(sc[0])
    if slot.work is None:
        slot.work = torch.empty(eng.lib.dcpt_nafnet_workspace_bytes(eng.plan, N, H, W), dtype=torch.uint8, device=dev)
    work = slot.work
    packed = eng.packed_for(params)
    pp, gp = pv.pp, slot.gp
    dfp = None
    if mask[1] and any(mask[1]):
        dfp = _l.ptr_array([s_.data_ptr() if m else 0 for s_, m in zip(slot.dfeats, mask[1])])
    dptr = _p(slot.dout) if mask[0] else None

    def run():
        slot.flat.zero_()
        _l.check(eng.lib.dcpt_nafnet_bwd(eng.plan, pp, _p(packed), _p(slot.saved), _p(slot.inp), dptr, dfp, gp, _p(work), N, H, W,
                                         _stream()), "nafnet_bwd")
    if mask not in slot.bgraphs:
        run()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=_capture_stream(slot.inp.device), capture_error_mode="thread_local"):
            run()
        slot.bgraphs[mask] = g
    else:
        slot.bgraphs[mask].replay()
    slot.release()
    if sc is not None:
        slot.flat.mul_(sc[1])
    if eng.grad_sync is not None:
        eng.grad_sync(slot.flat)            # data-parallel wrapper: ONE mean all-reduce of the flat gradient buffer
    # A second backward node of the same step (DCPT: pixel pass + hooked pass through one net_g): autograd would add its 664
    # gradients into the first node's one tensor at a time (664 tiny kernels and ~5 ms of autograd-thread time).  When every
    # parameter's .grad still IS the view of the flat buffer the first node handed out, ONE add over the flat buffer does the
    # same and the node reports "no gradient" for the parameters (None), so autograd has nothing left to accumulate.
    acc = getattr(eng, "_acc", None)
    if acc is not None and _accum_in_place() and acc[0] is pv and _grads_alias(pv, acc[1], slot.shapes):
        acc[1].add_(slot.flat)
        return [None] * len(slot.shapes)
    flat = slot.flat.clone()            # autograd owns the returned gradients; the slot's buffer is rewritten next step
    eng._acc = (pv, flat)
    return [flat[o:o + n].view(shp) for o, n, shp in slot.shapes]


def _accum_in_place():
    # OFF by default: measured neutral on the C4 step (46.57 vs 46.59, 47.07 vs 46.96 ms same-box) - autograd's per-tensor
    # accumulation runs on the autograd thread while the GPU is busy with the second node's graph
    return os.getenv("DCPT_GRAD_ACCUM_FLAT", "0") == "1"


def _grads_alias(pv, flat, shapes):
    """True when every Parameter's .grad is exactly the view of `flat` a previous backward node returned for it (autograd kept
    the views: nothing cloned, replaced, zeroed to None or re-pointed since)."""
    base, esz = flat.data_ptr(), flat.element_size()
    for p, (o, n, _) in zip(pv.params, shapes):
        g = p.grad
        if g is None or g.data_ptr() != base + o * esz or g.numel() != n or g.dtype != flat.dtype:
            return False
    return True


def _keep(need_grad):
    """1 = training forward (every saved tensor is written); DCPT_INFER_KEEP=1 forces it for inference passes too (A/B switch)."""
    return int(bool(need_grad) or os.getenv("DCPT_INFER_KEEP", "0") == "1")


_CAPTURE_STREAMS = {}


def _capture_stream(dev):
    """Stream the forward / backward graphs are captured on: a HIGH-priority stream, so that the captured chain kernels outrank
    the library's weight-gradient side stream (default priority - the lowest CUDA has) whenever both have thread blocks
    pending; the side stream then only fills SM time the chain leaves idle (r03g, same box: 21.07-21.26 -> 20.91 ms per
    step).  DCPT_GRAPH_PRIORITY=0 captures on torch's default capture stream."""
    if os.getenv("DCPT_GRAPH_PRIORITY", "1") == "0":
        return None
    key = torch.device(dev).index if torch.device(dev).index is not None else torch.cuda.current_device()
    if key not in _CAPTURE_STREAMS:
        _CAPTURE_STREAMS[key] = torch.cuda.Stream(device=key, priority=-1)
    return _CAPTURE_STREAMS[key]


class _null_ctx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _NAFNetFunction(torch.autograd.Function):
    """out, feat_0..feat_{n-1} = NAFNet(inp; params).  feats are NHWC storage viewed as logical NCHW."""

    @staticmethod
    def forward(ctx, engine, inp, hook, want_feats, need_grad, *params):
        # need_grad is decided by the caller: grad mode is always OFF inside Function.forward, and a forward whose
        # activations are needed by a later backward must own its saved-activation arena (DCPT runs two forwards
        # before one backward).
        ctx.set_materialize_grads(False)  # unused outputs (the pixel pass's decoder features in DCPT) arrive as None, not zeros
        pv = _param_view(engine, params)
        dparams = pv.dparams
        inp_c = inp.detach().contiguous().float()
        res = None
        with (training_pass() if need_grad else _null_ctx()):   # (no weight-fingerprint sync on training passes, params.py)
            if engine.use_graphs and inp_c.is_cuda and not torch.cuda.is_current_stream_capturing():
                res = _graph_forward(engine, pv, inp_c, hook, want_feats, need_grad)
            if res is not None:
                out, feats, saved = res
                if need_grad:
                    try:
                        weakref.finalize(ctx, saved.release, saved.token)   # a forward whose backward never runs must not pin the slot
                    except TypeError:
                        pass
            else:
                out, feats, saved = engine.forward(dparams, inp_c, hook=hook, want_feats=want_feats, keep_for_backward=need_grad)
        ctx.engine, ctx.hook, ctx.n_feats = engine, hook, len(feats) if feats else 0
        ctx.inp, ctx.saved, ctx.params, ctx.pv = inp_c, saved, dparams, pv
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
        with torch.cuda.device(ctx.inp.device), training_pass():
            if isinstance(ctx.saved, _GraphSlot):
                grads = _graph_backward(eng, ctx.pv, ctx.saved, None if ctx.hook else dout, dfe)
            else:
                grads = eng.backward(ctx.params, ctx.inp, ctx.saved, None if ctx.hook else dout, dfe)
        ctx.saved = None
        return (None, None, None, None, None) + tuple(grads)


def nafnet_apply(engine, inp, params, hook=False, want_feats=False):
    need_grad = torch.is_grad_enabled() and (inp.requires_grad or any(p.requires_grad for p in params))
    if inp.is_cuda and params[0].device != inp.device:
        raise _l.DcptError(f"input on {inp.device} but the network's parameters on {params[0].device}")
    # the C library launches on the CURRENT device: make it the tensors' device for the forward and, through autograd's
    # device guard of the node, for the backward (a model living on a non-current device)
    with torch.cuda.device(inp.device) if inp.is_cuda else _null_ctx():
        res = _NAFNetFunction.apply(engine, inp, hook, want_feats, need_grad, *params)
    out = None if hook else res[0]
    return out, list(res[1:])
