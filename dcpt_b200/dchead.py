"""Host-side engine for the degradation-classifier head ``PromptIR_NoImg_DC`` (reference:
basicsr/archs/degrad_classify_arch.py:558-641) on the sm_100a kernels.

Trunk tensors are bf16 NHWC; every convolution runs on the tcgen05 GEMM engine (1x1: ``dcpt_gemm_bf16``; dense 3x3:
``dcpt_conv3x3_fwd`` implicit GEMM; weight gradients: split-K MN-major GEMM / ``dcpt_conv3x3_wgrad``); LayerNorm+ReLU,
feature mixing, max-pool and the pooled classifier are CUDA-core kernels of ``csrc/dchead.cu``.  The Python here only
sequences C-ABI calls and owns the buffers (the reference's host side is Python too); ``dchead_apply`` wraps a whole
forward / backward in ONE ``autograd.Function`` so it composes with the NAFNet function under ``loss.backward()``.
"""
import ctypes as C
import os
import weakref

import torch

from . import lib as _l
from .ops import _p, _stream, gemm
from .params import LRUCache, PackedCacheKey, fp16_grad_scale, training_pass



def _bottleneck_names(prefix):
    return [prefix + n for n in ("conv1.weight", "conv1.norm.weight", "conv1.norm.bias", "conv2.weight", "conv2.norm.weight",
                                 "conv2.norm.bias", "conv3.weight", "conv3.norm.weight", "conv3.norm.bias")]


class DCHeadEngine:
    def __init__(self, feature_dims, num_res_blocks=2, num_classes=3, img_embed=False):
        """img_embed: the `PromptIR_DC` variant (degrad_classify_arch.py:480-556): the trunk starts from
        conv_embed(lq) = LayerNorm(Conv2d(3, f0, 7, stride 2, pad 3)(lq)) instead of nothing, so features[i] live at
        H/2^(i+1) (token backbones); `PromptIR_NoImg_DC` (:558-641) ignores lq."""
        self.lib = _l.load_library()
        self.dims = list(feature_dims)
        self.nb = num_res_blocks
        self.k = num_classes
        self.embed = bool(img_embed)
        self.grad_sync = None      # set by dcpt_b200.dist.FlatGradDataParallel: callable(flat fp32 gradient buffer)
        # parameter order == reference named_parameters() (degrad_classify_arch.py:497-547 / :577-620)
        names = ["mixing_weights"]
        if self.embed:
            names += ["conv_embed.0.weight", "conv_embed.0.bias", "conv_embed.1.weight", "conv_embed.1.bias"]
        for i in range(len(self.dims)):
            for j in range(self.nb):
                names += _bottleneck_names(f"bottleneck_layers.{i}.{j}.")
        names += [f"downsample_layers.{i}.0.weight" for i in range(len(self.dims))]
        for j in range(self.nb):
            names += _bottleneck_names(f"last_stage.{j}.")
        names += ["fc.weight", "fc.bias"]
        self.names = names
        self.index = {n: i for i, n in enumerate(names)}
        self._packed = None
        self._packed_key = PackedCacheKey()
        # CUDA-graph replay of the whole head forward / backward (192 launches through ctypes per step otherwise: measured 10.8 ms
        # of host enqueue for 11.2 ms of device time at the C4 size - the GPU waits for Python); DCPT_DCHEAD_GRAPH=0: eager
        self.use_graphs = os.getenv("DCPT_CUDA_GRAPH", "1") != "0" and os.getenv("DCPT_DCHEAD_GRAPH", "1") != "0"
        self._gslots = LRUCache(can_evict=lambda s_: not s_.busy)
        self._seen = LRUCache(cap=64)

    # ---- packed bf16 operand cache ---------------------------------------------------------------
    def _pack(self, params):
        if self._packed is not None and not self._packed_key.stale(params):  # dcpt_b200/params.py
            return self._packed
        dev = params[0].device
        P = lambda n: params[self.index[n]]
        # re-packing after a parameter update writes into the SAME tensors (captured graphs hold their addresses)
        old = self._packed if (self._packed is not None and next(iter(self._packed.values()))[0].device == dev) else {}
        pk = old

        def buf(name, k, n):
            if name in old and old[name][k] is not None and old[name][k].numel() == n:
                return old[name][k]
            return torch.empty(n, dtype=_l.operand_dtype(), device=dev)

        def mat(name):
            w = P(name)
            O, I = w.shape[0], w.shape[1]
            a = buf(name, 0, O * I).view(O, I)
            b = buf(name, 1, O * I).view(I, O)
            _l.check(self.lib.dcpt_pack_matrix(_p(w), _p(a), O, I, 0, _stream()), "pack_matrix")
            _l.check(self.lib.dcpt_pack_matrix(_p(w), _p(b), O, I, 1, _stream()), "pack_matrix")
            pk[name] = (a, b)

        def conv3(name):
            w = P(name)
            O, I = w.shape[0], w.shape[1]
            a = buf(name, 0, self.lib.dcpt_conv3x3_packed_elems(O, I, 0))
            b = buf(name, 1, self.lib.dcpt_conv3x3_packed_elems(O, I, 1))
            _l.check(self.lib.dcpt_conv3x3_pack(_p(w), _p(a), O, I, 0, _stream()), "conv3x3_pack")
            _l.check(self.lib.dcpt_conv3x3_pack(_p(w), _p(b), O, I, 1, _stream()), "conv3x3_pack")
            pk[name] = (a, b)

        for n in self.names:
            if n.endswith("conv2.weight"):
                conv3(n)
            elif n.endswith("conv1.weight") or n.endswith("conv3.weight") or n.startswith("downsample_layers"):
                mat(n)
        if self.embed:     # conv_embed weight as a GEMM operand [f0, 160]: weight.view(f0, 147) zero-padded (dcpt_im2col7x7s2 columns)
            w = P("conv_embed.0.weight")
            wp = torch.zeros(w.shape[0], 160, dtype=torch.float32, device=dev)
            wp[:, :147] = w.reshape(w.shape[0], 147)
            a = buf("conv_embed.0.weight", 0, w.shape[0] * 160).view(w.shape[0], 160)
            _l.check(self.lib.dcpt_pack_matrix(_p(wp), _p(a), w.shape[0], 160, 0, _stream()), "pack_matrix")
            pk["conv_embed.0.weight"] = (a, None)
        self._packed = pk
        return pk

    # ---- one BottleneckBlock (degrad_classify_arch.py:227-243) ------------------------------------
    def _ln(self, x, w, b, resid, relu):
        M, Cc = x.shape
        y = torch.empty_like(x)
        stats = torch.empty(M, 2, dtype=torch.float32, device=x.device)
        _l.check(self.lib.dcpt_ln_act_fwd(_p(x), _p(w), _p(b), _p(resid), _p(y), _p(stats), M, Cc, int(relu), 1e-6, _stream()),
                 "ln_act_fwd")
        return y, stats

    def _ln_bwd(self, dy, y, x, stats, w, gw, gb, want_dres, relu=True):
        M, Cc = x.shape
        dx = torch.empty_like(x)                                   # bf16: operand of the following dgrad / wgrad GEMMs
        dres = torch.empty(M, Cc, dtype=torch.float32, device=x.device) if want_dres else None
        _l.check(self.lib.dcpt_ln_act_bwd(_p(dy), _p(y), _p(x), _p(stats), _p(w), _p(dx), _p(dres), _p(gw), _p(gb), M, Cc, int(relu),
                                          _stream()), "ln_act_bwd")
        return dx, dres

    def _block_fwd(self, x, shp, params, pk, prefix):
        N, H, W, f = shp
        P = lambda n: params[self.index[prefix + n]]
        t1 = gemm(x, pk[prefix + "conv1.weight"][0])
        a, s1 = self._ln(t1, P("conv1.norm.weight"), P("conv1.norm.bias"), None, True)
        t2 = torch.empty_like(a)
        _l.check(self.lib.dcpt_conv3x3_fwd(_p(a), _p(pk[prefix + "conv2.weight"][0]), _p(t2), None, N, H, W, 2 * f, 2 * f, _stream()),
                 "conv3x3_fwd")
        b, s2 = self._ln(t2, P("conv2.norm.weight"), P("conv2.norm.bias"), None, True)
        t3 = gemm(b, pk[prefix + "conv3.weight"][0])
        out, s3 = self._ln(t3, P("conv3.norm.weight"), P("conv3.norm.bias"), x, True)
        return out, (x, t1, a, s1, t2, b, s2, t3, out, s3)

    def _block_bwd(self, dout, shp, saved, params, grads, pk, prefix, scratch):
        N, H, W, f = shp
        x, t1, a, s1, t2, b, s2, t3, out, s3 = saved
        P = lambda n: params[self.index[prefix + n]]
        G = lambda n: grads[self.index[prefix + n]]
        dt3, dres = self._ln_bwd(dout, out, t3, s3, P("conv3.norm.weight"), G("conv3.norm.weight"), G("conv3.norm.bias"), True)
        gemm(dt3, b, a_mn=True, b_mn=True, accumulate_into=G("conv3.weight").view(f, 2 * f), splits=0)
        db = gemm(dt3, pk[prefix + "conv3.weight"][1], out_dtype=torch.float32)      # gradients entering LN' stay fp32
        dt2, _ = self._ln_bwd(db, b, t2, s2, P("conv2.norm.weight"), G("conv2.norm.weight"), G("conv2.norm.bias"), False)
        _l.check(self.lib.dcpt_conv3x3_wgrad(_p(dt2), _p(a), _p(scratch), _p(G("conv2.weight")), N, H, W, 2 * f, 2 * f, _stream()),
                 "conv3x3_wgrad")
        da = torch.empty(a.shape, dtype=torch.float32, device=a.device)
        _l.check(self.lib.dcpt_conv3x3_fwd(_p(dt2), _p(pk[prefix + "conv2.weight"][1]), None, _p(da), N, H, W, 2 * f, 2 * f,
                                           _stream()), "conv3x3_dgrad")
        dt1, _ = self._ln_bwd(da, a, t1, s1, P("conv1.norm.weight"), G("conv1.norm.weight"), G("conv1.norm.bias"), False)
        gemm(dt1, x, a_mn=True, b_mn=True, accumulate_into=G("conv1.weight").view(2 * f, f), splits=0)
        return gemm(dt1, pk[prefix + "conv1.weight"][1], resid=dres, out_dtype=torch.float32)   # + shortcut gradient, fused

    # ---- whole head ---------------------------------------------------------------------------------
    def forward(self, params, feats, lq=None, pk=None):
        """feats[i]: fp32 NHWC [N, H>>i, W>>i, dims[i]] (contiguous).  Returns (logits fp32 [N, K], ctx).
        lq (fp32 NCHW) is read only by the img_embed (PromptIR_DC) variant.  pk: the packed operands when the caller already
        validated them (graph capture: the staleness check reads back a fingerprint and must stay outside)."""
        for p in params:
            if not p.is_cuda:
                raise _l.DcptError("dcpt_b200 has no CPU path: move the classifier head to a CUDA device")
        pk = self._pack(params) if pk is None else pk
        mw = torch.softmax(params[0].detach().float(), dim=0).contiguous()     # 1-D softmax of len(dims) scalars (:633)
        ctx = {"mw": mw, "stages": [], "feats": feats}
        z = None
        if self.embed:      # lq_feats = conv_embed(lq) (:549): 7x7 stride-2 conv as im2col + GEMM (+ bias), then channel LayerNorm
            if lq is None or not lq.is_cuda:
                raise _l.DcptError("PromptIR_DC needs the degraded image `lq` on the CUDA device")
            lq = lq.detach().contiguous().float()
            N, _, H, W = lq.shape
            Ho, Wo = (H - 1) // 2 + 1, (W - 1) // 2 + 1
            patches = torch.empty(N * Ho * Wo, 160, dtype=_l.operand_dtype(), device=lq.device)
            _l.check(self.lib.dcpt_im2col7x7s2(_p(lq), _p(patches), N, H, W, _stream()), "im2col7x7s2")
            t0 = gemm(patches, pk["conv_embed.0.weight"][0], bias=params[self.index["conv_embed.0.bias"]])
            z, s0 = self._ln(t0, params[self.index["conv_embed.1.weight"]], params[self.index["conv_embed.1.bias"]], None, False)
            ctx["embed"] = (patches, t0, z, s0)
            if tuple(feats[0].shape[:3]) != (N, Ho, Wo):
                raise _l.DcptError(f"PromptIR_DC: features[0] must be {Ho}x{Wo} (conv_embed halves the resolution), got {tuple(feats[0].shape)}")
        for i, f in enumerate(feats):
            N, H, W, Cc = f.shape
            assert Cc == self.dims[i] and f.dtype == torch.float32 and f.is_contiguous()
            zin = torch.empty(N * H * W, Cc, dtype=_l.operand_dtype(), device=f.device)
            _l.check(self.lib.dcpt_mix_fwd(_p(z), _p(f), _p(mw[i:i + 1]), _p(zin), f.numel(), _stream()), "mix_fwd")
            x, blocks = zin, []
            for j in range(self.nb):
                x, sv = self._block_fwd(x, (N, H, W, Cc), params, pk, f"bottleneck_layers.{i}.{j}.")
                blocks.append(sv)
            wname = f"downsample_layers.{i}.0.weight"
            t = gemm(x, pk[wname][0])
            Cn = t.shape[1]
            y = torch.empty(N * (H // 2) * (W // 2), Cn, dtype=_l.operand_dtype(), device=f.device)
            _l.check(self.lib.dcpt_maxpool2_relu_fwd(_p(t), _p(y), N, H // 2, W // 2, Cn, _stream()), "maxpool_fwd")
            ctx["stages"].append((blocks, x, t, (N, H, W, Cc, Cn)))
            z = y
        N, H, W, Cc, Cn = ctx["stages"][-1][3]
        shp = (N, H // 2, W // 2, Cn)
        last = []
        x = z
        for j in range(self.nb):
            x, sv = self._block_fwd(x, shp, params, pk, f"last_stage.{j}.")
            last.append(sv)
        HW = shp[1] * shp[2]
        pooled = torch.empty(N, Cn, dtype=torch.float32, device=x.device)
        logits = torch.empty(N, self.k, dtype=torch.float32, device=x.device)
        _l.check(self.lib.dcpt_meanpool_fc_fwd(_p(x), _p(params[self.index["fc.weight"]]), _p(params[self.index["fc.bias"]]),
                                               _p(pooled), _p(logits), N, HW, Cn, self.k, _stream()), "meanpool_fc_fwd")
        ctx.update(last=last, shp=shp, pooled=pooled, xlast=x)
        return logits, ctx

    def backward(self, params, ctx, dlogits, pk=None, sync=True):
        """Returns (dfeats list of fp32 NHWC, grads list in parameter order).  sync=False: the caller runs grad_sync itself on
        self._last_flat (graph replay: a collective cannot sit inside the captured graph)."""
        pk = self._pack(params) if pk is None else pk
        dev = dlogits.device
        offs, off = [], 0
        for p in params:                                   # one flat buffer (256-byte aligned views): one all-reduce under DP
            offs.append(off)
            off += (p.numel() + 63) // 64 * 64
        flat = torch.zeros(off, dtype=torch.float32, device=dev)
        grads = [flat[o:o + p.numel()].view(p.shape) for o, p in zip(offs, params)]
        maxw = max(2 * d for d in self.dims)
        scratch = torch.empty(self.lib.dcpt_conv3x3_packed_elems(maxw, maxw, 0), dtype=torch.float32, device=dev)
        N, h, w, Cn = ctx["shp"]
        dx = torch.empty(N * h * w, Cn, dtype=torch.float32, device=dev)
        sc = fp16_grad_scale([dlogits])                    # IEEE-half operand build only (params.py)
        if sc is not None:
            dlogits = dlogits * sc[0]
        _l.check(self.lib.dcpt_meanpool_fc_bwd(_p(dlogits.contiguous().float()), _p(ctx["pooled"]), _p(params[self.index["fc.weight"]]),
                                               _p(grads[self.index["fc.weight"]]), _p(grads[self.index["fc.bias"]]), _p(dx), N, h * w,
                                               Cn, self.k, _stream()), "meanpool_fc_bwd")
        for j in reversed(range(self.nb)):
            dx = self._block_bwd(dx, ctx["shp"], ctx["last"][j], params, grads, pk, f"last_stage.{j}.", scratch)
        dmw = torch.zeros(len(self.dims), dtype=torch.float32, device=dev)
        dfeats = [None] * len(self.dims)
        for i in reversed(range(len(self.dims))):
            blocks, xs, t, (N, H, W, Cc, Cn) = ctx["stages"][i]
            dt = torch.empty_like(t)
            _l.check(self.lib.dcpt_maxpool2_relu_bwd(_p(t), _p(dx), _p(dt), N, H // 2, W // 2, Cn, _stream()), "maxpool_bwd")
            wname = f"downsample_layers.{i}.0.weight"
            gemm(dt, xs, a_mn=True, b_mn=True, accumulate_into=grads[self.index[wname]].view(Cn, Cc), splits=0)
            dx = gemm(dt, pk[wname][1], out_dtype=torch.float32)
            for j in reversed(range(self.nb)):
                dx = self._block_bwd(dx, (N, H, W, Cc), blocks[j], params, grads, pk, f"bottleneck_layers.{i}.{j}.", scratch)
            f = ctx["feats"][i]
            dfeats[i] = torch.empty_like(f)
            _l.check(self.lib.dcpt_mix_bwd(_p(dx), _p(f), _p(ctx["mw"][i:i + 1]), _p(dfeats[i]), _p(dmw[i:i + 1]), f.numel(), _stream()),
                     "mix_bwd")
            # dx (= d z_in) is also the gradient of the previous stage's pooled output (z = prev + mw * feat)
        if self.embed:      # dx = d(z_in of stage 0) = d(conv_embed output): LayerNorm', then the conv's weight / bias gradients
            patches, t0, z0, s0 = ctx["embed"]
            gi = self.index
            dt0, _ = self._ln_bwd(dx, z0, t0, s0, params[gi["conv_embed.1.weight"]], grads[gi["conv_embed.1.weight"]],
                                  grads[gi["conv_embed.1.bias"]], False, relu=False)
            f0 = t0.shape[1]
            gw = torch.zeros(f0, 160, dtype=torch.float32, device=dev)
            gemm(dt0, patches, a_mn=True, b_mn=True, accumulate_into=gw, splits=0)
            grads[gi["conv_embed.0.weight"]].add_(gw[:, :147].reshape(grads[gi["conv_embed.0.weight"]].shape))
            grads[gi["conv_embed.0.bias"]].add_(dt0.float().sum(0))
        mw = ctx["mw"]
        grads[0].copy_((mw * (dmw - (dmw * mw).sum())).view(grads[0].shape))   # softmax backward on len(dims) scalars
        if sc is not None:
            flat.mul_(sc[1])
            torch._foreach_mul_(dfeats, sc[1])
        self._last_flat, self._last_offs = flat, offs
        if sync and self.grad_sync is not None:
            self.grad_sync(flat)
        return dfeats, grads

    # ---- CUDA-graph replay -----------------------------------------------------------------------------
    def graph_forward(self, params, feats, lq):
        """Replay of `forward` on static buffers, or None (first sighting of a shape, slot busy, graphs off): the caller then
        launches eagerly.  Returns (logits, slot); the slot is the backward's context and is busy until `graph_backward`."""
        if not self.use_graphs or torch.cuda.is_current_stream_capturing():
            return None
        key = (tuple(tuple(f.shape) for f in feats), None if lq is None else tuple(lq.shape), feats[0].device,
               tuple(p.data_ptr() for p in params))
        slot = self._gslots.get(key)
        if slot is None:
            if self._seen.get(key) is None:       # a shape earns a slot (static buffers + captures) when it comes back
                self._seen.put(key, True)
                return None
            slot = _HeadSlot([torch.empty_like(f) for f in feats], None if lq is None else torch.empty_like(lq))
            self._gslots.put(key, slot)
        if slot.busy:
            return None
        pk = self._pack(params)                    # eager: staleness check / re-pack; the packed tensors' addresses are stable
        for s_, f in zip(slot.feats, feats):
            s_.copy_(f)
        if slot.lq is not None:
            slot.lq.copy_(lq)
        if slot.fgraph is None:
            self.forward(params, slot.feats, lq=slot.lq, pk=pk)     # eager once: one-time initialisation inside the library
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, capture_error_mode="thread_local"):
                slot.logits, slot.ctx = self.forward(params, slot.feats, lq=slot.lq, pk=pk)
            slot.fgraph, slot.pk = g, pk
            g.replay()                             # a capture records, it does not execute: fill the captured tensors
        elif slot.pk is not pk:                    # the weights were re-packed into new tensors: the captures point at the old ones
            self._gslots.d.pop(key, None)
            return None
        else:
            slot.fgraph.replay()
        slot.busy = True
        slot.token += 1
        return slot.logits.clone(), slot

    def graph_backward(self, params, slot, dlogits):
        pk = slot.pk
        if slot.dlogits is None:
            slot.dlogits = torch.empty_like(dlogits, dtype=torch.float32).contiguous()
        slot.dlogits.copy_(dlogits)
        if slot.bgraph is None:
            self.backward(params, slot.ctx, slot.dlogits, pk=pk, sync=False)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g, pool=slot.fgraph.pool(), capture_error_mode="thread_local"):
                slot.dfeats, slot.grads = self.backward(params, slot.ctx, slot.dlogits, pk=pk, sync=False)
            slot.bgraph, slot.flat, slot.offs = g, self._last_flat, self._last_offs
            g.replay()
        else:
            slot.bgraph.replay()
        slot.busy = False
        if self.grad_sync is not None:
            self.grad_sync(slot.flat)
        flat = slot.flat.clone()        # autograd owns what it is handed; the slot's buffers are rewritten by the next replay
        grads = [flat[o:o + p.numel()].view(p.shape) for o, p in zip(slot.offs, params)]
        return [d.clone() for d in slot.dfeats], grads


class _HeadSlot:
    """Static inputs, captured graphs and their outputs for one (feature shapes, parameter storage) signature of the head."""

    def __init__(self, feats, lq):
        self.feats, self.lq = feats, lq
        self.fgraph = self.bgraph = self.pk = None
        self.logits = self.ctx = self.dlogits = self.dfeats = self.grads = self.flat = self.offs = None
        self.busy = False
        self.token = 0          # use counter: a finalizer of an OLD autograd node must not release a newer forward's claim

    def release(self, token):
        if self.token == token:
            self.busy = False


class _DCHeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, engine, n_feats, lq, need_grad, *args):
        feats, params = args[:n_feats], args[n_feats:]
        # features arrive as logical NCHW (channels_last memory from the NAFNet function, or plain NCHW): make NHWC fp32
        fh = [f.detach().permute(0, 2, 3, 1).contiguous().float() for f in feats]
        dparams = [p.detach().contiguous() for p in params]
        with (training_pass() if need_grad else _NullCtx()):      # no weight-fingerprint sync on training passes (params.py)
            res = engine.graph_forward(dparams, fh, lq) if need_grad else None
            if res is None:
                logits, c = engine.forward(dparams, fh, lq=lq)
            else:
                logits, c = res
                try:
                    weakref.finalize(ctx, c.release, c.token)     # a forward whose backward never runs must not pin the slot
                except TypeError:
                    pass
        ctx.engine, ctx.c, ctx.params, ctx.n_feats = engine, c, dparams, n_feats
        return logits

    @staticmethod
    def backward(ctx, dlogits):
        with training_pass():
            if isinstance(ctx.c, _HeadSlot):
                dfeats, grads = ctx.engine.graph_backward(ctx.params, ctx.c, dlogits)
            else:
                dfeats, grads = ctx.engine.backward(ctx.params, ctx.c, dlogits)
        ctx.c = None
        return (None, None, None, None) + tuple(d.permute(0, 3, 1, 2) for d in dfeats) + tuple(grads)   # (lq is data: no gradient)


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def dchead_apply(engine, feats, params, lq=None):
    need_grad = torch.is_grad_enabled() and (any(f.requires_grad for f in feats) or any(p.requires_grad for p in params))
    return _DCHeadFunction.apply(engine, len(feats), lq if engine.embed else None, need_grad, *feats, *params)
