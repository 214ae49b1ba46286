"""NAFNet on the B200 hot path — drop-in for the reference's ``basicsr/archs/nafnet_arch.py``.

Same class names, ctor kwargs, sub-module / parameter names and ``state_dict`` shapes as the
reference (nafnet_arch.py:56-64 LayerNorm2d, :77-80 SimpleGate, :83-186 NAFBlock, :189-274
NAFNetBaseline), so ``build_network(opt['network_g'])`` and ``load_state_dict(strict=True)`` work on
the reference's yml files and checkpoints.  The modules are parameter containers: ``forward`` hands
the whole network (or one block) to the sm_100a kernels through the C ABI — there is no PyTorch
implementation of the math here and no CPU path.
"""
import torch
import torch.nn as nn

from basicsr.utils.registry import ARCH_REGISTRY
from dcpt_b200.lib import DcptError
from dcpt_b200.nafnet import NAFNetEngine, nafnet_apply
from dcpt_b200.ops import NAFBLOCK_PARAM_ORDER, NAFBlockOp


class LayerNorm2d(nn.Module):
    """Channel LayerNorm parameters (nafnet_arch.py:56-64): weight/bias of shape (C,), eps 1e-6."""

    def __init__(self, channels, eps=1e-6):
        super().__init__()
        self.register_parameter("weight", nn.Parameter(torch.ones(channels)))
        self.register_parameter("bias", nn.Parameter(torch.zeros(channels)))
        self.eps = eps

    def forward(self, x):
        """Standalone use: x fp32 NCHW on CUDA -> LN(x) (bf16-rounded, returned as fp32). Inference only."""
        from dcpt_b200 import ops
        if torch.is_grad_enabled() and (x.requires_grad or self.weight.requires_grad):
            raise DcptError("standalone LayerNorm2d is inference-only; training runs through NAFBlock / NAFNetBaseline")
        N, C, H, W = x.shape
        rows = x.permute(0, 2, 3, 1).reshape(-1, C).contiguous().float()
        out, _ = ops.layernorm2d_fwd(rows, self.weight.detach(), self.bias.detach(), self.eps)
        return out.float().reshape(N, H, W, C).permute(0, 3, 1, 2)


class SimpleGate(nn.Module):
    """x[:, :C] * x[:, C:] (nafnet_arch.py:77-80); fused into the conv epilogues on the hot path."""

    def forward(self, x):
        a, b = x.chunk(2, dim=1)
        return a * b


class _NAFBlockFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, *params):
        op = NAFBlockOp([p.detach().contiguous() for p in params])
        xh = x.detach().permute(0, 2, 3, 1).contiguous().float()
        out, saved, _ = op.forward(xh)
        ctx.op, ctx.xh, ctx.saved = op, xh, saved
        return out.permute(0, 3, 1, 2)

    @staticmethod
    def backward(ctx, dout):
        dx, grads = ctx.op.backward(ctx.xh, ctx.saved, dout.permute(0, 2, 3, 1).contiguous().float())
        return (dx.permute(0, 3, 1, 2),) + tuple(grads)


class NAFBlock(nn.Module):
    """Parameter layout of the reference NAFBlock (nafnet_arch.py:84-163)."""

    def __init__(self, c, DW_Expand=2, FFN_Expand=2, drop_out_rate=0.0):
        super().__init__()
        if DW_Expand != 2 or FFN_Expand != 2:
            raise DcptError("the sm_100a NAFBlock kernels are built for DW_Expand = FFN_Expand = 2 (SimpleGate halves)")
        if drop_out_rate > 0.0:
            raise DcptError("drop_out_rate > 0 is not on the hot path (NAFNetBaseline never sets it, nafnet_arch.py:229)")
        dw = c * DW_Expand
        self.conv1 = nn.Conv2d(c, dw, 1, bias=True)
        self.conv2 = nn.Conv2d(dw, dw, 3, padding=1, groups=dw, bias=True)
        self.conv3 = nn.Conv2d(dw // 2, c, 1, bias=True)
        self.sca = nn.Sequential(nn.AdaptiveAvgPool2d(1), nn.Conv2d(dw // 2, dw // 2, 1, bias=True))
        self.sg = SimpleGate()
        ffn = FFN_Expand * c
        self.conv4 = nn.Conv2d(c, ffn, 1, bias=True)
        self.conv5 = nn.Conv2d(ffn // 2, c, 1, bias=True)
        self.norm1 = LayerNorm2d(c)
        self.norm2 = LayerNorm2d(c)
        self.dropout1 = nn.Identity()
        self.dropout2 = nn.Identity()
        self.beta = nn.Parameter(torch.zeros((1, c, 1, 1)), requires_grad=True)
        self.gamma = nn.Parameter(torch.zeros((1, c, 1, 1)), requires_grad=True)

    def ordered_params(self):
        named = dict(self.named_parameters())
        return [named[k] for k in NAFBLOCK_PARAM_ORDER]

    def forward(self, inp):
        """Standalone block (fp32 NCHW CUDA tensor) with autograd, one C-ABI call each way."""
        return _NAFBlockFunction.apply(inp, *self.ordered_params())


@ARCH_REGISTRY.register()
class NAFNetBaseline(nn.Module):
    """nafnet_arch.py:189-274.  ``window_size`` is accepted and ignored by the arch, as in the
    reference (SRModel.pre_test reads it from the yml)."""

    def __init__(self, img_channel=3, width=16, middle_blk_num=1, enc_blk_nums=[], dec_blk_nums=[], window_size=8):
        super().__init__()
        self.intro = nn.Conv2d(img_channel, width, 3, padding=1, bias=True)
        self.ending = nn.Conv2d(width, img_channel, 3, padding=1, bias=True)
        self.encoders = nn.ModuleList()
        self.middle_blks = nn.ModuleList()
        self.ups = nn.ModuleList()
        self.downs = nn.ModuleList()
        chan = width
        for num in enc_blk_nums:
            self.encoders.append(nn.Sequential(*[NAFBlock(chan) for _ in range(num)]))
            self.downs.append(nn.Conv2d(chan, 2 * chan, 2, 2))
            chan *= 2
        self.middle_blks = nn.Sequential(*[NAFBlock(chan) for _ in range(middle_blk_num)])
        for i, num in enumerate(dec_blk_nums):
            self.ups.append(nn.Sequential(nn.Conv2d(chan, chan * 2, 1, bias=False), nn.PixelShuffle(2)))
            chan //= 2
            setattr(self, f"decoder{i}", nn.Sequential(*[NAFBlock(chan) for _ in range(num)]))
        self._cfg = (img_channel, width, middle_blk_num, tuple(enc_blk_nums), tuple(dec_blk_nums))
        self._engine = None

    # -- engine -----------------------------------------------------------------------------
    def engine(self):
        if self._engine is None:
            ic, w, mid, enc, dec = self._cfg
            self._engine = NAFNetEngine(ic, w, mid, list(enc), list(dec))
            shapes = [tuple(p.shape) for p in self.parameters()]
            want = [tuple(d for d in s[:len(sh)]) for s, sh in zip(self._engine.shapes, shapes)]
            if [int(torch.tensor(s).prod()) for s in self._engine.shapes] != [p.numel() for p in self.parameters()]:
                raise DcptError(f"parameter layout mismatch between module and C plan: {shapes[:4]} vs {want[:4]}")
        return self._engine

    def _decoder_hook_targets(self):
        """Sub-modules with forward hooks that the fused forward must serve, per decoder level: (modules, block index).
        DCPT registers on every module whose name contains ``hook_names`` and has exactly one dot
        (degradation_classification_pretrain_model.py:64-67): ``decoder{i}.0`` - the level's FIRST block - for an unwrapped
        network, ``module.decoder{i}`` - the whole level - under DDP.  One tap per level: the container (= its last block) or
        one block ``decoder{i}.{j}``; the engine materialises that block's output and routes the gradient back into it."""
        n_dec = len(self._cfg[4])
        targets, blocks = [], []
        for i in range(n_dec):
            seq = getattr(self, f"decoder{i}")
            last = len(seq) - 1
            hooked = [j for j in range(len(seq)) if len(seq[j]._forward_hooks) > 0]
            if len(seq._forward_hooks) > 0 and last >= 0 and last not in hooked:
                hooked.append(last)
            if len(set(hooked)) > 1:
                raise DcptError(f"forward hooks on several blocks of decoder{i} ({sorted(set(hooked))}): the fused forward "
                                "materialises one feature per decoder level")
            j = hooked[0] if hooked else -1
            mods = ([seq] if len(seq._forward_hooks) > 0 and (j == last or last < 0) else []) + \
                   ([seq[j]] if j >= 0 and len(seq[j]._forward_hooks) > 0 else [])
            targets.append(mods)
            blocks.append(j if j != last else -1)
        return targets, blocks

    def train(self, mode=True):
        """nn.Module.train without the walk over ~700 parameter-container sub-modules when the mode does not change (the step
        classes call net_g.train() every iteration, sr_model.py:133: 1.3 ms of host time before the first launch of the step;
        nothing in these containers depends on .training)."""
        if self.training == bool(mode):
            return self
        return super().train(mode)

    def prepack(self):
        """Re-pack the 16-bit operand images of the weights NOW instead of at the next forward.  The step classes call it right
        after optimizer.step(): the ~300 small pack launches then queue up behind the optimizer kernels while the GPU is still
        busy, instead of costing 1.5 ms of host time in front of the next iteration's first kernel."""
        params = self._param_list()
        if params and params[0].is_cuda:
            from dcpt_b200.nafnet import _param_view
            eng = self.engine()
            eng.packed_for(_param_view(eng, params).dparams)

    def _param_list(self):
        """list(self.parameters()) without the module-tree walk (0.4 ms per call for 664 tensors, paid while the GPU idles
        before the graph launch): the (module, name) slots are cached, the Parameter objects are read fresh every call."""
        slots = self.__dict__.get("_pslots")
        if slots is None:
            slots = [(m, n) for m in self.modules() for n in m._parameters if m._parameters[n] is not None]
            self.__dict__["_pslots"] = slots
        return [m._parameters[n] for m, n in slots]

    def zero_grad(self, set_to_none=True):
        """nn.Module.zero_grad over the cached parameter slots (same semantics; 0.85 -> 0.1 ms for 664 tensors)."""
        for p in self._param_list():
            if p.grad is not None:
                if set_to_none:
                    p.grad = None
                else:
                    p.grad.detach_()
                    p.grad.requires_grad_(False)
                    p.grad.zero_()

    def forward(self, inp, hook=False):
        params = self._param_list()
        targets, blocks = self._decoder_hook_targets()
        want_feats = any(len(t) > 0 for t in targets)
        eng = self.engine()
        if want_feats:
            eng.set_hook_blocks(blocks)
        out, feats = nafnet_apply(eng, inp, params, hook=bool(hook), want_feats=want_feats)
        if want_feats:
            for mods, f in zip(targets, feats):
                for m in mods:
                    for fn in list(m._forward_hooks.values()):
                        fn(m, (None,), f)
        if not hook:
            return out
        return None


@ARCH_REGISTRY.register()
class NAFNet(NAFNetBaseline):
    """nafnet_arch.py:277-288 - ``Local_Base`` + ``NAFNetBaseline``: the test-time local converter (TLC).  The reference
    replaces every ``AdaptiveAvgPool2d(1)`` of the SCA branches by ``AvgPool2d(base_size = 1.5 x train_size)``
    (arch_util.py:436-455) and fixes the per-level kernels with one forward on a ``train_size`` dummy input; here the same
    kernels are computed in closed form and handed to the C plan (``dcpt_nafnet_set_tlc``).  Parameters and ``state_dict``
    are those of ``NAFNetBaseline``.  Inference only (the reference constructs it in ``eval()`` mode)."""

    def __init__(self, *args, train_size=(1, 3, 128, 128), fast_imp=False, **kwargs):
        super().__init__(*args, **kwargs)
        if fast_imp:
            raise DcptError("NAFNet(fast_imp=True) (the non-equivalent strided approximation, arch_util.py:355-377) is not built")
        self.train_size = tuple(train_size)
        H, W = self.train_size[-2], self.train_size[-1]
        bh, bw = int(H * 1.5), int(W * 1.5)                                   # nafnet_arch.py:284-285
        n_levels = len(self._cfg[3]) + 1
        # arch_util.py:341-347: kernel = feature-map size of the dummy forward * base_size // train_size, per level
        self.tlc_kernels = [((H >> l) * bh // H, (W >> l) * bw // W) for l in range(n_levels)]
        self.eval()

    def engine(self):
        eng = super().engine()
        if not eng.tlc:
            eng.set_tlc(self.tlc_kernels)
        return eng

    def forward(self, inp, hook=False):
        if torch.is_grad_enabled() and (inp.requires_grad or any(p.requires_grad for p in self.parameters())):
            raise DcptError("NAFNet (TLC) is a test-time converter: call it under torch.no_grad() (SRModel.test does)")
        return super().forward(inp, hook=hook)
