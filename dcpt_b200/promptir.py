"""Host-side engine for the PromptIR network (reference: basicsr/archs/promptir_arch.py:267-518), inference.

``PromptIREngine`` owns the C plan (``dcpt_promptir_create``), the packed 16-bit operand cache and the per-shape workspaces /
captured CUDA graphs; the ``basicsr`` mirror's ``PromptIR`` module calls ``forward``.  No PyTorch implementation of the math
lives here and there is no CPU path.  Training PromptIR is not on this path: the module raises when gradients are required."""
import torch

from . import lib as _l
from .ops import _p, _stream
from .restormer import RestormerEngine


class PromptIREngine(RestormerEngine):
    _ABI = "dcpt_promptir"

    def __init__(self, inp_channels=3, out_channels=3, dim=48, num_blocks=(4, 6, 6, 8), num_refinement_blocks=4, heads=(1, 2, 4, 8),
                 ffn_expansion_factor=2.66, bias=False, ln_with_bias=True):
        super().__init__(inp_channels, out_channels, dim, num_blocks, num_refinement_blocks, heads, ffn_expansion_factor, bias,
                         ln_with_bias, attn_softmax=True)

    def _run(self, params, packed, inp, out, work):
        N, _, H, W = inp.shape
        pp = _l.ptr_array([p.data_ptr() for p in params])
        _l.check(self.lib.dcpt_promptir_fwd(self.plan, pp, _p(packed), _p(inp), _p(out), _p(work), N, H, W, _stream()), "promptir_fwd")

    def forward(self, params, inp):
        """inp fp32 NCHW [N,3,H,W] (H, W multiples of 8) -> restored image, same shape."""
        self._check_params(params)
        if not inp.is_cuda:
            raise _l.DcptError("dcpt_b200 has no CPU path: input is on %s" % inp.device)
        inp = inp.contiguous().float()
        N, _, H, W = inp.shape
        dev = inp.device
        packed = self.packed_for(params)
        nbytes = lambda: self.lib.dcpt_promptir_workspace_bytes(self.plan, N, H, W)     # noqa: E731
        if self.use_graphs and not torch.cuda.is_current_stream_capturing():
            # a shape earns a graph entry (static buffers + workspace + capture) when it comes back; first sighting is eager
            gkey = (N, H, W, dev, tuple(p.data_ptr() for p in params))
            ent = self._graphs.get(gkey)
            if ent is not None or self._seen.get(gkey) is not None:
                if ent is None:
                    ent = {"inp": torch.empty_like(inp), "out": torch.empty_like(inp), "graph": None,
                           "work": torch.empty(nbytes(), dtype=torch.uint8, device=dev)}
                    self._graphs.put(gkey, ent)
                ent["inp"].copy_(inp)
                if ent["graph"] is None:
                    self._run(params, packed, ent["inp"], ent["out"], ent["work"])   # eager once: one-time library initialisation
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g, capture_error_mode="thread_local"):
                        self._run(params, packed, ent["inp"], ent["out"], ent["work"])
                    ent["graph"] = g
                else:
                    ent["graph"].replay()
                return ent["out"].clone()
            self._seen.put(gkey, True)
        work = self._work.setdefault((N, H, W, dev), lambda: torch.empty(nbytes(), dtype=torch.uint8, device=dev))
        out = torch.empty_like(inp)
        self._run(params, packed, inp, out, work)
        return out
