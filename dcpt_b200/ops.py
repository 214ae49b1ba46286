"""Tensor-level wrappers over the C ABI (torch is used for device memory and streams only).

Layout contract: activations are NHWC, flattened to ``[M = N*H*W, C]``; fp32 for the residual
stream, bf16 for branch tensors.  Every function launches on ``torch.cuda.current_stream()``.
"""
import ctypes as C

import torch

from . import lib as _l


def _lib():
    return _l.load_library()


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    if t is None:
        return None
    assert t.is_cuda and t.is_contiguous(), "dcpt_b200 ops need contiguous CUDA tensors"
    return C.c_void_p(t.data_ptr())


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _l.DcptError("dcpt_b200 has no CPU path: tensor is on %s" % t.device)


def layernorm2d_fwd(x, weight, bias, eps=1e-6):
    """x fp32 [M, C] -> (out bf16 [M, C], stats fp32 [M, 2]).  nafnet_arch.py:27-35."""
    _need_cuda(x, weight, bias)
    M, Cc = x.shape
    out = torch.empty(M, Cc, dtype=_l.operand_dtype(), device=x.device)
    stats = torch.empty(M, 2, dtype=torch.float32, device=x.device)
    _l.check(_lib().dcpt_layernorm2d_fwd(_p(x), _p(weight), _p(bias), _p(out), _p(stats), M, Cc, eps, _stream()),
             "layernorm2d_fwd")
    return out, stats


def layernorm2d_bwd(dn, x, stats, weight, dres=None, want_mirror=True):
    """Returns (dx fp32, dx_bf16, dweight, dbias, colsum).  nafnet_arch.py:38-53."""
    _need_cuda(dn, x, stats, weight, dres)
    M, Cc = x.shape
    dx = torch.empty_like(x)
    dxb = torch.empty(M, Cc, dtype=_l.operand_dtype(), device=x.device) if want_mirror else None
    dw = torch.zeros(Cc, dtype=torch.float32, device=x.device)
    db = torch.zeros_like(dw)
    cs = torch.zeros_like(dw)
    _l.check(_lib().dcpt_layernorm2d_bwd(_p(dn), _p(x), _p(stats), _p(weight), _p(dres), _p(dx), _p(dxb), _p(dw), _p(db),
                                         _p(cs), M, Cc, _stream()), "layernorm2d_bwd")
    return dx, dxb, dw, db, cs


def gemm(A, B, *, a_mn=False, b_mn=False, bias=None, resid=None, out_dtype=None, splits=1, accumulate_into=None,
         impl=0):
    """D[M,N] = A * B^T with bf16 operands.  K-major: A [M,K], B [N,K]; MN-major: A [K,M], B [K,N]."""
    _need_cuda(A, B)
    if a_mn:
        K, M = A.shape
    else:
        M, K = A.shape
    N = B.shape[1] if b_mn else B.shape[0]
    out_f32 = out_bf16 = None
    acc = 0
    if accumulate_into is not None:
        out_f32, acc = accumulate_into, 1
    elif splits > 1:
        out_f32 = torch.zeros(M, N, dtype=torch.float32, device=A.device)
    elif out_dtype == torch.float32:
        out_f32 = torch.empty(M, N, dtype=torch.float32, device=A.device)
    else:
        out_bf16 = torch.empty(M, N, dtype=_l.operand_dtype(), device=A.device)
    _l.check(_lib().dcpt_gemm_bf16(_p(A), A.stride(0), int(a_mn), _p(B), B.stride(0), int(b_mn), M, N, K, _p(out_f32),
                                   _p(out_bf16), N, _p(bias), _p(resid), splits, acc, impl, _stream()), "gemm_bf16")
    return out_f32 if out_f32 is not None else out_bf16


def gemm_ex(desc: "_l.GemmDesc", impl=0):
    _l.check(_lib().dcpt_gemm_ex(C.byref(desc), impl, _stream()), "gemm_ex")


def dwconv3x3_gate_fwd(u, weight, bias):
    """u bf16 [N,H,W,2C] -> (g bf16 [N,H,W,C], pool fp32 [N,C]).  nafnet_arch.py:96-104,:171-172."""
    _need_cuda(u, weight, bias)
    N, H, W, C2 = u.shape
    Cc = C2 // 2
    g = torch.empty(N, H, W, Cc, dtype=_l.operand_dtype(), device=u.device)
    pool = torch.zeros(N, Cc, dtype=torch.float32, device=u.device)
    _l.check(_lib().dcpt_dwconv3x3_gate_fwd(_p(u), _p(weight), _p(bias), _p(g), _p(pool), N, H, W, Cc, _stream()),
             "dwconv3x3_gate_fwd")
    return g, pool


NAFBLOCK_PARAM_ORDER = ["beta", "gamma", "conv1.weight", "conv1.bias", "conv2.weight", "conv2.bias", "conv3.weight",
                        "conv3.bias", "sca.1.weight", "sca.1.bias", "conv4.weight", "conv4.bias", "conv5.weight",
                        "conv5.bias", "norm1.weight", "norm1.bias", "norm2.weight", "norm2.bias"]


class NAFBlockOp:
    """One NAFBlock (nafnet_arch.py:83-186) through the C ABI: pack / fwd / bwd on NHWC fp32 tensors."""

    def __init__(self, params):
        """params: list of 18 contiguous fp32 CUDA tensors in NAFBLOCK_PARAM_ORDER."""
        assert len(params) == 18
        _need_cuda(*params)
        self.params = [p.detach() for p in params]
        self.C = self.params[0].numel()
        self._pp = _l.ptr_array([p.data_ptr() for p in self.params])
        self.packed = torch.empty(_lib().dcpt_nafblock_packed_bytes(self.C), dtype=torch.uint8, device=params[0].device)
        self.pack()

    def pack(self):
        _l.check(_lib().dcpt_nafblock_pack(self._pp, _p(self.packed), self.C, _stream()), "nafblock_pack")

    def forward(self, x, want_mirror=False):
        N, H, W, Cc = x.shape
        assert Cc == self.C and x.dtype == torch.float32
        out = torch.empty_like(x)
        mirror = torch.empty(x.shape, dtype=_l.operand_dtype(), device=x.device) if want_mirror else None
        saved = torch.empty(_lib().dcpt_nafblock_saved_bytes(N, H, W, Cc), dtype=torch.uint8, device=x.device)
        _l.check(_lib().dcpt_nafblock_fwd(self._pp, _p(self.packed), _p(x), _p(out), _p(mirror), _p(saved), N, H, W, Cc,
                                          _stream()), "nafblock_fwd")
        return out, saved, mirror

    def backward(self, x, saved, dout):
        N, H, W, Cc = x.shape
        dev = x.device
        dout = dout.contiguous()
        dout_b = dout.to(_l.operand_dtype())
        dout_cs = dout.reshape(-1, Cc).sum(0).float().contiguous()
        dx = torch.empty_like(x)
        grads = [torch.zeros_like(p) for p in self.params]
        gp = _l.ptr_array([g.data_ptr() for g in grads])
        ws = torch.empty(_lib().dcpt_nafblock_workspace_bytes(N, H, W, Cc), dtype=torch.uint8, device=dev)
        _l.check(_lib().dcpt_nafblock_bwd(self._pp, _p(self.packed), _p(saved), _p(x), _p(dout), _p(dout_b), _p(dout_cs),
                                          _p(dx), None, None, gp, _p(ws), N, H, W, Cc, _stream()), "nafblock_bwd")
        return dx, grads
