"""GPU parity tests for the individual sm_100a kernels, called through the C ABI
(dcpt_b200.ops -> libdcpt_sm100.so) and checked against the CPU oracle / plain torch fp32 math.

Tolerances (written next to each check):
  * fp32 outputs of a GEMM with bf16 operands and fp32 accumulation: rel-L2 <= 2e-5 against an
    fp32 matmul of the SAME bf16-rounded operands (only summation order differs);
  * bf16 outputs: one final rounding, rel-L2 <= 4e-3 (bf16 eps = 2^-8);
  * LayerNorm statistics / fp32 elementwise math: rel-L2 <= 1e-5.
"""
import os

import numpy as np
import pytest
import torch

from dcpt_b200.lib import operand_dtype as OPD  # bf16 by default; fp16 for the DCPT_OPERAND=fp16 parity build

pytestmark = pytest.mark.gpu

from oracle import nafnet_oracle as O  # noqa: E402


def rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def ops():
    from dcpt_b200 import ops as _ops
    _ops._lib()
    return _ops


def bf(t):
    return t.to(OPD())


# ------------------------------------------------------------------ GEMM engine
GEMM_SHAPES = [(128, 64, 64), (256, 128, 64), (384, 256, 128), (1000, 512, 512), (4096, 1024, 512), (130, 72, 40),
               (64, 16, 8), (16384, 128, 64), (40000, 264, 72), (33000, 512, 256)]


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("M,N,K", GEMM_SHAPES)
def test_gemm_store(ops, impl, M, N, K):
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = bf(torch.randn(M, K, device="cuda", generator=g))
    B = bf(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    ref = A.float() @ B.float().t()
    out = ops.gemm(A, B, out_dtype=torch.float32, impl=impl)
    assert rel(out, ref) < 2e-5
    out = ops.gemm(A, B, bias=bias, resid=resid, out_dtype=torch.float32, impl=impl)
    assert rel(out, ref + bias + resid) < 2e-5
    out = ops.gemm(A, B, bias=bias, impl=impl)
    assert out.dtype == OPD() and rel(out.float(), ref + bias) < 4e-3


@pytest.mark.parametrize("M,N,K,mirror", [(1000, 512, 512, False), (16384, 512, 512, False), (4099, 256, 256, True), (33000, 128, 64, False),
                                          (70000, 64, 64, True), (130, 72, 40, False), (300, 24, 16, False), (257, 8, 8, False),
                                          (2048, 384, 128, False), (16500, 512, 1024, True)])
def test_gemm_store_fused_layernorm(ops, M, N, K, mirror):
    """STORE epilogue with the consumer's LayerNorm2d fused (EpiParams::ln_*): out = A B^T + bias + resid (fp32), and the channel
    LayerNorm of those rows (nafnet_arch.py:27-35) as bf16 + (mean, rstd), against fp32 torch math on the GEMM's own fp32 output:
    statistics 1e-5 (fp32 sums in another order), normalised rows one bf16 rounding (4e-3).  N = 512 exercises the whole-slab
    schedule (two accumulator buffers per row slab), N % 32 != 0 the column masking, M % 128 != 0 the row clipping."""
    from dcpt_b200.lib import GemmDesc
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = bf(torch.randn(M, K, device="cuda", generator=g))
    B = bf(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g) + 0.7          # un-centred rows: the shifted sums must cope
    lw = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    lb = 0.3 * torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda")
    ln = torch.full((M, N), float("nan"), device="cuda", dtype=OPD())
    stats = torch.full((M, 2), float("nan"), device="cuda")
    mir = torch.empty(M, N, device="cuda", dtype=OPD()) if mirror else None
    d = GemmDesc()
    for k, v in dict(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=0, out_f32=out, out_bf16=mir, ldo=N, bias=bias, resid=resid,
                     ldr=N, ln_weight=lw, ln_bias=lb, ln_out=ln, ld_ln=N, ln_stats=stats, ln_eps=1e-6).items():
        setattr(d, k, v.data_ptr() if torch.is_tensor(v) else v)
    ops.gemm_ex(d)
    ref = A.float() @ B.float().t() + bias + resid
    assert rel(out, ref) < 2e-5
    if mirror:
        assert rel(mir.float(), ref) < 4e-3
    x = out.double()
    mu = x.mean(1)
    var = (x - mu[:, None]).pow(2).mean(1)
    rstd = 1.0 / (var + 1e-6).sqrt()
    assert rel(stats[:, 0], mu) < 1e-5 and rel(stats[:, 1], rstd) < 1e-5
    ref_ln = (x - mu[:, None]) * rstd[:, None] * lw.double() + lb.double()
    assert torch.isfinite(ln.float()).all()
    assert rel(ln.float(), ref_ln) < 4e-3
    # bit-level agreement with the standalone LayerNorm kernel on the same fp32 rows (both round the same fp32 formula once)
    n_ref, st_ref = ops.layernorm2d_fwd(out, lw, lb)
    assert rel(stats, st_ref) < 1e-5
    assert (ln.float() - n_ref.float()).abs().max() <= 2 ** -7 * ref_ln.abs().max()


@pytest.mark.parametrize("M,N,K", [(1024, 48, 48), (4096, 96, 96), (260, 192, 192), (2048, 384, 384)])
def test_gemm_store_fused_biasfree_layernorm(ops, M, N, K):
    """The BiasFree form of the fused LayerNorm epilogue (Restormer's BiasFree_LayerNorm, restormer_arch.py:26-40: x / sqrt(var + eps)
    * weight, var about the mean, no centring, no bias) at Restormer's widths.  (The per-image-B GEMM it rides on in the network
    is exercised by the TransformerBlock goldens in tests/test_gpu_restormer.py.)"""
    from dcpt_b200.lib import GemmDesc
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = bf(torch.randn(M, K, device="cuda", generator=g))
    B = bf(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    resid = torch.randn(M, N, device="cuda", generator=g) + 0.5
    lw = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    out = torch.full((M, N), float("nan"), device="cuda")
    ln = torch.full((M, N), float("nan"), device="cuda", dtype=OPD())
    stats = torch.full((M, 2), float("nan"), device="cuda")
    d = GemmDesc()
    for k, v in dict(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=0, out_f32=out, ldo=N, resid=resid, ldr=N, ln_weight=lw,
                     ln_out=ln, ld_ln=N, ln_stats=stats, ln_eps=1e-6, ln_nocenter=1).items():
        setattr(d, k, v.data_ptr() if torch.is_tensor(v) else v)
    ops.gemm_ex(d)
    ref = A.float() @ B.float().t() + resid
    assert rel(out, ref) < 2e-5
    x = out.double()
    mu = x.mean(1)
    rstd = 1.0 / ((x - mu[:, None]).pow(2).mean(1) + 1e-6).sqrt()
    assert rel(stats[:, 0], mu) < 1e-5 and rel(stats[:, 1], rstd) < 1e-5
    assert rel(ln.float(), x * rstd[:, None] * lw.double()) < 4e-3


@pytest.mark.parametrize("M,N,K,dres,mirror", [(1000, 512, 1024, True, True), (16384, 512, 1024, True, True), (4099, 256, 512, True, False),
                                                (33000, 128, 256, False, True), (70000, 64, 128, True, True), (130, 72, 40, True, True),
                                                (300, 24, 16, False, False), (2048, 384, 128, True, True)])
def test_gemm_fused_layernorm_backward(ops, M, N, K, dres, mirror):
    """STORE epilogue with the LayerNorm BACKWARD fused (EpiParams::lnb_*): the GEMM result is dn = d(loss)/d(LN output) and the
    epilogue applies LayerNormFunction.backward (nafnet_arch.py:38-53) + the residual gradient, against double-precision torch
    math on the fp32 product of the same operands.  g and xhat travel between the two epilogue passes as 16-bit values (the
    separate ln_bwd kernel reads a 16-bit dn too): dx within one operand rounding of the LN part; the column sums to 2e-3."""
    from dcpt_b200.lib import GemmDesc
    from tol import tol
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    A = bf(torch.randn(M, K, device="cuda", generator=g))
    B = bf(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    x = torch.randn(M, N, device="cuda", generator=g) * 1.3 + 0.4
    w = 1.0 + 0.2 * torch.randn(N, device="cuda", generator=g)
    dr = torch.randn(M, N, device="cuda", generator=g) if dres else None
    xd = x.double()
    mu = xd.mean(1, keepdim=True)
    rstd = 1.0 / ((xd - mu).pow(2).mean(1, keepdim=True) + 1e-6).sqrt()
    stats = torch.cat([mu, rstd], 1).float().contiguous()
    out = torch.full((M, N), float("nan"), device="cuda")
    mir = torch.full((M, N), float("nan"), device="cuda", dtype=OPD()) if mirror else None
    dw, db, cs = (torch.zeros(N, device="cuda") for _ in range(3))
    d = GemmDesc()
    for k, v in dict(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=0, out_f32=out, out_bf16=mir, ldo=N, lnb_x=x, ld_lnb=N,
                     lnb_stats=stats, lnb_weight=w, lnb_dres=dr, lnb_dweight=dw, lnb_dbias=db, lnb_colsum=cs).items():
        setattr(d, k, (v.data_ptr() if torch.is_tensor(v) else v) if v is not None else None)
    ops.gemm_ex(d)
    dn = (A.float() @ B.float().t()).double()
    xhat = (xd - mu) * rstd
    gg = dn * w.double()
    ref = (gg - xhat * (gg * xhat).mean(1, keepdim=True) - gg.mean(1, keepdim=True)) * rstd
    ln_part = rel(out.double() - (dr.double() if dres else 0.0), ref)
    if dres:
        ref = ref + dr.double()
    assert torch.isfinite(out).all()
    assert ln_part < tol(4e-3, 6e-4), ln_part
    assert rel(out, ref) < tol(4e-3, 6e-4)
    if mirror:
        assert rel(mir.float(), ref) < tol(5e-3, 8e-4)
    assert rel(dw, (dn * xhat).sum(0)) < 2e-3 and rel(db, dn.sum(0)) < 2e-3 and rel(cs, ref.sum(0)) < tol(5e-3, 2e-3)


@pytest.mark.parametrize("M,N,K", [(1000, 512, 512), (130, 72, 40), (33000, 96, 48)])
def test_gemm_store_both_outputs(ops, M, N, K):
    """fp32 output + bf16 mirror + residual in one launch (the level-boundary GEMMs of the networks)."""
    from dcpt_b200.lib import GemmDesc
    g = torch.Generator(device="cuda").manual_seed(M + N)
    A = bf(torch.randn(M, K, device="cuda", generator=g))
    B = bf(torch.randn(N, K, device="cuda", generator=g) / K ** 0.5)
    bias = torch.randn(N, device="cuda", generator=g)
    resid = torch.randn(M, N, device="cuda", generator=g)
    o32 = torch.empty(M, N, device="cuda")
    o16 = torch.empty(M, N, device="cuda", dtype=OPD())
    d = GemmDesc()
    for k, v in dict(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=0, out_f32=o32, out_bf16=o16, ldo=N, bias=bias,
                     resid=resid, ldr=N).items():
        setattr(d, k, v.data_ptr() if torch.is_tensor(v) else v)
    ops.gemm_ex(d)
    ref = A.float() @ B.float().t() + bias + resid
    assert rel(o32, ref) < 2e-5 and rel(o16.float(), ref) < 4e-3
    # bf16 output with a residual and no fp32 output
    d.out_f32 = None
    o16.zero_()
    ops.gemm_ex(d)
    assert rel(o16.float(), ref) < 4e-3


WGRAD_SHAPES = [(128, 64, 4096), (256, 128, 1000), (1024, 512, 16384), (64, 64, 65536), (16, 8, 200), (512, 2048, 4096)]


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("O_,I_,Mpx", WGRAD_SHAPES)
def test_gemm_wgrad_mn_major(ops, impl, O_, I_, Mpx):
    """dW[O,I] = dY[Mpx,O]^T X[Mpx,I]: MN-major operands, split-K, fp32 atomics (accumulates)."""
    if impl == 1 and O_ * I_ * Mpx > 2 ** 31:
        pytest.skip("too slow on CUDA cores")
    g = torch.Generator(device="cuda").manual_seed(O_ + I_)
    dY = bf(torch.randn(Mpx, O_, device="cuda", generator=g))
    X = bf(torch.randn(Mpx, I_, device="cuda", generator=g))
    ref = dY.float().t() @ X.float()
    for splits in (1, 7):
        acc = torch.ones(O_, I_, device="cuda")
        ops.gemm(dY, X, a_mn=True, b_mn=True, splits=splits, accumulate_into=acc, impl=impl)
        # fp32 accumulation over Mpx terms in a different order than torch: error grows ~ sqrt(Mpx) * 2^-24
        assert rel(acc - 1.0, ref) < max(3e-5, 6e-7 * Mpx ** 0.5), splits


def _desc(ops, **kw):
    from dcpt_b200.lib import GemmDesc
    d = GemmDesc()
    for k, v in kw.items():
        if torch.is_tensor(v):
            v = v.data_ptr()
        setattr(d, k, v)
    return d


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("M,C", [(256, 64), (1000, 16), (512, 512), (130, 8)])
def test_gemm_gate_epilogue(ops, impl, M, C):
    """conv4 + SimpleGate (nafnet_arch.py:180-181): pair-interleaved weights, x4 and sg outputs."""
    g = torch.Generator(device="cuda").manual_seed(C)
    n2 = bf(torch.randn(M, C, device="cuda", generator=g))
    W4 = bf(torch.randn(2 * C, C, device="cuda", generator=g) / C ** 0.5)
    b4 = torch.randn(2 * C, device="cuda", generator=g)
    p = torch.arange(2 * C, device="cuda")
    orig = ((p % 16) // 8) * C + (p // 16) * 8 + (p % 8)  # packed row -> original out-channel
    W4p, b4p = W4[orig].contiguous(), b4[orig].contiguous()
    x4 = torch.empty(M, 2 * C, dtype=OPD(), device="cuda")
    sg = torch.empty(M, C, dtype=OPD(), device="cuda")
    d = _desc(ops, M=M, N=2 * C, K=C, A=n2, lda=C, B=W4p, ldb=C, splits=1, epilogue=1, out_bf16=x4, ldo=2 * C, bias=b4p,
              out2_bf16=sg, ldo2=C, C=C)
    ops.gemm_ex(d, impl)
    ref = n2.float() @ W4.float().t() + b4
    assert rel(x4.float(), ref) < 4e-3
    x4r = x4.float()
    assert rel(sg.float(), x4r[:, :C] * x4r[:, C:]) < 4e-3


@pytest.mark.parametrize("M,C", [(256, 64), (1000, 32), (512, 512), (33000, 128), (130, 96), (4099, 1024)])
def test_gemm_gate32_epilogue(ops, M, C):
    """conv4 + SimpleGate with 32-wide pair packing and the TMA-tiled epilogue (epilogue id 7, C % 32 == 0)."""
    g = torch.Generator(device="cuda").manual_seed(C + 7)
    n2 = bf(torch.randn(M, C, device="cuda", generator=g))
    W4 = bf(torch.randn(2 * C, C, device="cuda", generator=g) / C ** 0.5)
    b4 = torch.randn(2 * C, device="cuda", generator=g)
    p = torch.arange(2 * C, device="cuda")
    orig = ((p % 64) // 32) * C + (p // 64) * 32 + (p % 32)  # packed row -> original out-channel
    W4p, b4p = W4[orig].contiguous(), b4[orig].contiguous()
    x4 = torch.full((M, 2 * C), float("nan"), dtype=OPD(), device="cuda")
    sg = torch.full((M, C), float("nan"), dtype=OPD(), device="cuda")
    d = _desc(ops, M=M, N=2 * C, K=C, A=n2, lda=C, B=W4p, ldb=C, splits=1, epilogue=7, out_bf16=x4, ldo=2 * C, bias=b4p,
              out2_bf16=sg, ldo2=C, C=C)
    ops.gemm_ex(d, 0)
    ref = n2.float() @ W4.float().t() + b4
    assert rel(x4.float(), ref) < 4e-3
    x4r = x4.float()
    assert rel(sg.float(), x4r[:, :C] * x4r[:, C:]) < 4e-3


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("M,C", [(256, 64), (1000, 16), (512, 512), (33000, 128), (130, 96)])
def test_gemm_gate_bwd_epilogue(ops, impl, M, C):
    g = torch.Generator(device="cuda").manual_seed(C + 1)
    dout = bf(torch.randn(M, C, device="cuda", generator=g))
    W5t = bf(torch.randn(C, C, device="cuda", generator=g) / C ** 0.5)  # [in, out]
    x4 = bf(torch.randn(M, 2 * C, device="cuda", generator=g))
    dx4 = torch.empty(M, 2 * C, dtype=OPD(), device="cuda")
    d = _desc(ops, M=M, N=C, K=C, A=dout, lda=C, B=W5t, ldb=C, splits=1, epilogue=2, out_bf16=dx4, ldo=2 * C, aux_bf16=x4,
              ldaux=2 * C, C=C)
    ops.gemm_ex(d, impl)
    dsg = dout.float() @ W5t.float().t()
    ref = torch.cat([dsg * x4.float()[:, C:], dsg * x4.float()[:, :C]], 1)
    assert rel(dx4.float(), ref) < 4e-3


@pytest.mark.parametrize("impl", [0, 1], ids=["tcgen05", "simt"])
@pytest.mark.parametrize("N,H,W,Cin", [(2, 8, 8, 64), (1, 5, 7, 32), (2, 16, 16, 256)])
def test_gemm_pixshuf_epilogue(ops, impl, N, H, W, Cin):
    """ups[i] = 1x1 conv (no bias) + PixelShuffle(2), then + skip (nafnet_arch.py:238-242,264-265)."""
    import torch.nn.functional as F
    g = torch.Generator(device="cuda").manual_seed(Cin)
    x = bf(torch.randn(N, H, W, Cin, device="cuda", generator=g))
    Wu = bf(torch.randn(2 * Cin, Cin, device="cuda", generator=g) / Cin ** 0.5)
    Cseg = Cin // 2
    skip = torch.randn(N, 2 * H, 2 * W, Cseg, device="cuda", generator=g)
    p = torch.arange(2 * Cin, device="cuda")
    orig = (p % Cseg) * 4 + p // Cseg
    Wp = Wu[orig].contiguous()
    out = torch.empty_like(skip)
    mirror = torch.empty(skip.shape, dtype=OPD(), device="cuda")
    d = _desc(ops, M=N * H * W, N=2 * Cin, K=Cin, A=x, lda=Cin, B=Wp, ldb=Cin, splits=1, epilogue=3, out_f32=out,
              out_bf16=mirror, resid=skip, H=H, W=W, Cseg=Cseg)
    ops.gemm_ex(d, impl)
    y = F.conv2d(x.float().permute(0, 3, 1, 2), Wu.float()[:, :, None, None])
    ref = F.pixel_shuffle(y, 2).permute(0, 2, 3, 1) + skip
    assert rel(out, ref) < 2e-5
    assert rel(mirror.float(), ref) < 4e-3


# ------------------------------------------------------------------ LayerNorm
@pytest.mark.parametrize("M,C", [(70, 24), (1000, 64), (4096, 128), (777, 512), (300, 1024), (5, 8), (4099, 512), (16384, 512)])
def test_layernorm_fwd_bwd(ops, M, C):
    g = torch.Generator().manual_seed(C)
    x = torch.randn(M, C, generator=g) * 1.7 + 0.4
    w = 1 + 0.2 * torch.randn(C, generator=g)
    b = 0.2 * torch.randn(C, generator=g)
    dn = torch.randn(M, C, generator=g).to(OPD())
    dres = torch.randn(M, C, generator=g)
    x4 = x.t().reshape(1, C, M, 1)  # NCHW view of the same rows
    y, y_hat, var = O.layernorm2d_fwd(x4, w, b)
    dx_ref, dw_ref, db_ref = O.layernorm2d_bwd(dn.float().t().reshape(1, C, M, 1), y_hat, var, w)
    out, stats = ops.layernorm2d_fwd(x.cuda(), w.cuda(), b.cuda())
    assert rel(out.float(), y[0, :, :, 0].t()) < 4e-3                      # bf16 output
    assert rel(stats[:, 0], x.mean(1)) < 1e-5
    assert rel(stats[:, 1], 1 / torch.sqrt(x.var(1, unbiased=False) + 1e-6)) < 1e-5
    dx, dxb, dw, db, cs = ops.layernorm2d_bwd(dn.cuda(), x.cuda(), stats, w.cuda(), dres.cuda())
    ref = dx_ref[0, :, :, 0].t() + dres
    assert rel(dx, ref) < 1e-5
    assert rel(dxb.float(), ref) < 4e-3
    assert rel(dw, dw_ref) < 1e-4 and rel(db, db_ref) < 1e-4
    assert rel(cs, ref.sum(0)) < 1e-4
    if M >= 2048:  # no residual gradient (dres = NULL) through the same (wide-row, bulk-copy pipelined) kernel
        dx2 = ops.layernorm2d_bwd(dn.cuda(), x.cuda(), stats, w.cuda(), None)[0]
        assert rel(dx2, dx_ref[0, :, :, 0].t()) < 1e-5


def test_layernorm_golden(ops, golden_dir):
    z = np.load(os.path.join(golden_dir, "layernorm2d.npz"))
    x = torch.from_numpy(z["x"])
    N, C, H, W = x.shape
    rows = lambda t: torch.as_tensor(t).permute(0, 2, 3, 1).reshape(-1, C).contiguous().cuda()
    out, stats = ops.layernorm2d_fwd(rows(x), torch.from_numpy(z["weight"]).cuda(), torch.from_numpy(z["bias"]).cuda())
    assert rel(out.float(), rows(z["y"])) < 4e-3
    dy = rows(z["dy"]).to(OPD())
    dx, _, dw, db, _ = ops.layernorm2d_bwd(dy, rows(x), stats, torch.from_numpy(z["weight"]).cuda())
    # dy was rounded to bf16 (the kernel's input type) -> compare against the oracle on the rounded dy
    dyr = dy.float().reshape(N, H, W, C).permute(0, 3, 1, 2).cpu()
    _, y_hat, var = O.layernorm2d_fwd(x, torch.from_numpy(z["weight"]), torch.from_numpy(z["bias"]))
    dx_ref, dw_ref, db_ref = O.layernorm2d_bwd(dyr, y_hat, var, torch.from_numpy(z["weight"]))
    assert rel(dx, rows(dx_ref)) < 1e-5 and rel(dw, dw_ref) < 1e-4 and rel(db, db_ref) < 1e-4
    assert rel(dx, rows(z["dx"])) < 4e-3  # vs the reference's own output (bf16-rounded dy)


# ------------------------------------------------------------------ dw3x3 + SimpleGate
@pytest.mark.parametrize("N,H,W,C", [(2, 9, 7, 16), (1, 16, 16, 64), (2, 5, 33, 8), (1, 3, 3, 512)])
def test_dwconv_gate_fwd(ops, N, H, W, C):
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(C + H)
    u = torch.randn(N, H, W, 2 * C, generator=g).to(OPD())
    w2 = torch.randn(2 * C, 1, 3, 3, generator=g) / 3
    b2 = torch.randn(2 * C, generator=g) * 0.1
    v = F.conv2d(u.float().permute(0, 3, 1, 2), w2, b2, padding=1, groups=2 * C)
    gref = O.simple_gate(v).permute(0, 2, 3, 1)
    gk, pool = ops.dwconv3x3_gate_fwd(u.cuda(), w2.cuda(), b2.cuda())
    assert rel(gk.float(), gref) < 4e-3
    assert rel(pool, gk.float().sum((1, 2))) < 1e-5     # pool sums the stored (rounded) g
