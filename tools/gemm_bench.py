"""Micro-benchmark of the tcgen05 GEMM engine on the NAFNet-w64 shapes (CUDA events, L2 flushed between launches).
    python tools/gemm_bench.py [--ncu]     (--ncu: one launch per shape, for profiling)
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dcpt_b200 import ops  # noqa: E402
from dcpt_b200.lib import GemmDesc  # noqa: E402

dev = torch.device("cuda", 0)
ncu = "--ncu" in sys.argv
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def bench(name, fn, flops, nbytes, iters=10):
    fn()
    torch.cuda.synchronize()
    if ncu:
        return
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    t = ts[len(ts) // 2] * 1e-3
    print(f"{name:44s} {t * 1e6:8.1f} us  {flops / t / 1e12:7.1f} TFLOP/s  {nbytes / t / 1e9:7.0f} GB/s", flush=True)


def desc(**kw):
    d = GemmDesc()
    for k, v in kw.items():
        setattr(d, k, v.data_ptr() if torch.is_tensor(v) else v)
    return d


def run(M, N, K, kind):
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    if kind == "store_bf16":
        out = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        d = desc(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=0, out_bf16=out, ldo=N)
        nb = 2 * (M * K + N * K + M * N)
    elif kind == "store_resid":
        out = torch.empty(M, N, dtype=torch.float32, device=dev)
        res = torch.randn(M, N, device=dev)
        bias = torch.randn(N, device=dev)
        d = desc(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=0, out_f32=out, ldo=N, resid=res, ldr=N, bias=bias)
        nb = 2 * (M * K + N * K) + 8 * M * N
    elif kind == "gate":
        C = N // 2
        x4 = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        sg = torch.empty(M, C, dtype=torch.bfloat16, device=dev)
        bias = torch.randn(N, device=dev)
        d = desc(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=1, out_bf16=x4, ldo=N, out2_bf16=sg, ldo2=C, C=C, bias=bias)
        nb = 2 * (M * K + N * K) + 3 * M * N
    elif kind == "gate_bwd":
        x4 = torch.randn(M, 2 * N, device=dev).bfloat16()
        dx4 = torch.empty(M, 2 * N, dtype=torch.bfloat16, device=dev)
        d = desc(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=2, out_bf16=dx4, ldo=2 * N, aux_bf16=x4, ldaux=2 * N, C=N)
        nb = 2 * (M * K + N * K) + 8 * M * N
    elif kind.startswith("wgrad"):
        splits = int(kind[5:])
        A = torch.randn(K, M, device=dev).bfloat16()
        B = torch.randn(K, N, device=dev).bfloat16()
        out = torch.zeros(M, N, device=dev)
        d = desc(M=M, N=N, K=K, A=A, lda=M, a_mn=1, B=B, ldb=N, b_mn=1, splits=splits, epilogue=4, out_f32=out, ldo=N)
        nb = 2 * (M * K + N * K) + 4 * M * N * splits
    keep = (A, B, d)
    bench(f"{kind} {M}x{N}x{K}", lambda: ops.gemm_ex(d), 2.0 * M * N * K, nb)
    return keep


def run_ln(M, N, K):
    """STORE + fp32 residual + the consumer's LayerNorm fused (EpiParams::ln_*): conv3 -> norm2 / conv5 -> next norm1."""
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, dtype=torch.float32, device=dev)
    res = torch.randn(M, N, device=dev)
    bias, lw, lb = torch.randn(N, device=dev), torch.ones(N, device=dev), torch.zeros(N, device=dev)
    ln = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    st = torch.empty(M, 2, device=dev)
    d = desc(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=0, out_f32=out, ldo=N, resid=res, ldr=N, bias=bias, ln_weight=lw,
             ln_bias=lb, ln_out=ln, ld_ln=N, ln_stats=st, ln_eps=1e-6)
    keep = (A, B, out, res, bias, lw, lb, ln, st, d)
    bench(f"store_resid+LN {M}x{N}x{K}", lambda: ops.gemm_ex(d), 2.0 * M * N * K, 2 * (M * K + N * K) + 10 * M * N)
    return keep


def run_lnb(M, N, K):
    """dgrad GEMM with the LayerNorm backward fused (EpiParams::lnb_*): conv4 dgrad -> norm2', conv1 dgrad -> norm1'."""
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    x = torch.randn(M, N, device=dev)
    dres = torch.randn(M, N, device=dev)
    stats = torch.stack([x.mean(1), 1.0 / (x.var(1, unbiased=False) + 1e-6).sqrt()], 1).contiguous()
    w = torch.ones(N, device=dev)
    out = torch.empty(M, N, device=dev)
    mir = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
    dw, db, cs = (torch.zeros(N, device=dev) for _ in range(3))
    d = desc(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=0, out_f32=out, out_bf16=mir, ldo=N, lnb_x=x, ld_lnb=N, lnb_stats=stats,
             lnb_weight=w, lnb_dres=dres, lnb_dweight=dw, lnb_dbias=db, lnb_colsum=cs)
    keep = (A, B, x, dres, stats, w, out, mir, dw, db, cs, d)
    bench(f"dgrad+LN-bwd {M}x{N}x{K}", lambda: ops.gemm_ex(d), 2.0 * M * N * K, 2 * (M * K + N * K) + 18 * M * N)
    return keep


if ncu:   # one launch per shape, in this order: tools/ncu_traffic.py pairs the capture's launches with this manifest
    import json
    man = [dict(tag="gemm_tc<256,store_tma,k>", shape="16384x512x512", note="conv3 / conv5 at C=512: fp32 residual in, fp32 out"),
           dict(tag="gemm_tc<256,store_tma,k>+LN", shape="16384x512x512", note="same + fused LayerNorm of the output rows (bf16 n + stats)"),
           dict(tag="gemm_tc<256,store_tma,k>", shape="16384x1024x512", note="conv1 at C=512: bf16 out, 128-byte-row store boxes"),
           dict(tag="gemm_tc<256,store_tma,k>", shape="16384x512x1024", note="conv4 / conv1 dgrad at C=512: bf16 out"),
           dict(tag="gemm_tc<256,gate_tma,k>", shape="16384x1024x512", note="conv4 + SimpleGate epilogue"),
           dict(tag="gemm_tc<256,gate_bwd_tma,k>", shape="16384x512x512", note="conv5 dgrad + SimpleGate backward"),
           dict(tag="gemm_tc<256,lnbwd_tma,k>", shape="16384x512x1024", note="conv4 / conv1 dgrad at C=512 + fused LayerNorm backward (fp32 x, dres in; fp32 + bf16 dx out)")]
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(man, open(os.path.join(ROOT, "gpurun_out", "gemm_bench_manifest.json"), "w"), indent=1)
    keep_all = [run(16384, 512, 512, "store_resid"), run_ln(16384, 512, 512), run(16384, 1024, 512, "store_bf16"),
                run(16384, 512, 1024, "store_bf16"), run(16384, 1024, 512, "gate"), run(16384, 512, 512, "gate_bwd"),
                run_lnb(16384, 512, 1024)]
    sys.exit(0)

shapes = [(16384, 512, 512, "store_resid"), (16384, 512, 512, "store_bf16"), (16384, 1024, 512, "store_bf16"),
          (16384, 512, 1024, "store_bf16"), (16384, 1024, 512, "gate"), (16384, 512, 512, "gate_bwd"),
          (1024, 512, 16384, "wgrad10"), (512, 512, 16384, "wgrad19"), (8192, 8192, 8192, "store_bf16"),
          (1048576, 64, 64, "store_resid"), (1048576, 128, 64, "store_bf16"), (262144, 128, 128, "store_resid")]
for s in shapes:
    run(*s)
run_ln(16384, 512, 512)
run_ln(1048576, 64, 64)
run_lnb(16384, 512, 1024)
run_lnb(262144, 128, 256)
