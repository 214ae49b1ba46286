"""Per-CTA event timeline of one tcgen05 GEMM launch (debug build: `python -m dcpt_b200.build --trace`, run with
DCPT_LIB=dcpt_b200/libdcpt_sm100_trace.so).  Answers "where do the 25 us of a 16384x512x512 launch go" with device clocks
instead of guesses: kernel entry, end of the prologue, first operands landed, MMAs of each tile issued/retired, epilogue done.

    DCPT_LIB=dcpt_b200/libdcpt_sm100_trace.so python tools/gemm_trace.py [M N K]
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dcpt_b200 import ops  # noqa: E402
from dcpt_b200.lib import GemmDesc, load_library  # noqa: E402

NAMES = ["entry", "prologue done", "first operands landed (MMA warp)", "tile0 MMAs issued", "last tile MMAs issued",
         "tile0 accumulator ready (epi warp)", "tile0 epilogue done", "last tile accumulator ready", "last tile epilogue done",
         "stores drained", "exit", "LN: pass 1 of the slab done", "LN: statistics exchanged", "LN: pass 2 done"]


def main():
    args = [a for a in sys.argv[1:] if not a.startswith("--")]
    M, N, K = (int(v) for v in args[:3]) if len(args) >= 3 else (16384, 512, 512)
    with_ln = "--ln" in sys.argv
    dev = torch.device("cuda", 0)
    lib = load_library()
    lib.dcpt_debug_set_trace.argtypes = [ctypes.c_void_p]
    A = torch.randn(M, K, device=dev).bfloat16()
    B = torch.randn(N, K, device=dev).bfloat16()
    out = torch.empty(M, N, dtype=torch.float32, device=dev)
    res = torch.randn(M, N, device=dev)
    bias = torch.randn(N, device=dev)
    d = GemmDesc()
    for k, v in dict(M=M, N=N, K=K, A=A, lda=K, B=B, ldb=K, splits=1, epilogue=0, out_f32=out, ldo=N, resid=res, ldr=N, bias=bias).items():
        setattr(d, k, v.data_ptr() if torch.is_tensor(v) else v)
    if with_ln:   # fused LayerNorm of the output rows (EpiParams::ln_*)
        lw, lb = torch.ones(N, device=dev), torch.zeros(N, device=dev)
        ln = torch.empty(M, N, dtype=torch.bfloat16, device=dev)
        stats = torch.empty(M, 2, device=dev)
        for k, v in dict(ln_weight=lw, ln_bias=lb, ln_out=ln, ld_ln=N, ln_stats=stats).items():
            setattr(d, k, v.data_ptr() if torch.is_tensor(v) else v)
        d.ln_eps = 1e-6
    tr = torch.zeros(148 * 16 * 2, dtype=torch.int64, device=dev)
    for _ in range(3):
        ops.gemm_ex(d)
    torch.cuda.synchronize()
    # back-to-back launch rate (includes the launch gaps)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(20):
        ops.gemm_ex(d)
    e1.record()
    torch.cuda.synchronize()
    print(f"GEMM {M}x{N}x{K} store+resid{'+LN' if with_ln else ''}: {e0.elapsed_time(e1) / 20 * 1e3:.2f} us per launch back to back (eager)")
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(20):
            ops.gemm_ex(d)
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    print(f"                               {e0.elapsed_time(e1) / 20 * 1e3:.2f} us per launch inside a CUDA graph of 20")
    assert lib.dcpt_debug_set_trace(ctypes.c_void_p(tr.data_ptr())) == 0
    ops.gemm_ex(d)
    ops.gemm_ex(d)
    torch.cuda.synchronize()
    t = tr.view(148, 16, 2).cpu()
    gt, ck = t[:, :, 0].double(), t[:, :, 1].double()
    used = gt[:, 0] > 0
    t0 = gt[used, 0].min()
    print(f"{int(used.sum())} CTAs traced; times in us relative to the first CTA's entry (globaltimer); clock64 deltas in cycles")
    print(f"{'event':44s} {'min':>8s} {'median':>8s} {'max':>8s}   cycles since entry (median)")
    for s, name in enumerate(NAMES):
        ok = used & (gt[:, s] > 0)
        if ok.sum() == 0:
            continue
        v = (gt[ok, s] - t0) / 1e3
        c = (ck[ok, s] - ck[ok, 0])
        print(f"{name:44s} {v.min():8.2f} {v.median():8.2f} {v.max():8.2f}   {c.median():10.0f}")
    two = used & (gt[:, 7] > 0)
    print(f"CTAs with 2 tiles: {int(two.sum())}, with 1 tile: {int((used & ~two).sum())}")
    lib.dcpt_debug_set_trace(None)


if __name__ == "__main__":
    main()
