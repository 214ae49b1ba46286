"""Per-kernel timing of ONE NAFBlock forward + backward through the C ABI (the unit the step is made of: the 29 blocks at
C = 512 are 52 % of the NAFNet-w64 step).  A 10-second alternative to a full bench.py run when iterating on one kernel.

    python tools/block_bench.py [--C 512] [--N 16] [--H 32] [--W 32] [--iters 20]

Prints the CUDA-event time of the block forward, the block backward (median over --iters, launches back to back, the
block's ~100 MB working set stays in L2 as it does inside the network) and the per-kernel table of one profiled pass
(dcpt_prof_*: events around every launch; tiny kernels read ~8 us there because of the event-bracketed eager launch).
"""
import argparse
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from dcpt_b200 import ops  # noqa: E402
from dcpt_b200.lib import load_library  # noqa: E402


def prof_rows(lib):
    buf = ctypes.create_string_buffer(1 << 16)
    lib.dcpt_prof_dump(buf, len(buf))
    rows = []
    for ln in buf.value.decode().splitlines():
        tag, n, ms, fl, by = ln.split("\t")
        rows.append((tag, int(n), float(ms), float(fl), float(by)))
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--C", type=int, default=512)
    ap.add_argument("--N", type=int, default=16)
    ap.add_argument("--H", type=int, default=32)
    ap.add_argument("--W", type=int, default=32)
    ap.add_argument("--iters", type=int, default=20)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    lib = load_library()
    C = a.C
    g = torch.Generator(device=dev).manual_seed(0)
    rn = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).contiguous()  # noqa: E731
    shapes = {"beta": (1, C, 1, 1), "gamma": (1, C, 1, 1), "conv1.weight": (2 * C, C, 1, 1), "conv1.bias": (2 * C,),
              "conv2.weight": (2 * C, 1, 3, 3), "conv2.bias": (2 * C,), "conv3.weight": (C, C, 1, 1), "conv3.bias": (C,),
              "sca.1.weight": (C, C, 1, 1), "sca.1.bias": (C,), "conv4.weight": (2 * C, C, 1, 1), "conv4.bias": (2 * C,),
              "conv5.weight": (C, C, 1, 1), "conv5.bias": (C,), "norm1.weight": (C,), "norm1.bias": (C,), "norm2.weight": (C,),
              "norm2.bias": (C,)}
    params = []
    for k in ops.NAFBLOCK_PARAM_ORDER:
        shp = shapes[k]
        is_conv = len(shp) == 4 and k.endswith("weight")
        params.append(rn(*shp, sc=1.0 / (shp[1] * shp[2] * shp[3]) ** 0.5 if is_conv else 0.3))
    blk = ops.NAFBlockOp(params)
    x = rn(a.N, a.H, a.W, C)
    dout = rn(a.N, a.H, a.W, C, sc=1e-3)

    def timed(fn):
        for _ in range(3):
            fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(a.iters):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); fn(); e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        ts.sort()
        return ts[len(ts) // 2]

    out, saved, _ = blk.forward(x)
    ms_f = timed(lambda: blk.forward(x))
    ms_b = timed(lambda: blk.backward(x, saved, dout))
    M = a.N * a.H * a.W
    fl = (12.0 * C * C + 36.0 * C) * M
    print(f"NAFBlock C={C} M={M}: fwd {ms_f * 1e3:.1f} us ({fl / ms_f / 1e9:.0f} TFLOP/s), bwd {ms_b * 1e3:.1f} us "
          f"({2 * fl / ms_b / 1e9:.0f} TFLOP/s)   [eager launches incl. host; inside the network the block is graph-replayed]")
    lib.dcpt_prof_enable(2)
    _, saved, _ = blk.forward(x)
    blk.backward(x, saved, dout)
    rows = prof_rows(lib)
    lib.dcpt_prof_enable(0)
    rows.sort(key=lambda r: -r[2])
    tot = sum(r[2] for r in rows) or 1.0
    print(f"{'kernel':52s} {'n':>3s} {'us/launch':>10s} {'share':>6s} {'TFLOP/s':>8s} {'GB/s':>7s}")
    for tag, n, ms, flops, by in rows:
        print(f"{tag[:52]:52s} {n:3d} {ms * 1e3 / n:10.1f} {ms / tot:6.3f} {flops / (ms * 1e-3 + 1e-12) / 1e12:8.1f} "
              f"{by / (ms * 1e-3 + 1e-12) / 1e9:7.0f}")


if __name__ == "__main__":
    main()
