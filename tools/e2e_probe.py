"""Where does the end-to-end step (public API, host buffers) spend its time beyond the device-resident step?
Prints wall ms per step for variants of bench.py's e2e loop and the host enqueue time of each stage."""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from basicsr.archs import build_network
import bench as BM

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = build_network(dict(type="NAFNetBaseline", window_size=16, **BM.CFG)).to(dev)
B = 16
h_inp = torch.rand(B, 3, 256, 256).pin_memory()
h_gt = torch.rand(B, 3, 256, 256).pin_memory()
d_inp, d_gt = h_inp.to(dev), h_gt.to(dev)
stages = {}

def step(h2d=True, item=True, prof=False):
    def mark(name, t0):
        if prof:
            stages[name] = stages.get(name, 0.0) + (time.perf_counter() - t0) * 1e3
        return time.perf_counter()
    t = time.perf_counter()
    lq = h_inp.to(dev, non_blocking=True) if h2d else d_inp
    tgt = h_gt.to(dev, non_blocking=True) if h2d else d_gt
    t = mark("h2d_enqueue", t)
    net.zero_grad(set_to_none=True)
    t = mark("zero_grad", t)
    out = net(lq)
    t = mark("forward_enqueue", t)
    loss = torch.nn.functional.l1_loss(out, tgt)
    t = mark("loss_enqueue", t)
    loss.backward()
    t = mark("backward_enqueue", t)
    if item:
        v = float(loss.item())
        t = mark("item_wait", t)
        return v
    return loss

def timed(name, n=10, **kw):
    for _ in range(3):
        step(**kw)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(n):
        step(**kw)
    torch.cuda.synchronize()
    print(f"{name}: {(time.perf_counter() - t0) * 1e3 / n:.3f} ms/step", flush=True)

timed("e2e (h2d + item)")
timed("no h2d")
timed("no item (sync at end only)", item=False)
timed("no h2d, no item", h2d=False, item=False)
stages.clear()
for _ in range(10):
    step(prof=True)
print({k: round(v / 10, 3) for k, v in stages.items()})
# GPU-side timeline of one step
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(3):
        step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
ev.sort(key=lambda e: e.time_range.start)
# GPU idle gaps > 50 us and the kernels around them, for the last step
if ev:
    t_end = ev[-1].time_range.end
    last = [e for e in ev if e.time_range.start > t_end - 27000]
    busy = sum(e.time_range.end - e.time_range.start for e in last)
    print(f"last ~27 ms window: {len(last)} GPU events, busy {busy / 1e3:.2f} ms")
    prev = None
    for e in last:
        if prev is not None and e.time_range.start - prev.time_range.end > 50:
            print(f"  gap {e.time_range.start - prev.time_range.end:.0f} us between '{prev.name[:50]}' and '{e.name[:50]}'")
        prev = e
    big = sorted(last, key=lambda e: -(e.time_range.end - e.time_range.start))[:8]
    for e in big:
        print(f"  {e.time_range.end - e.time_range.start:.0f} us  {e.name[:90]}")
