"""Same-box A/B of the NAFNet-w64 inference forward (C2: 16 x 3 x 256 x 256) with / without the stores only a backward reads."""
import os, sys, subprocess, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1:
    import torch
    from basicsr.archs import build_network
    import bench as BM
    net = build_network(dict(type="NAFNetBaseline", window_size=16, **BM.CFG)).cuda().eval()
    x = torch.rand(16, 3, 256, 256, device="cuda")
    with torch.no_grad():
        for _ in range(5):
            net(x)
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        for _ in range(30):
            net(x)
        b.record()
        torch.cuda.synchronize()
    print("RESULT", a.elapsed_time(b) / 30)
else:
    for r in range(3):
        for keep in ("0", "1"):
            out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, DCPT_INFER_KEEP=keep), capture_output=True, text=True).stdout
            ms = [l for l in out.splitlines() if l.startswith("RESULT")]
            print("keep_all_stores=" + keep, ms[-1] if ms else out[-300:])
