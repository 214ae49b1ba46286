"""Same-box A/B of Restormer / PromptIR 128 x 128 inference (b1, b16) under an environment switch: python tools/restormer_ab.py VAR=VALUE"""
import os, sys, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    import torch
    from basicsr.archs import build_network
    net = build_network(dict(type="Restormer", window_size=8)).cuda().eval()
    res = []
    for B in (1, 16):
        x = torch.rand(B, 3, 128, 128, device="cuda")
        with torch.no_grad():
            for _ in range(5):
                net(x)
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(20):
                net(x)
            b.record()
            torch.cuda.synchronize()
        res.append(a.elapsed_time(b) / 20)
    print("RESULT b1 %.3f ms  b16 %.3f ms" % tuple(res))
else:
    kv = sys.argv[1] if len(sys.argv) > 1 else "X=0"
    k, v = kv.split("=", 1)
    for r in range(2):
        for env in ({}, {k: v}):
            out = subprocess.run([sys.executable, __file__, "child"], env=dict(os.environ, **env), capture_output=True, text=True).stdout
            print(("default" if not env else kv), [l for l in out.splitlines() if l.startswith("RESULT")][-1:])
