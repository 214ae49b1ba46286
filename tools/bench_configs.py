"""Timings of the other BASELINE.json configs on one B200 (the bench.py line covers configs[1] fwd+bwd):
  C2  NAFNet-w64 inference, batch 16 x 3 x 256 x 256              (configs[1], forward only)
  C3  Restormer (dim 48, [4,6,6,8] + 4 refinement) inference, 128 x 128 tiles, batch 1 and 16   (configs[2])
  C4  DCPT pretrain step: NAFNet-w64 on gt + hooked NAFNet-w64 on lq + classifier head, 8 x 3 x 256 x 256, one backward (configs[3])
  C5  fine-tune step, NAFNet-w64, 512 x 512 patches, batch 4, fwd + L1 + bwd        (configs[4])
All through the public API (registry-built modules), device-resident inputs, CUDA events, median of `--iters`.
    python tools/bench_configs.py [--iters 10] [--out gpurun_out/configs.json]
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.nn.functional as F  # noqa: E402
from basicsr.archs import build_network  # noqa: E402
from oracle import nafnet_oracle as O  # noqa: E402  (synthetic weight generators only)
from oracle import restormer_oracle as RO  # noqa: E402
from oracle import dchead_oracle as D  # noqa: E402

CFG = dict(width=64, enc_blk_nums=[1, 1, 1, 28], middle_blk_num=1, dec_blk_nums=[1, 1, 1, 1])


def timeit(fn, iters, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "configs.json"))
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev).manual_seed(0)
    res = {}

    net = build_network(dict(type="NAFNetBaseline", window_size=16, **CFG)).to(dev)
    net.load_state_dict(O.random_nafnet_state_dict(seed=0, **CFG), strict=True)
    x = torch.rand(16, 3, 256, 256, device=dev, generator=g)
    with torch.no_grad():
        ms = timeit(lambda: net(x), args.iters)
    res["C2_nafnet_w64_infer_b16_256"] = {"ms": round(ms, 3), "MPix/s": round(16 * 65536 / ms / 1e3, 2)}

    # C5: fine-tune step on 512 x 512 patches
    x5 = torch.rand(4, 3, 512, 512, device=dev, generator=g)
    t5 = torch.rand(4, 3, 512, 512, device=dev, generator=g)

    def ft_step():
        net.zero_grad(set_to_none=True)
        F.l1_loss(net(x5), t5).backward()
    ms = timeit(ft_step, args.iters)
    res["C5_nafnet_w64_finetune_b4_512"] = {"ms": round(ms, 3), "MPix/s": round(4 * 512 * 512 / ms / 1e3, 2)}

    # C4: DCPT pretrain step (degradation_classification_pretrain_model.py:133-169)
    dims = [64, 128, 256, 512]
    head = build_network(dict(type="PromptIR_NoImg_DC", feature_dims=dims, num_res_blocks=2, num_classes=5)).to(dev)
    head.load_state_dict(D.random_dchead_state_dict(dims, 2, 5, seed=1), strict=True)
    hook_outputs = []
    hooks = [m.register_forward_hook(lambda mod, i, o: hook_outputs.append(o)) for n, m in net.named_modules()
             if "decoder" in n and n.count(".") == 1]
    gt = torch.rand(8, 3, 256, 256, device=dev, generator=g)
    lq = torch.rand(8, 3, 256, 256, device=dev, generator=g)
    labels = torch.randint(0, 5, (8,), device=dev, generator=g)

    def dcpt_step():
        net.zero_grad(set_to_none=True); head.zero_grad(set_to_none=True)
        pix = net(gt, hook=False)
        hook_outputs.clear()
        l_pix = F.l1_loss(pix, gt)
        net(lq, hook=True)
        cls = head(lq, hook_outputs[::-1])
        (l_pix + F.cross_entropy(cls, labels)).backward()
        hook_outputs.clear()
    ms = timeit(dcpt_step, args.iters)
    res["C4_dcpt_pretrain_step_b8_256"] = {"ms": round(ms, 3), "MPix/s": round(8 * 65536 / ms / 1e3, 2),
                                           "note": "2 backbone forwards + classifier head + one backward, per GPU"}
    for h in hooks:
        h.remove()
    del net, head
    torch.cuda.empty_cache()

    rcfg = dict(dim=48, num_blocks=[4, 6, 6, 8], num_refinement_blocks=4, heads=[1, 2, 4, 8])
    rnet = build_network(dict(type="Restormer", window_size=8, **rcfg)).to(dev)
    for B in (1, 16):
        xr = torch.rand(B, 3, 128, 128, device=dev, generator=g)
        with torch.no_grad():
            ms = timeit(lambda: rnet(xr), args.iters)
        res[f"C3_restormer_infer_b{B}_128"] = {"ms": round(ms, 3), "MPix/s": round(B * 128 * 128 / ms / 1e3, 3),
                                               "TFLOP/s": round(B * 77.44e9 / (ms * 1e-3) / 1e12, 1)}
    # PromptIR (options/all_in_one/test/test_PromptIR_5d.yml: default ctor, 35.4 M parameters), inference on 128 x 128 tiles
    pnet = build_network(dict(type="PromptIR", window_size=8)).to(dev)
    for B in (1, 16):
        xp = torch.rand(B, 3, 128, 128, device=dev, generator=g)
        with torch.no_grad():
            ms = timeit(lambda: pnet(xp), args.iters)
        res[f"promptir_infer_b{B}_128"] = {"ms": round(ms, 3), "MPix/s": round(B * 128 * 128 / ms / 1e3, 3)}
    del pnet
    torch.cuda.empty_cache()
    # Restormer fine-tune step (not a BASELINE config; recorded for the training path): fwd + L1 + bwd, 128 x 128, batch 4
    xr = torch.rand(4, 3, 128, 128, device=dev, generator=g)
    tr = torch.rand(4, 3, 128, 128, device=dev, generator=g)

    def r_step():
        rnet.zero_grad(set_to_none=True)
        F.l1_loss(rnet(xr), tr).backward()
    ms = timeit(r_step, max(3, args.iters // 2))
    res["restormer_train_step_b4_128"] = {"ms": round(ms, 3), "MPix/s": round(4 * 128 * 128 / ms / 1e3, 3),
                                          "note": "public API (autograd Function), forward-with-save and backward replayed from CUDA graphs"}
    # DCPT pretrain step with the Restormer backbone (hook_names: decoder -> decoder_level{3,2,1}.body), 128 x 128, batch 4
    d = rcfg["dim"]
    rhead = build_network(dict(type="PromptIR_NoImg_DC", feature_dims=[2 * d, 2 * d, 4 * d], num_res_blocks=2, num_classes=5)).to(dev)
    r_hooked = []
    rhooks = [m.register_forward_hook(lambda mod, i, o: r_hooked.append(o)) for n, m in rnet.named_modules()
              if "decoder" in n and n.count(".") == 1]
    rl = torch.randint(0, 5, (4,), device=dev, generator=g)

    def r_dcpt_step():
        rnet.zero_grad(set_to_none=True); rhead.zero_grad(set_to_none=True)
        pix = rnet(tr, hook=False)
        r_hooked.clear()
        l_pix = F.l1_loss(pix, tr)
        rnet(xr, hook=True)
        cls = rhead(xr, r_hooked[::-1])
        (l_pix + F.cross_entropy(cls, rl)).backward()
        r_hooked.clear()
    ms = timeit(r_dcpt_step, max(3, args.iters // 2))
    res["restormer_dcpt_pretrain_step_b4_128"] = {"ms": round(ms, 3), "MPix/s": round(4 * 128 * 128 / ms / 1e3, 3),
                                                  "note": "2 Restormer forwards (one hooked) + classifier head + one backward; backbone passes replayed from CUDA graphs"}
    for h in rhooks:
        h.remove()
    print(json.dumps(res, indent=1))
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    json.dump(res, open(args.out, "w"), indent=1)


if __name__ == "__main__":
    main()
