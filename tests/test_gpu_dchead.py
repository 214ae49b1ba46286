"""GPU parity tests of the degradation-classifier head (PromptIR_NoImg_DC) and its kernels, through the C ABI, against the
CPU oracle / the reference's golden vectors.

Tolerances: trunk tensors are bf16 (one rounding per stored tensor, fp32 accumulation).
  implicit-GEMM 3x3 conv, bf16 out: rel-L2 <= 4e-3 vs fp32 conv of the same bf16 operands; fp32 out / wgrad: <= 3e-5
  head logits: <= 1e-2 rel-L2 (10 LN+ReLU blocks deep in bf16; the oracle's rounding hook predicts ~4e-3)
  head gradients: <= 4e-2 (bf16 gradient tensors through 10 blocks), median <= 2e-2
"""
import ctypes as C
import os

import numpy as np
import pytest
import torch

from dcpt_b200.lib import operand_dtype as OPD  # bf16 by default; fp16 for the DCPT_OPERAND=fp16 parity build
from tol import FP16, report, tol  # noqa: F401
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

from oracle import dchead_oracle as D  # noqa: E402


def rel(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return float((a - b).norm() / (b.norm() + 1e-30))


@pytest.fixture(scope="module")
def lib():
    from dcpt_b200.lib import load_library
    return load_library()


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


@pytest.mark.parametrize("N,H,W,Cin,Cout", [(2, 16, 16, 64, 64), (1, 9, 13, 16, 24), (2, 32, 32, 128, 128), (1, 8, 8, 256, 512),
                                            (3, 4, 4, 1024, 256), (1, 20, 36, 72, 40)])
def test_conv3x3_fwd_dgrad_wgrad(lib, N, H, W, Cin, Cout):
    from dcpt_b200.lib import check
    g = torch.Generator(device="cuda").manual_seed(Cin + Cout + H)
    x = torch.randn(N, H, W, Cin, device="cuda", generator=g).to(OPD())
    w = torch.randn(Cout, Cin, 3, 3, device="cuda", generator=g) / (3 * Cin ** 0.5)
    wq = w.to(OPD()).float()
    wp = torch.empty(lib.dcpt_conv3x3_packed_elems(Cout, Cin, 0), dtype=OPD(), device="cuda")
    wd = torch.empty(lib.dcpt_conv3x3_packed_elems(Cout, Cin, 1), dtype=OPD(), device="cuda")
    check(lib.dcpt_conv3x3_pack(_p(w), _p(wp), Cout, Cin, 0, _st()), "pack")
    check(lib.dcpt_conv3x3_pack(_p(w), _p(wd), Cout, Cin, 1, _st()), "pack")
    # forward
    y32 = torch.empty(N, H, W, Cout, device="cuda")
    y16 = torch.empty(N, H, W, Cout, dtype=OPD(), device="cuda")
    check(lib.dcpt_conv3x3_fwd(_p(x), _p(wp), _p(y16), _p(y32), N, H, W, Cin, Cout, _st()), "conv fwd")
    ref = F.conv2d(x.float().permute(0, 3, 1, 2), wq, padding=1).permute(0, 2, 3, 1)
    assert rel(y32, ref) < 3e-5
    assert rel(y16.float(), ref) < 4e-3
    # dgrad = conv of dy with the flipped/transposed operand
    dy = torch.randn(N, H, W, Cout, device="cuda", generator=g).to(OPD())
    dx = torch.empty(N, H, W, Cin, device="cuda")
    check(lib.dcpt_conv3x3_fwd(_p(dy), _p(wd), None, _p(dx), N, H, W, Cout, Cin, _st()), "conv dgrad")
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    wr = wq.clone().requires_grad_(True)
    F.conv2d(xr, wr, padding=1).backward(dy.float().permute(0, 3, 1, 2))
    assert rel(dx, xr.grad.permute(0, 2, 3, 1)) < 3e-5
    # wgrad (accumulates)
    dw = torch.ones(Cout, Cin, 3, 3, device="cuda")
    scratch = torch.empty(lib.dcpt_conv3x3_packed_elems(Cout, Cin, 0), device="cuda")
    check(lib.dcpt_conv3x3_wgrad(_p(dy), _p(x), _p(scratch), _p(dw), N, H, W, Cin, Cout, _st()), "conv wgrad")
    assert rel(dw - 1.0, wr.grad) < 1e-4


@pytest.mark.parametrize("M,C,relu,use_res", [(300, 64, 1, 1), (1000, 128, 1, 0), (77, 24, 0, 0), (512, 1024, 1, 1)])
def test_ln_act(lib, M, C, relu, use_res):
    from dcpt_b200.lib import check
    g = torch.Generator().manual_seed(C)
    x = (torch.randn(M, C, generator=g) * 1.3 + 0.2).to(OPD())
    res = torch.randn(M, C, generator=g).to(OPD()) if use_res else None
    w, b = 1 + 0.2 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    dy = torch.randn(M, C, generator=g).to(OPD())
    xr = x.float().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    rr = res.float().requires_grad_(True) if use_res else None
    y = D.layernorm_cf(xr.t().reshape(1, C, M, 1), wr, br)[0, :, :, 0].t()
    if use_res:
        y = y + rr
    if relu:
        y = F.relu(y)
    xc, yc = x.cuda(), torch.empty(M, C, dtype=OPD(), device="cuda")
    wc, bc, rc, dyc = w.cuda(), b.cuda(), (res.cuda() if use_res else None), dy.float().cuda()   # keep device copies alive
    stats = torch.empty(M, 2, device="cuda")
    check(lib.dcpt_ln_act_fwd(_p(xc), _p(wc), _p(bc), _p(rc), _p(yc), _p(stats), M, C, relu, 1e-6, _st()), "ln_act_fwd")
    assert rel(yc.float(), y) < 4e-3
    dx = torch.empty(M, C, dtype=OPD(), device="cuda")
    dres = torch.empty(M, C, dtype=torch.float32, device="cuda") if use_res else None
    dw, db = torch.zeros(C, device="cuda"), torch.zeros(C, device="cuda")
    check(lib.dcpt_ln_act_bwd(_p(dyc), _p(yc), _p(xc), _p(stats), _p(wc), _p(dx), _p(dres), _p(dw), _p(db), M, C, relu, _st()),
          "ln_act_bwd")
    # compare against the same mask the kernel used
    mk = (yc.float().cpu() > 0).float() if relu else torch.ones(M, C)
    xr.grad = None; wr.grad = None; br.grad = None
    lin = D.layernorm_cf(xr.t().reshape(1, C, M, 1), wr, br)[0, :, :, 0].t()
    lin.backward(dy.float() * mk)
    assert rel(dx.float(), xr.grad) < 6e-3
    assert rel(dw, wr.grad) < 1e-3 and rel(db, br.grad) < 1e-3
    if use_res:
        assert rel(dres.float(), dy.float() * mk) < 1e-6


def test_maxpool_mix_meanpool(lib):
    from dcpt_b200.lib import check
    g = torch.Generator().manual_seed(0)
    N, Ho, Wo, Cc, K = 2, 5, 7, 24, 5
    x = torch.randn(N, 2 * Ho, 2 * Wo, Cc, generator=g).to(OPD())
    xr = x.float().permute(0, 3, 1, 2).requires_grad_(True)
    y = F.relu(F.max_pool2d(xr, 2, 2))
    dy = torch.randn(N, Ho, Wo, Cc, generator=g).to(OPD())
    y.backward(dy.float().permute(0, 3, 1, 2))
    yc = torch.empty(N, Ho, Wo, Cc, dtype=OPD(), device="cuda")
    xg, dyg = x.cuda(), dy.float().cuda()
    check(lib.dcpt_maxpool2_relu_fwd(_p(xg), _p(yc), N, Ho, Wo, Cc, _st()), "maxpool")
    assert torch.equal(yc.cpu().float(), y.detach().permute(0, 2, 3, 1))
    dxc = torch.empty_like(x, device="cuda")
    check(lib.dcpt_maxpool2_relu_bwd(_p(xg), _p(dyg), _p(dxc), N, Ho, Wo, Cc, _st()), "maxpool bwd")
    assert rel(dxc.float(), xr.grad.permute(0, 2, 3, 1)) < 1e-6
    # mix
    prev = torch.randn(N * Ho * Wo, Cc, generator=g).to(OPD())
    feat = torch.randn(N, Ho, Wo, Cc, generator=g)
    mw = torch.tensor([0.37])
    z = torch.empty(N * Ho * Wo, Cc, dtype=OPD(), device="cuda")
    prevg, featg, mwg = prev.cuda(), feat.cuda(), mw.cuda()
    check(lib.dcpt_mix_fwd(_p(prevg), _p(featg), _p(mwg), _p(z), feat.numel(), _st()), "mix")
    assert rel(z.float(), prev.float() + 0.37 * feat.reshape(-1, Cc)) < 4e-3
    dz = torch.randn(N * Ho * Wo, Cc, generator=g).to(OPD())
    dfeat, dmw, dzg = torch.empty_like(feat, device="cuda"), torch.zeros(1, device="cuda"), dz.float().cuda()
    check(lib.dcpt_mix_bwd(_p(dzg), _p(featg), _p(mwg), _p(dfeat), _p(dmw), feat.numel(), _st()), "mix bwd")
    assert rel(dfeat, 0.37 * dz.float().reshape(feat.shape)) < 1e-6
    assert abs(float(dmw) - float((dz.float() * feat.reshape(-1, Cc)).sum())) < 1e-3 * float(dz.float().abs().sum())
    # mean pool + fc
    xm = torch.randn(N, Ho * Wo, Cc, generator=g).to(OPD())
    w, b = torch.randn(K, Cc, generator=g), torch.randn(K, generator=g)
    pooled, logits = torch.empty(N, Cc, device="cuda"), torch.empty(N, K, device="cuda")
    xmg, wg, bg = xm.cuda(), w.cuda(), b.cuda()
    check(lib.dcpt_meanpool_fc_fwd(_p(xmg), _p(wg), _p(bg), _p(pooled), _p(logits), N, Ho * Wo, Cc, K, _st()), "fc")
    xmr = xm.float().requires_grad_(True)
    wr, br = w.clone().requires_grad_(True), b.clone().requires_grad_(True)
    lr = F.linear(xmr.mean(1), wr, br)
    assert rel(logits, lr) < 1e-5
    dl = torch.randn(N, K, generator=g)
    lr.backward(dl)
    dw, db = torch.zeros(K, Cc, device="cuda"), torch.zeros(K, device="cuda")
    dxm, dlg = torch.empty(xm.shape, device="cuda"), dl.cuda()
    check(lib.dcpt_meanpool_fc_bwd(_p(dlg), _p(pooled), _p(wg), _p(dw), _p(db), _p(dxm), N, Ho * Wo, Cc, K, _st()), "fc bwd")
    assert rel(dw, wr.grad) < 1e-5 and rel(db, br.grad) < 1e-5 and rel(dxm, xmr.grad) < 1e-5


class ReplayActs(D.Acts):
    """Replays the ReLU masks and max-pool winners recorded by the CUDA forward (from its saved tensors)."""

    def __init__(self, ctx, dims, nb):
        self.m, self.pool = {}, {}

        def nchw(t, shp):
            N, H, W, Cc = shp
            return t.float().cpu().reshape(N, H, W, Cc).permute(0, 3, 1, 2)

        def blocks(prefix, saved, shp):
            N, H, W, f = shp
            for j, (x, t1, a, s1, t2, b, s2, t3, out, s3) in enumerate(saved):
                p = f"{prefix}{j}."
                self.m[p + "a"] = (nchw(a, (N, H, W, 2 * f)) > 0).float()
                self.m[p + "b"] = (nchw(b, (N, H, W, 2 * f)) > 0).float()
                self.m[p + "out"] = (nchw(out, (N, H, W, f)) > 0).float()

        for i, (blk, xs, t, (N, H, W, Cc, Cn)) in enumerate(ctx["stages"]):
            blocks(f"bottleneck_layers.{i}.", blk, (N, H, W, Cc))
            tv = nchw(t, (N, H, W, Cn))
            mx, idx = F.max_pool2d(tv, 2, 2, return_indices=True)
            self.pool[f"downsample_layers.{i}"] = (idx, (mx > 0).float())
        blocks("last_stage.", ctx["last"], ctx["shp"])

    def relu(self, t, name):
        return t * self.m[name]

    def pool_relu(self, t, name):
        idx, pos = self.pool[name]
        return t.flatten(2).gather(2, idx.flatten(2)).view_as(idx) * pos


def test_dchead_golden(golden_dir):
    """Registry-built head vs the reference's golden logits, and its backward vs the oracle's.

    The gradient comparison REPLAYS the CUDA forward's ReLU masks / max-pool winners in the oracle.  Reason: with ~24
    ReLU layers, rounding activations to bf16 flips ~0.2 % of the ReLU decisions per layer, which moves early-layer
    gradients by 20-40 % relative to fp32 (measured with the oracle alone at widths 8..256, logits move by 1e-3) - a
    property of ANY bf16-operand implementation of this head, the reference under autocast included.  On a fixed
    piecewise-linear branch the network is smooth and kernels can be compared tightly; the deviation from the fp32
    golden gradients is printed for the record."""
    from basicsr.archs import build_network
    from oracle.nafnet_oracle import bf16_ste
    z = np.load(os.path.join(golden_dir, "dchead.npz"))
    dims = z["dims"].tolist()
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}
    head = build_network(dict(type="PromptIR_NoImg_DC", feature_dims=dims, num_res_blocks=2, num_classes=5)).cuda()
    head.load_state_dict(sd, strict=True)
    feats_cpu = [torch.from_numpy(z[f"feat{i}"]) for i in range(len(dims))]
    labels = torch.from_numpy(z["labels"])
    eng = head._engine
    params = [p.detach() for p in head.parameters()]
    fh = [f.permute(0, 2, 3, 1).contiguous().cuda() for f in feats_cpu]
    logits, ctx = eng.forward(params, fh)
    assert rel(logits, z["logits"]) < 1e-2                                  # vs the reference's fp32 logits
    dlogits = (torch.softmax(logits, 1) - F.one_hot(labels.cuda(), 5).float()) / logits.shape[0]   # d CE(mean) / d logits
    dfeats, grads = eng.backward(params, ctx, dlogits)
    # oracle on the same branch, same storage roundings
    lh = {k: (v.to(OPD()).float() if v.dim() == 4 else v).clone().requires_grad_(True) for k, v in sd.items()}
    fe = [f.clone().requires_grad_(True) for f in feats_cpu]
    o_logits = D.dchead_fwd(fe, lh, q=bf16_ste, acts=ReplayActs(ctx, dims, 2))
    assert rel(logits, o_logits) < 4e-3
    o_logits.backward(dlogits.cpu())
    for i in range(len(dims)):
        assert rel(dfeats[i].permute(0, 3, 1, 2), fe[i].grad) < 3e-2, (i, rel(dfeats[i].permute(0, 3, 1, 2), fe[i].grad))
    errs = {k: rel(g, lh[k].grad) for k, g in zip(eng.names, grads)}
    print("grad err vs replayed oracle: max %.3g median %.3g | vs fp32 golden max %.3g" %
          (max(errs.values()), float(np.median(list(errs.values()))), max(rel(g, z["g." + k]) for k, g in zip(eng.names, grads))))
    assert max(errs.values()) < 4e-2, {k: v for k, v in errs.items() if v > 2e-2}
    assert float(np.median(list(errs.values()))) < 1.5e-2
    # Against the reference's own fp32 golden gradients (no replay): per parameter group
    eg = {k: rel(g, z["g." + k]) for k, g in zip(eng.names, grads)}
    groups = {}
    for k, v in eg.items():
        gname = ".".join(k.split(".")[:2]) if k[0] in "bl" else k.split(".")[0]
        groups.setdefault(gname, []).append(v)
    for gname, vs in groups.items():
        report(f"DC head grads vs fp32 golden [{gname}]", worst=max(vs), median=float(np.median(vs)))
    # Asserted where few ReLU / max-pool decisions lie downstream, so that flipped decisions cannot dominate:
    # the classifier and the last stage.  Earlier stages sit behind up to 24 ReLU layers: every 16-bit rounding of an
    # activation near zero flips a decision of the piecewise-linear network, and their deviation is the reported, branch-
    # dependent figure above (bf16 build 0.16-0.65; the replayed-branch comparison above is the kernel check for them).
    assert max(groups["fc"]) < tol(1e-2, 1e-3), groups["fc"]   # measured 2.9e-3 / 2.6e-4
    assert max(groups["last_stage.0"] + groups["last_stage.1"]) < tol(8e-2, 3e-2)   # measured 4.0e-2 (bf16 build) / 1.8e-2 (fp16 build)
    report("DC head vs fp32 golden", logits=rel(logits, z["logits"]),
           dfeats_worst=max(rel(dfeats[i].permute(0, 3, 1, 2), torch.from_numpy(z[f"dfeat{i}"])) for i in range(len(dims)))
           if "dfeat0" in z.files else float("nan"))


def test_dcpt_pretrain_step():
    """DCPTModel.optimize_parameters (models/degradation_classification_pretrain_model.py:133-169): pixel forward on gt,
    hooked forward on lq, classifier on the hooked decoder features, L1 + CE, ONE backward, both optimizers step.
    Checked against the same step on the CPU oracle."""
    from basicsr.archs import build_network
    from oracle import nafnet_oracle as O
    cfg = dict(width=16, enc_blk_nums=[1, 1], middle_blk_num=1, dec_blk_nums=[1, 1])
    dims = [16, 32]                                            # decoder features reversed: fine -> coarse
    sd_g = O.random_nafnet_state_dict(seed=0, **cfg)
    sd_h = D.random_dchead_state_dict(dims, 2, 5, seed=1)
    net = build_network(dict(type="NAFNetBaseline", **cfg)).cuda()
    head = build_network(dict(type="PromptIR_NoImg_DC", feature_dims=dims, num_res_blocks=2, num_classes=5)).cuda()
    net.load_state_dict(sd_g, strict=True)
    head.load_state_dict(sd_h, strict=True)
    g = torch.Generator().manual_seed(3)
    gt, lq = torch.rand(2, 3, 32, 32, generator=g), torch.rand(2, 3, 32, 32, generator=g)
    labels = torch.tensor([4, 1])
    hook_outputs = []
    hooks = [m.register_forward_hook(lambda mod, i, o: hook_outputs.append(o)) for n, m in net.named_modules()
             if "decoder" in n and n.count(".") == 1]                       # :65-68
    opt_g = torch.optim.AdamW(net.parameters(), lr=1e-3)
    opt_h = torch.optim.AdamW(head.parameters(), lr=1e-3)
    opt_g.zero_grad(); opt_h.zero_grad()
    pix = net(gt.cuda(), hook=False)                                        # :140
    hook_outputs.clear()                                                    # :141
    # pixel loss: MSE here instead of the reference's L1 - sign(out - gt) is discontinuous, and on a random-init net with a
    # random target a handful of sign flips (|out - gt| < 1e-3) moves the pure-noise ending.weight gradient by 10-30 %
    l_pix = ((pix - gt.cuda()) ** 2).mean()
    assert net(lq.cuda(), hook=True) is None                                # :154
    kept = list(hook_outputs)
    for h in kept:
        h.retain_grad()
    cls = head(lq.cuda(), hook_outputs[::-1])                               # :155
    l_cls = F.cross_entropy(cls, labels.cuda())
    (l_pix + l_cls).backward()                                              # :163
    # Oracle.  The classifier's gradient w.r.t. the decoder features is taken from the CUDA run (it is only comparable on
    # a fixed ReLU branch, see test_dchead_golden); NAFNet itself is smooth, so given those feature gradients its
    # parameter gradients must match the oracle's: loss_oracle = L1(pix) + sum_i <feat_i, dfeat_i(cuda)>.
    dfe = [h.grad.detach().cpu() for h in kept]
    lg = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
    o_pix = O.nafnet_fwd(gt, lg, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"])
    fe = []
    O.nafnet_fwd(lq, lg, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"], hook=True, decoder_feats=fe)
    with torch.no_grad():
        o_cls = D.dchead_fwd(fe[::-1], sd_h)
    (((o_pix - gt) ** 2).mean() + sum((f * d).sum() for f, d in zip(fe, dfe))).backward()
    assert rel(pix, o_pix) < 2e-3 and rel(cls, o_cls) < 1e-2
    eg = {k: rel(p.grad, lg[k].grad) for k, p in net.named_parameters()}
    print("nafnet grads max %.3g median %.3g" % (max(eg.values()), float(np.median(list(eg.values())))))
    assert max(eg.values()) < 6e-2 and float(np.median(list(eg.values()))) < 2e-2, sorted(eg.items(), key=lambda kv: -kv[1])[:5]
    for k, p in head.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0, k
    opt_g.step(); opt_h.step()                                              # :164-165
    for h in hooks:
        h.remove()


def test_dcpt_step_full_width_one_image():
    """BASELINE.json configs[3] (C4) at full WIDTH on one image: DCPTModel.optimize_parameters (models/
    degradation_classification_pretrain_model.py:133-169) with NAFNet-w64 (36 blocks) + PromptIR_NoImg_DC over the decoder
    features f = [64, 128, 256, 512] at 256x256, the reference's L1 pixel loss + cross entropy, ONE backward - against the fp32
    CPU oracle.  As in test_dcpt_pretrain_step the classifier's feature gradients are taken from the CUDA run (the head is
    only comparable on a fixed ReLU branch), so the check is: logits and pixel output vs oracle, the L1 loss value, and all 664
    backbone gradients given (L1 gradient at the reference's output + those feature gradients)."""
    from basicsr.archs import build_network
    from oracle import nafnet_oracle as O
    cfg = dict(width=64, enc_blk_nums=[1, 1, 1, 28], middle_blk_num=1, dec_blk_nums=[1, 1, 1, 1])
    dims = [64, 128, 256, 512]
    sd_g = O.random_nafnet_state_dict(seed=0, **cfg)
    sd_h = D.random_dchead_state_dict(dims, 2, 5, seed=1)
    net = build_network(dict(type="NAFNetBaseline", **cfg)).cuda()
    head = build_network(dict(type="PromptIR_NoImg_DC", feature_dims=dims, num_res_blocks=2, num_classes=5)).cuda()
    net.load_state_dict(sd_g, strict=True)
    head.load_state_dict(sd_h, strict=True)
    g = torch.Generator().manual_seed(7)
    gt = torch.rand(1, 3, 256, 256, generator=g)
    lq = (gt + 0.1 * torch.randn(gt.shape, generator=g)).clamp(0, 1)
    labels = torch.tensor([3])
    hook_outputs = []
    hooks = [m.register_forward_hook(lambda mod, i, o: hook_outputs.append(o)) for n, m in net.named_modules()
             if "decoder" in n and n.count(".") == 1]                       # :64-67
    assert len(hooks) == 4
    # oracle first: its pixel output defines the L1 gradient both sides use
    lg = {k: v.clone().requires_grad_(True) for k, v in sd_g.items()}
    o_pix = O.nafnet_fwd(gt, lg, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"])
    fe = []
    O.nafnet_fwd(lq, lg, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"], hook=True, decoder_feats=fe)
    with torch.no_grad():
        o_cls = D.dchead_fwd(fe[::-1], sd_h)
    d_pix = torch.sign(o_pix.detach() - gt) / o_pix.numel()
    # CUDA step
    pix = net(gt.cuda(), hook=False)                                        # :140
    hook_outputs.clear()                                                    # :141
    l_pix = F.l1_loss(pix, gt.cuda())                                       # :145-148 (L1Loss, mean)
    assert net(lq.cuda(), hook=True) is None                                # :154
    kept = list(hook_outputs)
    for h in kept:
        h.retain_grad()
    cls = head(lq.cuda(), hook_outputs[::-1])                               # :155
    l_cls = F.cross_entropy(cls, labels.cuda())                             # :158-161
    # one backward (:163); the pixel branch enters with the L1 gradient evaluated at the reference's output
    torch.autograd.backward([pix, l_cls], [d_pix.cuda(), torch.ones((), device="cuda")])
    dfe = [h.grad.detach().cpu() for h in kept]
    (o_pix * d_pix).sum().backward(retain_graph=True)
    sum((f * d).sum() for f, d in zip(fe, dfe)).backward()
    e_pix, e_cls = rel(pix, o_pix), rel(cls, o_cls)
    l_ref = float((o_pix.detach() - gt).abs().mean())
    eg = {k: rel(p.grad, lg[k].grad) for k, p in net.named_parameters()}
    vals = np.array(list(eg.values()))
    report("DCPT step, NAFNet-w64 + head f=[64..512], 1x256x256 (C4 width)", pix=e_pix, logits=e_cls, l1_loss_rel=abs(float(l_pix) - l_ref) / l_ref,
           grads_median=float(np.median(vals)), grads_p95=float(np.percentile(vals, 95)), grads_worst=float(vals.max()))
    assert e_pix < tol(5e-3) and e_cls < tol(2e-2, 3e-3) and abs(float(l_pix) - l_ref) < tol(2e-3) * l_ref
    assert float(np.median(vals)) < tol(2e-2, 3e-3) and float(np.percentile(vals, 95)) < tol(6e-2, 1e-2), sorted(eg.items(), key=lambda kv: -kv[1])[:5]
    for k, p in head.named_parameters():
        assert p.grad is not None and torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0, k
    for h in hooks:
        h.remove()


def test_promptir_dc_img_golden(golden_dir):
    """`PromptIR_DC` (degrad_classify_arch.py:480-556, row a11): registry-built, strict load of the reference's keys,
    conv_embed = 7x7 stride-2 conv (im2col + tcgen05 GEMM) + LayerNorm on lq, then the shared trunk - logits and the gradients of
    the late layers / fc / conv_embed against the REAL reference's fp32 golden (tests/golden/dchead_img.npz)."""
    from basicsr.archs import build_network
    z = np.load(os.path.join(golden_dir, "dchead_img.npz"))
    dims = z["dims"].tolist()
    sd = {k[2:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("p.")}
    head = build_network(dict(type="PromptIR_DC", feature_dims=dims, num_res_blocks=2, num_classes=5)).cuda()
    head.load_state_dict(sd, strict=True)
    lq = torch.from_numpy(z["lq"]).cuda()
    feats = [torch.from_numpy(z[f"feat{i}"]).cuda().requires_grad_(True) for i in range(len(dims))]
    logits = head(lq, feats)
    F.cross_entropy(logits, torch.from_numpy(z["labels"]).cuda()).backward()
    e_log = rel(logits, z["logits"])
    eg = {k: rel(p.grad, z["g." + k]) for k, p in head.named_parameters()}
    groups = {}
    for k, v in eg.items():
        gname = ".".join(k.split(".")[:2]) if k[0] in "bl" else k.split(".")[0]
        groups.setdefault(gname, []).append(v)
    for gname, vs in groups.items():
        report(f"PromptIR_DC grads vs fp32 golden [{gname}]", worst=max(vs), median=float(np.median(vs)))
    report("PromptIR_DC vs fp32 golden", logits=e_log, dfeats_worst=max(rel(f.grad, z[f"dfeat{i}"]) for i, f in enumerate(feats)))
    assert e_log < tol(1e-2, 1e-3)
    assert max(groups["fc"]) < tol(1e-2, 1e-3) and max(groups["last_stage.0"] + groups["last_stage.1"]) < tol(8e-2, 3e-2)
    assert all(torch.isfinite(p.grad).all() and float(p.grad.abs().max()) > 0 for p in head.parameters())
    # the embedding path in isolation: head(lq, zero features) depends on lq only through conv_embed
    with torch.no_grad():
        a = head(lq, [torch.zeros_like(f) for f in feats])
        b = head(lq.flip(3), [torch.zeros_like(f) for f in feats])
        from oracle import dchead_oracle as D
        ref = D.dchead_fwd([torch.zeros(f.shape) for f in feats], sd, lq=torch.from_numpy(z["lq"]))
    assert rel(a, ref) < tol(1e-2, 1e-3) and rel(a, b) > 1e-3


def test_dchead_graph_replay_matches_eager_and_survives_abandoned_forwards():
    """The head's CUDA-graph replay (dcpt_b200/dchead.py graph_forward / graph_backward) against eager launches of the same engine
    code on the same weights: logits, feature gradients and parameter gradients; forwards whose backward never runs (their
    autograd nodes are dropped) must release the slot; after an in-place weight update the replays use the re-packed weights."""
    from basicsr.archs import build_network
    dims = [16, 32]
    sd = D.random_dchead_state_dict(dims, 2, 5, seed=11)
    heads = []
    for graphs in (True, False):
        h = build_network(dict(type="PromptIR_NoImg_DC", feature_dims=dims, num_res_blocks=2, num_classes=5)).cuda()
        h.load_state_dict(sd, strict=True)
        h.engine().use_graphs = graphs
        heads.append(h)
    g = torch.Generator().manual_seed(6)
    lq = torch.rand(2, 3, 32, 32, generator=g).cuda()
    labels = torch.tensor([1, 3]).cuda()

    def feats():
        gg = torch.Generator().manual_seed(7)
        return [torch.randn(2, c, 32 >> i, 32 >> i, generator=gg).cuda().requires_grad_(True) for i, c in enumerate(dims)]

    for _ in range(4):                                  # first sighting eager, then capture, then replays - all abandoned
        out = heads[0](lq, feats())
        del out
    slots = list(heads[0].engine()._gslots.d.values())
    assert len(slots) == 1 and not slots[0].busy and slots[0].fgraph is not None
    for rnd in range(2):
        res = []
        for h in heads:
            h.zero_grad(set_to_none=True)
            fs = feats()
            logits = h(lq, fs)
            F.cross_entropy(logits, labels).backward()
            res.append((logits.detach(), [f.grad for f in fs], [p.grad.clone() for p in h.parameters()]))
        assert slots[0].bgraph is not None and not slots[0].busy
        assert rel(res[0][0], res[1][0]) < 1e-5
        for a, b in zip(res[0][1], res[1][1]):
            assert rel(a, b) < 1e-4
        worst = max(rel(a, b) for a, b in zip(res[0][2], res[1][2]) if float(b.abs().max()) > 0)
        assert worst < 2e-3, worst                      # split-K atomics order
        with torch.no_grad():                           # an "optimizer step": in-place update, version counters bumped
            for h in heads:
                for p in h.parameters():
                    p.mul_(1.01)
