"""GPU parity of the fused parameter update (dcpt_optim_* through dcpt_b200.optim.FusedAdam) — SURVEY.md §8(f) row 1.

Compared with: the golden vectors made by torch.optim.Adam / AdamW + clip_grad_norm_ + the reference's model_ema
(tests/golden/optim_step.npz), the numpy oracle on ragged / unaligned tensors, and at the full NAFNet-w64 size through the
oracle and through size-independent properties.  Tolerance 2e-6 of the tensor's max magnitude: fp32 arithmetic whose only
freedom against torch is FMA contraction and the summation order of the gradient norm."""
import copy
import os
import sys

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden_optim as MG  # noqa: E402
from oracle import optim_oracle as OO  # noqa: E402

TOL = 2e-6


def err(a, b):
    a = torch.as_tensor(a).detach().double().cpu()
    b = torch.as_tensor(np.asarray(b)).double()
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-30))


def test_golden_cases(golden_dir):
    from dcpt_b200.optim import get_optimizer
    z = np.load(os.path.join(golden_dir, "optim_step.npz"))
    for name, (cls, kw, clip, decay, steps) in MG.CASES.items():
        params, grads = MG.tensors(sum(map(ord, name)))
        net = [torch.nn.Parameter(p.clone().cuda()) for p in params]
        ema = [p.clone().cuda() for p in params]
        kw = dict(kw)
        opt = get_optimizer(cls, net, kw.pop("lr"), **kw)                    # base_model.py:120-139
        norms = []
        for t in range(steps):
            for p, g in zip(net, grads[t]):
                p.grad = g.clone().cuda()
            n = opt.step(grad_clip=clip, ema_params=ema if decay > 0 else None, ema_decay=decay)   # sr_model.py:166-174
            if clip:
                norms.append(float(n))
        worst = 0.0
        for i, p in enumerate(net):
            st = opt.state[p]
            assert float(st["step"]) == steps
            worst = max(worst, err(p, z[f"{name}|p{i}"]), err(st["exp_avg"], z[f"{name}|m{i}"]),
                        err(st["exp_avg_sq"], z[f"{name}|v{i}"]), err(ema[i], z[f"{name}|e{i}"]))
        print(f"{name}: worst max-normalised error {worst:.2e}; total_norm {norms} vs {z[name + '|norms'].tolist()}")
        assert worst < TOL
        np.testing.assert_allclose(norms, z[f"{name}|norms"], rtol=2e-6)


def test_ragged_unaligned_and_missing_grads_vs_oracle():
    """Sizes around the 8192-element chunk and the 4-element vector, parameters that are views at odd offsets (4-byte
    aligned only), a zero-size tensor, a parameter without gradient (skipped, like torch), two steps."""
    from dcpt_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(3)
    sizes = [1, 2, 3, 4, 5, 8191, 8192, 8193, 16384 + 7, 0, 100003]
    big = torch.randn(sum(sizes) + 3 * len(sizes) + 8, generator=g).cuda()
    net, off = [], 1                                            # off = 1: every view starts 4 bytes off a 16-byte boundary
    for n in sizes:
        net.append(torch.nn.Parameter(big[off:off + n]))
        off += n + 3
    frozen = torch.nn.Parameter(torch.randn(10, generator=g).cuda())          # never gets a gradient
    ema = [p.detach().clone() for p in net] + [frozen.detach().clone()]
    P = [p.detach().cpu().numpy().copy() for p in net]
    E = [p.copy() for p in P]
    M, V = [np.zeros_like(p) for p in P], [np.zeros_like(p) for p in P]
    kw = dict(lr=1e-2, betas=(0.8, 0.95), eps=1e-6, weight_decay=0.1)
    opt = FusedAdam(net + [frozen], decoupled_weight_decay=False, **kw)
    frozen0 = frozen.detach().clone()
    for t in range(2):
        G = [torch.randn(n, generator=g) * 0.3 for n in sizes]
        for p, gr in zip(net, G):
            p.grad = gr.cuda()
        n_gpu = opt.step(grad_clip=0.5, ema_params=ema, ema_decay=0.9)
        n_cpu = OO.train_update(P, [x.numpy() for x in G], M, V, E, t + 1, grad_clip=0.5, ema_decay=0.9, decoupled=False, **kw)
        assert abs(float(n_gpu) - float(n_cpu)) < 2e-6 * float(n_cpu)
    for i, p in enumerate(net):
        if sizes[i] == 0:
            continue
        assert err(p, P[i]) < TOL and err(opt.state[p]["exp_avg"], M[i]) < TOL and err(opt.state[p]["exp_avg_sq"], V[i]) < TOL, i
        assert err(ema[i], E[i]) < TOL, i
    assert torch.equal(frozen, frozen0) and torch.equal(ema[-1], frozen0) and len(opt.state[frozen]) == 0
    # the 3-float gaps between the views were not touched
    ref = torch.randn(big.numel(), generator=torch.Generator().manual_seed(3))
    off = 1
    for n in sizes:
        assert torch.equal(big[off + n:off + n + 3].cpu(), ref[off + n:off + n + 3])
        off += n + 3


def test_state_dict_interchange_with_torch():
    """A torch.optim.AdamW state (the reference's .state resume file, base_model.py:413-430) loads into FusedAdamW and the
    next step equals torch's own next step; and the other way round."""
    from dcpt_b200.optim import FusedAdamW
    g = torch.Generator().manual_seed(9)
    shapes = [(33, 7), (129,), (5000,)]
    a = [torch.nn.Parameter(torch.randn(s, generator=g).cuda()) for s in shapes]
    b = [torch.nn.Parameter(p.detach().clone()) for p in a]
    kw = dict(lr=1e-3, betas=(0.9, 0.99), weight_decay=0.05)
    ta, fb = torch.optim.AdamW(a, **kw), FusedAdamW(b, **kw)
    grads = [[torch.randn(s, generator=g).cuda() for s in shapes] for _ in range(4)]

    def step(opt, ps, t):
        for p, gr in zip(ps, grads[t]):
            p.grad = gr.clone()
        opt.step()
    step(ta, a, 0); step(ta, a, 1)
    fb.load_state_dict(copy.deepcopy(ta.state_dict()))           # torch -> fused (state_dict() hands out references)
    for p, q in zip(a, b):
        q.data.copy_(p.data)
    step(ta, a, 2); step(fb, b, 2)
    assert max(err(q, p.detach().cpu().numpy()) for p, q in zip(a, b)) < TOL
    ta2 = torch.optim.AdamW(a, **kw)
    ta2.load_state_dict(copy.deepcopy(fb.state_dict()))          # fused -> torch
    step(ta2, a, 3); step(fb, b, 3)
    assert max(err(q, p.detach().cpu().numpy()) for p, q in zip(a, b)) < TOL
    assert float(fb.state[b[0]]["step"]) == 4


def test_full_size_nafnet_w64_update():
    """The 664 parameter tensors of NAFNet-w64 (67.9 M parameters, BASELINE configs[1]'s network): one AdamW + clip + EMA
    update against the numpy oracle, plus size-independent properties: a zero gradient with wd = 0 leaves p unchanged and
    moves ema toward p exactly; lr = 0 leaves p unchanged while m, v still move; clipping to max_norm scales the update's
    first moment by max_norm / total_norm."""
    from basicsr.archs import build_network
    from dcpt_b200.optim import FusedAdamW
    torch.manual_seed(0)
    net = build_network(dict(type="NAFNetBaseline", width=64, enc_blk_nums=[1, 1, 1, 28], middle_blk_num=1, dec_blk_nums=[1, 1, 1, 1])).cuda()
    params = list(net.parameters())
    assert len(params) == 664 and sum(p.numel() for p in params) == 67888835
    ema = [p.detach().clone() + 0.01 for p in params]
    g = torch.Generator(device="cuda").manual_seed(1)
    for p in params:
        p.grad = torch.randn(p.shape, generator=g, device="cuda") * 1e-3
    P = [p.detach().cpu().numpy().copy() for p in params]
    E = [e.cpu().numpy().copy() for e in ema]
    M, V = [np.zeros_like(p) for p in P], [np.zeros_like(p) for p in P]
    kw = dict(lr=1e-3, betas=(0.9, 0.9), eps=1e-8, weight_decay=1e-3)
    opt = FusedAdamW(params, **kw)
    n_gpu = opt.step(grad_clip=0.01, ema_params=ema, ema_decay=0.999)
    n_cpu = OO.train_update(P, [p.grad.cpu().numpy() for p in params], M, V, E, 1, grad_clip=0.01, ema_decay=0.999, decoupled=True, **kw)
    assert abs(float(n_gpu) - float(n_cpu)) < 2e-6 * float(n_cpu)
    worst = max(max(err(p, P[i]), err(opt.state[p]["exp_avg"], M[i]), err(opt.state[p]["exp_avg_sq"], V[i]), err(ema[i], E[i]))
                for i, p in enumerate(params))
    print(f"NAFNet-w64 full-size update vs oracle: worst max-normalised error {worst:.2e}, total_norm {float(n_gpu):.6f}")
    assert worst < TOL
    # clipped first moment: m = (1 - b1) * g * max_norm / (norm + 1e-6)
    i = 100
    want = 0.1 * params[i].grad * (0.01 / (float(n_gpu) + 1e-6))
    assert err(opt.state[params[i]]["exp_avg"], want.cpu().numpy()) < 1e-5
    # properties on fresh optimizers
    p0 = [p.detach().clone() for p in params]
    for p in params:
        p.grad.zero_()
    o2 = FusedAdamW(params, lr=1e-3, weight_decay=0.0)
    e2 = [torch.zeros_like(p) for p in params]
    o2.step(ema_params=e2, ema_decay=0.5)
    assert all(torch.equal(p, q) for p, q in zip(params, p0))                 # g = 0, wd = 0: p += -lr * 0 / eps
    assert all(torch.equal(e, 0.5 * q) for e, q in zip(e2, p0))               # ema = 0 * 0.5 + 0.5 * p, exact in fp32
    for p in params:
        p.grad.fill_(1e-3)
    o3 = FusedAdamW(params, lr=0.0, weight_decay=0.3)
    o3.step()
    assert all(torch.equal(p, q) for p, q in zip(params, p0))                 # lr = 0: no decay, no step
    assert all(float(o3.state[p]["exp_avg"].flatten()[0]) > 0 for p in params[:5])


def _small_net(seed=0):
    from basicsr.archs import build_network
    from oracle import nafnet_oracle as O
    cfg = dict(width=16, enc_blk_nums=[1, 1], middle_blk_num=1, dec_blk_nums=[1, 1])
    sd = O.random_nafnet_state_dict(seed=seed, **cfg)
    net = build_network(dict(type="NAFNetBaseline", window_size=16, **cfg)).cuda()
    net.load_state_dict(sd, strict=True)
    return net, cfg


def _oracle_out(net, cfg, inp):
    from oracle import nafnet_oracle as O
    sd = {k: v.detach().cpu() for k, v in net.state_dict().items()}
    return O.nafnet_fwd(inp, sd, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"])


def _rel(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return float((a - b).norm() / b.norm())


def test_forward_after_fused_step_uses_updated_weights():
    """FusedAdam writes the parameters through raw pointers; the engines' packed bf16 operand cache must be
    rebuilt for the next forward (optim.py bumps the version counters).  forward -> FusedAdam.step -> forward, each forward
    against the oracle run on the weights the module holds at that moment."""
    from dcpt_b200.optim import FusedAdamW
    net, cfg = _small_net()
    g = torch.Generator().manual_seed(3)
    inp, gt = torch.rand(2, 3, 32, 32, generator=g), torch.rand(2, 3, 32, 32, generator=g)
    opt = FusedAdamW(net.parameters(), lr=5e-2, betas=(0.9, 0.9), weight_decay=0.0)   # a large step: stale weights would show
    out0 = net(inp.cuda())
    ref0 = _oracle_out(net, cfg, inp)
    assert _rel(out0, ref0) < 2e-3
    ((out0 - gt.cuda()) ** 2).mean().backward()
    opt.step()
    out1 = net(inp.cuda())
    ref1 = _oracle_out(net, cfg, inp)
    moved = _rel(ref1, ref0)
    e1 = _rel(out1, ref1)
    print(f"update moved the output by {moved:.2e}; forward after the step vs oracle with the updated weights: {e1:.2e}")
    assert moved > 2e-2, "the test's update is too small to detect stale packed weights"
    assert e1 < 2e-3
    # and the backward of that second forward uses the fresh weights as well (dgrad operands are packed too)
    net.zero_grad()
    ((out1 - gt.cuda()) ** 2).mean().backward()
    from oracle import nafnet_oracle as O
    leaves = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in net.state_dict().items()}
    ro = O.nafnet_fwd(inp, leaves, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"])
    ((ro - gt) ** 2).mean().backward()
    errs = sorted(_rel(p.grad, leaves[k].grad) for k, p in net.named_parameters())
    assert errs[len(errs) // 2] < 1.5e-2, errs[len(errs) // 2]


def test_no_grad_forward_sees_data_writes():
    """The reference's model_ema (base_model.py:86-95) writes `ema.data.mul_(decay).add_(p.data, alpha=1-decay)`, which bumps no
    version counter; net_g_ema is then run under no_grad (sr_model.py:176-185).  The no-grad weight fingerprint
    (dcpt_b200/params.py) must catch it."""
    ema_net, cfg = _small_net(seed=0)
    src_net, _ = _small_net(seed=1)
    g = torch.Generator().manual_seed(4)
    inp = torch.rand(1, 3, 32, 32, generator=g)
    with torch.no_grad():
        out0 = ema_net(inp.cuda())
    ref0 = _oracle_out(ema_net, cfg, inp)
    assert _rel(out0, ref0) < 2e-3
    versions = [p._version for p in ema_net.parameters()]
    for e, p in zip(ema_net.parameters(), src_net.parameters()):
        e.data.mul_(0.5).add_(p.data, alpha=0.5)                    # model_ema with decay 0.5
    assert versions == [p._version for p in ema_net.parameters()]   # the premise: torch did not notice
    with torch.no_grad():
        out1 = ema_net(inp.cuda())
    ref1 = _oracle_out(ema_net, cfg, inp)
    assert _rel(ref1, ref0) > 2e-2
    assert _rel(out1, ref1) < 2e-3


def test_grad_clip_over_several_param_groups():
    """clip_grad_norm_ over every parameter when the optimizer has several jobs (two param groups here): the per-plan norms
    are combined on the device."""
    from dcpt_b200.optim import FusedAdam
    g = torch.Generator().manual_seed(5)
    shapes = [(300,), (17, 9), (4096,), (33,)]
    P = [torch.randn(s, generator=g) for s in shapes]
    G = [torch.randn(s, generator=g) for s in shapes]
    net = [torch.nn.Parameter(p.clone().cuda()) for p in P]
    for p, gg in zip(net, G):
        p.grad = gg.clone().cuda()
    opt = FusedAdam([{"params": net[:2], "lr": 1e-2}, {"params": net[2:], "lr": 3e-2}], betas=(0.9, 0.99))
    n = float(opt.step(grad_clip=0.5))
    ref = [torch.nn.Parameter(p.clone()) for p in P]
    for p, gg in zip(ref, G):
        p.grad = gg.clone()
    ropt = torch.optim.Adam([{"params": ref[:2], "lr": 1e-2}, {"params": ref[2:], "lr": 3e-2}], betas=(0.9, 0.99))
    rn = float(torch.nn.utils.clip_grad_norm_(ref, 0.5))
    ropt.step()
    assert abs(n - rn) < 2e-6 * rn
    for a, b in zip(net, ref):
        assert err(a, b.detach().numpy()) < TOL
