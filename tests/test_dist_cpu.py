"""world_size-2 gloo test of the data-parallel host logic (dcpt_b200/dist.py) on CPU.

Property checked: the path shards by minibatch with one exchange step, so the mean-all-reduced
per-rank gradients equal the gradients of the full batch.  The math runs through the CPU oracle here (the CUDA
engine is exercised on the GPU box); what this covers is the sharding / all-reduce / loss-reduce plumbing.
"""
import os
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from dcpt_b200 import dist as D
    from oracle import nafnet_oracle as O
    assert D.init_dist("gloo", timeout_s=120) == (rank, world)
    cfg = dict(width=8, enc_blk_nums=[1, 1], middle_blk_num=1, dec_blk_nums=[1, 1])
    sd = O.random_nafnet_state_dict(seed=3 + rank, **cfg)               # ranks start different ...
    D.broadcast_params_(list(sd.values()), src=0)                        # ... and are synchronised
    g = torch.Generator().manual_seed(0)
    inp, gt = torch.rand(4, 3, 16, 16, generator=g), torch.rand(4, 3, 16, 16, generator=g)
    idx = D.shard_batch(4)
    assert idx == list(range(rank, 4, world))
    _, loss, grads = O.nafnet_fwd_bwd(inp[idx], gt[idx], sd, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"])
    flat = torch.cat([grads[k].reshape(-1) for k in sd])
    D.allreduce_mean_(flat)
    logged = D.reduce_loss_dict({"l_pix": loss})
    if rank == 0:
        _, loss_full, grads_full = O.nafnet_fwd_bwd(inp, gt, sd, cfg["enc_blk_nums"], cfg["middle_blk_num"], cfg["dec_blk_nums"])
        ref = torch.cat([grads_full[k].reshape(-1) for k in sd])
        err = float((flat - ref).norm() / ref.norm())
        q.put((err, abs(logged["l_pix"] - float(loss_full)), D.rank_seed(10), D.get_dist_info()))
    else:
        assert D.rank_seed(10) == 11
    torch.distributed.destroy_process_group()


def _worker_wrapper(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from basicsr.archs import build_network
    from dcpt_b200 import dist as D
    D.init_dist("gloo", timeout_s=120)
    torch.manual_seed(100 + rank)                                         # ranks construct different weights ...
    net = build_network(dict(type="NAFNetBaseline", width=8, enc_blk_nums=[1], middle_blk_num=1, dec_blk_nums=[1]))
    dp = D.FlatGradDataParallel(net)                                      # ... rank 0's are broadcast (as DDP does)
    assert dp.module is net and net.engine().grad_sync is not None
    chk = torch.cat([p.detach().reshape(-1) for p in net.parameters()])
    gathered = [torch.empty_like(chk) for _ in range(world)]
    torch.distributed.all_gather(gathered, chk)
    same = all(torch.equal(gathered[0], g) for g in gathered)
    flat = torch.full((1000,), float(rank + 1))
    net.engine().grad_sync(flat)                                          # what the engine calls after its backward kernels
    mean_ok = bool(torch.allclose(flat, torch.full((1000,), (1 + world) / 2)))
    flat2 = torch.full((10,), float(rank))
    with dp.no_sync():
        net.engine().grad_sync(flat2)
    skip_ok = bool(torch.equal(flat2, torch.full((10,), float(rank))))
    for p in net.parameters():
        p.grad = torch.ones_like(p)
    dp.zero_grad()
    zg_ok = all(p.grad is None for p in net.parameters())
    if rank == 0:
        q.put((same, mean_ok, skip_ok, zg_ok))
    torch.distributed.destroy_process_group()


def test_flat_grad_data_parallel_gloo_world2():
    """dcpt_b200.dist.FlatGradDataParallel: broadcast at construction, the engine's grad_sync hook = mean all-reduce,
    no_sync(), zero_grad; the CUDA engine itself is exercised by bench.py --gpus 2 on the GPU box."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29300 + os.getpid() % 300
    procs = [ctx.Process(target=_worker_wrapper, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    assert q.get(timeout=5) == (True, True, True, True)


def test_data_parallel_gloo_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29600 + os.getpid() % 300
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=240)
        assert p.exitcode == 0
    err, loss_err, seed0, info = q.get(timeout=5)
    assert err < 1e-5, err            # fp32: mean of half-batch grads == full-batch grads
    assert loss_err < 1e-6
    assert seed0 == 10 and info == (0, 2)


def _dcpt_opt(dist):
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from boundary_overlay_check import CFG, DIMS
    return {"name": "dp", "model_type": "DCPTModel", "scale": 1, "num_gpu": 0, "dist": dist, "is_train": True, "rank": 0, "world_size": 1,
            "network_g": dict(type="NAFNetBaseline", window_size=16, **CFG),
            "network_dc": dict(type="PromptIR_NoImg_DC", feature_dims=DIMS, num_res_blocks=2, num_classes=5), "hook_names": "decoder",
            "path": {"pretrain_network_g": None}, "train": {
                "optim_g": {"type": "AdamW", "lr": 1e-3, "weight_decay": 1e-4, "betas": [0.9, 0.9]},
                "optim_dc": {"type": "AdamW", "lr": 1e-3, "weight_decay": 1e-4, "betas": [0.9, 0.9]},
                "scheduler": {"type": "MultiStepLR", "milestones": [100], "gamma": 0.5},
                "pixel_opt": {"type": "L1Loss", "loss_weight": 1.0, "reduction": "mean"},
                "classify_opt": {"type": "CrossEntropyLoss", "loss_weight": 1.0}}}


def _dcpt_step(model, sd_g, sd_h, idx):
    model.get_bare_model(model.net_g).load_state_dict(sd_g, strict=True)
    model.get_bare_model(model.net_dc).load_state_dict(sd_h, strict=True)
    g = torch.Generator().manual_seed(21)
    gt, lq = torch.rand(4, 3, 32, 32, generator=g), torch.rand(4, 3, 32, 32, generator=g)
    labels = torch.tensor([4, 1, 0, 2])
    model.feed_data({"lq": lq[idx], "gt": gt[idx], "dataset_idx": labels[idx]})
    model.optimize_parameters(1)
    flat = lambda net: torch.cat([p.grad.reshape(-1) for p in model.get_bare_model(net).parameters()])  # noqa: E731
    return flat(model.net_g), flat(model.net_dc), model.get_current_log()


def _worker_dcpt_model(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    torch.set_num_threads(2)
    from basicsr.models import build_model
    from dcpt_b200 import dist as D
    from oracle import dchead_oracle as DH
    from oracle import nafnet_oracle as O
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from boundary_overlay_check import CFG, DIMS, stub_engines
    stub_engines()                                                        # no GPU here: the engines' math is the oracle's
    D.init_dist("gloo", timeout_s=120)
    sd_g = O.random_nafnet_state_dict(seed=3, **CFG)
    sd_h = DH.random_dchead_state_dict(DIMS, 2, 5, seed=4)
    calls = []
    real = D.allreduce_mean_
    D.allreduce_mean_ = lambda flat, group=None: (calls.append(flat.numel()), real(flat, group))[1]
    model = build_model(_dcpt_opt(True))
    wrapped = type(model.net_g).__name__ == "FlatGradDataParallel" and type(model.net_dc).__name__ == "FlatGradDataParallel"
    hooked = sorted(n for n, m in model.net_g.named_modules() if len(m._forward_hooks) > 0)
    gg, gh, log = _dcpt_step(model, sd_g, sd_h, D.shard_batch(4))
    # a flat-buffer network: views of one base are exchanged as ONE all-reduce of the base
    lin = torch.nn.Linear(600, 500)                                       # 1.2 MB: above the run threshold
    base = torch.full((300500 + 16,), float(rank + 1))                    # 16 elements of alignment padding between the two views
    lin.weight.grad = base[:300000].view(500, 600).detach()               # aliases without ._base, as autograd keeps them
    lin.bias.grad = base[300016:].detach()
    n_before = len(calls)
    D.exchange_accumulated_grads_([lin])
    base_ok = len(calls) == n_before + 1 and calls[-1] == 300516 and bool(torch.allclose(base, torch.full_like(base, (1 + world) / 2)))
    # the mixed case of the DCPT step: the node that ran first (hooked pass) did not reach `ending`, so two small gradients alias
    # the OTHER node's flat buffer - one whole-run all-reduce for the big buffer, one coalesced call for the leftovers
    mix = torch.nn.ParameterList([torch.nn.Parameter(torch.zeros(500, 600)), torch.nn.Parameter(torch.zeros(500)),
                                  torch.nn.Parameter(torch.zeros(7, 3)), torch.nn.Parameter(torch.zeros(7))])
    flat_a = torch.full((300500 + 16,), float(rank + 1))
    flat_b = torch.full((300500 + 64 + 21 + 64 + 7,), float(10 * (rank + 1)))
    mix[0].grad, mix[1].grad = flat_a[:300000].view(500, 600).detach(), flat_a[300016:].detach()
    mix[2].grad, mix[3].grad = flat_b[300564:300585].view(7, 3).detach(), flat_b[300649:300656].detach()
    n_before = len(calls)
    D.exchange_accumulated_grads_([mix])
    mixed_ok = (len(calls) == n_before + 2 and sorted(calls[-2:]) == [28, 300516]
                and bool(torch.allclose(mix[0].grad, torch.full((500, 600), (1 + world) / 2)))
                and bool(torch.allclose(mix[2].grad, torch.full((7, 3), 10 * (1 + world) / 2)))
                and bool(torch.allclose(mix[3].grad, torch.full((7,), 10 * (1 + world) / 2)))
                and float(flat_b[0]) == 10 * (rank + 1))           # the rest of the second buffer is left alone
    base_ok = base_ok and mixed_ok
    if rank == 0:
        D.allreduce_mean_ = real
        torch.distributed.destroy_process_group()
        single = build_model(_dcpt_opt(False))
        fg, fh, flog = _dcpt_step(single, sd_g, sd_h, list(range(4)))
        q.put((wrapped, hooked, len(calls) - 3, float((gg - fg).norm() / fg.norm()), float((gh - fh).norm() / fh.norm()),
               abs(log["l_pix"] - flog["l_pix"]), abs(log["l_classify"] - flog["l_classify"]), base_ok))   # n_calls: the model's two
    else:
        torch.distributed.destroy_process_group()


def test_dcpt_model_data_parallel_gloo_world2():
    """The DCPTModel mirror under ``dist: true`` (world 2, gloo): both networks wrapped, hooks found through the wrapper's
    ``module.`` prefix exactly as with the reference's DDP (one-dot rule -> the decoder containers), the two net_g backward
    nodes + the classifier exchanged ONCE per network after the backward, and the averaged gradients / logged losses equal to
    the single-process step on the full batch."""
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29900 + os.getpid() % 90
    procs = [ctx.Process(target=_worker_dcpt_model, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=400)
        assert p.exitcode == 0
    wrapped, hooked, n_calls, eg, eh, el, ec, base_ok = q.get(timeout=5)
    assert wrapped and base_ok
    assert hooked == ["module.decoder0", "module.decoder1", "module.decoder2", "module.decoder3"]
    assert n_calls == 2, n_calls                       # one exchange per network for the whole two-pass step
    assert eg < 1e-4 and eh < 1e-4, (eg, eh)
    assert el < 1e-6 and ec < 1e-6
