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
