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
