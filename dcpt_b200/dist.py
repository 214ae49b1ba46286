"""Data-parallel host logic (one process per GPU; reference: basicsr/utils/dist_util.py:11-82 and the DDP wrap in
basicsr/models/base_model.py:100-118).  The hot path shards by minibatch only: each rank runs the
same kernels on its own images and the single exchange step is the gradient all-reduce (mean).  These helpers work
on any torch.distributed backend (NCCL on the B200 box; gloo in the CPU tests)."""
import os

import torch
import torch.distributed as dist


def get_dist_info():
    """(rank, world_size) — dist_util.py:61-72."""
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def init_dist(backend="nccl", timeout_s=600):
    """`--launcher pytorch` initialisation (dist_util.py:22-26): RANK / WORLD_SIZE / MASTER_* from the environment,
    device = rank % device_count."""
    import datetime
    rank = int(os.environ["RANK"])
    kw = {}
    if backend == "nccl":
        n = torch.cuda.device_count()
        torch.cuda.set_device(rank % n)
        kw["device_id"] = torch.device("cuda", rank % n)
    dist.init_process_group(backend=backend, timeout=datetime.timedelta(seconds=timeout_s), **kw)
    return get_dist_info()


def rank_seed(manual_seed, rank=None):
    """Per-rank RNG seed = manual_seed + rank (basicsr/utils/options.py:142)."""
    if rank is None:
        rank = get_dist_info()[0]
    return manual_seed + rank


def shard_batch(n_items, rank=None, world=None):
    """Rank-strided sample indices (EnlargedSampler, basicsr/data/data_sampler.py:30-43)."""
    r, w = get_dist_info()
    rank = r if rank is None else rank
    world = w if world is None else world
    return list(range(rank, n_items, world))


def allreduce_mean_(flat):
    """In-place mean all-reduce of the flat fp32 gradient buffer (DDP's bucketed all-reduce, base_model.py:111-115)."""
    rank, world = get_dist_info()
    if world == 1:
        return flat
    if flat.is_cuda:
        dist.all_reduce(flat, op=dist.ReduceOp.AVG)
    else:                                   # gloo has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        flat.div_(world)
    return flat


def broadcast_params_(params, src=0):
    """Make every rank start from rank `src`'s weights (what DDP does at construction)."""
    _, world = get_dist_info()
    if world > 1:
        for p in params:
            dist.broadcast(p.data if hasattr(p, "data") else p, src=src)
    return params


def reduce_loss_dict(loss_dict):
    """Average the logged losses onto rank 0 (basicsr/models/base_model.py:432-457) with ONE host sync."""
    rank, world = get_dist_info()
    keys = list(loss_dict.keys())
    vals = torch.stack([loss_dict[k].detach().float().reshape(()) for k in keys])
    if world > 1:
        dist.reduce(vals, dst=0)
        if rank == 0:
            vals /= world
    vals = vals.tolist()
    return {k: v for k, v in zip(keys, vals)}
