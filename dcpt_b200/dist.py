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


def enlarged_indices(dataset_size, num_replicas, rank, ratio=1, epoch=0):
    """The index list ``EnlargedSampler.__iter__`` yields (basicsr/data/data_sampler.py:8-48): a seeded permutation of
    ``ceil(size * ratio / replicas) * replicas`` positions folded onto the dataset, every ``replicas``-th entry from ``rank``."""
    import math
    num_samples = math.ceil(dataset_size * ratio / num_replicas)
    total = num_samples * num_replicas
    g = torch.Generator()
    g.manual_seed(epoch)
    idx = [v % dataset_size for v in torch.randperm(total, generator=g).tolist()]
    idx = idx[rank:total:num_replicas]
    assert len(idx) == num_samples
    return idx


def allreduce_mean_(flat, group=None):
    """In-place mean all-reduce of the flat fp32 gradient buffer (DDP's bucketed all-reduce, base_model.py:111-115)."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat
    world = dist.get_world_size(group)
    if world == 1:
        return flat
    if flat.is_cuda:
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:                                   # gloo has no AVG
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.div_(world)
    return flat


def allreduce_grads_overlapped_(engine, flat, group=None, comm_stream=None):
    """Mean all-reduce of a NAFNet engine's flat gradient buffer, overlapped with the tail of the backward that is still
    running: the slice that was final when ``engine.split_event`` fired (deepest encoder level + middle + ups, ~90 % of the
    bytes) is reduced on ``comm_stream`` as soon as the event fires, the rest on the current stream after the backward.
    Falls back to one plain all-reduce when the engine has no split event (or on CPU / a single rank)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat
    ev = getattr(engine, "split_event", None)
    if ev is None or not flat.is_cuda or comm_stream is None:
        return allreduce_mean_(flat, group)
    a, b = engine.early_grad_range()
    if not (0 <= a < b <= flat.numel()):
        return allreduce_mean_(flat, group)
    comm_stream.wait_event(ev)
    with torch.cuda.stream(comm_stream):
        allreduce_mean_(flat[a:b], group)
    if a > 0:
        allreduce_mean_(flat[:a], group)
    if b < flat.numel():
        allreduce_mean_(flat[b:], group)
    torch.cuda.current_stream().wait_stream(comm_stream)
    return flat


def broadcast_params_(params, src=0, group=None):
    """Make every rank start from rank `src`'s weights (what DDP does at construction)."""
    _, world = get_dist_info()
    if world > 1:
        with torch.no_grad():
            for p in params:
                # p.detach() shares p's version counter (p.data does not): the engines' packed-operand caches see the write
                dist.broadcast(p.detach(), src=src, group=group)
    return params


class FlatGradDataParallel(torch.nn.Module):
    """Data-parallel wrapper for the dcpt_b200 networks: what ``DistributedDataParallel(net, device_ids=[dev])`` does for the
    reference in ``BaseModel.model_to_device`` (base_model.py:100-118), specialised to this hot path.

    The networks' backward is ONE autograd node that writes every parameter gradient into one flat fp32 buffer, so the
    gradient exchange is ONE mean all-reduce of that buffer, issued by the engine right after the backward kernels (before
    autograd hands the views to ``.grad``).  torch's DDP also works (tested), but it cannot overlap anything with a single
    backward node either and pays for it: 664 per-parameter hooks, 664 copies into 25 MB buckets and 11 bucket all-reduces
    (measured at N = 2: 27.5 ms per step with DDP).  Same contract as DDP: parameters and buffers are broadcast from rank 0 at
    construction, gradients are averaged over the group, ``.module`` is the wrapped network, ``no_sync()`` skips the exchange
    (gradient accumulation).  Parameters a pass does not reach simply get no gradient on every rank alike
    (``find_unused_parameters`` has no meaning here)."""

    def __init__(self, module, process_group=None, broadcast=True, overlap=True):
        super().__init__()
        if not hasattr(module, "engine"):
            raise TypeError("FlatGradDataParallel wraps the dcpt_b200 networks (NAFNetBaseline, Restormer, PromptIR_NoImg_DC)")
        self.module = module
        self.process_group = process_group
        self._sync_enabled = True
        if broadcast:
            broadcast_params_(list(module.parameters()) + list(module.buffers()), group=process_group)
        eng = module.engine()
        eng.grad_sync = self._sync
        # overlap the all-reduce with the backward's tail where the engine supports it (NAFNet; bf16 build: the parity build
        # rescales the whole buffer after the backward) - DCPT_DP_OVERLAP=0 turns it off
        self._comm = None
        if overlap and hasattr(eng, "enable_grad_overlap") and torch.cuda.is_available() and os.getenv("DCPT_DP_OVERLAP", "1") != "0":
            from .lib import operand_dtype
            if operand_dtype() == torch.bfloat16 and get_dist_info()[1] > 1:
                eng.enable_grad_overlap(True)
                # high priority (DCPT_COMM_PRIORITY=0: default): the captured backward runs on a high-priority stream too, and the
                # few thread blocks NCCL needs must not wait behind the weight-gradient side stream's
                self._comm = torch.cuda.Stream(priority=-1 if os.getenv("DCPT_COMM_PRIORITY", "1") != "0" else 0)

    def _sync(self, flat):
        if self._sync_enabled:
            allreduce_grads_overlapped_(self.module.engine(), flat, group=self.process_group, comm_stream=self._comm)

    def no_sync(self):
        import contextlib

        @contextlib.contextmanager
        def ctx():
            old, self._sync_enabled = self._sync_enabled, False
            try:
                yield
            finally:
                self._sync_enabled = old
        return ctx()

    def zero_grad(self, set_to_none=True):
        return self.module.zero_grad(set_to_none=set_to_none)

    def forward(self, *args, **kwargs):
        return self.module(*args, **kwargs)


def _grad_segments(grads, max_gap=512):
    """Group gradient tensors into runs that sit back to back (alignment gaps <= max_gap bytes) inside ONE storage.
    Host cost matters here - the GPU idles while this runs - so it is two cheap calls per tensor and numpy for the rest."""
    import numpy as np
    ptr = np.fromiter((g.data_ptr() for g in grads), dtype=np.int64, count=len(grads))
    size = np.fromiter((g.numel() * g.element_size() for g in grads), dtype=np.int64, count=len(grads))
    order = np.argsort(ptr, kind="stable")
    p, e = ptr[order], ptr[order] + size[order]
    gap = p[1:] - e[:-1]
    cuts = np.flatnonzero((gap < 0) | (gap > max_gap)) + 1
    runs, stray = [], []
    for lo, hi in zip(np.concatenate(([0], cuts)), np.concatenate((cuts, [len(grads)]))):
        idx = order[lo:hi]
        first, last = grads[idx[0]], grads[idx[-1]]
        # every rank must cut the same runs: only the engines' flat buffers qualify (one storage, parameter order, >= 1 MiB);
        # allocations that merely happen to be neighbours on this rank never share a storage
        ok = hi - lo > 1 and int(e[hi - 1] - p[lo]) >= (1 << 20) and bool(np.all(np.diff(idx) == 1)) and first.dtype == last.dtype and \
            first.is_contiguous() and last.is_contiguous() and first.untyped_storage().data_ptr() == last.untyped_storage().data_ptr()
        if ok:
            n = int(e[hi - 1] - p[lo]) // first.element_size()
            runs.append((int(idx[0]), first.new_empty(0).set_(first.untyped_storage(), first.storage_offset(), (n,))))
        else:
            stray.extend(int(i) for i in idx)
    runs = [r for _, r in sorted(runs, key=lambda t: t[0])]        # rank-independent order: by first parameter index
    stray = [grads[i] for i in sorted(stray)]
    return runs, stray


def exchange_accumulated_grads_(nets, group=None):
    """Mean all-reduce of the ``.grad`` tensors of ``nets`` after a backward run under ``no_sync()`` - the exchange of a step
    whose networks are reached by SEVERAL backward nodes (the DCPT step runs net_g twice,
    degradation_classification_pretrain_model.py:141-162), where a per-node exchange would move the same buffer more than once.

    The engines hand autograd views of one flat fp32 buffer per backward node; autograd keeps (aliases of) the first node's
    views as ``.grad`` and accumulates later nodes into them in place.  So after the backward a network's gradients sit back to
    back in one storage - possibly two, when the node that ran first did not reach every parameter (the hooked pass never
    reaches ``ending``).  Each such run is all-reduced in place as one tensor; what is left over (isolated tensors, a network
    reached through plain autograd) is coalesced into one temporary and copied back.  Returns the number of all-reduce calls."""
    world = dist.get_world_size(group) if dist.is_available() and dist.is_initialized() else 1
    if world == 1:
        return 0
    calls = 0
    for net in nets:
        grads = [p.grad for p in net.parameters()]
        grads = [g for g in grads if g is not None]
        if not grads:
            continue
        runs, stray = _grad_segments(grads)
        for r in runs:
            allreduce_mean_(r, group)
            calls += 1
        if stray:
            flat = torch.cat([g.reshape(-1) for g in stray])
            allreduce_mean_(flat, group)
            calls += 1
            torch._foreach_copy_(stray, [c.view(g.shape) for c, g in zip(flat.split([g.numel() for g in stray]), stray)])
    return calls


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
