"""N-rank check + timing of the overlapped gradient exchange (dcpt_b200.dist.allreduce_grads_overlapped_): the same NAFNet-w64
step with (a) one plain all-reduce after the backward, (b) the split all-reduce issued eagerly after eager launches, (c) after a
CUDA-graph replay (external-event record node inside the graph, NCCL outside).  The reduced
gradient buffers must agree with (a) to the run-to-run level of two plain runs (split-K atomics make no two runs bit-identical).
    torchrun --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_overlap_check.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from bench import CFG  # noqa: E402
from dcpt_b200.dist import allreduce_grads_overlapped_, allreduce_mean_  # noqa: E402
from dcpt_b200.nafnet import NAFNetEngine  # noqa: E402
from oracle import nafnet_oracle as O  # noqa: E402  (weight generator only)

rank, local = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = int(os.getenv("B", "16"))
sd = O.random_nafnet_state_dict(seed=0, **CFG)
params = [v.to(dev).contiguous() for v in sd.values()]
eng = NAFNetEngine(3, CFG["width"], CFG["middle_blk_num"], CFG["enc_blk_nums"], CFG["dec_blk_nums"])
g = torch.Generator(device=dev).manual_seed(1 + rank)
inp = torch.rand(B, 3, 256, 256, device=dev, generator=g)
dout = torch.randn(B, 3, 256, 256, device=dev, generator=g) / inp.numel()
flat, grads = eng.alloc_flat_grads(params)
comm = torch.cuda.Stream()
a, b = eng.early_grad_range()


def step():
    flat.zero_()
    _, _, saved = eng.forward(params, inp)
    eng.backward(params, inp, saved, dout, grads=grads)


def dev_vs(ref):
    d = (flat - ref).abs()
    return float(d[a:b].max() / ref[a:b].abs().max()), float(torch.cat([d[:a], d[b:]]).max() / torch.cat([ref[:a], ref[b:]]).abs().max())


def timed(fn, n=20):
    for _ in range(3):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


class _Out(dict):
    def __setitem__(self, k, v):
        super().__setitem__(k, v)
        if rank == 0:
            print(f"{k:48s} {v}", flush=True)


out = _Out()
step(); allreduce_mean_(flat); torch.cuda.synchronize()
ref = flat.clone()
step(); allreduce_mean_(flat); torch.cuda.synchronize()
out["noise (plain vs plain)"] = dev_vs(ref)
g_plain = torch.cuda.CUDAGraph()
with torch.cuda.graph(g_plain):
    step()
out["t plain graph + AR"] = timed(lambda: (g_plain.replay(), allreduce_mean_(flat)))
out["t graph only (no AR)"] = timed(lambda: g_plain.replay())
# (b), (c): event recorded mid-backward, NCCL issued from the host after the launches
eng.enable_grad_overlap(True, external=True)
step(); allreduce_grads_overlapped_(eng, flat, comm_stream=comm); torch.cuda.synchronize()
out["eager overlapped vs plain"] = dev_vs(ref)
g_ext = torch.cuda.CUDAGraph()
step(); torch.cuda.synchronize()
with torch.cuda.graph(g_ext):
    step()
g_ext.replay(); allreduce_grads_overlapped_(eng, flat, comm_stream=comm); torch.cuda.synchronize()
out["graph (external event) overlapped vs plain"] = dev_vs(ref)
out["t graph(external event) + overlapped AR"] = timed(lambda: (g_ext.replay(), allreduce_grads_overlapped_(eng, flat, comm_stream=comm)))
noise = max(out["noise (plain vs plain)"])
ok = all(max(v) < 20 * noise + 1e-6 for k, v in out.items() if k.endswith("vs plain") and isinstance(v, tuple))
t = torch.tensor([int(ok)], device=dev)
dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0:
    print("all ranks ok:", bool(t.item()))
dist.destroy_process_group()
sys.exit(0 if ok else 1)
