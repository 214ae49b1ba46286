"""BASELINE.json configs[3] (C4) through the step class: DCPTModel.optimize_parameters with NAFNet-w64 + PromptIR_NoImg_DC
(f = [64,128,256,512]), 8 x 3 x 256 x 256 per GPU, L1 + cross entropy, two fused AdamW steps - one process per GPU
(python tools/dcpt_model_bench.py, or under torchrun for N > 1: both networks in FlatGradDataParallel, one gradient exchange
per network per step).  Prints one JSON line on rank 0; iteration tool, not the bench contract."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402


def main():
    steps = int(os.getenv("STEPS", "10"))
    warmup = int(os.getenv("WARMUP", "4"))
    batch = int(os.getenv("BATCH", "8"))
    world = int(os.getenv("WORLD_SIZE", "1"))
    rank = int(os.getenv("RANK", "0"))
    torch.cuda.set_device(int(os.getenv("LOCAL_RANK", "0")))
    from dcpt_b200 import dist as D
    if world > 1:
        D.init_dist("nccl")
    from basicsr.models import build_model
    adamw = {"type": "AdamW", "lr": 3e-4, "weight_decay": 1e-4, "betas": [0.9, 0.9]}
    opt = {"name": "c4", "model_type": "DCPTModel", "scale": 1, "num_gpu": 1, "dist": world > 1, "is_train": True, "rank": rank,
           "world_size": world, "hook_names": "decoder", "path": {"pretrain_network_g": None},
           "network_g": dict(type="NAFNetBaseline", width=64, enc_blk_nums=[1, 1, 1, 28], middle_blk_num=1, dec_blk_nums=[1, 1, 1, 1],
                             window_size=16),
           "network_dc": dict(type="PromptIR_NoImg_DC", feature_dims=[64, 128, 256, 512], num_res_blocks=2, num_classes=5),
           "train": {"optim_g": dict(adamw), "optim_dc": dict(adamw), "scheduler": {"type": "MultiStepLR", "milestones": [10 ** 6], "gamma": 0.5},
                     "pixel_opt": {"type": "L1Loss", "loss_weight": 1.0, "reduction": "mean"},
                     "classify_opt": {"type": "CrossEntropyLoss", "loss_weight": 1.0}}}
    torch.manual_seed(0)
    model = build_model(opt)
    calls = []
    if world > 1:
        real = D.allreduce_mean_
        D.allreduce_mean_ = lambda flat, group=None: (calls.append(flat.numel() * 4), real(flat, group))[1]
    xev = []
    if world > 1:
        real_x = D.exchange_accumulated_grads_

        def timed_exchange(nets, group=None):
            if os.getenv("EXCHANGE", "1") == "0":
                return 0
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            n = real_x(nets, group)
            b.record()
            xev.append((a, b))
            return n
        D.exchange_accumulated_grads_ = timed_exchange
    g = torch.Generator().manual_seed(D.rank_seed(0, rank))           # seed + rank (options.py:142)
    gt = torch.rand(batch, 3, 256, 256, generator=g).pin_memory()
    lq = torch.rand(batch, 3, 256, 256, generator=g).pin_memory()
    idx = torch.randint(0, 5, (batch,), generator=g).pin_memory()
    logs = []

    sev = []

    def step(it):
        ev = torch.cuda.Event(enable_timing=True)
        ev.record()
        sev.append(ev)
        model.update_learning_rate(it)
        model.feed_data({"lq": lq, "gt": gt, "dataset_idx": idx})
        model.optimize_parameters(it)
        logs.append(dict(model.get_current_log()))

    for it in range(1, warmup + 1):
        step(it)
    n_calls_per_step = len(calls) // max(warmup, 1)
    bytes_per_step = sum(calls) // max(warmup, 1)
    torch.cuda.synchronize()
    if world > 1:
        torch.distributed.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for it in range(warmup + 1, warmup + steps + 1):
        step(it)
    e1.record()
    torch.cuda.synchronize()
    wall = (time.perf_counter() - t0) * 1e3 / steps
    ms = torch.tensor([e0.elapsed_time(e1) / steps], device="cuda")
    chk = torch.stack([torch.cat([p.detach().reshape(-1) for p in model.get_bare_model(n).parameters()]).double().sum()
                       for n in (model.net_g, model.net_dc)])
    same = True
    pre = post = None
    if world > 1 and xev:
        mine = torch.tensor([sum(s.elapsed_time(a) for s, (a, b) in zip(sev[-steps:], xev[-steps:])) / steps,
                             sum(a.elapsed_time(b) for a, b in xev[-steps:]) / steps], device="cuda")
        allr = [torch.empty_like(mine) for _ in range(world)]
        torch.distributed.all_gather(allr, mine)
        pre = [round(float(t[0]), 3) for t in allr]
        post = [round(float(t[1]), 3) for t in allr]
    if world > 1:
        torch.distributed.all_reduce(ms, op=torch.distributed.ReduceOp.MAX)
        gathered = [torch.empty_like(chk) for _ in range(world)]
        torch.distributed.all_gather(gathered, chk)
        same = all(torch.equal(gathered[0], x) for x in gathered)
    if rank == 0:
        ms = float(ms)
        print(json.dumps({"config": "C4 DCPTModel.optimize_parameters (NAFNet-w64 + PromptIR_NoImg_DC f=[64..512])", "n_gpus": world,
                          "batch_per_gpu": batch, "ms_per_step": round(ms, 3), "wall_ms_per_step": round(wall, 3),
                          "images_per_s": round(world * batch / ms * 1e3, 2), "allreduce_calls_per_step": n_calls_per_step,
                          "allreduce_bytes_per_step": bytes_per_step,
                          "exchange_ms_per_step": round(sum(a.elapsed_time(b) for a, b in xev[-steps:]) / steps, 3) if xev else 0.0, "per_rank_ms_step_start_to_exchange": pre,
                          "per_rank_ms_in_exchange": post, "replicas_identical_after_run": same,
                          "first_log": logs[0], "last_log": logs[-1], "steps": steps, "warmup": warmup}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
