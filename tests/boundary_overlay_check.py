"""Boundary proof on the reference's OWN model classes (VERDICT r1 "Next" #8) - run by tests/test_boundary_reference_cpu.py in
a child interpreter (it replaces the `basicsr` package of the process).  TEST INFRASTRUCTURE; needs /root/reference.

What a maintainer does to adopt this repo is drop three files into the reference checkout:
`basicsr/archs/{nafnet,restormer,degrad_classify}_arch.py`.  This script builds exactly that tree (symlinks to the reference,
the three arch files from this repo), imports it, and drives the reference's real `SRModel` / `DCPTModel`
(models/sr_model.py, models/degradation_classification_pretrain_model.py) through it - once with the overlay, once with the
untouched reference - on CPU (`num_gpu: 0`).  There is no GPU here, so in the overlay run the engine entry points
(`nafnet_apply`, `dchead_apply`) are replaced by the CPU oracle's functional forward: the kernels are not what is being tested,
the plumbing around them is - strict `params_ema` loading, the one-dot hook rule, `hook=True` returning None, `pre_test` /
`post_test` padding, optimizer parameter order, the two-pass DCPT step.  Prints one JSON line per scenario."""
import json
import os
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from oracle import _ref_import as R  # noqa: E402

OURS = ["nafnet_arch.py", "restormer_arch.py", "degrad_classify_arch.py"]
CFG = dict(width=16, enc_blk_nums=[1, 1, 1, 2], middle_blk_num=1, dec_blk_nums=[1, 1, 1, 1])
DIMS = [16, 32, 64, 128]


def make_overlay(dst):
    """<dst>/basicsr = the reference tree (symlinks) with this repo's three arch files dropped in."""
    src = os.path.join(R.REFERENCE_ROOT, "basicsr")
    for dirpath, dirnames, filenames in os.walk(src):
        dirnames[:] = [d for d in dirnames if d != "__pycache__"]
        rel = os.path.relpath(dirpath, src)
        os.makedirs(os.path.join(dst, "basicsr", rel), exist_ok=True)
        for f in filenames:
            target = os.path.join(dirpath, f)
            if rel == "archs" and f in OURS:
                target = os.path.join(ROOT, "basicsr", "archs", f)
            os.symlink(target, os.path.join(dst, "basicsr", rel, f))
    return dst


def stub_engines():
    """No GPU in this container: route the two engine entry points of the dropped-in arch files through the CPU oracle."""
    import basicsr.archs.degrad_classify_arch as DA
    import basicsr.archs.nafnet_arch as NA
    from oracle import dchead_oracle as D
    from oracle import nafnet_oracle as O

    def nafnet_apply(engine, inp, params, hook=False, want_feats=False):
        keys = [k for k, _ in O.nafnet_state_dict_keys(engine.width, CFG["enc_blk_nums"], CFG["middle_blk_num"], CFG["dec_blk_nums"])]
        sd = dict(zip(keys, params))
        feats = [] if want_feats else None
        blocks = getattr(engine, "_hook_blocks", None)
        blk = None
        if want_feats and blocks and any(b >= 0 for b in blocks):
            blk = max(blocks)
        out = O.nafnet_fwd(inp, sd, CFG["enc_blk_nums"], CFG["middle_blk_num"], CFG["dec_blk_nums"], hook=hook, decoder_feats=feats,
                           decoder_feat_block=blk)
        return out, (feats or [])

    def dchead_apply(engine, feats, params):
        return D.dchead_fwd(list(feats), dict(zip(engine.names, params)))
    NA.nafnet_apply = nafnet_apply
    DA.dchead_apply = dchead_apply


def base_opt(model_type, is_train):
    return {"name": "boundary", "model_type": model_type, "scale": 1, "num_gpu": 0, "dist": False, "is_train": is_train, "rank": 0,
            "world_size": 1, "manual_seed": 0,
            "network_g": dict(type="NAFNetBaseline", window_size=16, **CFG),
            "path": {"pretrain_network_g": None, "strict_load_g": True, "param_key_g": "params_ema"}, "logger": {"print_freq": 1}}


def scenario_sr(basicsr, ckpt):
    """options/all_in_one/test/test_NAFNet_5d.yml's model section at a small width: SRModel, `param_key_g: params_ema`, strict
    load, pre_test (reflect pad to window_size) -> test -> post_test on a 100 x 120 image."""
    from basicsr.models import build_model
    from basicsr.models.base_model import BaseModel
    BaseModel.print_network = lambda self, *a, **k: None        # (the reference calls .cuda() there unconditionally)
    opt = base_opt("SRModel", False)
    opt["path"]["pretrain_network_g"] = ckpt
    model = build_model(opt)
    g = torch.Generator().manual_seed(11)
    lq = torch.rand(1, 3, 100, 120, generator=g)
    model.feed_data({"lq": lq})
    model.pre_test()
    padded = tuple(model.lq.shape)
    model.test()
    model.post_test()
    out = model.output
    return {"padded": padded, "out_shape": tuple(out.shape), "out": out.detach().flatten()[::997].tolist(), "out_norm": float(out.norm()),
            "n_params": sum(p.numel() for p in model.net_g.parameters()), "keys": list(model.net_g.state_dict().keys())[:6]}


def scenario_dcpt(basicsr, sd_g, sd_h):
    """DCPTModel.optimize_parameters: pixel pass on gt, hooked pass on lq (`hook_names: decoder`), classifier on the reversed
    hook outputs, L1 + CE, one backward, both AdamW steps."""
    from basicsr.models import build_model
    from basicsr.models.base_model import BaseModel
    BaseModel.print_network = lambda self, *a, **k: None
    opt = base_opt("DCPTModel", True)
    opt["network_dc"] = dict(type="PromptIR_NoImg_DC", feature_dims=DIMS, num_res_blocks=2, num_classes=5)
    opt["hook_names"] = "decoder"
    opt["train"] = {"optim_g": {"type": "AdamW", "lr": 1e-3, "weight_decay": 1e-4, "betas": [0.9, 0.9]},
                    "optim_dc": {"type": "AdamW", "lr": 1e-3, "weight_decay": 1e-4, "betas": [0.9, 0.9]},
                    "scheduler": {"type": "MultiStepLR", "milestones": [100], "gamma": 0.5},
                    "pixel_opt": {"type": "L1Loss", "loss_weight": 1.0, "reduction": "mean"},
                    "classify_opt": {"type": "CrossEntropyLoss", "loss_weight": 1.0}}
    model = build_model(opt)
    model.net_g.load_state_dict(sd_g, strict=True)
    model.net_dc.load_state_dict(sd_h, strict=True)
    hooked = [n for n, m in model.net_g.named_modules() if len(m._forward_hooks) > 0]
    g = torch.Generator().manual_seed(12)
    gt, lq = torch.rand(2, 3, 32, 32, generator=g), torch.rand(2, 3, 32, 32, generator=g)
    model.feed_data({"lq": lq, "gt": gt, "dataset_idx": torch.tensor([4, 1])})
    ret_hook = model.net_g(lq, hook=True)
    n_hook_out = len(model.hook_outputs)
    model.hook_outputs = list()
    model.optimize_parameters(1)
    log = {k: float(v) for k, v in model.get_current_log().items()}
    opt_shapes = [tuple(p.shape) for p in model.optimizer_g.param_groups[0]["params"]]
    named_shapes = [tuple(p.shape) for _, p in model.net_g.named_parameters()]
    after_g = torch.cat([p.detach().flatten() for p in model.net_g.parameters()])
    after_h = torch.cat([p.detach().flatten() for p in model.net_dc.parameters()])
    return {"hooked": hooked, "hook_true_returns_none": ret_hook is None, "n_hook_outputs": n_hook_out, "log": log,
            "optimizer_order_is_named_parameters_order": opt_shapes == named_shapes, "n_opt_params": len(opt_shapes),
            "after_g": after_g[::10007].tolist(), "after_g_norm": float(after_g.norm()), "after_h": after_h[::10007].tolist(),
            "after_h_norm": float(after_h.norm())}


def scenario_sr_train(basicsr, sd_g):
    """SRModel.optimize_parameters x 3 (sr_model.py:132-174): L1, top-level grad_clip, Adam, EMA 0.9, cosine-restart schedule
    advanced by update_learning_rate as basicsr/train.py does each iteration."""
    from basicsr.models import build_model
    from basicsr.models.base_model import BaseModel
    BaseModel.print_network = lambda self, *a, **k: None
    opt = base_opt("SRModel", True)
    opt["grad_clip"] = 0.05
    opt["train"] = {"ema_decay": 0.9, "optim_g": {"type": "Adam", "lr": 2e-3, "weight_decay": 0, "betas": [0.9, 0.99]},
                    "scheduler": {"type": "CosineAnnealingRestartLR", "periods": [2, 4], "restart_weights": [1, 0.5], "eta_min": [1e-5, 2e-5]},
                    "pixel_opt": {"type": "L1Loss", "loss_weight": 1.0, "reduction": "mean"}}
    model = build_model(opt)
    model.get_bare_model(model.net_g).load_state_dict(sd_g, strict=True)
    model.model_ema(0)
    g = torch.Generator().manual_seed(13)
    logs, lrs = [], []
    for it in range(1, 5):
        model.update_learning_rate(it, warmup_iter=-1)
        lrs.append(model.get_current_learning_rate()[0])
        model.feed_data({"lq": torch.rand(2, 3, 32, 32, generator=g), "gt": torch.rand(2, 3, 32, 32, generator=g)})
        model.optimize_parameters(it)
        logs.append(float(model.get_current_log()["l_pix"]))
    after = torch.cat([p.detach().flatten() for p in model.net_g.parameters()])
    ema = torch.cat([p.detach().flatten() for p in model.net_g_ema.parameters()])
    model.feed_data({"lq": torch.rand(1, 3, 32, 32, generator=g)})
    model.test()                                               # the EMA network answers (sr_model.py:176-180)
    return {"logs": logs, "lrs": lrs, "after": after[::10007].tolist(), "after_norm": float(after.norm()), "ema": ema[::10007].tolist(),
            "ema_norm": float(ema.norm()), "test_norm": float(model.output.norm())}


def scenario_schedules(basicsr):
    """Learning-rate trajectories of the two schedule families under warm-up (base_model.py:141-186; lr_scheduler.py)."""
    from basicsr.models.base_model import BaseModel
    out = {}
    for name, sched in (("multistep", {"type": "MultiStepLR", "milestones": [3, 6], "gamma": 0.5}),
                        ("cosine", {"type": "CosineAnnealingRestartLR", "periods": [4, 4, 6], "restart_weights": [1, 0.5, 0.25], "eta_min": [1e-6, 1e-6, 2e-6]})):
        m = BaseModel({"num_gpu": 0, "is_train": True, "dist": False, "train": {"scheduler": dict(sched)}})
        m.optimizers = [torch.optim.Adam([torch.nn.Parameter(torch.zeros(1))], lr=1e-3)]
        m.setup_schedulers()
        traj = []
        for it in range(1, 13):
            m.update_learning_rate(it, warmup_iter=3)
            traj.append(m.get_current_learning_rate()[0])
            m.optimizers[0].step()
        out[name] = traj
    return out


def run(root, overlay):
    if root == "mirror":                 # this repo's own basicsr package (arch files + the model / loss mirrors)
        import basicsr
    else:
        basicsr = R.import_reference(root)
    if overlay:
        stub_engines()
    from oracle import dchead_oracle as D
    from oracle import nafnet_oracle as O
    sd_g = O.random_nafnet_state_dict(seed=3, **CFG)
    sd_h = D.random_dchead_state_dict(DIMS, 2, 5, seed=4)
    with tempfile.TemporaryDirectory() as td:
        ckpt = os.path.join(td, "net_g.pth")
        torch.save({"params": {k: torch.zeros_like(v) for k, v in sd_g.items()}, "params_ema": sd_g}, ckpt)   # the yml loads params_ema
        sr = scenario_sr(basicsr, ckpt)
    dc = scenario_dcpt(basicsr, sd_g, sd_h)
    return {"sr": sr, "dcpt": dc, "sr_train": scenario_sr_train(basicsr, sd_g), "schedules": scenario_schedules(basicsr)}


if __name__ == "__main__":
    which = sys.argv[1]
    if which == "mirror":
        res = run("mirror", True)
    elif which == "overlay":
        with tempfile.TemporaryDirectory() as td:
            res = run(make_overlay(td), True)
    else:
        res = run(None, False)
    print("RESULT " + json.dumps(res))
