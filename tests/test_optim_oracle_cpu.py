"""CPU tests of the parameter-update row: the numpy oracle against golden vectors made by the real
torch.optim.Adam / AdamW + clip_grad_norm_ + the reference's model_ema loop (tests/golden/make_golden_optim.py), against the
live torch implementation, and the host-side behaviour of dcpt_b200.optim that needs no GPU."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden_optim as MG  # noqa: E402
from oracle import optim_oracle as OO  # noqa: E402

TOL = 2e-6  # fp32 element-wise arithmetic in a different association order than ATen's vectorised kernels


def _close(a, b, what):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    err = np.abs(a - b).max() / max(np.abs(b).max(), 1e-30)
    assert err < TOL, (what, err)


def run_oracle(name):
    cls, kw, clip, decay, steps = MG.CASES[name]
    params, grads = MG.tensors(sum(map(ord, name)))
    P = [p.numpy().copy() for p in params]
    E = [p.numpy().copy() for p in params]
    M = [np.zeros_like(p) for p in P]
    V = [np.zeros_like(p) for p in P]
    adam = dict(lr=kw["lr"], betas=kw.get("betas", (0.9, 0.999)), eps=kw.get("eps", 1e-8),
                weight_decay=kw.get("weight_decay", 1e-2 if cls == "AdamW" else 0.0), decoupled=cls == "AdamW")
    norms = []
    for t in range(steps):
        n = OO.train_update(P, [g.numpy() for g in grads[t]], M, V, E, t + 1, grad_clip=clip, ema_decay=decay, **adam)
        if n is not None:
            norms.append(float(n))
    return P, M, V, E, norms


@pytest.mark.parametrize("name", list(MG.CASES))
def test_oracle_matches_golden(golden_dir, name):
    z = np.load(os.path.join(golden_dir, "optim_step.npz"))
    P, M, V, E, norms = run_oracle(name)
    for i in range(len(P)):
        _close(P[i], z[f"{name}|p{i}"], f"p{i}")
        _close(M[i], z[f"{name}|m{i}"], f"m{i}")
        _close(V[i], z[f"{name}|v{i}"], f"v{i}")
        _close(E[i], z[f"{name}|e{i}"], f"e{i}")
    np.testing.assert_allclose(norms, z[f"{name}|norms"], rtol=1e-6)


def test_golden_is_what_torch_produces_here(golden_dir):
    """The committed fixture equals a fresh run of the generating script (torch CPU, this container's version)."""
    z = np.load(os.path.join(golden_dir, "optim_step.npz"))
    fresh = MG.run_case("adamw_clip_ema")
    for k, v in fresh.items():
        np.testing.assert_allclose(v, z[k], rtol=1e-6, atol=1e-12, err_msg=k)


def test_host_side_refuses_cpu_and_unsupported():
    from dcpt_b200.lib import DcptError
    from dcpt_b200 import optim as FO
    with pytest.raises(NotImplementedError):
        FO.get_optimizer("SGD", [torch.nn.Parameter(torch.zeros(3))], 1e-3)      # base_model.py:120-139 raises the same
    lib_there = os.path.exists(os.path.join(os.path.dirname(FO.__file__), "libdcpt_sm100.so"))
    if not lib_there:
        with pytest.raises(DcptError):
            FO.FusedAdamW([torch.nn.Parameter(torch.zeros(3))])
        return
    with pytest.raises(DcptError):
        FO.FusedAdam([torch.nn.Parameter(torch.zeros(3))], amsgrad=True)
    p = torch.nn.Parameter(torch.zeros(3))
    opt = FO.get_optimizer("AdamW", [p], 1e-3, weight_decay=0.0, betas=[0.9, 0.9])
    assert opt.step() is None                    # no gradients: nothing to do, as torch
    p.grad = torch.ones(3)
    with pytest.raises(DcptError):
        opt.step()                               # CPU parameter: no CPU path
    # param_groups / state_dict layout is torch.optim.AdamW's (the reference's resume files, base_model.py:413-430)
    ref = torch.optim.AdamW([torch.nn.Parameter(torch.zeros(3))], lr=1e-3, weight_decay=0.0, betas=(0.9, 0.9))
    mine, theirs = opt.state_dict()["param_groups"][0], ref.state_dict()["param_groups"][0]
    for k in ("lr", "betas", "eps", "weight_decay", "amsgrad", "maximize", "params"):
        assert tuple(mine[k]) == tuple(theirs[k]) if isinstance(mine[k], (list, tuple)) else mine[k] == theirs[k], k
