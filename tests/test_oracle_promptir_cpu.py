"""The PromptIR oracle (oracle/promptir_oracle.py) against the reference: the committed golden of the real ``PromptIR`` class
(tests/golden/promptir_net.npz) and, when /root/reference is present, the live module - parameter names / order / shapes,
PromptGenBlock alone, the whole forward.  Also: the mirror arch's state_dict contract and the C plan's parameter layout."""
import os
import sys

import numpy as np
import pytest
import torch

from oracle import promptir_oracle as PO
from oracle._ref_import import reference_available

sys.path.insert(0, os.path.join(os.path.dirname(__file__), "golden"))
import make_golden_promptir as MG  # noqa: E402


def rel(a, b):
    a, b = torch.as_tensor(np.asarray(a)).double(), torch.as_tensor(np.asarray(b)).double()
    return float((a - b).norm() / b.norm())


def test_promptir_net_golden(golden_dir):
    z = np.load(os.path.join(golden_dir, "promptir_net.npz"))
    sd = PO.random_promptir_state_dict(seed=MG.NET_SEED, **MG.NET_CFG)
    assert sum(v.numel() for v in sd.values()) == int(z["n_params"])
    assert abs(float(sum(v.double().sum() for v in sd.values())) - float(z["param_checksum"])) < 1e-6 * abs(float(z["param_checksum"]))
    for tag in ("a", "b"):
        with torch.no_grad():
            y = PO.promptir_fwd(torch.from_numpy(z["x_" + tag]), sd, MG.NET_CFG["num_blocks"], MG.NET_CFG["num_refinement_blocks"])
        assert rel(y, z["y_" + tag]) < 1e-4, tag


def test_promptir_mirror_contract():
    """The dropped-in arch file: same parameter names, order and shapes as the oracle's table (= the reference's, next test), and
    the C plan built from the ctor arguments lays its parameters out identically (CPU: plan construction only)."""
    from basicsr.archs import build_network
    for cfg in (dict(num_blocks=[1, 1, 1, 1], num_refinement_blocks=1), dict()):
        net = build_network(dict(type="PromptIR", window_size=8, **cfg))
        shapes = PO.promptir_param_shapes(**{k: tuple(v) if isinstance(v, list) else v for k, v in cfg.items()})
        assert [(k, tuple(v.shape)) for k, v in net.named_parameters()] == [(k, tuple(v)) for k, v in shapes.items()]
        assert net.engine().numels == [v.numel() for v in net.parameters()]
    assert sum(p.numel() for p in net.parameters()) == 35_376_967
    with pytest.raises(Exception, match="inference-only"):
        net(torch.rand(1, 3, 16, 16))
    with pytest.raises(Exception, match="no CPU path"), torch.no_grad():
        net(torch.rand(1, 3, 16, 16))
    with pytest.raises(Exception, match="dim 48"):
        build_network(dict(type="PromptIR", dim=32)).engine()
    with pytest.raises(Exception, match="bias=True"):
        build_network(dict(type="PromptIR", bias=True))


@pytest.mark.skipif(not reference_available(), reason="/root/reference is not present")
def test_promptir_oracle_vs_live_reference():
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    code = '''
import sys, json
sys.path.insert(0, %r)
import torch
from oracle._ref_import import import_reference
from oracle import promptir_oracle as PO
import_reference()
from basicsr.archs.promptir_arch import PromptIR, PromptGenBlock
cfg = dict(num_blocks=(2, 1, 1, 1), num_refinement_blocks=2, heads=(1, 2, 4, 8))
net = PromptIR(num_blocks=list(cfg["num_blocks"]), num_refinement_blocks=2, LayerNorm_type="BiasFree").eval()
shapes = PO.promptir_param_shapes(LayerNorm_type="BiasFree", **cfg)
same = [(k, tuple(v.shape)) for k, v in net.named_parameters()] == [(k, tuple(v)) for k, v in shapes.items()]
sd = PO.random_promptir_state_dict(seed=5, LayerNorm_type="BiasFree", **cfg)
net.load_state_dict(sd, strict=True)
x = torch.rand(1, 3, 48, 56)
with torch.no_grad():
    r_net = float((PO.promptir_fwd(x, sd, cfg["num_blocks"], 2, cfg["heads"]) - net(x)).norm() / net(x).norm())
    f = torch.randn(2, 192, 20, 12)
    pg = float((PO.prompt_gen(f, sd, "prompt2") - net.prompt2(f)).norm() / net.prompt2(f).norm())
print("RESULT " + json.dumps({"same": same, "net": r_net, "prompt": pg, "default_params": sum(p.numel() for p in PromptIR().parameters())}))
''' % root
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=900, cwd=root)
    assert r.returncode == 0, r.stderr[-2000:]
    import json
    res = json.loads([l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1][7:])
    assert res["same"] and res["net"] < 1e-4 and res["prompt"] < 1e-5 and res["default_params"] == 35_376_967, res
