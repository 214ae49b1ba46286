"""CPU-side checks of the drop-in boundary: the C-ABI library builds/loads, exports exactly what include/dcpt_ops.h
declares, argument errors come back as codes + messages (no exceptions, no CUDA needed), and the product package
never touches oracle/ or falls back to CPU."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from dcpt_b200.build import build
    build()
    from dcpt_b200.lib import load_library
    return load_library()


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "dcpt_ops.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(dcpt_[a-z0-9_]+)\s*\(", txt)))


def test_header_symbols_are_exported_and_bound(lib):
    from dcpt_b200.lib import LIB_PATH, PROTOTYPES
    declared = _declared_symbols()
    assert len(declared) >= 20
    out = subprocess.run(["nm", "-D", "--defined-only", LIB_PATH], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r" T (dcpt_[a-z0-9_]+)", out))
    missing = [s for s in declared if s not in exported]
    assert not missing, f"declared in dcpt_ops.h but not exported: {missing}"
    extra = sorted(exported - set(declared))
    assert not extra, f"exported but undeclared: {extra}"
    assert sorted(PROTOTYPES) == declared, "ctypes prototypes out of sync with the header"
    assert lib.dcpt_abi_version() == 1


def test_plan_and_size_queries_need_no_gpu(lib):
    enc = (ctypes.c_int * 4)(1, 1, 1, 28)
    dec = (ctypes.c_int * 4)(1, 1, 1, 1)
    plan = ctypes.c_void_p(lib.dcpt_nafnet_create(3, 64, 1, enc, 4, dec, 4))
    assert plan.value
    assert lib.dcpt_nafnet_num_params(plan) == 664                       # NAFNet-w64 has 664 tensors
    dims = (ctypes.c_int * 4)()
    total = sum(lib.dcpt_nafnet_param_shape(plan, i, dims) for i in range(664))
    assert total == 67_888_835                                            # SURVEY.md §2c parameter count
    assert lib.dcpt_nafnet_saved_bytes(plan, 16, 256, 256) > lib.dcpt_nafnet_saved_bytes(plan, 1, 256, 256) > 0
    assert lib.dcpt_nafnet_workspace_bytes(plan, 1, 256, 256) > 0 and lib.dcpt_nafnet_packed_bytes(plan) > 0
    lib.dcpt_nafnet_destroy(plan)
    assert lib.dcpt_nafblock_saved_bytes(2, 8, 8, 64) > 0 and lib.dcpt_nafblock_packed_bytes(64) > 0


def test_argument_errors_are_codes_not_exceptions(lib):
    enc = (ctypes.c_int * 1)(1)
    assert not lib.dcpt_nafnet_create(3, 12, 1, enc, 1, enc, 1)           # width % 8 != 0
    assert b"width" in lib.dcpt_last_error()
    rc = lib.dcpt_layernorm2d_fwd(None, None, None, None, None, 16, 12, 1e-6, None)   # C % 8 != 0 -> rejected before any launch
    assert rc == -2 and b"layernorm" in lib.dcpt_last_error()
    rc = lib.dcpt_gemm_bf16(None, 8, 0, None, 8, 0, 0, 8, 8, None, None, 8, None, None, 1, 0, 0, None)
    assert rc == -2


def test_product_never_imports_oracle_or_falls_back():
    for pkg in ("dcpt_b200", "basicsr"):
        for dirpath, _, files in os.walk(os.path.join(ROOT, pkg)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    txt = open(os.path.join(dirpath, f)).read()
                    assert not re.search(r"^\s*(from|import)\s+oracle\b", txt, flags=re.M), f"{pkg}/{f} imports oracle"
    code = ("import torch, sys; sys.path.insert(0, %r)\n"
            "from basicsr.archs import build_network\n"
            "net = build_network(dict(type='NAFNetBaseline', width=8, enc_blk_nums=[1], middle_blk_num=1, dec_blk_nums=[1]))\n"
            "try:\n    net(torch.rand(1, 3, 8, 8)); print('RAN')\n"
            "except Exception as e:\n    print(type(e).__name__)\n" % ROOT)
    r = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=300)
    assert "DcptError" in r.stdout and "RAN" not in r.stdout, r.stdout + r.stderr


def test_registry_surface():
    from basicsr.archs import build_network
    from basicsr.utils.registry import ARCH_REGISTRY
    from oracle.nafnet_oracle import nafnet_state_dict_keys
    assert "NAFNetBaseline" in ARCH_REGISTRY
    kw = dict(width=64, enc_blk_nums=[1, 1, 1, 28], middle_blk_num=1, dec_blk_nums=[1, 1, 1, 1], window_size=16)
    net = build_network(dict(type="NAFNetBaseline", **kw))                # options/all_in_one/test/test_NAFNet_5d.yml:50-56
    got = [(k, tuple(v.shape)) for k, v in net.named_parameters()]
    assert got == nafnet_state_dict_keys(64, [1, 1, 1, 28], 1, [1, 1, 1, 1])
    assert list(net.state_dict().keys()) == [k for k, _ in got]
    with pytest.raises(AssertionError):
        ARCH_REGISTRY.register(type(net))                                 # duplicate names rejected (registry.py:42-45)


def test_cached_parameter_slots_follow_the_module():
    """NAFNetBaseline._param_list (the per-call replacement for list(self.parameters())) returns the live Parameter objects
    in named_parameters() order, also after a Parameter was replaced, and zero_grad keeps nn.Module's semantics."""
    import torch
    from basicsr.archs import build_network
    net = build_network(dict(type="NAFNetBaseline", width=8, enc_blk_nums=[1, 1], middle_blk_num=1, dec_blk_nums=[1, 1]))
    a, b = net._param_list(), list(net.parameters())
    assert len(a) == len(b) and all(x is y for x, y in zip(a, b))
    net.intro.weight = torch.nn.Parameter(torch.zeros_like(net.intro.weight))     # replaced object is picked up
    assert all(x is y for x, y in zip(net._param_list(), net.parameters()))
    for p in net.parameters():
        p.grad = torch.ones_like(p)
    net.zero_grad(set_to_none=False)
    assert all(p.grad is not None and float(p.grad.abs().sum()) == 0 for p in net.parameters())
    net.zero_grad()
    assert all(p.grad is None for p in net.parameters())


def test_optim_plan_chunks_need_no_gpu(lib):
    numels = (ctypes.c_longlong * 5)(1, 8192, 8193, 0, 100003)
    plan = ctypes.c_void_p(lib.dcpt_optim_create(numels, 5))
    assert plan.value
    assert lib.dcpt_optim_num_chunks(plan) == 1 + 1 + 2 + 0 + 13                  # 8192-element chunks, never across tensors
    assert lib.dcpt_optim_workspace_bytes(plan) > 0
    assert lib.dcpt_optim_step(plan, None, 1, 1e-3, 0.9, 0.999, 1e-8, 0.0, 1, 0.0, 0.0, None) == -1   # null workspace -> code
    lib.dcpt_optim_destroy(plan)
    assert not lib.dcpt_optim_create(numels, 0) and b"optim_create" in lib.dcpt_last_error()


def test_graph_slot_ownership_tokens():
    """A slot released by the finalizer of an OLD autograd node must not be freed under a newer forward that owns it."""
    from dcpt_b200.nafnet import _GraphSlot
    from dcpt_b200.restormer import _TrainSlot
    for cls in (_GraphSlot, _TrainSlot):
        s = object.__new__(cls)
        s.busy, s.owner = False, None
        t1 = s.acquire()
        assert s.busy
        s.release()                      # backward of forward 1 ran
        assert not s.busy
        t2 = s.acquire()                 # forward 2 takes the slot
        s.release(t1)                    # ... then forward 1's node is garbage-collected: stale token, no effect
        assert s.busy and s.owner is t2
        s.release(t2)                    # forward 2 dropped without backward
        assert not s.busy


def test_restormer_module_hook_plumbing_without_gpu(monkeypatch):
    """Host logic of the Restormer mirror in a DCPT step (degradation_classification_pretrain_model.py:60-68, 140, 154): forward
    hooks registered on the one-dot `decoder_level{k}.body` modules fire in the order level 3, 2, 1 with the engine's feature
    outputs, hook=True returns None, and the parameters a hooked pass never reaches are named to the autograd glue.  The
    engine call is stubbed (no GPU here); the CUDA path is covered by tests/test_gpu_restormer.py."""
    import torch
    import basicsr.archs.restormer_arch as RA
    from dcpt_b200.lib import DcptError
    net = RA.Restormer(dim=16, num_blocks=[1, 1, 1, 1], num_refinement_blocks=1, heads=[1, 2, 4, 8])
    calls = []

    def fake_apply(engine, inp, params, hook=False, want_feats=False, dead=()):
        calls.append(dict(hook=hook, want_feats=want_feats, dead=list(dead), n=len(params)))
        N, _, H, W = inp.shape
        feats = [torch.full((N, 64, H // 4, W // 4), 3.0), torch.full((N, 32, H // 2, W // 2), 2.0), torch.full((N, 32, H, W), 1.0)]
        return (None if hook else inp + 1), (feats if want_feats else [])
    monkeypatch.setattr(RA, "restormer_apply", fake_apply)
    monkeypatch.setattr(RA.Restormer, "engine", lambda self: None)
    got = []
    names = [n for n, m in net.named_modules() if "decoder" in n and n.count(".") == 1]
    assert names == ["decoder_level3.body", "decoder_level2.body", "decoder_level1.body"]
    for n, m in net.named_modules():
        if n in names:
            m.register_forward_hook(lambda mod, i, o: got.append(float(o.flatten()[0])))
    x = torch.rand(1, 3, 16, 16)
    with pytest.raises(DcptError):
        net(x)                                             # CPU tensor in training mode: no CPU path
    monkeypatch.setattr(torch.Tensor, "is_cuda", property(lambda self: True))
    out = net(x, hook=False)
    assert out is not None and got == [3.0, 2.0, 1.0] and calls[-1]["want_feats"] and calls[-1]["dead"] == []
    got.clear()
    assert net(x, hook=True) is None and got == [3.0, 2.0, 1.0]
    keys = [k for k, _ in net.named_parameters()]
    dead = [keys[i] for i in calls[-1]["dead"]]
    assert dead and all(k.startswith(("refinement.", "output.")) for k in dead)
    assert len(dead) == sum(k.startswith(("refinement.", "output.")) for k in keys) and calls[-1]["n"] == len(keys)
