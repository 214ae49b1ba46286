"""Import the *real* reference (``/root/reference``) in the build container.

TEST INFRASTRUCTURE ONLY.  Used by ``tests/golden/make_golden.py`` and by the
CPU tests that pin the oracle against the reference when it is present.  The
GPU box has no ``/root/reference``; nothing run there may call this.

The reference imports five packages that are not installed offline
(SURVEY.md §8(c)); they are stubbed with the minimal behaviour the reference
uses:
  * ``skimage.io``                       (basicsr/utils/img_util.py:6)
  * ``fvcore.nn.weight_init.c2_msra_fill`` (basicsr/archs/degrad_classify_arch.py:3,213)
  * ``timm.models.layers`` / ``timm.utils.metrics.accuracy``
  * ``torchinfo.summary``                (basicsr/models/base_model.py:10)
"""
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("DCPT_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isdir(os.path.join(REFERENCE_ROOT, "basicsr", "archs"))


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def import_reference(root=None):
    """Returns the reference's ``basicsr`` package, imported from REFERENCE_ROOT (or from ``root``: a directory holding a
    ``basicsr`` tree, e.g. the reference checkout with this repo's arch files dropped in - tests/boundary_overlay_check.py)."""
    if not reference_available():
        raise RuntimeError(f"reference not present at {REFERENCE_ROOT}")
    root = root or REFERENCE_ROOT
    import torch
    from torch import nn

    if "skimage" not in sys.modules:
        sk = _stub("skimage")
        sk.io = _stub("skimage.io")
    if "fvcore" not in sys.modules:
        def c2_msra_fill(module):
            nn.init.kaiming_normal_(module.weight, mode="fan_out", nonlinearity="relu")
            if module.bias is not None:
                nn.init.constant_(module.bias, 0)
        fv = _stub("fvcore")
        fv.nn = _stub("fvcore.nn")
        fv.nn.weight_init = _stub("fvcore.nn.weight_init", c2_msra_fill=c2_msra_fill)
    if "timm" not in sys.modules:
        class DropPath(nn.Identity):
            def __init__(self, *a, **k):
                super().__init__()

        def to_2tuple(x):
            return tuple(x) if isinstance(x, (tuple, list)) else (x, x)

        def accuracy(output, target, topk=(1,)):
            maxk = min(max(topk), output.size(1))
            _, pred = output.topk(maxk, 1, True, True)
            pred = pred.t()
            correct = pred.eq(target.reshape(1, -1).expand_as(pred))
            return [correct[: min(k, maxk)].reshape(-1).float().sum(0) * 100.0 / target.size(0) for k in topk]

        tm = _stub("timm")
        tm.models = _stub("timm.models")
        tm.models.layers = _stub("timm.models.layers", DropPath=DropPath, to_2tuple=to_2tuple,
                                 trunc_normal_=nn.init.trunc_normal_)
        tm.utils = _stub("timm.utils")
        tm.utils.metrics = _stub("timm.utils.metrics", accuracy=accuracy)
    if "torchinfo" not in sys.modules:
        _stub("torchinfo", summary=lambda *a, **k: "")

    # our own repo also ships a ``basicsr`` mirror package: make sure the
    # reference wins inside this process.
    for k in [k for k in sys.modules if k == "basicsr" or k.startswith("basicsr.")]:
        del sys.modules[k]
    if root in sys.path:
        sys.path.remove(root)
    sys.path.insert(0, root)
    import basicsr  # noqa: F401  (the reference's)
    assert os.path.abspath(basicsr.__file__).startswith(os.path.abspath(root)), basicsr.__file__
    return basicsr


class reference_scope:
    """``with reference_scope() as basicsr:`` - the reference's package for the duration of the block, then this process goes back
    to whatever ``basicsr`` it had (this repo's mirror): the ``basicsr*`` entries of ``sys.modules`` and ``sys.path`` are restored
    on exit.  For in-process use inside a test session that also exercises the mirror package."""

    def __init__(self, root=None):
        self.root = root

    def __enter__(self):
        self._mods = {k: v for k, v in sys.modules.items() if k == "basicsr" or k.startswith("basicsr.")}
        self._path = list(sys.path)
        return import_reference(self.root)

    def __exit__(self, *exc):
        for k in [k for k in sys.modules if k == "basicsr" or k.startswith("basicsr.")]:
            del sys.modules[k]
        sys.modules.update(self._mods)
        sys.path[:] = self._path
        return False
