"""``basicsr.ops.dcpt_sm100`` — where the reference keeps its native ops (``basicsr/ops/<op>/``).
Thin alias of the in-tree extension: the kernels and the ctypes binding live in ``dcpt_b200``;
``BASICSR_JIT=True`` builds ``libdcpt_sm100.so`` on first use, like the reference's ops."""
from dcpt_b200 import ops  # noqa: F401
from dcpt_b200.lib import DcptError, load_library  # noqa: F401
from dcpt_b200.nafnet import NAFNetEngine, nafnet_apply  # noqa: F401
