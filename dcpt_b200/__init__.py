"""dcpt_b200 — B200 (sm_100a) native hot path for MILab-PKU/dcpt's image-restoration networks.

Holds only what the hot path needs: ``csrc/`` (hand-written CUDA kernels + the C ABI declared
in ``include/dcpt_ops.h``), the ctypes loader, and the host-side engines the ``basicsr`` mirror
package calls.  There is no CPU / PyTorch fallback: everything here fails loudly when
``libdcpt_sm100.so`` is missing or no CUDA device is present.
"""
from .lib import LIB_PATH, DcptError, load_library  # noqa: F401

__version__ = "0.1.0"
