"""BasicSR-compatible plugin surface for the B200-native DCPT hot path.

Mirrors the reference's package layout (MILab-PKU/dcpt ``basicsr/``) only as far as the hot path
needs it: the registries (``basicsr/utils/registry.py``), ``build_network`` (``basicsr/archs/__init__.py``)
and the arch classes whose ctor kwargs and ``state_dict`` keys are the compatibility contract, so the
reference's ``options/*.yml`` ``network_g`` sections and checkpoints load unchanged.  The module
bodies call the sm_100a kernels in ``dcpt_b200`` through the C ABI (``include/dcpt_ops.h``).
"""
