"""Arch registry population + ``build_network`` (reference: basicsr/archs/__init__.py:12-31):
every ``*_arch.py`` in this folder is imported so its ``@ARCH_REGISTRY.register()`` classes exist."""
import importlib
import os
from copy import deepcopy

from basicsr.utils import get_root_logger, scandir
from basicsr.utils.registry import ARCH_REGISTRY

__all__ = ["build_network"]

_here = os.path.dirname(os.path.abspath(__file__))
_arch_modules = [importlib.import_module(f"basicsr.archs.{os.path.splitext(os.path.basename(f))[0]}")
                 for f in scandir(_here) if f.endswith("_arch.py")]


def build_network(opt):
    opt = deepcopy(opt)
    network_type = opt.pop("type")
    net = ARCH_REGISTRY.get(network_type)(**opt)
    get_root_logger().info(f"Network [{net.__class__.__name__}] is created.")
    return net
