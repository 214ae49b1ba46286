"""Model registry population + ``build_model`` (reference: basicsr/models/__init__.py:12-36): every ``*_model.py`` here is
imported so its ``@MODEL_REGISTRY.register()`` class exists.  Two step classes are mirrored - the ones on the hot path
(SURVEY.md section 8 row a12): ``SRModel`` (fine-tune step, test / tiled test) and ``DCPTModel`` (the two-pass pretrain step)."""
import importlib
import os
from copy import deepcopy

from basicsr.utils import get_root_logger, scandir
from basicsr.utils.registry import MODEL_REGISTRY

__all__ = ["build_model"]

_here = os.path.dirname(os.path.abspath(__file__))
_model_modules = [importlib.import_module(f"basicsr.models.{os.path.splitext(os.path.basename(f))[0]}")
                  for f in scandir(_here) if f.endswith("_model.py")]


def build_model(opt):
    model = MODEL_REGISTRY.get(opt["model_type"])(deepcopy(opt))
    get_root_logger().info(f"Model [{model.__class__.__name__}] is created.")
    return model
