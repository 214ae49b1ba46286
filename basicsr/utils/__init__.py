import logging
import os

from .registry import ARCH_REGISTRY, DATASET_REGISTRY, LOSS_REGISTRY, METRIC_REGISTRY, MODEL_REGISTRY  # noqa: F401

_initialized = set()


def get_root_logger(logger_name="basicsr", log_level=logging.INFO, log_file=None):
    """Rank-0 INFO logger, other ranks ERROR (reference: basicsr/utils/logger.py:156-195)."""
    logger = logging.getLogger(logger_name)
    if logger_name in _initialized:
        return logger
    handler = logging.StreamHandler()
    handler.setFormatter(logging.Formatter("%(asctime)s %(levelname)s: %(message)s"))
    logger.addHandler(handler)
    logger.propagate = False
    rank = int(os.environ.get("RANK", "0"))
    logger.setLevel(log_level if rank == 0 else logging.ERROR)
    if log_file is not None and rank == 0:
        fh = logging.FileHandler(log_file, "w")
        fh.setFormatter(logging.Formatter("%(asctime)s %(levelname)s: %(message)s"))
        logger.addHandler(fh)
    _initialized.add(logger_name)
    return logger


def scandir(dir_path, suffix=None, recursive=False, full_path=False):
    """Yield file names under dir_path (reference: basicsr/utils/misc.py:60-99)."""
    root = dir_path

    def _scan(path):
        for entry in sorted(os.scandir(path), key=lambda e: e.name):
            if entry.is_file() and not entry.name.startswith("."):
                name = entry.path if full_path else os.path.relpath(entry.path, root)
                if suffix is None or name.endswith(suffix):
                    yield name
            elif entry.is_dir() and recursive:
                yield from _scan(entry.path)

    return _scan(dir_path)
