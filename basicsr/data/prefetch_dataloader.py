"""``CUDAPrefetcher`` under the reference's import path (basicsr/data/prefetch_dataloader.py:83-125); the implementation lives
in dcpt_b200/prefetch.py.  ``CPUPrefetcher`` / ``PrefetchDataLoader`` (:9-80) are plain host-side iterators and are not mirrored."""
from dcpt_b200.prefetch import CUDAPrefetcher  # noqa: F401

__all__ = ["CUDAPrefetcher"]
