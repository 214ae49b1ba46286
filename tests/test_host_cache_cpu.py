"""Host-side cache logic (dcpt_b200/params.py): no GPU, no compute calls."""
import torch

from dcpt_b200.params import LRUCache, PackedCacheKey


def test_lru_evicts_oldest_but_never_busy_entries():
    class Slot:
        def __init__(self, busy):
            self.busy = busy
    c = LRUCache(cap=2, can_evict=lambda slots: not any(s.busy for s in slots))
    c.put("a", [Slot(True)])
    c.put("b", [Slot(False)])
    c.put("c", [Slot(False)])          # over capacity: "a" is oldest but busy -> "b" goes
    assert "a" in c and "c" in c and "b" not in c
    c.get("a")                         # refresh "a"
    c.d["a"][0].busy = False
    c.put("d", [])
    assert "c" not in c and "a" in c and "d" in c and len(c) == 2
    assert c.setdefault("a", list) is c.get("a")


def test_packed_cache_key_tracks_versions_and_storage():
    p = [torch.nn.Parameter(torch.zeros(4)), torch.nn.Parameter(torch.ones(3))]
    d = [q.detach() for q in p]
    k = PackedCacheKey()
    assert k.stale(d)                  # first use: pack
    assert not k.stale(d)
    with torch.no_grad():
        p[1].add_(1.0)                 # an in-place torch write bumps the shared version counter
    assert k.stale(d) and not k.stale(d)
    torch.autograd.graph.increment_version(p)   # what FusedAdam.step does after its raw-pointer update
    assert k.stale(d) and not k.stale(d)
    p[0].data = torch.zeros(4)         # re-allocated storage (.to(), load with assign)
    assert k.stale([q.detach() for q in p])
    k.invalidate()
    assert k.stale([q.detach() for q in p])


def test_cuda_prefetcher_has_no_cpu_path():
    from dcpt_b200.prefetch import CUDAPrefetcher
    import pytest as _pytest
    with _pytest.raises(Exception, match="no CPU path"):
        CUDAPrefetcher([{"lq": torch.zeros(1)}], {"num_gpu": 0})


def test_frozen_weights_skips_the_fingerprint_after_the_first_check():
    """dcpt_b200.params.frozen_weights (the tile loop of dcpt_b200/tiling.py): per cache object the fingerprint is compared once
    inside the block, again outside it."""
    from dcpt_b200 import params as P
    a, b = P.PackedCacheKey(), P.PackedCacheKey()
    assert not P._fingerprint_already_checked(a) and not P._fingerprint_already_checked(a)      # outside: always check
    with P.frozen_weights():
        assert not P._fingerprint_already_checked(a)          # first forward of engine a: check
        assert P._fingerprint_already_checked(a)              # later ones: skip
        assert not P._fingerprint_already_checked(b)          # another engine checks for itself
        with P.frozen_weights():
            assert P._fingerprint_already_checked(a) and P._fingerprint_already_checked(b)
        assert P._fingerprint_already_checked(a)
    assert not P._fingerprint_already_checked(a)
    with P.frozen_weights():
        assert not P._fingerprint_already_checked(a)          # a new block starts over


def test_training_pass_marker_is_per_thread_and_nests():
    """dcpt_b200.params.training_pass: the marker that keeps the no-grad weight fingerprint (a host sync) off the forward /
    backward of gradient-needing calls - autograd runs backward on its own thread, so the marker must be thread-local."""
    import threading
    from dcpt_b200 import params as P
    assert not P.in_training_pass()
    seen = {}
    with P.training_pass():
        assert P.in_training_pass()
        with P.training_pass():
            assert P.in_training_pass()
        assert P.in_training_pass()
        t = threading.Thread(target=lambda: seen.setdefault("other", P.in_training_pass()))
        t.start()
        t.join()
    assert not P.in_training_pass() and seen["other"] is False
    # and a CPU parameter list never triggers the fingerprint path at all (no CUDA, no sync): stale() only looks at versions
    k = P.PackedCacheKey()
    ps = [torch.nn.Parameter(torch.zeros(3))]
    with torch.no_grad():
        assert k.stale(ps) and not k.stale(ps)
        ps[0].add_(1.0)
        assert k.stale(ps)
