"""Host-side rules of the resident-database cache (engine._ResidentCache): reuse only when the host array provably has
not changed.  No GPU: the builder is injected."""
import warnings

import numpy as np

from fastpyvectordb_b200 import engine as E


class _Fake:
    def __init__(self, arr, device):
        self.device = device
        self.snapshot = arr.copy()


def test_checksum_sees_any_single_element_edit():
    a = np.random.default_rng(0).standard_normal((1000, 37)).astype(np.float32)
    base = E._checksum(a)
    for pos in [(0, 0), (999, 36), (500, 17), (123, 1)]:
        b = a.copy()
        b[pos] = np.nextafter(b[pos], np.float32(10))
        assert E._checksum(b) != base
    b = a.copy()
    b[10] = 0
    b[20:30] *= -1
    assert E._checksum(b) != base
    assert E._checksum(a.copy()) == base
    assert E._checksum(np.zeros((0, 4), np.float32)) == E._checksum(np.zeros((0, 4), np.float32))
    codes = np.arange(13, dtype=np.uint8)                            # tail bytes (size not a multiple of 8)
    c2 = codes.copy(); c2[12] ^= 1
    assert E._checksum(codes) != E._checksum(c2)


def test_cache_reuse_rules(monkeypatch):
    built = []
    cache = E._ResidentCache(build=lambda arr, dev: built.append(1) or _Fake(arr, dev))
    a = np.ones((100, 8), np.float32)
    x = cache.get(a, "dev")
    assert cache.get(a, "dev") is x and len(built) == 1              # unchanged -> reused
    a[50, 3] = 2.0                                                    # in-place edit -> rebuilt, and sees the new data
    y = cache.get(a, "dev")
    assert y is not x and y.snapshot[50, 3] == 2.0 and len(built) == 2
    assert cache.get(a, "other") is not y                             # different device
    # large writeable arrays are never trusted; read-only ones are reused by identity
    monkeypatch.setattr(E, "FULL_CHECK_BYTES", 64)
    big = np.ones((100, 8), np.float32)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        p, q = cache.get(big, "dev"), cache.get(big, "dev")
    assert p is not q and any("register" in str(m.message) for m in w)
    big.setflags(write=False)
    assert cache.get(big, "dev") is cache.get(big, "dev")
