"""MemoryMappedVectors / ParallelCollection (parallel_search.py:427-952): on-disk format compatibility (CPU) and
search parity (GPU)."""
import importlib.util
import os
import struct
import sys

import numpy as np
import pytest

from fastpyvectordb_b200.mmap_store import MemoryMappedVectors, ParallelCollection
from oracle import oracle as O

REF = "/root/reference/parallel_search.py"


def _fill(path, n=300, d=24, seed=1):
    rng = np.random.default_rng(seed)
    vecs = rng.standard_normal((n, d)).astype(np.float32)
    st = MemoryMappedVectors(str(path), dimensions=d)
    st.create(n_vectors=n + 50)
    st.append_batch(vecs[:200], [f"v{i}" for i in range(200)], [{"i": i} for i in range(200)])
    for i in range(200, n):
        st.append(vecs[i], f"v{i}")
    st.close()
    return vecs


def test_file_format_matches_the_reference_layout(tmp_path):
    vecs = _fill(tmp_path)
    raw = open(tmp_path / "vectors.mmap", "rb").read()
    assert raw[:8] == b"PYVEC001"
    assert struct.unpack("<III", raw[8:20]) == (1, 300, 24)            # version, n_vectors, dimensions
    assert raw[20:64] == b"\x00" * 44
    payload = np.frombuffer(raw, np.float32, count=300 * 24, offset=64).reshape(300, 24)
    assert np.array_equal(payload, vecs)
    st = MemoryMappedVectors(str(tmp_path))
    assert len(st) == 300 and st.dimensions == 24
    assert np.array_equal(st.get(7), vecs[7]) and np.array_equal(st.get_range(10, 20), vecs[10:20])
    assert np.array_equal(st.get_batch([3, 1, 299]), vecs[[3, 1, 299]]) and np.array_equal(st.get_all(), vecs)
    with pytest.raises(IndexError):
        st.get(300)
    with pytest.raises(ValueError):
        st.append_batch(np.zeros((100, 24), np.float32))               # beyond the pre-allocated capacity


@pytest.mark.skipif(not os.path.exists(REF), reason="reference tree only exists in the build container")
def test_files_are_interchangeable_with_the_reference(tmp_path):
    spec = importlib.util.spec_from_file_location("ref_ps_mmap", REF)
    ref = importlib.util.module_from_spec(spec)
    sys.modules["ref_ps_mmap"] = ref
    spec.loader.exec_module(ref)
    vecs = _fill(tmp_path / "ours")
    theirs = ref.MemoryMappedVectors(str(tmp_path / "ours"))            # reference reads our file
    assert len(theirs) == 300 and np.array_equal(np.asarray(theirs.get_all()), vecs)
    w = ref.MemoryMappedVectors(str(tmp_path / "theirs"), dimensions=24)    # we read the reference's file
    w.create(400)
    w.append_batch(vecs, [f"v{i}" for i in range(300)])
    w.close()
    ours = MemoryMappedVectors(str(tmp_path / "theirs"))
    assert len(ours) == 300 and np.array_equal(ours.get_all(), vecs)


@pytest.mark.gpu
def test_search_resident_and_streamed_agree_with_oracle(tmp_path):
    vecs = _fill(tmp_path, n=5000, d=32)
    st = MemoryMappedVectors(str(tmp_path))
    q = np.random.default_rng(999).standard_normal(32).astype(np.float32)
    for metric in ("cosine", "l2", "ip"):
        ref = O.distances_single(q, vecs, metric)
        res = st.search_parallel(q, k=10, metric=metric)
        O.check_topk(ref, [r.index for r in res], [r.distance for r in res], 10, squared_near_zero=(metric == "l2"))
        streamed = MemoryMappedVectors(str(tmp_path))
        streamed.RESIDENT_BYTES, streamed.CHUNK_ROWS = 0, 700          # force the chunk-stream + merge route
        res2 = streamed.search_parallel(q, k=10, metric=metric)
        assert [(r.index, r.distance) for r in res2] == [(r.index, r.distance) for r in res]
        st2 = streamed.last_stream_stats                                # double-buffered pinned streaming, 8 chunks of 700 rows
        assert st2["chunks"] == 8 and st2["bytes"] == 5000 * 32 * 4 and st2["gb_per_s"] > 0


@pytest.mark.gpu
def test_parallel_collection_surface():
    rng = np.random.default_rng(5)
    vecs = rng.standard_normal((4000, 64)).astype(np.float32)
    col = ParallelCollection("c", 64)
    col.insert_batch(vecs, [f"id{i}" for i in range(4000)], [{"even": i % 2 == 0} for i in range(4000)])
    assert col.count() == 4000
    q = vecs[17] + 0.01 * rng.standard_normal(64).astype(np.float32)
    ref = O.distances_single(q, vecs, "cosine")
    res = col.search_parallel(q, k=5)
    assert res[0].id == "id17" and res[0].metadata == {"even": False}
    O.check_topk(ref, [r.index for r in res], [r.distance for r in res], 5)
    res = col.search_parallel(q, k=5, filter_fn=lambda m: m.get("even", False))
    assert all(r.index % 2 == 0 for r in res)
    O.check_topk(ref, [r.index for r in res], [r.distance for r in res], 5, valid=(np.arange(4000) % 2 == 0))
    hyb = col.search_hybrid(q, k=5, hnsw_candidates=200)
    assert hyb[0].id == "id17" and len(hyb) == 5
    assert all(abs(r.distance - ref[r.index]) < 1e-5 for r in hyb)      # re-ranked distances are exact
    with pytest.raises(NotImplementedError):
        col.search_hnsw(q)
