"""SearchPipeline (double-buffered H2D / search / D2H) returns exactly what the synchronous array API returns."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_pipeline_matches_search_arrays():
    import fastpyvectordb_b200 as fpv
    rng = np.random.default_rng(3)
    n, d, k = 30000, 128, 10
    db = rng.standard_normal((n, d)).astype(np.float32)
    batches = [rng.standard_normal((1 if i == 3 else 40 + 70 * i, d)).astype(np.float32) for i in range(6)]
    eng = fpv.ParallelSearchEngine()
    for metric in ("l2", "cosine"):
        pipe = fpv.SearchPipeline(eng, db, k=k, metric=metric)
        tickets, got = [], {}
        for i, b in enumerate(batches):
            tickets.append(pipe.submit(b))
            if i >= 1:                                               # read batch i-1 while batch i is in flight
                got[i - 1] = tuple(a.copy() for a in pipe.result(tickets[i - 1]))
        got[len(batches) - 1] = tuple(a.copy() for a in pipe.result(tickets[-1]))
        for i, b in enumerate(batches):
            idx, dist = eng.search_arrays(b, db, k, metric)
            assert np.array_equal(got[i][0], idx) and np.array_equal(got[i][1], dist), (metric, i)
        with pytest.raises(ValueError):
            pipe.result(tickets[0])                                  # its slot has been reused


def test_pipeline_accepts_pinned_tensors_and_custom_search():
    import torch
    import fastpyvectordb_b200 as fpv
    rng = np.random.default_rng(4)
    db = rng.standard_normal((20000, 64)).astype(np.float32)
    qs = torch.from_numpy(rng.standard_normal((300, 64)).astype(np.float32)).pin_memory()
    eng = fpv.ParallelSearchEngine()
    index = eng.register(db)
    calls = []

    def search(qd):
        calls.append(qd.shape[0])
        return eng.search_tensors(qd, index, 5, "ip")

    pipe = fpv.SearchPipeline(eng, k=5, metric="ip", search_fn=search, depth=3)
    t = [pipe.submit(qs) for _ in range(3)]
    ref_i, ref_d = eng.search_arrays(qs.numpy(), db, 5, "ip")
    for ti in t:
        i, dd = pipe.result(ti)
        assert np.array_equal(i, ref_i) and np.array_equal(dd, ref_d)
    assert calls == [300, 300, 300]


def test_pipeline_with_row_filter():
    import fastpyvectordb_b200 as fpv
    rng = np.random.default_rng(6)
    db = rng.standard_normal((25000, 96)).astype(np.float32)
    mask = rng.random(25000) < 0.1
    eng = fpv.ParallelSearchEngine()
    pipe = fpv.SearchPipeline(eng, db, k=7, metric="l2", filter_mask=mask)
    batches = [rng.standard_normal((n, 96)).astype(np.float32) for n in (2, 150)]
    tickets = [pipe.submit(b) for b in batches]
    for b, t in zip(batches, tickets):
        idx, dist = pipe.result(t)
        ref_i, ref_d = eng.search_arrays(b, db, 7, "l2", filter_mask=mask)
        assert np.array_equal(idx, ref_i) and np.array_equal(dist, ref_d) and mask[idx].all()
