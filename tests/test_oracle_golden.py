"""Pin the oracle (oracle/oracle.py) against outputs of the unmodified reference
(tests/golden/reference_outputs.npz, produced by tests/golden/make_golden.py)."""
import numpy as np
import pytest

import inputs as gi
from oracle import oracle as O


def _same_float(a, b, rtol=2e-6, atol=2e-6):
    # same NumPy/BLAS calls in the same order: expected to be bit-identical on the same host, the
    # tolerance only absorbs BLAS kernel / thread-count differences between hosts.
    np.testing.assert_allclose(np.asarray(a, np.float64), np.asarray(b, np.float64), rtol=rtol, atol=atol)


@pytest.mark.parametrize("case", gi.FLOAT_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("metric", ["cosine", "l2", "ip"])
def test_float_path(golden, case, metric):
    db, qs, mask = gi.float_inputs(case)
    tag = f"{case['name']}/{metric}"
    dist = np.stack([O.distances_single(q, db, metric) for q in qs])
    _same_float(dist, golden[tag + "/dist_single"])
    chunk = np.stack([O.distances_chunk(q, db, 7, metric) for q in qs[:2]])
    atol = 4e-3 if metric == "l2" else 2e-6   # explicit-diff form is exact at 0, the GEMV form is not
    _same_float(chunk, golden[tag + "/dist_chunk"], atol=atol)
    assert np.array_equal(chunk[..., 0], golden[tag + "/dist_chunk"][..., 0])

    k = case["k"]
    bi, bd = O.search_batch_parallel(qs, db, k, metric)
    _same_float(bd, golden[tag + "/batch_dist"])
    dm = O.distances_batch(qs, db, metric)
    for qi in range(len(qs)):
        # the reference's ids must be a valid tie-aware answer w.r.t. oracle distances and vice versa
        O.check_topk(dm[qi], golden[tag + "/batch_idx"][qi], golden[tag + "/batch_dist"][qi], k,
                     squared_near_zero=(metric == "l2"))
        O.check_topk(golden[tag + "/dist_single"][qi], *O.search_parallel(qs[qi], db, k, metric), k,
                     squared_near_zero=(metric == "l2"))
        mi, md = O.search_parallel(qs[qi], db, k, metric, filter_mask=mask)
        _same_float(md, golden[tag + "/masked_dist"][qi])
        O.check_topk(golden[tag + "/dist_single"][qi], mi, md, k, valid=mask,
                     squared_near_zero=(metric == "l2"))
        O.check_topk(dist[qi], golden[tag + "/masked_idx"][qi], golden[tag + "/masked_dist"][qi], k,
                     valid=mask, squared_near_zero=(metric == "l2"))
        ci, cd = O.search_chunked_parallel(qs[qi], db, k, metric, chunk_size=case["chunk"])
        _same_float(cd, golden[tag + "/chunked_dist"][qi], rtol=2e-5, atol=2e-5)
        O.check_topk(dist[qi], ci, cd, k, rtol=1e-4, squared_near_zero=(metric == "l2"))
        # canonical rule agrees with the reference on the distance multiset
        cidx, cdist = O.canonical_topk(dist[qi], k)
        _same_float(cdist, golden[tag + "/single_dist"][qi])


def test_merge_top_k(golden):
    blocks = gi.merge_inputs()
    for k in (5, 100):
        got = O.merge_top_k(blocks, k)
        ref = golden[f"merge/k{k}"]
        assert got.shape == ref.shape
        assert np.array_equal(got[:, 1], ref[:, 1])          # distance multiset, ascending
        allrows = np.vstack(blocks)
        for row in got:                                      # every (id, dist) pair is genuine
            assert allrows[int(row[0]), 1] == row[1]


@pytest.mark.parametrize("case", gi.SQ_CASES, ids=lambda c: c["name"])
def test_scalar_quantizer(golden, case):
    train, db, qs = gi.sq_inputs(case)
    tag = case["name"]
    lo, hi, scale = O.sq_train(train)
    assert np.array_equal(lo, golden[tag + "/min"]) and np.array_equal(hi, golden[tag + "/max"])
    assert np.array_equal(scale, golden[tag + "/scale"])
    codes = O.sq_encode(db, lo, scale)
    assert codes.dtype == np.uint8 and np.array_equal(codes, golden[tag + "/codes"])
    qcodes = np.stack([O.sq_encode(q, lo, scale)[0] for q in qs])
    assert np.array_equal(qcodes, golden[tag + "/qcodes"])
    assert np.array_equal(O.sq_decode(codes[:16], lo, scale), golden[tag + "/decode"])
    _same_float(np.stack([O.sq_distances_l2(q, codes, lo, scale) for q in qs]), golden[tag + "/l2"])
    _same_float(np.stack([O.sq_distances_dot(q, codes, lo, scale) for q in qs]), golden[tag + "/dot"])
    _same_float(np.stack([O.sq_distances_cosine(q, codes, lo, scale) for q in qs]), golden[tag + "/cosine"])


@pytest.mark.parametrize("case", gi.BQ_CASES, ids=lambda c: c["name"])
def test_binary_quantizer(golden, case):
    train, db, qs = gi.bq_inputs(case)
    tag = case["name"]
    if case["train"]:
        thr = O.bq_train(train, case["median"], case.get("threshold", 0.0))
        assert np.array_equal(thr, golden[tag + "/thresholds"])
        dims = case["d"]
    else:
        thr = np.full(case["d"], case.get("threshold", 0.0))
        dims = case["dims_attr"]
    codes = O.bq_encode(db, thr)
    assert np.array_equal(codes, golden[tag + "/codes"])
    qbits = np.stack([O.bq_encode(q, thr)[0] for q in qs])
    assert np.array_equal(qbits, golden[tag + "/qbits"])
    ham = np.stack([O.bq_hamming(qb, codes, dims) for qb in qbits])
    assert ham.dtype == np.float32 and np.array_equal(ham, golden[tag + "/hamming"])
    assert np.array_equal(ham, np.stack([O.bq_hamming_chunked(qb, codes, dims, rows=64) for qb in qbits]))
    for qi in range(len(qs)):
        O.check_topk(ham[qi], golden[tag + "/search_idx"][qi], golden[tag + "/search_dist"][qi],
                     case["k"], integer=False, rtol=0.0)
        cidx, cdist = O.canonical_topk(ham[qi], case["k"])
        O.check_topk(golden[tag + "/hamming"][qi], cidx, cdist, case["k"], integer=True)


@pytest.mark.parametrize("case", gi.PQ_CASES, ids=lambda c: c["name"])
def test_product_quantizer(golden, case):
    cb, db, qs = gi.pq_inputs(case)
    tag = case["name"]
    codes = O.pq_encode(db, cb)
    assert np.array_equal(codes, golden[tag + "/codes"])
    lut = np.stack([O.pq_lookup_table(q, cb) for q in qs])
    assert np.array_equal(lut, golden[tag + "/lut"])
    dist = np.stack([O.pq_distances_with_table(t, codes) for t in lut])
    assert np.array_equal(dist, golden[tag + "/dist"])          # sequential-in-m fp32: bit exact
    for qi in range(len(qs)):
        O.check_topk(dist[qi], golden[tag + "/search_idx"][qi], golden[tag + "/search_dist"][qi],
                     case["k"], rtol=0.0)


def test_kmeans_restatement(golden):
    data = gi.kmeans_inputs()
    np.random.seed(gi.KMEANS_SEED)
    got = O.pq_kmeans(data, gi.KMEANS_K, gi.KMEANS_ITERS)
    assert np.array_equal(got, golden["kmeans/centroids"])


def test_check_topk_rejects_wrong_answers():
    d = np.array([0.5, 0.1, 0.1, 0.9, 0.1, 0.3], np.float32)
    idx, dist = O.canonical_topk(d, 3)
    assert idx.tolist() == [1, 2, 4]
    O.check_topk(d, idx, dist, 3, integer=True)
    with pytest.raises(AssertionError):
        O.check_topk(d, np.array([1, 2, 5]), d[[1, 2, 5]], 3, integer=True)      # misses a better row
    with pytest.raises(AssertionError):
        O.check_topk(d, np.array([1, 2, 2]), d[[1, 2, 2]], 3, integer=True)      # duplicate
    with pytest.raises(AssertionError):
        O.check_topk(d, np.array([2, 1, 4]), d[[2, 1, 4]], 3, integer=True)      # tie order
    with pytest.raises(AssertionError):
        O.check_topk(d, idx, dist + 1e-3, 3)                                     # distance drift
    idx2, dist2 = O.canonical_topk(d, 2)
    assert idx2.tolist() == [1, 2]
    with pytest.raises(AssertionError):
        O.check_topk(d, np.array([1, 4]), d[[1, 4]], 2, integer=True)            # not lowest-index tie members
    O.check_topk(d, np.array([1, 4]), d[[1, 4]], 2)                              # fine for a float metric


def test_oracle_brute_force_distances_reproduce_the_reference_method(golden):
    """oracle.brute_force_distances is pinned against outputs of the UNMODIFIED vectordb_optimized.Collection
    .brute_force_search (run through tests/golden/hnswlib_shim.py by make_golden.py): recomputing the top-k from the
    oracle's distances gives the reference's ids and scores (a25 rested on an unpinned restatement)."""
    import inputs as gi
    from fastpyvectordb_b200.collection import Filter
    for case in gi.BRUTE_CASES:
        db, qs, ids, meta = gi.brute_inputs(case)
        for metric in ("cosine", "l2", "ip"):
            for fname in gi.BRUTE_FILTERS:
                flt = gi.brute_filter(Filter, fname)
                f = Filter.from_dict(flt) if isinstance(flt, dict) else flt
                valid = np.array([f.evaluate(m) for m in meta]) if f is not None else np.ones(len(meta), bool)
                tag = f"{case['name']}/{metric}/{fname}"
                for qi, q in enumerate(qs):
                    d = O.brute_force_distances(q, db, metric).astype(np.float64)
                    d = np.where(valid, d, np.inf)
                    cnt = int(golden[tag + "/count"][qi])
                    assert cnt == min(case["k"], int(valid.sum()))
                    want_s, want_i = golden[tag + "/score"][qi][:cnt], golden[tag + "/idx"][qi][:cnt]
                    # same arithmetic as the reference: identical scores (BLAS blocking may differ in the last ulp)
                    assert np.allclose(d[want_i], want_s, rtol=2e-6, atol=2e-7)
                    order = np.lexsort((np.arange(len(d)), d))[:cnt]
                    assert np.allclose(np.sort(d[order]), np.sort(want_s), rtol=2e-6, atol=2e-7)


def test_device_kmeans_host_logic_reproduces_reference_centroids(golden):
    """The host logic of pq_train (seeding, centroid update) is plain torch: run on CPU tensors -- with the assignment step,
    which the product does with its CUDA argmin kernel, restated here in torch -- it must land on the reference's centroids
    for the same np.random seed: the inverse-CDF seeding consumes the global np.random stream exactly like
    np.random.choice(n, p=...) does (quantization.py:486-494).  Without the stand-in the trainer refuses CPU tensors."""
    import inputs as gi
    import pytest
    import torch
    from fastpyvectordb_b200.pq_train import _kmeans

    def torch_assign(v, cent):                                   # argmin_k sum_d (x - c)^2, first minimum (quantization.py:497-498)
        m, k, dsub = cent.shape
        x = v.view(-1, m, dsub)
        return ((x[:, :, None, :] - cent[None]) ** 2).sum(-1).argmin(-1)

    data = gi.kmeans_inputs()
    with pytest.raises(RuntimeError):
        _kmeans(torch.from_numpy(data), gi.KMEANS_K, 1)
    np.random.seed(gi.KMEANS_SEED)
    cent = _kmeans(torch.from_numpy(data), gi.KMEANS_K, gi.KMEANS_ITERS, assign=torch_assign).numpy()
    assert np.abs(cent - golden["kmeans/centroids"]).max() < 1e-5
