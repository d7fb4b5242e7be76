"""GPU parity tests: the CUDA path (through the drop-in Python API -> C-ABI) against
 (a) the committed outputs of the unmodified reference (tests/golden/reference_outputs.npz) and
 (b) the oracle (oracle/oracle.py) run live on seeded inputs.

Bar (BASELINE.json north_star): integer / byte / index work bit-exact, ties broken by lowest index;
fp32 distances within 1e-5 relative (|d_gpu - d_ref| <= 1e-5 * max(|d_ref|, 1)); top-k id sets tie-aware identical.
"""
import numpy as np
import pytest
import torch

import inputs as gi
from oracle import oracle as O

pytestmark = pytest.mark.gpu

RTOL = 1e-5


@pytest.fixture(scope="module")
def fpv():
    import fastpyvectordb_b200 as m
    return m


@pytest.fixture(scope="module")
def engine(fpv):
    return fpv.ParallelSearchEngine()


def _close(got, ref, rtol=RTOL):
    got = np.asarray(got, np.float64)
    ref = np.asarray(ref, np.float64)
    lim = rtol * np.maximum(np.abs(ref), 1.0)
    bad = np.abs(got - ref) > lim
    assert not bad.any(), f"max err {np.abs(got - ref).max():.3e} (limit {lim.min():.1e}); {int(bad.sum())} bad"


# ------------------------------------------------------------------------------------------------ float path
@pytest.mark.parametrize("case", gi.FLOAT_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("metric", ["cosine", "l2", "ip"])
def test_float_search_against_reference_outputs(engine, golden, case, metric):
    db, qs, mask = gi.float_inputs(case)
    tag = f"{case['name']}/{metric}"
    ref_all = golden[tag + "/dist_single"]          # reference distances of every row, per query
    k = case["k"]
    sq0 = metric == "l2"
    # single-query API, object results
    for qi, q in enumerate(qs):
        res = engine.search_parallel(q, db, k=k, metric=metric)
        assert len(res) == min(k, len(db)) and type(res[0]).__name__ == "ParallelSearchResult"
        O.check_topk(ref_all[qi], [r.index for r in res], [r.distance for r in res], k, squared_near_zero=sq0)
        res = engine.search_parallel(q, db, k=k, metric=metric, filter_mask=mask)
        assert len(res) == min(k, int(mask.sum()))
        O.check_topk(ref_all[qi], [r.index for r in res], [r.distance for r in res], k, valid=mask, squared_near_zero=sq0)
        res = engine.search_chunked_parallel(q, db, k=k, metric=metric)
        O.check_topk(ref_all[qi], [r.index for r in res], [r.distance for r in res], k, squared_near_zero=sq0)
    # batch API
    res = engine.search_batch_parallel(qs, db, k=k, metric=metric)
    assert len(res) == len(qs)
    for qi, row in enumerate(res):
        O.check_topk(ref_all[qi], [r.index for r in row], [r.distance for r in row], k, squared_near_zero=sq0)
        # the reference's own batch output must be an acceptable answer by the same rule (sanity of the checker)
        O.check_topk(ref_all[qi], golden[tag + "/batch_idx"][qi], golden[tag + "/batch_dist"][qi], k, squared_near_zero=sq0)
    # 1-D query to the batch API is one query (parallel_search.py:262-263)
    one = engine.search_batch_parallel(qs[0], db, k=3, metric=metric)
    assert len(one) == 1 and len(one[0]) == min(3, len(db))


@pytest.mark.parametrize("n,d,q,k", [(20000, 384, 5, 10), (5000, 768, 33, 100), (3001, 100, 3, 1000), (257, 7, 9, 1024),
                                      (4096, 128, 64, 10), (100, 16, 2, 2000)])
@pytest.mark.parametrize("metric", ["cosine", "l2", "ip"])
def test_float_search_against_oracle(engine, n, d, q, k, metric):
    rng = np.random.default_rng(42)
    db = rng.standard_normal((n, d)).astype(np.float32)
    if metric == "cosine":
        db /= np.linalg.norm(db, axis=1, keepdims=True)
    qs = np.random.default_rng(999).standard_normal((q, d)).astype(np.float32)
    ref = O.distances_batch(qs, db, metric)
    idx, dist = engine.search_arrays(qs, db, k=k, metric=metric)
    assert idx.shape == (q, min(k, n)) and idx.dtype == np.int64 and dist.dtype == np.float32
    for qi in range(q):
        O.check_topk(ref[qi], idx[qi], dist[qi], k, squared_near_zero=(metric == "l2"))
    # results do not depend on how the batch is split (determinism / tie rule)
    idx1, dist1 = engine.search_arrays(qs[:1], db, k=k, metric=metric)
    assert np.array_equal(idx1[0], idx[0]) and np.array_equal(dist1[0], dist[0])


def test_float_edge_cases(engine):
    rng = np.random.default_rng(3)
    db = rng.standard_normal((50, 12)).astype(np.float32)
    q = rng.standard_normal(12).astype(np.float32)
    assert engine.search_parallel(q, np.zeros((0, 12), np.float32)) == []                       # empty database -> []
    assert engine.search_parallel(q, db, k=5, filter_mask=np.zeros(50, bool)) == []             # nothing permitted
    res = engine.search_parallel(q, db, k=500)                                                   # k >= N: all rows, sorted
    assert len(res) == 50 and sorted(r.index for r in res) == list(range(50))
    assert all(res[i].distance <= res[i + 1].distance for i in range(49))
    res = engine.search_parallel(q, db[:1], k=10)                                                # N == 1
    assert len(res) == 1 and res[0].index == 0
    with pytest.raises(ValueError):
        engine.search_parallel(np.zeros(5, np.float32), db)                                      # dimension mismatch
    # duplicates tie on distance -> lowest index first
    dup = np.repeat(db[:1], 40, axis=0)
    res = engine.search_parallel(q, dup, k=7, metric="ip")
    assert [r.index for r in res] == list(range(7))
    # unknown metric strings mean inner product, like the reference (parallel_search.py:96, 133)
    a = engine.search_parallel(q, db, k=5, metric="dot")
    b = engine.search_parallel(q, db, k=5, metric="ip")
    assert [r.index for r in a] == [r.index for r in b]
    # torch inputs stay on the device
    dist, idx, cnt = engine.search_tensors(torch.from_numpy(q).cuda(), torch.from_numpy(db).cuda(), k=5)
    assert dist.is_cuda and idx.dtype == torch.int64 and int(cnt[0]) == 5
    # an explicit resident handle gives the same answer as the host array
    handle = engine.register(db)
    c = engine.search_parallel(q, handle, k=5, metric="ip")
    assert [r.index for r in c] == [r.index for r in b]


def test_module_level_helpers_against_reference_outputs(golden):
    """_compute_distances_vectorized / _compute_distances_chunk / _merge_top_k (parallel_search.py:72-156)."""
    from fastpyvectordb_b200 import parallel_search as ps
    case = gi.FLOAT_CASES[0]
    db, qs, _mask = gi.float_inputs(case)
    for metric in ("cosine", "l2", "ip"):
        tag = f"{case['name']}/{metric}"
        d = ps._compute_distances_vectorized(qs[0], db, metric)
        assert d.shape == (len(db),) and d.dtype == np.float32
        ref = golden[tag + "/dist_single"][0]
        if metric == "l2":      # the GEMV form is noisy at true distance 0 (duplicates): compare squared there
            small = ref < 1e-2
            _close(d[~small], ref[~small])
            assert not small.any() or np.abs(d[small] ** 2 - ref[small] ** 2).max() < 1e-5
        else:
            _close(d, ref)
        ch = ps._compute_distances_chunk((qs[1], db, 7, metric))
        assert ch.shape == (len(db), 2) and ch.dtype == np.float64
        assert np.array_equal(ch[:, 0], golden[tag + "/dist_chunk"][1][:, 0])
        # the chunk form's L2 is the explicit-difference one (parallel_search.py:92-95): duplicates score exactly 0
        _close(ch[:, 1], golden[tag + "/dist_chunk"][1][:, 1])
        if metric == "l2":
            zero = golden[tag + "/dist_chunk"][1][:, 1] == 0.0
            assert np.array_equal(ch[:, 1] == 0.0, zero)
    blocks = gi.merge_inputs()
    for k in (5, 100):
        got = ps._merge_top_k(blocks, k)
        ref = golden[f"merge/k{k}"]
        assert got.shape == ref.shape and np.array_equal(got[:, 1], ref[:, 1])
        allrows = np.vstack(blocks)
        assert all(allrows[int(r[0]), 1] == r[1] for r in got)
        order = np.lexsort((got[:, 0], got[:, 1]))
        assert np.array_equal(order, np.arange(len(got)))           # (distance, index) order


def test_merge_and_rerank_kernels(fpv):
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(8)
    shards, q, k = 5, 7, 12
    dist = np.round(rng.random((shards, q, k)) * 16).astype(np.float32) / 16      # many ties
    dist.sort(axis=2)
    idx = rng.permutation(shards * q * k).reshape(shards, q, k).astype(np.int64)
    idx[2, :, 9:] = -1                                                            # a short shard
    od, oi, oc = ops.merge_topk(torch.from_numpy(dist).cuda(), torch.from_numpy(idx).cuda(), 20)
    od, oi, oc = od.cpu().numpy(), oi.cpu().numpy(), oc.cpu().numpy()
    for qi in range(q):
        d_all = dist[:, qi].reshape(-1)
        i_all = idx[:, qi].reshape(-1)
        keep = i_all >= 0
        order = np.lexsort((i_all[keep], d_all[keep]))[:20]
        assert np.array_equal(oi[qi], i_all[keep][order]) and np.array_equal(od[qi], d_all[keep][order])
        assert oc[qi] == 20
        # same multiset of distances as the reference's _merge_top_k restatement
        blocks = [np.column_stack([idx[s, qi][idx[s, qi] >= 0], dist[s, qi][idx[s, qi] >= 0]]) for s in range(shards)]
        assert np.array_equal(O.merge_top_k(blocks, 20)[:, 1], od[qi].astype(np.float64))
    # re-rank (parallel_search.py:919-934 pattern)
    db = rng.standard_normal((500, 48)).astype(np.float32)
    qs = rng.standard_normal((3, 48)).astype(np.float32)
    cand = np.stack([rng.choice(500, 100, replace=False) for _ in range(3)]).astype(np.int64)
    d, i, c = ops.rerank_f32(torch.from_numpy(qs).cuda(), torch.from_numpy(db).cuda(), torch.from_numpy(cand).cuda(), 10, "cosine")
    for qi in range(3):
        ri, rd = O.rerank_cosine(qs[qi], db, cand[qi], 10)
        full = O.distances_single(qs[qi], db, "cosine")
        valid = np.zeros(500, bool)
        valid[cand[qi]] = True
        O.check_topk(full, i[qi].cpu().numpy(), d[qi].cpu().numpy(), 10, valid=valid)
        _close(d[qi].cpu().numpy(), rd, rtol=2e-5)


# ------------------------------------------------------------------------------------------------ scalar quantizer
@pytest.mark.parametrize("case", gi.SQ_CASES, ids=lambda c: c["name"])
def test_scalar_quantizer_against_reference_outputs(fpv, golden, case):
    train, db, qs = gi.sq_inputs(case)
    tag = case["name"]
    sq = fpv.ScalarQuantizer().train(train)
    assert np.array_equal(sq.min_vals, golden[tag + "/min"]) and np.array_equal(sq.scale, golden[tag + "/scale"])
    codes = sq.encode(db)
    assert codes.dtype == np.uint8 and np.array_equal(codes, golden[tag + "/codes"])            # bit exact
    assert np.array_equal(np.stack([sq.encode_query(q) for q in qs]), golden[tag + "/qcodes"])
    assert np.array_equal(sq.decode(codes[:16]), golden[tag + "/decode"])
    for qi, q in enumerate(qs):
        _close(sq.distances_l2(q, codes), golden[tag + "/l2"][qi])
        _close(sq.distances_dot(q, codes), golden[tag + "/dot"][qi])
        _close(sq.distances_cosine(q, codes), golden[tag + "/cosine"][qi])
        for metric, key in (("l2", "/l2"), ("ip", "/dot"), ("cosine", "/cosine")):
            idx, dist = sq.search(q, codes, k=10, metric=metric)
            O.check_topk(golden[tag + key][qi], idx, dist, 10)
    with pytest.raises(ValueError):
        fpv.ScalarQuantizer().encode(db)


def test_scalar_quantizer_wide_rows(fpv):
    rng = np.random.default_rng(42)
    db = rng.standard_normal((3000, 1024)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    qs = np.random.default_rng(999).standard_normal((2, 1024)).astype(np.float32)
    qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    sq = fpv.ScalarQuantizer().train(db)
    lo, hi, scale = O.sq_train(db)
    codes = sq.encode(db)
    assert np.array_equal(codes, O.sq_encode(db, lo, scale))
    mask = rng.random(3000) < 0.25
    for q in qs:
        ref = O.sq_distances_l2(q, codes, lo, scale)
        _close(sq.distances_l2(q, codes), ref)
        idx, dist = sq.search(q, codes, k=100, metric="l2", filter_mask=mask)
        O.check_topk(ref, idx, dist, 100, valid=mask)
        _close(sq.distances_dot(q, codes), O.sq_distances_dot(q, codes, lo, scale))
        _close(sq.distances_cosine(q, codes), O.sq_distances_cosine(q, codes, lo, scale))


@pytest.mark.parametrize("d", [1024, 512, 208])
def test_scalar_quantizer_l2_large_scan_kernel(fpv, d):
    """>= 65536 rows take the shared-memory-staged L2 kernel (cp.async.bulk ring): same distances as the oracle's
    restatement of quantization.py:217-236, full distance row and fused top-k with a row filter, ragged last tile."""
    rng = np.random.default_rng(8)
    n = 70003
    codes = rng.integers(0, 256, size=(n, d), dtype=np.uint8)
    lo = (rng.random(d).astype(np.float32) - 0.5)
    scale = (rng.random(d).astype(np.float32) * 0.2 + 0.01)
    sq = fpv.ScalarQuantizer()
    sq.min_vals, sq.scale, sq.trained, sq.dimensions = lo, scale, True, d
    q = (lo + scale * rng.random(d).astype(np.float32)).astype(np.float32)
    codes[n - 1] = O.sq_encode(q[None, :], lo, scale)[0]                      # the last row of the ragged tile is the best
    ref = O.sq_distances_l2(q, codes, lo, scale)
    _close(sq.distances_l2(q, codes), ref)
    mask = rng.random(n) < 0.3
    mask[n - 1] = True
    idx, dist = sq.search(q, codes, k=100, metric="l2", filter_mask=mask)
    O.check_topk(ref, idx, dist, 100, valid=mask)
    assert idx[0] == n - 1 and dist[0] == 0.0


# ------------------------------------------------------------------------------------------------ binary quantizer
@pytest.mark.parametrize("case", gi.BQ_CASES, ids=lambda c: c["name"])
def test_binary_quantizer_against_reference_outputs(fpv, golden, case):
    train, db, qs = gi.bq_inputs(case)
    tag = case["name"]
    bq = fpv.BinaryQuantizer(threshold=case.get("threshold", 0.0))
    if case["train"]:
        bq.train(train, use_median=case["median"])
        assert np.array_equal(np.asarray(bq.thresholds), golden[tag + "/thresholds"])
    else:
        bq.dimensions = case["dims_attr"]
    codes = bq.encode(db)
    assert codes.dtype == np.uint8 and np.array_equal(codes, golden[tag + "/codes"])            # bit exact
    k = case["k"]
    for qi, q in enumerate(qs):
        qb = bq.encode_query(q)
        assert np.array_equal(qb, golden[tag + "/qbits"][qi])
        ham = bq.hamming_distances(qb, codes)
        assert ham.dtype == np.float32 and np.array_equal(ham, golden[tag + "/hamming"][qi])    # bit exact
        idx, dist = bq.search(q, codes, k=k)
        assert idx.dtype == np.int64 and dist.dtype == np.float32
        O.check_topk(golden[tag + "/hamming"][qi], idx, dist, k, integer=True)                  # lowest-index tie rule
        # the reference's own (arbitrary-order) answer has the same distance multiset
        assert np.array_equal(np.sort(dist), np.sort(golden[tag + "/search_dist"][qi]))


@pytest.mark.parametrize("d", [128, 256, 384, 512, 768, 1024, 2048, 4096, 40])
def test_hamming_all_code_widths(fpv, d):
    rng = np.random.default_rng(d)
    n = 4133                                                                 # ragged: not a multiple of 32
    db = rng.standard_normal((n, d)).astype(np.float32)
    bq = fpv.BinaryQuantizer().train(db[:1000])
    codes = bq.encode(db)
    thr = O.bq_train(db[:1000])
    assert np.array_equal(codes, O.bq_encode(db, thr))
    q = np.random.default_rng(999).standard_normal(d).astype(np.float32)
    ref = O.bq_hamming(O.bq_encode(q, thr)[0], codes, d)
    assert np.array_equal(bq.hamming_distances(bq.encode_query(q), codes), ref)
    mask = rng.random(n) < 0.3
    for k in (1, 100, 1024):
        idx, dist = bq.search(q, codes, k=k)
        O.check_topk(ref, idx, dist, k, integer=True)
        idx, dist = bq.search(q, codes, k=k, filter_mask=mask)
        O.check_topk(ref, idx, dist, k, integer=True, valid=mask)
    idx, dist = bq.search(q, codes, k=3000)                                  # k beyond the fused selector
    O.check_topk(ref, idx, dist, 3000, integer=True)


@pytest.mark.parametrize("nbytes,q", [(128, 7), (64, 4), (16, 2), (256, 3), (96, 5)])
def test_hamming_query_batches_share_one_pass(nbytes, q):
    """Batched queries (the kernel loads every code once per 4 queries) give exactly the per-query answers."""
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(nbytes)
    n = 6007
    codes = rng.integers(0, 256, (n, nbytes), dtype=np.uint8)
    qb = rng.integers(0, 256, (q, nbytes), dtype=np.uint8)
    mask = rng.random(n) < 0.4
    dc, dq = torch.from_numpy(codes).cuda(), torch.from_numpy(qb).cuda()
    words = ops.pack_mask(torch.from_numpy(mask).cuda())
    dist, idx, cnt, allrows = ops.hamming(dq, dc, 50, nbytes * 8 - 3, words, 0, want_all=True)
    for qi in range(q):
        ref = O.bq_hamming(qb[qi], codes, nbytes * 8 - 3)
        assert np.array_equal(allrows[qi].cpu().numpy(), ref)
        O.check_topk(ref, idx[qi].cpu().numpy(), dist[qi].cpu().numpy(), 50, integer=True, valid=mask)
        d1, i1, _, _ = ops.hamming(dq[qi:qi + 1].contiguous(), dc, 50, nbytes * 8 - 3, words, 0)
        assert torch.equal(i1[0], idx[qi]) and torch.equal(d1[0], dist[qi])


# ------------------------------------------------------------------------------------------------ product quantizer
@pytest.mark.parametrize("case", gi.PQ_CASES, ids=lambda c: c["name"])
def test_product_quantizer_against_reference_outputs(fpv, golden, case):
    cb, db, qs = gi.pq_inputs(case)
    tag = case["name"]
    pq = fpv.ProductQuantizer(case["d"], case["m"], case["kc"])
    with pytest.raises(ValueError):
        pq.encode(db)
    pq.codebooks, pq.trained = cb, True
    codes = pq.encode(db)
    assert codes.dtype == np.uint8 and np.array_equal(codes, golden[tag + "/codes"])            # bit exact
    k = case["k"]
    mask = np.random.default_rng(11).random(len(db)) < 0.25
    for qi, q in enumerate(qs):
        lut = pq.build_lookup_table(q)
        assert lut.shape == (case["m"], case["kc"]) and np.array_equal(lut, golden[tag + "/lut"][qi])   # bit exact
        dist_all = pq.distances_with_table(lut, codes)
        assert np.array_equal(dist_all, golden[tag + "/dist"][qi])                                 # bit exact
        for fast in (False, True):
            # exact-order kernel: bit-identical distances, lowest-index tie rule; rotated (conflict-free) kernel:
            # same rows up to fp32 rounding of the sum order
            pq.fast_search = fast
            idx, dist = pq.search(q, codes, k=k)
            O.check_topk(golden[tag + "/dist"][qi], idx, dist, k, integer=not fast, rtol=1e-5)
            idx, dist = pq.search(q, codes, k=k, filter_mask=mask)
            O.check_topk(golden[tag + "/dist"][qi], idx, dist, k, integer=not fast, rtol=1e-5, valid=mask)
    with pytest.raises(ValueError):
        fpv.ProductQuantizer(100, 48)


def test_product_quantizer_larger(fpv):
    rng = np.random.default_rng(7)
    cb = (rng.standard_normal((48, 256, 16)) / np.sqrt(768)).astype(np.float32)
    codes = np.random.default_rng(42).integers(0, 256, (50000, 48), dtype=np.uint8)
    pq = fpv.ProductQuantizer(768, 48, 256)
    pq.codebooks, pq.trained = cb, True
    mask = np.random.default_rng(11).random(50000) < 0.25
    for seed in (999, 1000):
        q = np.random.default_rng(seed).standard_normal(768).astype(np.float32)
        q /= np.linalg.norm(q)
        ref = O.pq_distances_with_table(O.pq_lookup_table(q, cb), codes)
        for k in (10, 100):
            for fast in (False, True):
                pq.fast_search = fast
                idx, dist = pq.search(q, codes, k=k, filter_mask=mask)
                O.check_topk(ref, idx, dist, k, integer=not fast, rtol=1e-5, valid=mask)
                idx, dist = pq.search(q, codes, k=k)
                O.check_topk(ref, idx, dist, k, integer=not fast, rtol=1e-5)
    # other subspace counts of the rotated kernel: 16 (one half block), 32, 64, 96 and a ragged row count
    for m_sub in (16, 32, 64, 96):
        dsub = 4
        cbm = (rng.standard_normal((m_sub, 256, dsub)) / np.sqrt(m_sub * dsub)).astype(np.float32)
        cm = rng.integers(0, 256, (7777, m_sub), dtype=np.uint8)
        pqm = fpv.ProductQuantizer(m_sub * dsub, m_sub, 256)
        pqm.codebooks, pqm.trained = cbm, True
        q = rng.standard_normal(m_sub * dsub).astype(np.float32)
        ref = O.pq_distances_with_table(O.pq_lookup_table(q, cbm), cm)
        idx, dist = pqm.search(q, cm, k=100)
        O.check_topk(ref, idx, dist, 100, rtol=1e-5)


def test_quantized_scan_then_exact_rerank_pipeline(fpv, engine):
    """BASELINE configs[3]: uint8-scalar / binary scan for 100 candidates, then exact fp32 re-rank of those rows
    (the search_hybrid pattern, parallel_search.py:919-934).  Every stage is checked against its oracle stage."""
    rng = np.random.default_rng(42)
    n, d = 20000, 256
    db = rng.standard_normal((n, d)).astype(np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    q = np.random.default_rng(999).standard_normal(d).astype(np.float32)
    q /= np.linalg.norm(q)
    exact = O.distances_single(q, db, "cosine")
    sq = fpv.ScalarQuantizer().train(db)
    codes = sq.encode(db)
    cand, cdist = sq.search(q, codes, k=100, metric="l2")
    lo, hi, scale = O.sq_train(db)
    O.check_topk(O.sq_distances_l2(q, O.sq_encode(db, lo, scale), lo, scale), cand, cdist, 100)
    idx, dist = engine.rerank(q, db, cand, k=10, metric="cosine")
    ri, rd = O.rerank_cosine(q, db, cand, 10)
    valid = np.zeros(n, bool)
    valid[cand] = True
    O.check_topk(exact, idx[0], dist[0], 10, valid=valid)
    _close(dist[0], rd, rtol=2e-5)
    true_top, _ = O.canonical_topk(exact, 10)
    assert len(set(idx[0]) & set(true_top)) >= 9                  # uint8 scan + re-rank recovers the exact top-10
    bq = fpv.BinaryQuantizer().train(db)
    bcand, _ = bq.search(q, bq.encode(db), k=100)
    idx, dist = engine.rerank(q, db, bcand, k=10, metric="cosine")
    valid = np.zeros(n, bool)
    valid[bcand] = True
    O.check_topk(exact, idx[0], dist[0], 10, valid=valid)


def test_pq_train_quality(fpv):
    rng = np.random.default_rng(0)
    centers = rng.standard_normal((16, 32)).astype(np.float32) * 3
    data = (centers[rng.integers(0, 16, 4000)] + 0.1 * rng.standard_normal((4000, 32))).astype(np.float32)
    np.random.seed(1)
    pq = fpv.ProductQuantizer(32, 4, 16).train(data, n_iter=10)
    assert pq.codebooks.shape == (4, 16, 8) and pq.trained
    codes = pq.encode(data)
    recon = np.concatenate([pq.codebooks[m][codes[:, m]] for m in range(4)], axis=1)
    err_gpu = float(((recon - data) ** 2).sum(axis=1).mean())
    # reference-algorithm restatement on the same data: quantisation error must be comparable
    np.random.seed(1)
    cb_ref = np.stack([O.pq_kmeans(data[:, m * 8:(m + 1) * 8], 16, 10) for m in range(4)])
    codes_ref = O.pq_encode(data, cb_ref)
    recon_ref = np.concatenate([cb_ref[m][codes_ref[:, m]] for m in range(4)], axis=1)
    err_ref = float(((recon_ref - data) ** 2).sum(axis=1).mean())
    assert err_gpu <= 1.5 * err_ref + 1e-3, (err_gpu, err_ref)


def test_gpu_kmeans_reproduces_the_reference_centroids(golden):
    """The device k-means draws the reference's random numbers (one randint + K-1 inverse-CDF uniforms from np.random,
    quantization.py:486-494) and evaluates the same Lloyd steps, so with the same seed it lands on the reference's
    centroids (committed golden ``kmeans/centroids``) up to fp32 summation order."""
    from fastpyvectordb_b200.pq_train import _kmeans
    data = gi.kmeans_inputs()
    np.random.seed(gi.KMEANS_SEED)
    cent = _kmeans(torch.from_numpy(data).cuda(), gi.KMEANS_K, gi.KMEANS_ITERS).cpu().numpy()
    ref = golden["kmeans/centroids"]
    assert cent.shape == ref.shape
    assert np.abs(cent - ref).max() < 1e-4, np.abs(cent - ref).max()


# ------------------------------------------------------------------------------------------------ residency / threads
def test_inplace_edit_of_host_database_is_seen(engine):
    """ADVICE r1 / VERDICT weak #9: the stateless API must answer for the caller's CURRENT data.  One element edited in
    place (at a position a sampled fingerprint would miss) must change the answer."""
    rng = np.random.default_rng(5)
    db = rng.standard_normal((20000, 64)).astype(np.float32)
    q = rng.standard_normal(64).astype(np.float32)
    first = engine.search_parallel(q, db, k=3, metric="l2")
    row = 12345
    assert row not in [r.index for r in first]
    db[row] = q                                                     # in-place: same object, same pointer, same shape
    again = engine.search_parallel(q, db, k=3, metric="l2")
    assert again[0].index == row and again[0].distance < 1e-3
    db[row, 7] += 0.5                                               # a single element
    third = engine.search_parallel(q, db, k=3, metric="l2")
    assert third[0].index == row and abs(third[0].distance - 0.5) < 1e-4
    # quantizer code matrices follow the same rule
    import fastpyvectordb_b200 as fpv
    bq = fpv.BinaryQuantizer(64)
    codes = bq.encode(db)
    idx0, _ = bq.search(q, codes, k=1)
    codes[777] = bq.encode_query(q)
    idx1, d1 = bq.search(q, codes, k=1)
    assert idx1[0] == 777 and d1[0] == 0.0 and idx0[0] != 777


def test_concurrent_searches_from_host_threads(engine):
    """pinned staging buffers are per thread and every op enqueues its launches under the device lock, so
    searches issued from several host threads return their own answers."""
    import threading
    rng = np.random.default_rng(8)
    db = rng.standard_normal((30000, 96)).astype(np.float32)
    handle = engine.register(db)
    batches = [rng.standard_normal((n, 96)).astype(np.float32) for n in (1, 3, 17, 64, 130, 2, 33, 256)]
    expect = [engine.search_arrays(b, handle, k=20, metric="l2") for b in batches]
    errors = []

    def work(t):
        try:
            for rep in range(6):
                j = (t + rep) % len(batches)
                idx, dist = engine.search_arrays(batches[j], handle, k=20, metric="l2")
                if not (np.array_equal(idx, expect[j][0]) and np.array_equal(dist, expect[j][1])):
                    errors.append((t, rep, j))
        except Exception as exc:                                    # pragma: no cover
            errors.append((t, repr(exc)))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(8)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:5]


def test_quantizer_large_k_respects_the_filter():
    """k > MAX_K with a filter must not return rejected rows; a short torch mask is an error."""
    import fastpyvectordb_b200 as fpv
    rng = np.random.default_rng(2)
    x = rng.standard_normal((3000, 32)).astype(np.float32)
    bq = fpv.BinaryQuantizer(32)
    codes = bq.encode(x)
    mask = rng.random(3000) < 0.5
    idx, dist = bq.search(x[0], codes, k=2000, filter_mask=mask)
    assert len(idx) == min(2000, int(mask.sum())) and mask[idx].all() and np.all(np.diff(dist) >= 0)
    with pytest.raises(ValueError):
        bq.search(x[0], codes, k=5, filter_mask=torch.ones(100, dtype=torch.bool))


# ------------------------------------------------------------------------------------------------ PQ two-pass filter
@pytest.mark.parametrize("case", ["nomask", "mask25", "sample_all_rejected", "k256"])
def test_pq_large_scan_two_pass_filter(case):
    """Scans of >= 1M rows take the two-pass form of fpv_pq_adc_packed_topk (selector kernel on a sample -> tau ->
    pure filter -> merge; csrc/fpv_pq.cu).  Same answer as the exact-order kernel up to the fp32 summation order, and
    an overflowing candidate list (here: a filter that rejects every sample row) falls back on the device."""
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(21)
    n, m = 1_200_000 + (7 if case == "k256" else 0), 48          # k256: with a last partial group of 32 rows
    codes = torch.from_numpy(rng.integers(0, 256, (n, m), dtype=np.uint8)).cuda()
    cb = torch.from_numpy((rng.standard_normal((m, 256, 16)) / np.sqrt(768)).astype(np.float32)).cuda()
    q = torch.from_numpy(rng.standard_normal((2, 768)).astype(np.float32)).cuda()
    lut = ops.pq_build_lut(cb, q)
    k = 256 if case == "k256" else 100
    mask = None
    if case == "mask25":
        mask = torch.from_numpy(rng.random(n) < 0.25).cuda()
    elif case == "sample_all_rejected":
        mask = torch.zeros(n, dtype=torch.bool, device="cuda")
        mask[600_000:] = True                                   # 600K permitted rows, none of them in the sample
    words = ops.pack_mask(mask) if mask is not None else None
    assert ops.pq_adc_packed_supported(2, n, m, 256, k)
    d1, i1, c1 = ops.pq_adc_packed(lut, ops.pq_pack(codes), k, words)
    d0, i0, c0, _ = ops.pq_adc(lut, codes, k, words)             # exact-order kernel (bit-identical to the reference)
    assert (c1 == k).all() and (c0 == k).all()
    assert torch.allclose(d1, d0, rtol=2e-6, atol=0)
    same = (i1 == i0).float().mean().item()
    assert same >= 0.95, same                                    # neighbours may swap where two sums differ by an ulp
    host_codes, host_lut = codes.cpu().numpy(), lut.cpu().numpy()
    valid = mask.cpu().numpy() if mask is not None else None
    for qi in range(2):
        ref = O.pq_distances_with_table(host_lut[qi], host_codes)
        O.check_topk(ref, i1[qi].cpu().numpy(), d1[qi].cpu().numpy(), k, valid=valid, rtol=1e-5)


@pytest.mark.parametrize("case", ["q4", "q2_mask25", "q7", "q5_kc200_m32", "q4_flat_table", "q3_m96_k10"])
def test_pq_query_batches_share_one_pass(case):
    """Q >= 2 large scans run ONE pass per group of four queries over fixed-point tables (pq_adc_quad_kernel: four
    u16 entries per 64-bit lookup) and re-score the survivors with the one-query kernel's fp32 arithmetic: the answer
    must be bit-identical to scanning once per query, whatever the batch size, table shape or bitmask; a table the
    fixed-point form cannot bound (all entries equal) falls back on the device."""
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(5)
    nq = int(case[1])
    n, m, kc, k = 1_100_000, 48, 256, 100
    if case == "q7":
        n += 13                                                  # a last partial group of 32 rows (row-major tail of the packed copy)
    if "kc200" in case:
        m, kc = 32, 200
    if "m96" in case:
        m, k = 96, 10
    dsub = 8
    codes = torch.from_numpy(rng.integers(0, 256, (n, m), dtype=np.uint8)).cuda()       # codes >= kc are clamped
    cb = torch.from_numpy((rng.standard_normal((m, kc, dsub)) / np.sqrt(m * dsub)).astype(np.float32)).cuda()
    if "flat" in case:
        cb[:] = 0.25                                             # every table entry of a subspace equal: range 0
    q = torch.from_numpy(rng.standard_normal((nq, m * dsub)).astype(np.float32)).cuda()
    lut = ops.pq_build_lut(cb, q)
    words = ops.pack_mask(torch.from_numpy(rng.random(n) < 0.25).cuda()) if "mask" in case else None
    packed = ops.pq_pack(codes)
    d4, i4, c4 = ops.pq_adc_packed(lut, packed, k, words)
    for qi in range(nq):
        d1, i1, c1 = ops.pq_adc_packed(lut[qi:qi + 1].contiguous(), packed, k, words)
        assert torch.equal(c4[qi:qi + 1], c1)
        assert torch.equal(i4[qi:qi + 1], i1), (case, qi)
        assert torch.equal(d4[qi:qi + 1], d1), (case, qi)
    if "flat" not in case:                                        # and against the reference arithmetic
        host_codes = np.minimum(codes.cpu().numpy(), kc - 1)
        ref = O.pq_distances_with_table(lut[0].cpu().numpy(), host_codes)
        valid = None
        if words is not None:
            valid = np.unpackbits(words.cpu().numpy().view(np.uint8), bitorder="little")[:n].astype(bool)
        O.check_topk(ref, i4[0].cpu().numpy(), d4[0].cpu().numpy(), k, valid=valid, rtol=1e-5)


def test_product_quantizer_search_batch_equals_search(fpv):
    """ProductQuantizer.search_batch (one pass per four queries over >= 2^20 codes) == search() per query, bit for bit."""
    rng = np.random.default_rng(31)
    n, d, m = 1_100_000, 64, 16
    pq = fpv.ProductQuantizer(d, num_subspaces=m, num_centroids=256)
    pq.codebooks = (rng.standard_normal((m, 256, d // m)) * 0.3).astype(np.float32)
    pq.trained = True
    codes = torch.from_numpy(rng.integers(0, 256, (n, m), dtype=np.uint8)).cuda()
    qs = (rng.standard_normal((6, d)) * 0.3).astype(np.float32)
    mask = rng.random(n) < 0.4
    idx, dist = pq.search_batch(qs, codes, k=50, filter_mask=mask)
    assert idx.shape == (6, 50)
    for qi in range(6):
        i1, d1 = pq.search(qs[qi], codes, k=50, filter_mask=mask)
        assert np.array_equal(idx[qi].cpu().numpy(), i1.cpu().numpy()) and np.array_equal(dist[qi].cpu().numpy(), d1.cpu().numpy()), qi
