"""CPU oracle for the FastPyVectorDB search hot path (TEST INFRASTRUCTURE ONLY).

This file is a NumPy restatement of the arithmetic in the reference's
``parallel_search.py`` and ``quantization.py``.  It is the *checker* for the CUDA
path: only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline /
``--impl reference`` legs may import it.  Nothing under ``fastpyvectordb_b200/`` may
import it, and the product path has no CPU fallback.

Parity pinning: the reference ships no test or golden vector for this path
(SURVEY.md §8c), so the oracle is pinned against the reference *itself*:
``tests/golden/make_golden.py`` imports ``/root/reference/parallel_search.py`` and
``/root/reference/quantization.py`` in the build container, runs them on seeded
inputs and commits the outputs as ``tests/golden/*.npz``;
``tests/test_oracle_golden.py`` replays the same inputs through this file.

All functions return arrays (never result objects); the tie rule of the build
(distance, then lowest index) is exposed separately as :func:`canonical_topk`
because the reference's own order among ties is arbitrary (argpartition).

Every function cites the reference ``file:line`` it follows.
"""
from __future__ import annotations

import numpy as np

EPS_SEARCH = 1e-10   # parallel_search.py:87-88,121-123,271,275
EPS_SQ_COS = 1e-8    # quantization.py:168-169


# --------------------------------------------------------------------------------------
# float distances (parallel_search.py)
# --------------------------------------------------------------------------------------
def distances_single(query, vectors, metric="cosine"):
    """1 x N distances; follows _compute_distances_vectorized (parallel_search.py:105-134)."""
    query = np.asarray(query, dtype=np.float32).reshape(-1)
    vectors = np.ascontiguousarray(vectors, dtype=np.float32)
    if metric == "cosine":
        qn = query / (np.linalg.norm(query) + EPS_SEARCH)
        vn = np.sqrt(np.einsum("ij,ij->i", vectors, vectors)) + EPS_SEARCH
        return 1.0 - np.dot(vectors, qn) / vn
    if metric == "l2":
        qq = np.dot(query, query)
        vv = np.einsum("ij,ij->i", vectors, vectors)
        return np.sqrt(np.maximum(qq + vv - 2 * np.dot(vectors, query), 0))
    return -np.dot(vectors, query)


def distances_chunk(query, chunk, start_idx, metric):
    """(n,2) float64 [global_idx, dist]; follows _compute_distances_chunk (parallel_search.py:72-102)."""
    query = np.asarray(query, dtype=np.float32).reshape(-1)
    chunk = np.asarray(chunk, dtype=np.float32)
    if metric == "cosine":
        qn = query / (np.linalg.norm(query) + EPS_SEARCH)
        rows = chunk / (np.linalg.norm(chunk, axis=1, keepdims=True) + EPS_SEARCH)
        d = 1.0 - np.dot(rows, qn)
    elif metric == "l2":
        delta = chunk - query
        d = np.sqrt(np.sum(delta ** 2, axis=1))
    else:
        d = -np.dot(chunk, query)
    ids = np.arange(start_idx, start_idx + len(chunk))
    return np.column_stack([ids, d])


def distances_batch(queries, vectors, metric="cosine"):
    """Q x N distance matrix; follows search_batch_parallel (parallel_search.py:268-290)."""
    queries = np.ascontiguousarray(queries, dtype=np.float32)
    vectors = np.ascontiguousarray(vectors, dtype=np.float32)
    if queries.ndim == 1:
        queries = queries.reshape(1, -1)
    if metric == "cosine":
        qn = np.sqrt(np.einsum("ij,ij->i", queries, queries, optimize=True)) + EPS_SEARCH
        vn = np.sqrt(np.einsum("ij,ij->i", vectors, vectors, optimize=True)) + EPS_SEARCH
        return 1.0 - np.dot(queries / qn[:, None], (vectors / vn[:, None]).T)
    if metric == "l2":
        qq = np.einsum("ij,ij->i", queries, queries)[:, None]
        vv = np.einsum("ij,ij->i", vectors, vectors)[None, :]
        return np.sqrt(np.maximum(qq + vv - 2 * np.dot(queries, vectors.T), 0))
    return -np.dot(queries, vectors.T)


def reference_order_topk(dist, k):
    """The reference's own selection: argpartition then argsort (parallel_search.py:228-233,
    299-303; quantization.py:388-392, 591-595).  Order among equal distances is arbitrary."""
    n = len(dist)
    if k < n:
        part = np.argpartition(dist, k)[:k]
        return part[np.argsort(dist[part])]
    return np.argsort(dist)


def canonical_topk(dist, k, valid=None):
    """Build's deterministic rule on top of reference distances: (distance, lowest index).
    ``valid`` (bool mask) drops rows the way filter_mask does (parallel_search.py:212-217)."""
    dist = np.asarray(dist)
    ids = np.arange(len(dist))
    if valid is not None:
        ids = ids[np.asarray(valid, dtype=bool)]
    order = np.lexsort((ids, dist[ids]))
    sel = ids[order[: min(k, len(ids))]]
    return sel.astype(np.int64), dist[sel]


def search_parallel(query, vectors, k=10, metric="cosine", filter_mask=None):
    """(idx, dist) in the reference's order; follows search_parallel (parallel_search.py:209-244)."""
    query = np.asarray(query, dtype=np.float32).flatten()
    vectors = np.ascontiguousarray(vectors, dtype=np.float32)
    keep = None
    rows = vectors
    if filter_mask is not None:
        keep = np.where(filter_mask)[0]
        rows = vectors[keep]
    if len(rows) == 0:
        return np.zeros(0, np.int64), np.zeros(0, np.float32)
    d = distances_single(query, rows, metric)
    top = reference_order_topk(d, min(k, len(rows)))
    ids = top if keep is None else keep[top]
    return ids.astype(np.int64), d[top]


def search_batch_parallel(queries, vectors, k=10, metric="cosine"):
    """(idx[Q,k'], dist[Q,k']); follows search_batch_parallel (parallel_search.py:259-311)."""
    dm = distances_batch(queries, vectors, metric)
    kk = min(k, dm.shape[1])
    idx = np.empty((dm.shape[0], kk), np.int64)
    dist = np.empty((dm.shape[0], kk), dm.dtype)
    for i in range(dm.shape[0]):
        top = reference_order_topk(dm[i], kk)[:kk]
        idx[i], dist[i] = top, dm[i][top]
    return idx, dist


def merge_top_k(results_list, k):
    """k-way merge of (n_i,2) float64 [idx, dist] blocks; follows _merge_top_k (parallel_search.py:137-156)."""
    stacked = np.vstack(results_list)
    if len(stacked) <= k:
        return stacked[np.argsort(stacked[:, 1])]
    top = stacked[np.argpartition(stacked[:, 1], k)[:k]]
    return top[np.argsort(top[:, 1])]


def search_chunked_parallel(query, vectors, k=10, metric="cosine", chunk_size=50000):
    """Row chunks -> local top-k -> merge; follows search_chunked_parallel (parallel_search.py:326-368)."""
    query = np.asarray(query, dtype=np.float32).flatten()
    vectors = np.ascontiguousarray(vectors, dtype=np.float32)
    n = len(vectors)
    if n <= chunk_size:
        return search_parallel(query, vectors, k, metric)
    blocks = []
    for start in range(0, n, chunk_size):
        d = distances_single(query, vectors[start:start + chunk_size], metric)
        kk = min(k, len(d))
        top = np.argpartition(d, kk)[:kk] if kk < len(d) else np.arange(len(d))
        blocks.append(np.column_stack([top + start, d[top]]))
    merged = merge_top_k(blocks, k)
    return merged[:, 0].astype(np.int64), merged[:, 1]


def rerank_cosine(query, vectors, candidate_ids, k):
    """Exact cosine re-rank of gathered candidates; follows ParallelCollection.search_hybrid
    (parallel_search.py:919-934): normalise query and candidate rows (eps 1e-10), 1 - dot, sort."""
    query = np.asarray(query, dtype=np.float32).flatten()
    cand = np.asarray(candidate_ids, dtype=np.int64)
    rows = np.ascontiguousarray(vectors, dtype=np.float32)[cand]
    qn = query / (np.linalg.norm(query) + EPS_SEARCH)
    rn = rows / (np.linalg.norm(rows, axis=1, keepdims=True) + EPS_SEARCH)
    d = 1.0 - np.dot(rn, qn)
    order = np.lexsort((cand, d))[: min(k, len(cand))]
    return cand[order], d[order]


def brute_force_distances(query, vectors, metric="cosine"):
    """vectordb_optimized.Collection.brute_force_search distance block
    (vectordb_optimized.py:668-684): cosine WITHOUT epsilon, l2 via norm(V - q), ip = -dot."""
    query = np.asarray(query, dtype=np.float32)
    vectors = np.asarray(vectors, dtype=np.float32)
    if metric == "cosine":
        qn = query / np.linalg.norm(query)
        vn = vectors / np.linalg.norm(vectors, axis=1, keepdims=True)
        return 1 - np.dot(vn, qn)
    if metric == "l2":
        return np.linalg.norm(vectors - query, axis=1)
    return -np.dot(vectors, query)


# --------------------------------------------------------------------------------------
# scalar (uint8) quantizer (quantization.py:64-276)
# --------------------------------------------------------------------------------------
def sq_train(vectors):
    """(min, max, scale); follows ScalarQuantizer.train (quantization.py:91-106)."""
    vectors = np.asarray(vectors, dtype=np.float32)
    if vectors.ndim == 1:
        vectors = vectors.reshape(1, -1)
    lo = vectors.min(axis=0)
    hi = vectors.max(axis=0)
    scale = hi - lo
    scale = np.where(scale == 0, 1.0, scale)
    return lo, hi, scale


def sq_encode(vectors, min_vals, scale):
    """uint8 codes by truncation; follows ScalarQuantizer.encode (quantization.py:118-126)."""
    vectors = np.asarray(vectors, dtype=np.float32)
    if vectors.ndim == 1:
        vectors = vectors.reshape(1, -1)
    unit = (vectors - min_vals) / scale
    return np.clip(unit * 255, 0, 255).astype(np.uint8)


def sq_decode(codes, min_vals, scale):
    """follows ScalarQuantizer.decode (quantization.py:136-139)."""
    return codes.astype(np.float32) / 255.0 * scale + min_vals


def sq_distances_l2(query, codes, min_vals, scale):
    """follows distances_l2 -> _sq_distances_l2_vectorized (quantization.py:151-152, 223-236)."""
    qc = sq_encode(np.asarray(query, np.float32).reshape(1, -1), min_vals, scale)[0]
    delta = qc.astype(np.int16) - codes.astype(np.int16)
    scaled = delta.astype(np.float32) * (scale / 255.0)
    return np.sqrt(np.sum(scaled ** 2, axis=1))


def sq_distances_dot(query, codes, min_vals, scale):
    """follows distances_dot -> _sq_distances_dot_vectorized (quantization.py:180-181, 245-251)."""
    qc = sq_encode(np.asarray(query, np.float32).reshape(1, -1), min_vals, scale)[0]
    qr = qc.astype(np.float32) / 255.0 * scale + min_vals
    dr = codes.astype(np.float32) / 255.0 * scale + min_vals
    return -np.dot(dr, qr)


def sq_distances_cosine(query, codes, min_vals, scale):
    """follows distances_cosine (quantization.py:161-174); the ``norms`` argument is ignored there."""
    qc = sq_encode(np.asarray(query, np.float32).reshape(1, -1), min_vals, scale)[0]
    dr = sq_decode(codes, min_vals, scale)
    qr = sq_decode(qc.reshape(1, -1), min_vals, scale)[0]
    qn = qr / (np.linalg.norm(qr) + EPS_SQ_COS)
    dn = dr / (np.linalg.norm(dr, axis=1, keepdims=True) + EPS_SQ_COS)
    return 1.0 - np.dot(dn, qn)


def sq_limb_dots_int64(weights_i32, codes):
    """Exact integer cross term used by the int8 tensor-core route: one int64
    dot per row between an int32 fixed-point query weight vector and the uint8 codes."""
    return codes.astype(np.int64) @ np.asarray(weights_i32, dtype=np.int64)


# --------------------------------------------------------------------------------------
# binary quantizer / Hamming (quantization.py:282-407)
# --------------------------------------------------------------------------------------
def bq_train(vectors, use_median=True, threshold=0.0):
    """per-dimension thresholds; follows BinaryQuantizer.train (quantization.py:315-327)."""
    vectors = np.asarray(vectors, dtype=np.float32)
    if vectors.ndim == 1:
        vectors = vectors.reshape(1, -1)
    if use_median:
        return np.median(vectors, axis=0)
    return np.full(vectors.shape[1], threshold)


def bq_encode(vectors, thresholds):
    """(v > thr) packed MSB-first; follows BinaryQuantizer.encode (quantization.py:336-350)."""
    vectors = np.asarray(vectors, dtype=np.float32)
    if vectors.ndim == 1:
        vectors = vectors.reshape(1, -1)
    return np.packbits((vectors > thresholds).astype(np.uint8), axis=1)


def bq_hamming(query_bits, db_bits, dimensions=None):
    """popcount(q XOR d) over the first ``dimensions`` bits as float32; follows
    BinaryQuantizer.hamming_distances (quantization.py:364-374)."""
    bits = np.unpackbits(np.bitwise_xor(query_bits, db_bits), axis=1)
    if dimensions:
        bits = bits[:, :dimensions]
    return bits.sum(axis=1).astype(np.float32)


def bq_hamming_chunked(query_bits, db_bits, dimensions=None, rows=262144):
    """Same as :func:`bq_hamming`, driven over row chunks so the N x D byte temporary stays small."""
    out = np.empty(len(db_bits), np.float32)
    for s in range(0, len(db_bits), rows):
        out[s:s + rows] = bq_hamming(query_bits, db_bits[s:s + rows], dimensions)
    return out


# --------------------------------------------------------------------------------------
# product quantizer (quantization.py:414-615)
# --------------------------------------------------------------------------------------
def pq_encode(vectors, codebooks):
    """(N,M) uint8, first-min argmin per subspace; follows ProductQuantizer.encode (quantization.py:520-539)."""
    vectors = np.asarray(vectors, dtype=np.float32)
    if vectors.ndim == 1:
        vectors = vectors.reshape(1, -1)
    m_sub, _, dsub = codebooks.shape
    codes = np.zeros((len(vectors), m_sub), np.uint8)
    for m in range(m_sub):
        sub = vectors[:, m * dsub:(m + 1) * dsub]
        d = np.sum((sub[:, np.newaxis] - codebooks[m]) ** 2, axis=2)
        codes[:, m] = np.argmin(d, axis=1)
    return codes


def pq_lookup_table(query, codebooks):
    """(M,K) fp32 squared distances; follows build_lookup_table (quantization.py:551-562)."""
    query = np.asarray(query, dtype=np.float32).flatten()
    m_sub, k_cent, dsub = codebooks.shape
    table = np.zeros((m_sub, k_cent), np.float32)
    for m in range(m_sub):
        table[m] = np.sum((codebooks[m] - query[m * dsub:(m + 1) * dsub]) ** 2, axis=1)
    return table


def pq_distances_with_table(table, codes):
    """sqrt(sum_m table[m, code]) accumulated sequentially in m, fp32; follows
    distances_with_table (quantization.py:571-578)."""
    acc = np.zeros(len(codes), np.float32)
    for m in range(table.shape[0]):
        acc += table[m, codes[:, m]]
    return np.sqrt(acc)


def pq_kmeans(data, k, n_iter, rng=np.random):
    """k-means++ seeding then Lloyd; follows ProductQuantizer._kmeans (quantization.py:482-508).
    ``rng`` is the global ``np.random`` module in the reference."""
    n = len(data)
    cent = np.zeros((k, data.shape[1]), np.float32)
    cent[0] = data[rng.randint(n)]
    for i in range(1, k):
        d = np.min([np.sum((data - c) ** 2, axis=1) for c in cent[:i]], axis=0)
        cent[i] = data[rng.choice(n, p=d / d.sum())]
    for _ in range(n_iter):
        d = np.array([np.sum((data - c) ** 2, axis=1) for c in cent])
        assign = np.argmin(d, axis=0)
        for j in range(k):
            members = assign == j
            if members.any():
                cent[j] = data[members].mean(axis=0)
    return cent


# --------------------------------------------------------------------------------------
# tie-aware comparison helpers
# --------------------------------------------------------------------------------------
def check_topk(ref_dist_all, got_idx, got_dist, k, *, rtol=1e-5, integer=False, valid=None,
               squared_near_zero=False):
    """Tie-aware check of one query's result against the reference distances of ALL rows.

    * every returned id is a permitted row and appears once;
    * returned distances match ``ref_dist_all[id]`` (exactly when ``integer``; otherwise within
      ``rtol * max(|d|, 1)``; with ``squared_near_zero`` distances below 1e-2 are compared squared,
      because the reference's q^2+v^2-2qv form is itself noisy at true distance 0);
    * every returned id has ref distance <= kth + tol, and every row with ref distance < kth - tol
      is present;
    * the list is ordered by (distance, index) up to tol.
    Returns None or raises AssertionError with a description."""
    ref = np.asarray(ref_dist_all)
    got_idx = np.asarray(got_idx).astype(np.int64)
    got_dist = np.asarray(got_dist)
    permitted = np.ones(len(ref), bool) if valid is None else np.asarray(valid, bool)
    n_ok = int(permitted.sum())
    kk = min(k, n_ok)
    assert len(got_idx) >= kk, f"returned {len(got_idx)} < expected {kk}"
    got_idx, got_dist = got_idx[:kk], got_dist[:kk]
    if kk == 0:
        return
    assert got_idx.min() >= 0 and got_idx.max() < len(ref), "id out of range"
    assert len(np.unique(got_idx)) == kk, "duplicate ids"
    assert permitted[got_idx].all(), "filtered row returned"
    rd = ref[got_idx].astype(np.float64)
    gd = got_dist.astype(np.float64)
    if integer:
        assert np.array_equal(rd, gd), f"integer distances differ: {rd[:5]} vs {gd[:5]}"
        tol = 0.0
    else:
        err = np.abs(rd - gd)
        lim = rtol * np.maximum(np.abs(rd), 1.0)
        if squared_near_zero:
            small = rd < 1e-2
            err = np.where(small, np.abs(rd * rd - gd * gd), err)
        bad = err > lim
        assert not bad.any(), f"distance mismatch: max err {err.max():.3e} at {got_idx[bad][:5]}"
        tol = rtol * max(abs(float(np.sort(ref[permitted])[kk - 1])), 1.0)
    ref_ok = np.where(permitted, ref, np.inf)
    kth = np.partition(ref_ok, kk - 1)[kk - 1]
    assert (rd <= kth + tol).all(), "returned a row worse than the k-th reference distance"
    must = np.where(ref_ok < kth - tol)[0]
    missing = np.setdiff1d(must, got_idx)
    assert len(missing) == 0, f"missing rows that beat the k-th distance: {missing[:5]}"
    if integer:
        order = np.lexsort((got_idx, gd))
        assert np.array_equal(order, np.arange(kk)), "not ordered by (distance, index)"
        # lowest-index rule inside the boundary tie group
        tie_rows = np.where(ref_ok == kth)[0]
        need = kk - int((ref_ok < kth).sum())
        assert np.array_equal(np.sort(got_idx[rd == kth]), tie_rows[:need]), \
            "boundary tie group is not the lowest-index members"
    else:
        assert (np.diff(gd) >= -tol).all(), "distances not ascending"
