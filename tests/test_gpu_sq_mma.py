"""uint8 scalar-quantizer L2 scan on the int8 tensor cores (csrc/fpv_sq_mma.cu) against
 (a) exact integer arithmetic: the three limb dot products per row are bit-identical to np.dot in int64
     (oracle.sq_limb_dots_int64 -- the "int8 dot products must be bit-exact" bar of BASELINE.json's north_star),
 (b) the SIMT scan kernel (same distances and ids, bit for bit: the tensor cores only filter, the survivors are re-scored
     with the scan's arithmetic), and
 (c) the oracle restatement of ScalarQuantizer.distances_l2 (quantization.py:145-152, 217-236) within 1e-5."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _quantizer(d, seed, uniform_scale=False):
    import fastpyvectordb_b200 as fpv
    rng = np.random.default_rng(seed)
    sq = fpv.ScalarQuantizer(d)
    lo = (-0.2 - 0.3 * rng.random(d)).astype(np.float32)
    hi = (0.2 + 0.5 * rng.random(d)).astype(np.float32)
    if uniform_scale:
        lo[:], hi[:] = -0.1, 0.1
    sq.min_vals, sq.max_vals, sq.scale, sq.trained = lo, hi, (hi - lo).astype(np.float32), True
    return sq


@pytest.mark.parametrize("n,d", [(70000, 1024), (66000, 208), (65536, 16)])
def test_limb_dots_are_exact_integers(n, d):
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(1)
    sq = _quantizer(d, 2)
    codes = torch.from_numpy(rng.integers(0, 256, (n, d), dtype=np.uint8)).cuda()
    qcode = torch.from_numpy(rng.integers(0, 256, (1, d), dtype=np.uint8)).cuda()
    scale = torch.from_numpy(sq.scale).cuda()
    limbs, mma, simt = ops.sq_mma_limb_dots(qcode, codes, scale)
    limbs, mma, simt = limbs.cpu().numpy(), mma.cpu().numpy(), simt.cpu().numpy()
    host = codes.cpu().numpy()
    for l in range(3):
        want = O.sq_limb_dots_int64(limbs[l, :d].astype(np.int32), host)
        assert np.array_equal(mma[l].astype(np.int64), want), f"limb {l}: tensor-core dot differs from int64 np.dot"
        assert np.array_equal(simt[l].astype(np.int64), want)
    # the limbs are the 24-bit fixed point of a_j = (scale_j/255)^2 * qcode_j against alpha = max a_j / (2^24 - 1)
    s = (sq.scale / np.float32(255.0)).astype(np.float32).astype(np.float64)
    a = s * s * qcode.cpu().numpy()[0].astype(np.float64)
    A = limbs[0, :d].astype(np.int64) + 256 * limbs[1, :d].astype(np.int64) + 65536 * limbs[2, :d].astype(np.int64)
    alpha = a.max() / 16777215.0
    assert np.abs(A * alpha - a).max() <= alpha * 0.51 + 1e-6 * a.max()
    assert (limbs[:, d:] == 0).all()


@pytest.mark.parametrize("n,d,q,k,mask", [(70000, 1024, 5, 100, False), (131072, 128, 21, 10, True), (65600, 208, 1, 100, False),
                                           (90000, 512, 16, 1000, True)])
def test_tensor_core_scan_equals_simt_scan_and_oracle(n, d, q, k, mask):
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(7)
    sq = _quantizer(d, 3)
    x = (rng.standard_normal((n, d)) * 0.15).astype(np.float32)
    x[1234] = x[77]                                                     # an exact duplicate: tie broken by index
    codes = sq.encode(x)
    qs = (rng.standard_normal((q, d)) * 0.15).astype(np.float32)
    m = (rng.random(n) < 0.3) if mask else None
    idx, dist = sq.search_batch(qs, codes, k=k, filter_mask=m)
    assert idx.shape == (q, k) and dist.dtype == np.float32
    flags = ops.sq_mma_last_flags(q, n, d, k, torch.device("cuda", 0))
    assert int(flags.sum()) == 0, "no query of this well-spread data may need the SIMT fallback"
    sq.tensor_core_scan = False
    for qi in range(q):
        i1, d1 = sq.search(qs[qi], codes, k=k, filter_mask=m)
        assert np.array_equal(idx[qi], i1) and np.array_equal(dist[qi], d1), f"query {qi}"
    sq.tensor_core_scan = True
    for qi in range(min(q, 3)):
        ref = O.sq_distances_l2(qs[qi], codes, sq.min_vals, sq.scale)
        O.check_topk(ref, idx[qi], dist[qi], k, valid=m, rtol=1e-5)
    # the single-query API takes the same path for large code matrices
    i0, d0 = sq.search(qs[0], codes, k=k, filter_mask=m)
    assert np.array_equal(i0, idx[0]) and np.array_equal(d0, dist[0])


def test_tie_heavy_codes_fall_back_to_the_exact_scan():
    """Low-dimensional binary-ish codes: thousands of rows share the k-th distance, the certified window does not fit
    and the query is answered by the SIMT scan on the device -- same answer, flagged."""
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(11)
    n, d = 80000, 16
    sq = _quantizer(d, 5, uniform_scale=True)
    codes = rng.integers(0, 2, (n, d), dtype=np.uint8) * 200
    codes[:, 4:] = 0                                     # 16 distinct distances, ~5000 rows each: the top-50 all tie
    qs = np.zeros((3, d), np.float32)
    idx, dist = sq.search_batch(qs, codes, k=50)
    flags = ops.sq_mma_last_flags(3, n, d, 50, torch.device("cuda", 0))
    sq.tensor_core_scan = False
    for qi in range(3):
        i1, d1 = sq.search(qs[qi], codes, k=50)
        assert np.array_equal(idx[qi], i1) and np.array_equal(dist[qi], d1)
    assert int(flags.sum()) >= 1


@pytest.mark.parametrize("metric", ["dot", "cosine"])
@pytest.mark.parametrize("n,d,q,k,mask", [(70000, 1024, 5, 100, False), (131072, 128, 21, 10, True), (65600, 208, 1, 100, False),
                                           (90000, 512, 16, 1000, True)])
def test_dot_and_cosine_on_the_tensor_cores_equal_the_simt_scan_and_oracle(n, d, q, k, mask, metric):
    """distances_dot / distances_cosine batches: the signed weights are shifted so that the limbs stay unsigned
    (sum_j a_j b_j = alpha T - c sum_j b_j), cosine multiplies by the per-row inverse norm; the certified window is
    re-scored by the scan's own loop, so ids and distances equal the SIMT scan bit for bit."""
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(17)
    sq = _quantizer(d, 4)
    x = (rng.standard_normal((n, d)) * 0.15).astype(np.float32)
    x[4321] = x[99]                                                     # an exact duplicate: tie broken by index
    codes = sq.encode(x)
    qs = (rng.standard_normal((q, d)) * 0.15).astype(np.float32)
    m = (rng.random(n) < 0.3) if mask else None
    idx, dist = sq.search_batch(qs, codes, k=k, filter_mask=m, metric=metric)
    assert idx.shape == (q, k) and dist.dtype == np.float32
    flags = ops.sq_mma_last_flags(q, n, d, k, torch.device("cuda", 0))
    assert int(flags.sum()) == 0, "no query of this well-spread data may need the SIMT fallback"
    sq.tensor_core_scan = False
    for qi in range(q):
        i1, d1 = sq.search(qs[qi], codes, k=k, metric=metric, filter_mask=m)
        assert np.array_equal(idx[qi], i1) and np.array_equal(dist[qi], d1), f"query {qi}"
    sq.tensor_core_scan = True
    ref_fn = O.sq_distances_dot if metric == "dot" else O.sq_distances_cosine
    for qi in range(min(q, 3)):
        ref = ref_fn(qs[qi], codes, sq.min_vals, sq.scale)
        O.check_topk(ref, idx[qi], dist[qi], k, valid=m, rtol=1e-5)
    i0, d0 = sq.search(qs[0], codes, k=k, metric=metric, filter_mask=m)
    assert np.array_equal(i0, idx[0]) and np.array_equal(d0, dist[0])


def test_cosine_with_a_zero_row_falls_back():
    """A row that decodes to (almost) the zero vector makes the cosine error bound useless (1 / |row| ~ 1e8): every
    query is answered by the SIMT scan on the device -- same answer, flagged."""
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(23)
    n, d = 70000, 64
    sq = _quantizer(d, 6)
    sq.min_vals[:] = 0.0                                                # code 0 decodes to exactly 0
    sq.scale[:] = sq.max_vals
    x = (np.abs(rng.standard_normal((n, d))) * 0.15).astype(np.float32)
    x[500] = 0.0
    codes = sq.encode(x)
    qs = (np.abs(rng.standard_normal((2, d))) * 0.15).astype(np.float32)
    idx, dist = sq.search_batch(qs, codes, k=20, metric="cosine")
    flags = ops.sq_mma_last_flags(2, n, d, 20, torch.device("cuda", 0))
    sq.tensor_core_scan = False
    for qi in range(2):
        i1, d1 = sq.search(qs[qi], codes, k=20, metric="cosine")
        assert np.array_equal(idx[qi], i1) and np.array_equal(dist[qi], d1)
    assert int(flags.sum()) == 2
