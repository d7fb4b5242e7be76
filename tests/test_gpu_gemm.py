"""GPU parity tests of the tensor-core batch path (csrc/fpv_gemm_topk.cu) against the oracle.

The path must return the SAME exact fp32 results as the scan path: the approximate TF32/BF16 pass only filters,
a certificate decides per query whether the exact re-rank of the candidates is provably complete, and queries
that fail it are recomputed by the exact scan on the device."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def mods():
    import fastpyvectordb_b200 as fpv
    from fastpyvectordb_b200 import engine_gemm, ops
    return fpv, engine_gemm, ops


def _data(n, d, q, unit, seed=42):
    rng = np.random.default_rng(seed)
    db = rng.standard_normal((n, d)).astype(np.float32)
    qs = np.random.default_rng(999).standard_normal((q, d)).astype(np.float32)
    if unit:
        db /= np.linalg.norm(db, axis=1, keepdims=True)
        qs /= np.linalg.norm(qs, axis=1, keepdims=True)
    else:
        db *= rng.uniform(0.5, 2.0, size=(n, 1)).astype(np.float32)
    return db, qs


def _run(mods, db, qs, k, metric, mode):
    fpv, engine_gemm, ops = mods
    index = fpv.GpuIndex(db)
    q = torch.from_numpy(qs).cuda()
    assert engine_gemm.available(index, len(qs), k)
    dist, idx, cnt = engine_gemm.search(q, index, k, metric, mode=mode)
    torch.cuda.synchronize()
    flags = ops.gemm_last_flags(len(qs), len(db), db.shape[1], k, 0 if mode == "tf32" else 1, index.device)
    return dist.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy(), flags.cpu().numpy()


@pytest.mark.parametrize("mode", ["tf32", "bf16"])
@pytest.mark.parametrize("metric", ["cosine", "l2", "ip"])
@pytest.mark.parametrize("n,d,q,k,unit", [(8192, 64, 128, 10, True), (20000, 384, 200, 10, True),
                                          (30011, 768, 130, 100, False), (5000, 100, 17, 256, False),
                                          # CTA-pair kernel corner cases: a ragged last tile (133 rows) whose second
                                          # half is almost entirely out of bounds; an odd number of query blocks padded to whole pairs
                                          (4229, 64, 257, 50, True), (70000, 128, 600, 100, True)])
def test_gemm_path_matches_oracle(mods, n, d, q, k, unit, metric, mode):
    if mode == "bf16" and d % 8:
        pytest.skip("bf16 pass needs d % 8 == 0")
    db, qs = _data(n, d, q, unit)
    dist, idx, cnt, flags = _run(mods, db, qs, k, metric, mode)
    ref = O.distances_batch(qs, db, metric)
    assert (cnt == k).all()
    for qi in range(q):
        O.check_topk(ref[qi], idx[qi], dist[qi], k, squared_near_zero=(metric == "l2"))
    # the certificate is expected to hold for (almost) every query on random data
    assert flags.mean() <= 0.25, f"{int(flags.sum())}/{q} queries fell back to the exact scan"
    # bit-identical to the fp32 scan path: every exact kernel uses the same summation order, so an answer does not
    # depend on which path (or batch size) produced it
    fpv, engine_gemm, ops = mods
    index = fpv.GpuIndex(db)
    sd, si, sc = ops.scan_f32_topk(torch.from_numpy(qs[:24]).cuda(), index.rows, k, metric, None, index.row_sq, 0)
    assert np.array_equal(si.cpu().numpy(), idx[:24]) and np.array_equal(sd.cpu().numpy(), dist[:24])


@pytest.mark.parametrize("keep_frac", [0.5, 0.02, 0.0005])
@pytest.mark.parametrize("mode", ["tf32", "bf16"])
def test_gemm_path_with_row_filter(mods, keep_frac, mode):
    """filter_mask of search_parallel (parallel_search.py:212-217) applied in the tensor-core epilogue: only permitted
    rows compete; a mask that leaves fewer than k rows returns exactly those (count < k, padding (inf, -1))."""
    fpv, engine_gemm, ops = mods
    n, d, q, k = 40000, 128, 200, 50
    db, qs = _data(n, d, q, True, seed=7)
    mask = np.random.default_rng(11).random(n) < keep_frac
    mask[n - 1] = True
    n_ok = int(mask.sum())
    index = fpv.GpuIndex(db)
    words = ops.pack_mask(torch.from_numpy(mask).cuda())
    dist, idx, cnt = engine_gemm.search(torch.from_numpy(qs).cuda(), index, k, "l2", mode=mode, mask_words=words)
    dist, idx, cnt = dist.cpu().numpy(), idx.cpu().numpy(), cnt.cpu().numpy()
    assert (cnt == min(k, n_ok)).all()
    ref = O.distances_batch(qs, db, "l2")
    for qi in range(q):
        kk = int(cnt[qi])
        O.check_topk(ref[qi], idx[qi, :kk], dist[qi, :kk], k, valid=mask, squared_near_zero=True)
        assert (idx[qi, kk:] == -1).all() and np.isinf(dist[qi, kk:]).all()
    # the engine routes a filtered batch to the same path and agrees with the filtered fp32 scan bit for bit
    eng = fpv.ParallelSearchEngine()
    ed, ei, ec = eng.search_tensors(qs, index, k, "l2", filter_mask=mask)
    assert np.array_equal(ei.cpu().numpy(), idx) and np.array_equal(ed.cpu().numpy(), dist)
    sd, si, sc = ops.scan_f32_topk(torch.from_numpy(qs[:8]).cuda(), index.rows, k, "l2", words, index.row_sq, 0)
    assert np.array_equal(si.cpu().numpy(), idx[:8]) and np.array_equal(sd.cpu().numpy(), dist[:8])


@pytest.mark.parametrize("metric", ["l2", "ip", "cosine"])
def test_bf16_error_bound_holds_for_aligned_rounding_errors(mods, metric):
    """The BF16 filter's error bound is the Cauchy-Schwarz bound on MEASURED rounding errors.  Worst case for it: the
    part of every row that bf16 rounds away points exactly along the query, so |q.(v - bf16(v))| = |q| |v - bf16(v)|.
    Rows are a bf16-exact base plus eps_i * c * q_dir with |residual element| below half an ulp of the base element:
    the shadow copy is the base alone, the exact scores differ by up to c |q| and only the exact re-rank can order
    them.  A bound that is too small would drop true neighbours here."""
    fpv, engine_gemm, ops = mods
    rng = np.random.default_rng(21)
    n, d, q, k = 20000, 64, 130, 20
    qs = rng.standard_normal((q, d)).astype(np.float32)
    qdir = (qs[0] / np.linalg.norm(qs[0])).astype(np.float32)
    qdir = np.clip(qdir, -0.3, 0.3)
    mag = rng.integers(9, 31, size=(n, d)).astype(np.float32) / 16.0          # 0.5625 .. 1.875 in steps of 1/16: bf16 exact
    base = mag * rng.choice([-1.0, 1.0], size=(n, d)).astype(np.float32)
    eps = rng.uniform(-1.0, 1.0, size=(n, 1)).astype(np.float32)
    db = (base + eps * np.float32(5e-3) * qdir[None, :]).astype(np.float32)   # |residual| <= 1.5e-3 < 2^-9
    shadow = torch.from_numpy(db).cuda().to(torch.bfloat16).float().cpu().numpy()
    assert np.array_equal(shadow, base), "construction: bf16 must round the residual away"
    dist, idx, cnt, flags = _run(mods, db, qs, k, metric, "bf16")
    ref = O.distances_batch(qs, db, metric)
    for qi in range(q):
        O.check_topk(ref[qi], idx[qi], dist[qi], k, squared_near_zero=(metric == "l2"))
    assert flags.mean() <= 0.25


@pytest.mark.parametrize("mode", ["bf16", "tf32", None])
def test_gemm_c1_shape_cosine_top10(mods, mode):
    """BASELINE configs[0]: 100k x 384 unit rows, 1000 queries, cosine top-10 -- in the operand format the product
    dispatch picks at this shape (None -> engine_gemm's own choice: bf16) and in both explicit ones."""
    fpv, engine_gemm, ops = mods
    db, qs = _data(100_000, 384, 1000, True)
    if mode is None:
        index = fpv.GpuIndex(db)
        assert engine_gemm._effective_mode(None, index, 10, len(qs)) == "bf16"
        eng = fpv.ParallelSearchEngine()
        idx, dist = eng.search_arrays(qs, index, k=10, metric="cosine")
        flags = engine_gemm.last_fallback_fraction(index, len(qs), 10)
    else:
        dist, idx, cnt, flags = _run(mods, db, qs, 10, "cosine", mode)
        flags = flags.mean()
    ref = O.distances_batch(qs, db, "cosine")
    for qi in range(len(qs)):
        O.check_topk(ref[qi], idx[qi], dist[qi], 10)
    assert flags <= 0.05


def test_gemm_falls_back_when_certificate_fails(mods):
    """All rows identical: every approximate value ties, nothing can be certified, the device-side exact scan
    answers and the tie rule (lowest index) still holds.  Sorted data: candidate buffers overflow -> same."""
    fpv, engine_gemm, ops = mods
    rng = np.random.default_rng(1)
    row = rng.standard_normal(64).astype(np.float32)
    db = np.repeat(row[None, :], 6000, axis=0)
    qs = rng.standard_normal((40, 64)).astype(np.float32)
    dist, idx, cnt, flags = _run(mods, db, qs, 10, "ip", "tf32")
    assert flags.all()
    assert (idx == np.arange(10)[None, :]).all()
    ref = O.distances_batch(qs, db, "ip")
    for qi in range(len(qs)):
        O.check_topk(ref[qi], idx[qi], dist[qi], 10)
    # rows sorted from worst to best for one direction: later slabs always beat the threshold
    base = rng.standard_normal(64).astype(np.float32)
    base /= np.linalg.norm(base)
    n = 40000
    noise = rng.standard_normal((n, 64)).astype(np.float32) * 0.01
    db2 = (np.linspace(0.1, 3.0, n, dtype=np.float32)[:, None] * base[None, :] + noise).astype(np.float32)
    qs2 = np.repeat(base[None, :], 20, axis=0) + rng.standard_normal((20, 64)).astype(np.float32) * 0.01
    dist, idx, cnt, flags = _run(mods, db2, qs2.astype(np.float32), 10, "ip", "tf32")
    ref = O.distances_batch(qs2.astype(np.float32), db2, "ip")
    for qi in range(len(qs2)):
        O.check_topk(ref[qi], idx[qi], dist[qi], 10)


def test_engine_dispatches_large_batches_to_gemm(mods):
    fpv, engine_gemm, ops = mods
    from fastpyvectordb_b200 import _native
    db, qs = _data(16384, 128, 64, True)
    eng = fpv.ParallelSearchEngine()
    idx_b, dist_b = eng.search_arrays(qs, db, k=10, metric="l2")           # batch >= GEMM_MIN_BATCH -> tensor-core path
    idx_s, dist_s = eng.search_arrays(qs[:8], db, k=10, metric="l2")       # a smaller batch of the same queries
    ref = O.distances_batch(qs, db, "l2")
    for qi in range(len(qs)):
        O.check_topk(ref[qi], idx_b[qi], dist_b[qi], 10, squared_near_zero=True)
    for qi in range(8):
        O.check_topk(ref[qi], idx_s[qi], dist_s[qi], 10, squared_near_zero=True)
