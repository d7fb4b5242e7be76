"""Seeded input generators shared by make_golden.py (reference side) and the tests (oracle / CUDA side).

Generators follow the reference's benchmark harness: DB ``default_rng(42).standard_normal`` row
normalised, queries ``default_rng(999)`` (examples/benchmark_parallel.py:212-217, 329-330), plus raw
(un-normalised) variants so that L2 / IP are not degenerate with cosine (README.md:360-361).
"""
import numpy as np

FLOAT_CASES = [
    dict(name="f32_unit", n=1500, d=64, q=6, k=10, chunk=400, unit=True, seed=42),
    dict(name="f32_raw_odd", n=777, d=37, q=4, k=100, chunk=300, unit=False, seed=43),
    dict(name="f32_kgeN", n=40, d=16, q=3, k=64, chunk=50000, unit=False, seed=44),
]

SQ_CASES = [
    dict(name="sq_a", n=600, d=48, q=4, seed=51),
    dict(name="sq_clip", n=300, d=20, q=3, seed=52, out_of_range=True),
]

BQ_CASES = [
    dict(name="bq_median", n=900, d=128, q=4, k=25, seed=61, train=True, median=True),
    dict(name="bq_odd", n=500, d=100, q=3, k=10, seed=62, train=True, median=False, threshold=0.05),
    dict(name="bq_untrained", n=300, d=70, q=2, k=400, seed=63, train=False, median=False, dims_attr=70),
]

PQ_CASES = [
    dict(name="pq_m8", n=400, d=64, m=8, kc=256, q=3, k=10, seed=71),
    dict(name="pq_m48", n=300, d=768, m=48, kc=256, q=2, k=100, seed=72),
    dict(name="pq_small_k", n=200, d=24, m=6, kc=16, q=2, k=300, seed=73),
]

KMEANS_SEED, KMEANS_K, KMEANS_ITERS = 1234, 8, 3

# vectordb_optimized.Collection.brute_force_search (vectordb_optimized.py:650-721), run through tests/golden/hnswlib_shim.py
BRUTE_CASES = [
    dict(name="brute_unit", n=1200, d=48, q=5, k=10, unit=True, seed=81),
    dict(name="brute_raw", n=700, d=33, q=4, k=100, unit=False, seed=82),
]
BRUTE_FILTERS = {"none": None, "cat": {"category": "tech"}, "price_and_cat": "and"}   # "and": built by brute_filter()


def _unit(x):
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def float_inputs(case):
    rng = np.random.default_rng(case["seed"])
    db = rng.standard_normal((case["n"], case["d"])).astype(np.float32)
    qs = np.random.default_rng(999).standard_normal((case["q"], case["d"])).astype(np.float32)
    if case["unit"]:
        db, qs = _unit(db), _unit(qs)
    else:
        db *= rng.uniform(0.2, 3.0, size=(case["n"], 1)).astype(np.float32)
    # a few exact duplicates / a zero row exercise ties and the eps terms
    db[5] = db[3]
    db[11] = db[3]
    if not case["unit"]:
        db[7] = 0.0
    mask = np.random.default_rng(11).random(case["n"]) < 0.25
    return np.ascontiguousarray(db), np.ascontiguousarray(qs), mask


def merge_inputs():
    rng = np.random.default_rng(5)
    blocks = []
    base = 0
    for n in (10, 3, 25, 1, 17):
        ids = np.arange(base, base + n, dtype=np.float64)
        d = np.round(rng.random(n) * 8) / 8  # plenty of ties
        blocks.append(np.column_stack([ids, d]))
        base += n
    return blocks


def sq_inputs(case):
    rng = np.random.default_rng(case["seed"])
    db = rng.standard_normal((case["n"], case["d"])).astype(np.float32)
    db[:, 3] = 0.25                      # constant dimension -> scale 1.0 branch
    qs = np.random.default_rng(999).standard_normal((case["q"], case["d"])).astype(np.float32)
    if case.get("out_of_range"):
        qs *= 4.0                        # forces clipping in encode_query
    return db, db, qs


def bq_inputs(case):
    rng = np.random.default_rng(case["seed"])
    db = rng.standard_normal((case["n"], case["d"])).astype(np.float32)
    db[9] = db[2]
    qs = np.random.default_rng(999).standard_normal((case["q"], case["d"])).astype(np.float32)
    return db, db, qs


def pq_inputs(case):
    rng = np.random.default_rng(case["seed"])
    dsub = case["d"] // case["m"]
    cb = (rng.standard_normal((case["m"], case["kc"], dsub)) / np.sqrt(case["d"])).astype(np.float32)
    db = _unit(rng.standard_normal((case["n"], case["d"])).astype(np.float32))
    qs = _unit(np.random.default_rng(999).standard_normal((case["q"], case["d"])).astype(np.float32))
    return cb, db, qs


def kmeans_inputs():
    return np.random.default_rng(81).standard_normal((120, 6)).astype(np.float32)


def brute_inputs(case):
    """(db, queries, ids, metadata list) of a brute_force_search case; row 3 is duplicated at 5 and 11 (exact ties and an
    exactly-zero L2 distance when a query equals a stored row: the last query IS row 3)."""
    rng = np.random.default_rng(case["seed"])
    db = rng.standard_normal((case["n"], case["d"])).astype(np.float32)
    qs = np.random.default_rng(999).standard_normal((case["q"], case["d"])).astype(np.float32)
    if case["unit"]:
        db, qs = _unit(db), _unit(qs)
    else:
        db *= rng.uniform(0.2, 3.0, size=(case["n"], 1)).astype(np.float32)
    db[5] = db[3]
    db[11] = db[3]
    qs[-1] = db[3]
    cats = ["tech", "science", "art", "sport"]
    meta = [{"category": cats[int(rng.integers(0, 4))], "price": float(np.float32(rng.random() * 100)), "rank": int(rng.integers(0, 10))}
            for _ in range(case["n"])]
    ids = [f"v{i:05d}" for i in range(case["n"])]
    return np.ascontiguousarray(db), np.ascontiguousarray(qs), ids, meta


def brute_filter(FilterCls, name):
    """The filter object of BRUTE_FILTERS[name], built with the given Filter class (the reference's or ours)."""
    if name == "none":
        return None
    if name == "cat":
        return {"category": "tech"}
    return FilterCls.and_([FilterCls.eq("category", "science"), FilterCls.gt("price", 30.0)])
