#!/usr/bin/env python
"""Generate tests/golden/*.npz by running the UNMODIFIED reference in the build container.

    python tests/golden/make_golden.py            # needs /root/reference (read-only)

The reference has no tests or golden vectors for its search hot path, so the
oracle is pinned against outputs of the reference itself.  The reference cannot travel to the
GPU box; these small fixtures (and this script) are what is committed.  Inputs are regenerated
from the seeds stored in each file (``tests/golden/inputs.py``), outputs are stored verbatim.
"""
import importlib.util
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import inputs as gi  # noqa: E402

REF = os.environ.get("FPV_REFERENCE", "/root/reference")


def _load(name):
    spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF, name + ".py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["ref_" + name] = mod
    spec.loader.exec_module(mod)
    return mod


def main():
    ps = _load("parallel_search")
    qz = _load("quantization")
    out = {}

    # ---- float path ---------------------------------------------------------------------
    for case in gi.FLOAT_CASES:
        db, qs, mask = gi.float_inputs(case)
        eng = ps.ParallelSearchEngine(n_workers=2, chunk_size=case["chunk"])
        tag = case["name"]
        for metric in ("cosine", "l2", "ip"):
            out[f"{tag}/{metric}/dist_single"] = np.stack(
                [ps._compute_distances_vectorized(q, db, metric) for q in qs])
            out[f"{tag}/{metric}/dist_chunk"] = np.stack(
                [ps._compute_distances_chunk((q, db, 7, metric)) for q in qs[:2]])
            res = eng.search_batch_parallel(qs, db, k=case["k"], metric=metric)
            out[f"{tag}/{metric}/batch_idx"] = np.array([[r.index for r in row] for row in res], np.int64)
            out[f"{tag}/{metric}/batch_dist"] = np.array([[r.distance for r in row] for row in res], np.float64)
            res = [eng.search_parallel(q, db, k=case["k"], metric=metric) for q in qs]
            out[f"{tag}/{metric}/single_idx"] = np.array([[r.index for r in row] for row in res], np.int64)
            out[f"{tag}/{metric}/single_dist"] = np.array([[r.distance for r in row] for row in res], np.float64)
            res = [eng.search_parallel(q, db, k=case["k"], metric=metric, filter_mask=mask) for q in qs]
            out[f"{tag}/{metric}/masked_idx"] = np.array([[r.index for r in row] for row in res], np.int64)
            out[f"{tag}/{metric}/masked_dist"] = np.array([[r.distance for r in row] for row in res], np.float64)
            res = [eng.search_chunked_parallel(q, db, k=case["k"], metric=metric) for q in qs]
            out[f"{tag}/{metric}/chunked_idx"] = np.array([[r.index for r in row] for row in res], np.int64)
            out[f"{tag}/{metric}/chunked_dist"] = np.array([[r.distance for r in row] for row in res], np.float64)
    blocks = gi.merge_inputs()
    out["merge/k5"] = ps._merge_top_k(blocks, 5)
    out["merge/k100"] = ps._merge_top_k(blocks, 100)

    # ---- scalar quantizer ---------------------------------------------------------------
    for case in gi.SQ_CASES:
        train, db, qs = gi.sq_inputs(case)
        sq = qz.ScalarQuantizer().train(train)
        tag = case["name"]
        codes = sq.encode(db)
        out[f"{tag}/min"], out[f"{tag}/max"], out[f"{tag}/scale"] = sq.min_vals, sq.max_vals, sq.scale
        out[f"{tag}/codes"] = codes
        out[f"{tag}/qcodes"] = np.stack([sq.encode_query(q) for q in qs])
        out[f"{tag}/decode"] = sq.decode(codes[:16])
        out[f"{tag}/l2"] = np.stack([sq.distances_l2(q, codes) for q in qs])
        out[f"{tag}/dot"] = np.stack([sq.distances_dot(q, codes) for q in qs])
        out[f"{tag}/cosine"] = np.stack([sq.distances_cosine(q, codes) for q in qs])

    # ---- binary quantizer ---------------------------------------------------------------
    for case in gi.BQ_CASES:
        train, db, qs = gi.bq_inputs(case)
        bq = qz.BinaryQuantizer(threshold=case.get("threshold", 0.0))
        if case["train"]:
            bq.train(train, use_median=case["median"])
        else:
            bq.dimensions = case["dims_attr"]
        tag = case["name"]
        codes = bq.encode(db)
        out[f"{tag}/codes"] = codes
        if bq.thresholds is not None:
            out[f"{tag}/thresholds"] = np.asarray(bq.thresholds)
        out[f"{tag}/qbits"] = np.stack([bq.encode_query(q) for q in qs])
        out[f"{tag}/hamming"] = np.stack([bq.hamming_distances(bq.encode_query(q), codes) for q in qs])
        res = [bq.search(q, codes, k=case["k"]) for q in qs]
        out[f"{tag}/search_idx"] = np.stack([r[0] for r in res]).astype(np.int64)
        out[f"{tag}/search_dist"] = np.stack([r[1] for r in res])

    # ---- product quantizer --------------------------------------------------------------
    for case in gi.PQ_CASES:
        cb, db, qs = gi.pq_inputs(case)
        pq = qz.ProductQuantizer(case["d"], case["m"], case["kc"])
        pq.codebooks = cb
        pq.trained = True
        tag = case["name"]
        codes = pq.encode(db)
        out[f"{tag}/codes"] = codes
        out[f"{tag}/lut"] = np.stack([pq.build_lookup_table(q) for q in qs])
        out[f"{tag}/dist"] = np.stack([pq.distances_with_table(pq.build_lookup_table(q), codes) for q in qs])
        res = [pq.search(q, codes, k=case["k"]) for q in qs]
        out[f"{tag}/search_idx"] = np.stack([r[0] for r in res]).astype(np.int64)
        out[f"{tag}/search_dist"] = np.stack([r[1] for r in res])
    # k-means (global np.random state, quantization.py:458,486,494)
    data = gi.kmeans_inputs()
    np.random.seed(gi.KMEANS_SEED)
    pq = qz.ProductQuantizer(data.shape[1], 1, gi.KMEANS_K)
    out["kmeans/centroids"] = pq._kmeans(data, gi.KMEANS_K, gi.KMEANS_ITERS)

    # ---- vectordb_optimized.Collection.brute_force_search, the UNMODIFIED reference method ---------------------------
    # (the module hard-imports the third-party hnswlib, which cannot be installed here: a label -> vector shim stands in
    # for it; brute_force_search itself only reads the stored rows back through get_items)
    import tempfile
    from pathlib import Path
    import hnswlib_shim
    sys.modules.setdefault("hnswlib", hnswlib_shim)
    vo = _load("vectordb_optimized")
    for case in gi.BRUTE_CASES:
        db, qs, ids, meta = gi.brute_inputs(case)
        for metric in ("cosine", "l2", "ip"):
            with tempfile.TemporaryDirectory() as tmp:
                cfg = vo.CollectionConfig(name="g", dimensions=case["d"], metric=vo.DistanceMetric(metric), max_elements=case["n"] + 10)
                col = vo.Collection(cfg, Path(tmp))
                col.insert_batch(db, ids=ids, metadata_list=meta)
                for fname in gi.BRUTE_FILTERS:
                    flt = gi.brute_filter(vo.Filter, fname)
                    rows, scores, counts = [], [], []
                    for q in qs:
                        res = col.brute_force_search(q, k=case["k"], filter=flt)
                        counts.append(len(res))
                        r = [int(x.id[1:]) for x in res] + [-1] * (case["k"] - len(res))
                        sc = [x.score for x in res] + [np.inf] * (case["k"] - len(res))
                        rows.append(r); scores.append(sc)
                    tag = f"{case['name']}/{metric}/{fname}"
                    out[tag + "/idx"] = np.array(rows, np.int64)
                    out[tag + "/score"] = np.array(scores, np.float64)
                    out[tag + "/count"] = np.array(counts, np.int64)

    np.savez_compressed(os.path.join(HERE, "reference_outputs.npz"), **out)
    total = sum(v.nbytes for v in out.values())
    print(f"wrote {len(out)} arrays, {total/1e6:.2f} MB raw")


if __name__ == "__main__":
    main()
