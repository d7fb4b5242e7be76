"""Collection / Filter layer: the vectorised filter compiler agrees with the per-row reference semantics (CPU), and
brute_force_search / search / query agree with the oracle restatement of vectordb_optimized.py:650-721 (GPU)."""
import numpy as np
import pytest

from fastpyvectordb_b200.collection import (Collection, CollectionConfig, DistanceMetric, DocumentCollection, Filter,
                                            FilterCondition, FilterOp, VectorDB, _Columns)
from oracle import oracle as O
import inputs as gi


def _metadata(n, seed=0):
    rng = np.random.default_rng(seed)
    cats = ["tech", "science", "art", "sport"]
    rows = []
    for i in range(n):
        m = {"category": cats[int(rng.integers(0, 4))], "price": float(rng.random() * 100), "rank": int(rng.integers(0, 10))}
        if i % 7 == 0:
            del m["price"]                                       # a missing field never matches
        if i % 11 == 0:
            m["tags"] = "alpha-beta" if i % 2 else "gamma"
        rows.append(m)
    return rows


FILTERS = [
    Filter.eq("category", "tech"),
    Filter.ne("category", "tech"),
    Filter.gt("price", 50), Filter.gte("rank", 5), Filter.lt("price", 10.5), Filter.lte("rank", 0),
    Filter.in_("category", ["art", "sport"]), Filter.nin("rank", [1, 2, 3]),
    Filter.contains("tags", "beta"), Filter.regex("category", "^s"),
    Filter.and_([Filter.eq("category", "science"), Filter.gt("price", 20)]),
    Filter.or_([Filter.eq("category", "art"), Filter.lt("rank", 2)]),
    Filter.not_(Filter.gt("price", 30)),
    Filter.from_dict({"category": "sport", "rank": 3}),
    Filter.from_dict({}),
    Filter(lambda m: m.get("rank", 0) % 2 == 0),
    Filter.eq("missing_field", 1),
]


@pytest.mark.parametrize("flt", FILTERS, ids=lambda f: f.kind)
def test_filter_mask_equals_per_row_evaluate(flt):
    rows = _metadata(500)
    mask = flt.mask(_Columns(rows))
    expect = np.array([flt.evaluate(m) for m in rows])
    assert mask.dtype == bool and np.array_equal(mask, expect)


def test_filter_mask_with_list_tuple_and_huge_int_values():
    """a list / tuple expected value must be compared as ONE value per row (not broadcast against the
    column), and ints beyond 2**53 must not be compared through float64."""
    rows = [{"tags": ["a"], "big": 2 ** 60}, {"tags": "a", "big": 2 ** 60 + 1}, {"tags": ["a", "b"], "big": 5},
            {"tags": ["a"], "big": 2 ** 60}, {"tags": ("a",)}, {}]
    for n_rows in (len(rows), 1):
        sub = rows[:n_rows]
        for flt in (Filter.eq("tags", ["a"]), Filter.ne("tags", ["a"]), Filter.eq("tags", ("a",)), Filter.eq("tags", "a"),
                    Filter.eq("tags", ["a", "b", "c", "d", "e", "f"][:n_rows]), Filter.eq("big", 2 ** 60),
                    Filter.ne("big", 2 ** 60), Filter.gt("big", 2 ** 60), Filter.from_dict({"tags": ["a"]})):
            expect = np.array([flt.evaluate(m) for m in sub], dtype=bool)
            assert np.array_equal(flt.mask(_Columns(sub)), expect), (flt.kind, getattr(flt.cond, "value", None), n_rows)


def test_filter_condition_semantics():
    c = FilterCondition("x", FilterOp.GT, 3)
    assert c.evaluate({"x": 4}) and not c.evaluate({"x": 3}) and not c.evaluate({})
    assert FilterCondition("s", FilterOp.CONTAINS, "ell").evaluate({"s": "hello"})
    assert FilterCondition("s", FilterOp.REGEX, r"^h.*o$").evaluate({"s": "hello"})


def test_collection_batch_bookkeeping_without_a_gpu():
    """Host-side id / metadata / row bookkeeping (vectordb_optimized.py:441-504, 737-739) needs no device: the engine
    is only touched by searches."""
    c = Collection(CollectionConfig(name="t", dimensions=4), engine=object())
    rows = np.eye(4, dtype=np.float32).repeat(2, axis=0)
    c.insert_batch(rows, ids=[f"v{i}" for i in range(8)], metadata_list=[{"i": i} for i in range(8)])
    assert c.delete_batch(["v1", "v6", "nope", "v1"]) == 2 and c.count() == 6
    assert c.list_ids() == ["v0", "v2", "v3", "v4", "v5", "v7"]
    got = c.get_batch(["v0", "nope", "v7"], include_vectors=True)
    assert got[1] is None and got[0]["metadata"] == {"i": 0} and np.array_equal(got[2]["vector"], rows[7])
    plain = c.get_batch(["v2", "nope"])
    assert plain[0].id == "v2" and plain[0].metadata == {"i": 2} and plain[1] is None
    assert c.get("v3", include_vector=True).vector.tolist() == rows[3].tolist()      # rows were compacted consistently
    assert c.delete_batch([]) == 0 and c.delete_batch(["gone"]) == 0
    c.set_ef_search(77)
    assert c.config.ef_search == 77


def test_document_collection_crud_without_a_gpu():
    """get / peek / update / upsert / delete of fastpyvectordb/client.py:161-182, 276-445 (host-side bookkeeping; the
    ``where`` filters go through the vectorised compiler)."""
    dc = DocumentCollection(Collection(CollectionConfig(name="t", dimensions=4), engine=object()))
    dc.add(ids=["a", "b", "c", "d"], embeddings=np.eye(4, dtype=np.float32),
           metadatas=[{"cat": "x", "n": 1}, {"cat": "y", "n": 2}, {"cat": "x", "n": 3}, {"cat": "y", "n": 4}],
           documents=["A", "B", "C", "D"])
    assert len(dc) == 4 and dc.get(ids=["b", "zz"]).documents == ["B"]
    g = dc.get(where={"cat": "x"}, include=["metadatas", "embeddings"])
    assert g.ids == ["a", "c"] and g.documents == [None, None] and g.metadatas[1] == {"cat": "x", "n": 3}
    assert np.array_equal(g.embeddings[1], np.eye(4, dtype=np.float32)[2])
    assert dc.get(where={"cat": "y"}, limit=1, offset=1).ids == ["d"] and dc.peek(2).ids == ["a", "b"]
    dc.update(["b"], metadatas=[{"n": 20}], documents=["B2"], embeddings=[[0, 0, 0, 9]])
    got = dc.get(ids="b", include=["documents", "metadatas", "embeddings"])
    assert got.documents == ["B2"] and got.metadatas == [{"cat": "y", "n": 20}] and got.embeddings[0].tolist() == [0, 0, 0, 9]
    with pytest.raises(ValueError):
        dc.update(["nope"], metadatas=[{}])
    dc.upsert(ids=["a", "e"], embeddings=np.ones((2, 4), np.float32), metadatas=[{"cat": "z"}, {"cat": "z"}])
    assert dc.count() == 5 and sorted(dc.get(where={"cat": "z"}).ids) == ["a", "e"]
    assert dc.delete(where={"cat": "y"}) == 2 and dc.delete(ids="c") == 1 and dc.count() == 2
    with pytest.raises(ValueError):
        dc.delete()


@pytest.mark.gpu
@pytest.mark.parametrize("metric", ["cosine", "l2", "ip"])
def test_collection_exact_search_against_oracle(metric):
    n, d = 3000, 48
    rng = np.random.default_rng(42)
    vecs = rng.standard_normal((n, d)).astype(np.float32)
    meta = _metadata(n, 1)
    db = VectorDB()
    col = db.create_collection("docs", dimensions=d, metric=metric)
    ids = col.insert_batch(vecs, [f"id{i}" for i in range(n)], meta)
    assert ids[:2] == ["id0", "id1"] and col.count() == n and len(col) == n
    q = np.random.default_rng(999).standard_normal(d).astype(np.float32)
    ref = O.brute_force_distances(q, vecs, metric)
    for flt in (None, {"category": "tech"}, Filter.and_([Filter.gt("price", 40), Filter.ne("category", "art")])):
        res = col.brute_force_search(q, k=10, filter=flt)
        f = Filter.from_dict(flt) if isinstance(flt, dict) else flt
        valid = np.array([f.evaluate(m) for m in meta]) if f is not None else None
        got_idx = [int(r.id[2:]) for r in res]
        O.check_topk(ref, got_idx, [r.score for r in res], 10, valid=valid, rtol=1e-5, squared_near_zero=(metric == "l2"))
        assert all(r.metadata is meta[i] or r.metadata == meta[i] for r, i in zip(res, got_idx))
        res2 = col.search(q, k=10, filter=flt)
        assert [r.id for r in res2] == [r.id for r in res]
    assert col.brute_force_search(q, k=5, filter={"category": "nope"}) == []
    with pytest.raises(ValueError):
        col.search(np.zeros(d + 1, np.float32))
    # writes invalidate the resident copy
    col.upsert(q, "id7", {"category": "tech", "price": 1.0, "rank": 0})
    top = col.search(q, k=1)[0]
    assert top.id == "id7"
    assert col.delete("id7") and col.get("id7") is None and col.count() == n - 1
    assert col.search(q, k=1)[0].id != "id7"
    with pytest.raises(ValueError):
        col.insert(q, "id8")
    assert db.list_collections() == ["docs"] and db.get_collection("docs") is col


@pytest.mark.gpu
def test_document_collection_query_surface():
    n, d = 500, 32
    rng = np.random.default_rng(3)
    vecs = rng.standard_normal((n, d)).astype(np.float32)
    col = Collection(CollectionConfig("c", d, DistanceMetric.COSINE))
    docs = DocumentCollection(col)
    docs.add([f"d{i}" for i in range(n)], vecs, [{"category": "tech" if i % 2 else "art"} for i in range(n)],
             [f"text {i}" for i in range(n)])
    out = docs.query(query_embeddings=vecs[:3].tolist(), n_results=5, where={"category": "tech"})
    assert len(out.ids) == 3 and all(len(r) == 5 for r in out.ids)
    assert out.ids[1][0] == "d1" and abs(out.distances[1][0]) < 1e-5          # a row is its own nearest neighbour
    assert all(m == {"category": "tech"} for row in out.metadatas for m in row)     # "_document" is hidden
    assert out.documents[1][0] == "text 1" and out.embeddings is None
    with pytest.raises(ValueError):
        docs.query()
    out = docs.query(query_embeddings=[vecs[0]], n_results=2, include=["embeddings", "distances"])
    assert out.embeddings is not None and out.documents == [[None, None]] and out.metadatas == [[{}, {}]]


@pytest.mark.gpu
@pytest.mark.parametrize("case", gi.BRUTE_CASES, ids=lambda c: c["name"])
@pytest.mark.parametrize("metric", ["cosine", "l2", "ip"])
def test_brute_force_search_against_reference_outputs(golden, case, metric):
    """Collection.brute_force_search against the committed outputs of the reference's OWN method
    (vectordb_optimized.py:650-721, generated by tests/golden/make_golden.py through the hnswlib shim): same ids (tie
    aware), scores within 1e-5, same result counts under the filters; a stored duplicate of the query scores exactly 0
    under L2, as it does in the reference."""
    db, qs, ids, meta = gi.brute_inputs(case)
    col = Collection(CollectionConfig(name="g", dimensions=case["d"], metric=DistanceMetric(metric)))
    col.insert_batch(db, ids=ids, metadata_list=meta)
    ref_all = [O.brute_force_distances(q, db, metric) for q in qs]
    for fname in gi.BRUTE_FILTERS:
        flt = gi.brute_filter(Filter, fname)
        f = Filter.from_dict(flt) if isinstance(flt, dict) else flt
        valid = np.array([f.evaluate(m) for m in meta]) if f is not None else None
        tag = f"{case['name']}/{metric}/{fname}"
        for qi, q in enumerate(qs):
            res = col.brute_force_search(q, k=case["k"], filter=flt)
            cnt = int(golden[tag + "/count"][qi])
            assert len(res) == cnt
            got_idx = [int(r.id[1:]) for r in res]
            got = np.array([r.score for r in res])
            want = golden[tag + "/score"][qi][:cnt]
            # scores position by position (both lists are ascending): 1e-5 relative, near-zero L2 compared squared
            lim = 1e-5 * np.maximum(np.abs(want), 1.0)
            if metric == "l2":
                small = want < 1e-2
                assert np.all(np.abs(got[~small] - want[~small]) <= lim[~small])
                assert np.all(np.abs(got[small] ** 2 - want[small] ** 2) <= 1e-5)
                assert np.array_equal(got == 0.0, want == 0.0)
            else:
                assert np.all(np.abs(got - want) <= lim + (2e-6 if metric == "cosine" else 0.0))
            O.check_topk(ref_all[qi], got_idx, got, case["k"], valid=valid, rtol=1e-5, squared_near_zero=(metric == "l2"))
            # the reference's own answer passes the same checker against the oracle restatement (pins the oracle)
            ridx = golden[tag + "/idx"][qi][:cnt]
            O.check_topk(ref_all[qi], ridx, want, case["k"], valid=valid, rtol=1e-5, squared_near_zero=(metric == "l2"))
