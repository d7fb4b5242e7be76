"""``Collection`` / ``VectorDB`` served by the exact GPU engine — the caller side of the hot path.

The reference's ``vectordb_optimized.Collection`` stores vectors inside an ``hnswlib`` index and answers
``search`` approximately; its exact path is ``brute_force_search`` (vectordb_optimized.py:650-721: NumPy distances,
a per-row Python ``Filter.evaluate`` loop, ``np.where(mask, d, inf)``, argpartition).  Here the collection keeps an
(N, D) fp32 matrix resident in HBM and BOTH ``search`` and ``brute_force_search`` are the exact fused GPU search,
so the call surface that sits above the path keeps working without ``hnswlib``:

    Collection.insert / insert_batch / upsert / get / get_batch / delete / delete_batch / count / list_ids
                                                                        (vectordb_optimized.py:337-505)
    Collection.search / search_batch / brute_force_search / set_ef_search               (:507-739)
    Filter / FilterCondition / FilterOp / SearchResult / CollectionConfig / DistanceMetric (:40-200)
    DocumentCollection.query(query_embeddings=..., n_results, where, include) -> QueryResult
    DocumentCollection.add / upsert / get / peek / update / delete / count -> GetResult
                                                                        (fastpyvectordb/client.py:90-445)

Filters are compiled to a row bitmask with vectorised column predicates (one NumPy comparison per condition)
instead of N Python calls, and the mask is applied inside the kernel.  Score conventions are the exact path's
(cosine ``1 - cos``, L2 with the square root, ``-dot``).  Storage engine, HNSW parameters and persistence are out of
scope: ``M`` / ``ef_*`` are accepted and ignored, ``base_path`` is not written to.
"""
from __future__ import annotations

import re
import threading
import uuid
from dataclasses import dataclass, field
from enum import Enum
from typing import Any, Dict, List, Optional, Sequence, Union

import numpy as np

from .engine import GpuIndex, ParallelSearchEngine


@dataclass
class SearchResult:
    """A single search result (vectordb_optimized.py:40-46)."""
    id: str
    score: float
    metadata: dict = field(default_factory=dict)
    vector: Optional[np.ndarray] = None


class DistanceMetric(Enum):
    COSINE = "cosine"
    EUCLIDEAN = "l2"
    DOT_PRODUCT = "ip"


class FilterOp(Enum):
    EQ = "eq"
    NE = "ne"
    GT = "gt"
    GTE = "gte"
    LT = "lt"
    LTE = "lte"
    IN = "in"
    NIN = "nin"
    CONTAINS = "contains"
    REGEX = "regex"


_MISSING = object()


def _scalar_test(op: FilterOp, actual, expected) -> bool:
    if op is FilterOp.EQ:
        return actual == expected
    if op is FilterOp.NE:
        return actual != expected
    if op is FilterOp.GT:
        return actual > expected
    if op is FilterOp.GTE:
        return actual >= expected
    if op is FilterOp.LT:
        return actual < expected
    if op is FilterOp.LTE:
        return actual <= expected
    if op is FilterOp.IN:
        return actual in expected
    if op is FilterOp.NIN:
        return actual not in expected
    if op is FilterOp.CONTAINS:
        return expected in str(actual)
    if op is FilterOp.REGEX:
        return bool(re.search(expected, str(actual)))
    return False


@dataclass
class FilterCondition:
    """field <op> value; a row without the field never matches (vectordb_optimized.py:73-105)."""
    field: str
    op: FilterOp
    value: Any

    def evaluate(self, metadata: dict) -> bool:
        actual = metadata.get(self.field, _MISSING)
        if actual is _MISSING:
            return False
        return bool(_scalar_test(self.op, actual, self.value))


class Filter:
    """Composable filter expression with the reference's constructors (vectordb_optimized.py:108-184).  Kept as an
    expression tree (not closures) so that it can be compiled to a vectorised row mask."""

    def __init__(self, kind: str = "true", cond: FilterCondition = None, children: Sequence["Filter"] = (), fn=None):
        self.kind, self.cond, self.children, self.fn = kind, cond, list(children), fn
        if callable(kind):                      # Filter(callable) like the reference's constructor
            self.kind, self.fn = "fn", kind

    # ---- evaluation on one metadata dict (reference semantics) ----
    def evaluate(self, metadata: dict) -> bool:
        k = self.kind
        if k == "true":
            return True
        if k == "cond":
            return self.cond.evaluate(metadata)
        if k == "and":
            return all(c.evaluate(metadata) for c in self.children)
        if k == "or":
            return any(c.evaluate(metadata) for c in self.children)
        if k == "not":
            return not self.children[0].evaluate(metadata)
        return bool(self.fn(metadata))

    # ---- vectorised evaluation over a whole collection ----
    def mask(self, columns: "_Columns") -> np.ndarray:
        k = self.kind
        n = columns.n
        if k == "true":
            return np.ones(n, bool)
        if k == "cond":
            return columns.test(self.cond)
        if k == "and":
            out = np.ones(n, bool)
            for c in self.children:
                out &= c.mask(columns)
            return out
        if k == "or":
            out = np.zeros(n, bool)
            for c in self.children:
                out |= c.mask(columns)
            return out
        if k == "not":
            return ~self.children[0].mask(columns)
        return np.fromiter((bool(self.fn(m)) for m in columns.rows), bool, n)

    # ---- constructors ----
    @staticmethod
    def _c(fieldname, op, value):
        return Filter("cond", FilterCondition(fieldname, op, value))

    @staticmethod
    def eq(field: str, value: Any) -> "Filter":
        return Filter._c(field, FilterOp.EQ, value)

    @staticmethod
    def ne(field: str, value: Any) -> "Filter":
        return Filter._c(field, FilterOp.NE, value)

    @staticmethod
    def gt(field: str, value: Any) -> "Filter":
        return Filter._c(field, FilterOp.GT, value)

    @staticmethod
    def gte(field: str, value: Any) -> "Filter":
        return Filter._c(field, FilterOp.GTE, value)

    @staticmethod
    def lt(field: str, value: Any) -> "Filter":
        return Filter._c(field, FilterOp.LT, value)

    @staticmethod
    def lte(field: str, value: Any) -> "Filter":
        return Filter._c(field, FilterOp.LTE, value)

    @staticmethod
    def in_(field: str, values: list) -> "Filter":
        return Filter._c(field, FilterOp.IN, values)

    @staticmethod
    def nin(field: str, values: list) -> "Filter":
        return Filter._c(field, FilterOp.NIN, values)

    @staticmethod
    def contains(field: str, substring: str) -> "Filter":
        return Filter._c(field, FilterOp.CONTAINS, substring)

    @staticmethod
    def regex(field: str, pattern: str) -> "Filter":
        return Filter._c(field, FilterOp.REGEX, pattern)

    @staticmethod
    def and_(filters: list) -> "Filter":
        return Filter("and", children=filters)

    @staticmethod
    def or_(filters: list) -> "Filter":
        return Filter("or", children=filters)

    @staticmethod
    def not_(filter_: "Filter") -> "Filter":
        return Filter("not", children=[filter_])

    @staticmethod
    def from_dict(d: dict) -> "Filter":
        if not d:
            return Filter("true")
        return Filter.and_([Filter.eq(k, v) for k, v in d.items()])


class _Columns:
    """Column view of the per-row metadata dicts, built lazily per field and dropped when rows change."""

    _NUMERIC = (int, float, np.integer, np.floating)

    def __init__(self, rows: List[dict]):
        self.rows = rows
        self.n = len(rows)
        self._cache: Dict[str, tuple] = {}

    def _column(self, name: str):
        hit = self._cache.get(name)
        if hit is None:
            vals = [m.get(name, _MISSING) for m in self.rows]
            present = np.fromiter((v is not _MISSING for v in vals), bool, self.n)
            numeric = all(self._float64_exact(v) for v in vals if v is not _MISSING)
            arr = None
            if numeric and present.any():
                arr = np.array([v if v is not _MISSING else np.nan for v in vals], dtype=np.float64)
            hit = (vals, present, arr)
            self._cache[name] = hit
        return hit

    @classmethod
    def _float64_exact(cls, v) -> bool:
        """True when comparing ``v`` against the float64 column is the same as comparing the original Python values."""
        if isinstance(v, bool) or not isinstance(v, cls._NUMERIC):
            return False
        if isinstance(v, (int, np.integer)):
            return abs(int(v)) <= 2 ** 53
        return True

    def test(self, cond: FilterCondition) -> np.ndarray:
        vals, present, arr = self._column(cond.field)
        op, exp = cond.op, cond.value
        if arr is not None and self._float64_exact(exp) and op in (
                FilterOp.EQ, FilterOp.NE, FilterOp.GT, FilterOp.GTE, FilterOp.LT, FilterOp.LTE):
            with np.errstate(invalid="ignore"):
                res = {FilterOp.EQ: arr == exp, FilterOp.NE: arr != exp, FilterOp.GT: arr > exp, FilterOp.GTE: arr >= exp,
                       FilterOp.LT: arr < exp, FilterOp.LTE: arr <= exp}[op]
            return res & present
        # vectorised object compare only for SCALAR expected values: a list / tuple / array would broadcast against the
        # column instead of being compared as one value per row (Filter.evaluate semantics, vectordb_optimized.py:139-156)
        if (op is FilterOp.EQ or op is FilterOp.NE) and (exp is None or isinstance(exp, (str, bytes, bool, int, float,
                                                                                      np.integer, np.floating, np.bool_))):
            try:
                obj = np.empty(self.n, dtype=object)
                obj[:] = vals
                res = obj == exp if op is FilterOp.EQ else obj != exp
                return np.asarray(res, bool) & present
            except Exception:
                pass
        out = np.zeros(self.n, bool)
        for i, v in enumerate(vals):
            if v is not _MISSING:
                try:
                    out[i] = bool(_scalar_test(op, v, exp))
                except TypeError:
                    out[i] = False
        return out


@dataclass
class CollectionConfig:
    """Configuration for a collection (vectordb_optimized.py:191-200); the HNSW knobs are accepted and ignored."""
    name: str
    dimensions: int
    metric: DistanceMetric = DistanceMetric.COSINE
    M: int = 16
    ef_construction: int = 200
    ef_search: int = 50
    max_elements: int = 1_000_000


class Collection:
    """In-memory collection whose every search is the exact GPU search (see module docstring)."""

    def __init__(self, config: CollectionConfig, base_path=None, device=None, engine: ParallelSearchEngine = None):
        self.config = config
        self.base_path = base_path
        self._lock = threading.RLock()                    # writes are serialised, reads lock-free (vectordb_optimized.py:224)
        self._engine = engine or ParallelSearchEngine(device=device)
        self._rows = np.zeros((0, config.dimensions), np.float32)
        self._ids: List[str] = []
        self._row_of: Dict[str, int] = {}
        self._meta: List[dict] = []
        self._index: Optional[GpuIndex] = None            # device copy, rebuilt lazily after writes (the _rebuild_cache analogue)
        self._columns: Optional[_Columns] = None

    # ------------------------------------------------------------------ writes
    def _touch(self):
        self._index = None
        self._columns = None

    def insert(self, vector: np.ndarray, id: str = None, metadata: dict = None) -> str:
        return self.insert_batch(np.asarray(vector, np.float32).reshape(1, -1), [id] if id is not None else None,
                                 [metadata] if metadata is not None else None)[0]

    def insert_batch(self, vectors: np.ndarray, ids: List[str] = None, metadata_list: List[dict] = None) -> List[str]:
        vectors = np.asarray(vectors, dtype=np.float32)
        if vectors.ndim != 2 or vectors.shape[1] != self.config.dimensions:
            raise ValueError(f"Vectors have {vectors.shape[-1]} dimensions, expected {self.config.dimensions}")
        n = len(vectors)
        ids = list(ids) if ids is not None else [str(uuid.uuid4()) for _ in range(n)]
        metadata_list = list(metadata_list) if metadata_list is not None else [{} for _ in range(n)]
        if len(ids) != n or len(metadata_list) != n:
            raise ValueError("ids / metadata_list length must match the number of vectors")
        with self._lock:
            for i in ids:
                if i in self._row_of:
                    raise ValueError(f"ID '{i}' already exists. Use upsert to update.")
            if len(set(ids)) != n:
                raise ValueError("duplicate ids in batch")
            base = len(self._ids)
            self._rows = np.concatenate([self._rows, vectors], axis=0)
            for j, (i, m) in enumerate(zip(ids, metadata_list)):
                self._row_of[i] = base + j
                self._ids.append(i)
                self._meta.append(dict(m or {}))
            self._touch()
        return ids

    def upsert(self, vector: np.ndarray, id: str, metadata: dict = None) -> str:
        with self._lock:
            row = self._row_of.get(id)
            if row is None:
                return self.insert(vector, id, metadata)
            self._rows[row] = np.asarray(vector, np.float32).reshape(-1)
            if metadata is not None:
                self._meta[row] = dict(metadata)
            self._touch()
        return id

    def delete(self, id: str) -> bool:
        with self._lock:
            row = self._row_of.pop(id, None)
            if row is None:
                return False
            self._rows = np.delete(self._rows, row, axis=0)
            del self._ids[row]
            del self._meta[row]
            for j in range(row, len(self._ids)):
                self._row_of[self._ids[j]] = j
            self._touch()
        return True

    def delete_batch(self, ids: List[str]) -> int:
        """Delete several vectors, returns how many existed (vectordb_optimized.py:485-504).  One compaction of the
        row array for the whole batch."""
        with self._lock:
            rows = sorted({self._row_of[i] for i in ids if i in self._row_of})
            if not rows:
                return 0
            keep = np.ones(len(self._ids), bool)
            keep[rows] = False
            self._rows = self._rows[keep]
            self._ids = [x for x, k in zip(self._ids, keep) if k]
            self._meta = [x for x, k in zip(self._meta, keep) if k]
            self._row_of = {x: j for j, x in enumerate(self._ids)}
            self._touch()
        return len(rows)

    def set_ef_search(self, ef: int):
        """HNSW knob of the reference (vectordb_optimized.py:737-739); searches here are exact, the value is only kept."""
        self.config.ef_search = int(ef)

    # ------------------------------------------------------------------ reads
    def get_batch(self, ids: List[str], include_vectors: bool = False) -> List[Optional[dict]]:
        """vectordb_optimized.py:441-466: without vectors a list of ``get()`` results, with vectors a list of
        ``{"id", "metadata", "vector"}`` dicts; ``None`` for unknown ids."""
        if not include_vectors:
            return [self.get(i, False) for i in ids]
        out = []
        for i in ids:
            row = self._row_of.get(i)
            out.append(None if row is None else {"id": i, "metadata": self._meta[row], "vector": self._rows[row].copy()})
        return out

    def get(self, id: str, include_vector: bool = False) -> Optional[SearchResult]:
        row = self._row_of.get(id)
        if row is None:
            return None
        return SearchResult(id=id, score=0.0, metadata=self._meta[row],
                            vector=self._rows[row].copy() if include_vector else None)

    def count(self) -> int:
        return len(self._ids)

    def __len__(self) -> int:
        return self.count()

    def list_ids(self, limit: int = 100, offset: int = 0) -> List[str]:
        return self._ids[offset:offset + limit]

    def _resident(self) -> GpuIndex:
        idx = self._index
        if idx is None:
            with self._lock:
                if self._index is None:
                    self._index = GpuIndex(self._rows, self._engine.device)
                idx = self._index
        return idx

    def _row_mask(self, filter) -> Optional[np.ndarray]:
        if filter is None:
            return None
        if isinstance(filter, dict):
            filter = Filter.from_dict(filter)
        if self._columns is None:
            self._columns = _Columns(self._meta)
        return filter.mask(self._columns)

    def _search_arrays(self, queries: np.ndarray, k: int, filter):
        queries = np.asarray(queries, dtype=np.float32)
        if queries.ndim == 1:
            queries = queries.reshape(1, -1)
        if queries.shape[1] != self.config.dimensions:
            raise ValueError(f"Query has {queries.shape[1]} dimensions, expected {self.config.dimensions}")
        if not self._ids:
            return np.zeros((len(queries), 0), np.int64), np.zeros((len(queries), 0), np.float32)
        mask = self._row_mask(filter)
        return self._engine.search_arrays(queries, self._resident(), k, self.config.metric.value, mask)

    def _results(self, idx_row, dist_row, include_vectors=False) -> List[SearchResult]:
        out = []
        for i, d in zip(idx_row, dist_row):
            i = int(i)
            out.append(SearchResult(id=self._ids[i], score=float(d), metadata=self._meta[i],
                                    vector=self._rows[i].copy() if include_vectors else None))
        return out

    def brute_force_search(self, query: np.ndarray, k: int = 10, filter: Union[Filter, dict] = None) -> List[SearchResult]:
        """Exact search (vectordb_optimized.py:650-721): rows failing the filter are excluded (their distance is +inf
        there), at most min(k, permitted rows) results, ascending score."""
        idx, dist = self._search_arrays(query, k, filter)
        if not idx.shape[1]:
            return []
        if self.config.metric is DistanceMetric.EUCLIDEAN:
            # the reference scores L2 as norm(V - q) (vectordb_optimized.py:679-680): re-score the k winners from explicit
            # differences so that a stored duplicate of the query scores exactly 0, as it does there
            idx, dist = self._engine.rerank(np.asarray(query, np.float32).reshape(1, -1), self._resident(), idx[:1], idx.shape[1],
                                            "l2_diff")
        return self._results(idx[0], dist[0])

    def search(self, query: np.ndarray, k: int = 10, filter: Union[Filter, dict] = None, include_vectors: bool = False,
               ef_search: int = None) -> List[SearchResult]:
        """Signature of vectordb_optimized.py:507-510.  Exact instead of HNSW-approximate; ``ef_search`` is ignored."""
        idx, dist = self._search_arrays(query, k, filter)
        return self._results(idx[0], dist[0], include_vectors) if idx.shape[1] else []

    def search_batch(self, queries: np.ndarray, k: int = 10, filter: Union[Filter, dict] = None,
                     include_vectors: bool = False) -> List[List[SearchResult]]:
        """Batch search (vectordb_optimized.py:581-644): one fused GPU pass for the whole batch."""
        idx, dist = self._search_arrays(queries, k, filter)
        return [self._results(i, d, include_vectors) for i, d in zip(idx, dist)]


class VectorDB:
    """Registry of collections with the reference's method names (vectordb_optimized.py:747-818); in memory only."""

    def __init__(self, path: str = None, device=None):
        self.path = path
        self._device = device
        self._collections: Dict[str, Collection] = {}
        self._engine: Optional[ParallelSearchEngine] = None

    def _eng(self) -> ParallelSearchEngine:
        if self._engine is None:
            self._engine = ParallelSearchEngine(device=self._device)
        return self._engine

    def create_collection(self, name: str, dimensions: int, metric: str = "cosine", **kwargs) -> Collection:
        if name in self._collections:
            raise ValueError(f"Collection '{name}' already exists")
        cfg = CollectionConfig(name=name, dimensions=dimensions, metric=DistanceMetric(metric), **kwargs)
        col = Collection(cfg, self.path, engine=self._eng())
        self._collections[name] = col
        return col

    def get_collection(self, name: str) -> Collection:
        if name not in self._collections:
            raise ValueError(f"Collection '{name}' not found")
        return self._collections[name]

    def get_or_create_collection(self, name: str, dimensions: int, metric: str = "cosine", **kwargs) -> Collection:
        return self._collections.get(name) or self.create_collection(name, dimensions, metric, **kwargs)

    def list_collections(self) -> List[str]:
        return list(self._collections)

    def delete_collection(self, name: str) -> bool:
        return self._collections.pop(name, None) is not None


@dataclass
class QueryResult:
    """Result of DocumentCollection.query (fastpyvectordb/client.py:49-56)."""
    ids: List[List[str]]
    documents: List[List[Optional[str]]]
    metadatas: List[List[dict]]
    distances: List[List[float]]
    embeddings: Optional[List[List[np.ndarray]]] = None


@dataclass
class GetResult:
    """Result of DocumentCollection.get / peek (fastpyvectordb/client.py:60-66)."""
    ids: List[str]
    documents: List[Optional[str]]
    metadatas: List[dict]
    embeddings: Optional[List[np.ndarray]] = None


class DocumentCollection:
    """The high-level ``Collection.query`` surface of fastpyvectordb/client.py:184-274 on top of :class:`Collection`.
    Text embedding is upstream of the path: pass ``query_embeddings`` or give an ``embedding_function``."""

    def __init__(self, collection: Collection, embedding_function=None):
        self._collection = collection
        self._embed = embedding_function

    def add(self, ids: List[str], embeddings, metadatas: List[dict] = None, documents: List[str] = None):
        metas = [dict(m or {}) for m in (metadatas or [{} for _ in ids])]
        if documents is not None:
            for m, doc in zip(metas, documents):
                m["_document"] = doc
        self._collection.insert_batch(np.asarray(embeddings, np.float32), list(ids), metas)

    def count(self) -> int:
        return self._collection.count()

    def __len__(self) -> int:
        return self.count()

    def _embed_docs(self, documents):
        if self._embed is None:
            raise ValueError("documents need an embedding_function (or pass embeddings)")
        return np.asarray(self._embed(list(documents)), np.float32)

    def upsert(self, ids: List[str], embeddings=None, metadatas: List[dict] = None, documents: List[str] = None):
        """Add or replace by id (client.py:161-182)."""
        if ids is None:
            raise ValueError("IDs must be provided for upsert")
        if embeddings is None:
            embeddings = self._embed_docs(documents)
        self._collection.delete_batch(list(ids))
        self.add(ids, embeddings, metadatas, documents)
        return list(ids)

    @staticmethod
    def _public(meta: dict) -> dict:
        return {k: v for k, v in meta.items() if not k.startswith("_")}

    def get(self, ids=None, where: Optional[dict] = None, limit: Optional[int] = None, offset: int = 0,
            include: List[str] = None) -> GetResult:
        """By id, or by metadata filter over the whole collection (client.py:276-355; the filter is evaluated by the
        vectorised compiler instead of row by row)."""
        include = include or ["documents", "metadatas"]
        c = self._collection
        if ids is not None:
            rows = [c._row_of[i] for i in ([ids] if isinstance(ids, str) else list(ids)) if i in c._row_of]
        else:
            mask = c._row_mask(where) if where else None
            rows = list(range(len(c._ids))) if mask is None else np.flatnonzero(mask).tolist()
            rows = rows[offset:offset + limit] if limit else rows[offset:]
        metas = [c._meta[r] for r in rows]
        return GetResult(
            ids=[c._ids[r] for r in rows],
            documents=[m.get("_document") for m in metas] if "documents" in include else [None] * len(rows),
            metadatas=[self._public(m) for m in metas] if "metadatas" in include else [{} for _ in rows],
            embeddings=[c._rows[r].copy() for r in rows] if "embeddings" in include and rows else None)

    def peek(self, limit: int = 10) -> GetResult:
        return self.get(limit=limit)

    def update(self, ids: List[str], embeddings=None, metadatas: List[dict] = None, documents: List[str] = None):
        """Replace the vector and / or merge metadata of existing documents (client.py:357-394)."""
        c = self._collection
        new_vecs = self._embed_docs(documents) if (embeddings is None and documents is not None and self._embed) else None
        for i, id_ in enumerate(ids):
            existing = c.get(id_, include_vector=True)
            if existing is None:
                raise ValueError(f"Document with ID '{id_}' not found")
            if embeddings is not None:
                vec = np.asarray(embeddings[i], np.float32)
            elif new_vecs is not None:
                vec = new_vecs[i]
            else:
                vec = existing.vector
            meta = dict(existing.metadata)
            if metadatas is not None and i < len(metadatas):
                meta.update(metadatas[i])
            if documents is not None and i < len(documents):
                meta["_document"] = documents[i]
            c.upsert(vec, id_, meta)

    def delete(self, ids=None, where: Optional[dict] = None) -> int:
        """By id and / or by metadata filter; returns the number deleted (client.py:396-430)."""
        if ids is None and where is None:
            raise ValueError("Either ids or where must be provided")
        c = self._collection
        doomed = set([ids] if isinstance(ids, str) else (ids or []))
        if where is not None:
            mask = c._row_mask(where)
            doomed.update(c._ids[r] for r in np.flatnonzero(mask).tolist())
        return c.delete_batch(list(doomed))

    def query(self, query_texts=None, query_embeddings=None, n_results: int = 10, where: Optional[dict] = None,
              include: List[str] = None) -> QueryResult:
        if query_texts is None and query_embeddings is None:
            raise ValueError("Either query_texts or query_embeddings must be provided")
        include = include or ["documents", "metadatas", "distances"]
        if query_embeddings is not None:
            queries = np.array(query_embeddings, dtype=np.float32)
        else:
            if self._embed is None:
                raise ValueError("query_texts needs an embedding_function")
            queries = np.asarray(self._embed([query_texts] if isinstance(query_texts, str) else list(query_texts)), np.float32)
        rows = self._collection.search_batch(queries, k=n_results, filter=Filter.from_dict(where) if where else None,
                                             include_vectors="embeddings" in include)      # ONE batched GPU call
        ids = [[r.id for r in row] for row in rows]
        return QueryResult(
            ids=ids,
            documents=[[r.metadata.get("_document") if "documents" in include else None for r in row] for row in rows],
            metadatas=[[{k: v for k, v in r.metadata.items() if not k.startswith("_")} if "metadatas" in include else {}
                        for r in row] for row in rows],
            distances=[[r.score if "distances" in include else 0.0 for r in row] for row in rows],
            embeddings=[[r.vector for r in row] for row in rows] if "embeddings" in include else None)
