"""``MemoryMappedVectors`` and ``ParallelCollection`` — the callers that wrap the engine in the reference
(parallel_search.py:427-750 and :757-952), the data format on the feed side of the path.

``MemoryMappedVectors`` keeps the reference's on-disk format bit for bit (so files are interchangeable):
``vectors.mmap`` = 64-byte header (``b'PYVEC001'``, ``<III`` version / n_vectors / dimensions, zero padding) followed
by the row-major float32 payload; ``ids.json``; ``metadata.json``.  ``search_parallel`` streams the file to the GPU in
row chunks: every chunk is searched by the fused scan with ``id_base = chunk start`` (the reference's
``_compute_distances_chunk`` + ``_merge_top_k`` loop, :702-722) and the per-chunk lists are merged by the CUDA merge
kernel; a store that fits in HBM is uploaded once and stays resident until the next append.

``ParallelCollection`` keeps ``insert_batch`` / ``search_parallel`` (with ``filter_fn``) / ``search_hybrid`` / ``count``.
The HNSW index of the reference is out of scope, so the candidate stage of ``search_hybrid`` is the binary-quantizer
Hamming scan (same pattern: approximate candidates, then the exact cosine re-rank of :919-934) and ``search_hnsw``
raises.
"""
from __future__ import annotations

import json
import struct
import uuid
from pathlib import Path
from typing import Callable, Dict, List, Optional

import numpy as np
import torch

from . import ops
from .engine import GpuIndex, ParallelSearchEngine, ParallelSearchResult
from .quantizers import BinaryQuantizer
from .sharded import pack_candidates, unpack_candidates


class MemoryMappedVectors:
    MAGIC = b"PYVEC001"
    HEADER_SIZE = 64
    #: rows per streamed chunk when the store does not fit the resident budget
    CHUNK_ROWS = 1_000_000
    #: stores up to this many bytes are kept resident in HBM between searches
    RESIDENT_BYTES = 64 << 30

    def __init__(self, path: str, dimensions: int = None):
        self.path = Path(path)
        self.path.mkdir(parents=True, exist_ok=True)
        self.data_path = self.path / "vectors.mmap"
        self.meta_path = self.path / "metadata.json"
        self.ids_path = self.path / "ids.json"
        self._dimensions = dimensions
        self._n_vectors = 0
        self._capacity = 0
        self._mmap: Optional[np.memmap] = None
        self._ids: List[str] = []
        self._metadata: Dict[str, dict] = {}
        self._resident: Optional[GpuIndex] = None
        if self.data_path.exists():
            self._load()

    # ------------------------------------------------------------------ format
    @property
    def dimensions(self) -> int:
        return self._dimensions

    @property
    def n_vectors(self) -> int:
        return self._n_vectors

    def __len__(self) -> int:
        return self._n_vectors

    def _header(self, n: int) -> bytes:
        head = self.MAGIC + struct.pack("<III", 1, n, self._dimensions)
        return head + b"\x00" * (self.HEADER_SIZE - len(head))

    def _load(self):
        with open(self.data_path, "rb") as f:
            head = f.read(self.HEADER_SIZE)
        if head[:8] != self.MAGIC:
            raise ValueError(f"Invalid file format: {head[:8]}")
        _version, n, dims = struct.unpack("<III", head[8:20])
        self._n_vectors, self._dimensions = n, dims
        payload = self.data_path.stat().st_size - self.HEADER_SIZE
        self._capacity = payload // (4 * dims) if dims else 0
        if self._capacity > 0:
            self._mmap = np.memmap(self.data_path, dtype=np.float32, mode="r+", offset=self.HEADER_SIZE,
                                   shape=(self._capacity, dims))
        if self.ids_path.exists():
            self._ids = json.loads(self.ids_path.read_text())
        if self.meta_path.exists():
            self._metadata = json.loads(self.meta_path.read_text())

    def create(self, n_vectors: int, dimensions: int = None):
        """Pre-allocate room for ``n_vectors`` rows (parallel_search.py:516-558)."""
        if dimensions:
            self._dimensions = dimensions
        if not self._dimensions:
            raise ValueError("dimensions must be specified")
        self._n_vectors, self._capacity = 0, int(n_vectors)
        with open(self.data_path, "wb") as f:
            f.write(self._header(0))
            if n_vectors > 0:
                f.seek(self.HEADER_SIZE + n_vectors * self._dimensions * 4 - 1)
                f.write(b"\x00")
        self._mmap = np.memmap(self.data_path, dtype=np.float32, mode="r+", offset=self.HEADER_SIZE,
                               shape=(max(n_vectors, 1), self._dimensions)) if n_vectors > 0 else None
        self._ids, self._metadata, self._resident = [], {}, None

    def _write_count(self):
        with open(self.data_path, "r+b") as f:
            f.write(self._header(self._n_vectors))

    def append(self, vector: np.ndarray, id: str = None, metadata: dict = None) -> str:
        return self.append_batch(np.asarray(vector, np.float32).reshape(1, -1), [id] if id else None,
                                 [metadata] if metadata else None)[0]

    def append_batch(self, vectors: np.ndarray, ids: List[str] = None, metadata_list: List[dict] = None) -> List[str]:
        vectors = np.asarray(vectors, dtype=np.float32)
        if self._mmap is None:
            raise ValueError("Storage not created. Call create() first.")
        n = len(vectors)
        if self._n_vectors + n > self._capacity:
            raise ValueError("Storage full")
        if vectors.shape[1] != self._dimensions:
            raise ValueError(f"Vectors have {vectors.shape[1]} dimensions, expected {self._dimensions}")
        ids = list(ids) if ids is not None else [str(uuid.uuid4()) for _ in range(n)]
        self._mmap[self._n_vectors:self._n_vectors + n] = vectors
        self._ids.extend(ids)
        if metadata_list:
            for i, m in zip(ids, metadata_list):
                if m:
                    self._metadata[i] = m
        self._n_vectors += n
        self._write_count()
        self._resident = None
        return ids

    def get(self, idx: int) -> np.ndarray:
        if idx < 0 or idx >= self._n_vectors:
            raise IndexError(f"Index {idx} out of range")
        return np.array(self._mmap[idx])

    def get_batch(self, indices: List[int]) -> np.ndarray:
        return np.array(self._mmap[np.asarray(indices, dtype=np.int64)])

    def get_range(self, start: int, end: int) -> np.ndarray:
        return np.array(self._mmap[start:min(end, self._n_vectors)])

    def get_all(self) -> np.ndarray:
        if self._mmap is None or self._n_vectors == 0:
            return np.zeros((0, self._dimensions or 0), np.float32)
        return self._mmap[:self._n_vectors]

    def save_metadata(self):
        self.ids_path.write_text(json.dumps(self._ids))
        self.meta_path.write_text(json.dumps(self._metadata))

    def close(self):
        if self._mmap is not None:
            self._mmap.flush()
        self.save_metadata()
        self._mmap, self._resident = None, None

    # ------------------------------------------------------------------ search
    def search_parallel(self, query: np.ndarray, k: int = 10, metric: str = "cosine",
                        engine: ParallelSearchEngine = None) -> List[ParallelSearchResult]:
        """Exact top-k over the stored rows (parallel_search.py:684-727)."""
        engine = engine or ParallelSearchEngine()
        n = self._n_vectors
        if n == 0:
            return []
        if n * self._dimensions * 4 <= self.RESIDENT_BYTES:
            if self._resident is None:
                self._resident = GpuIndex(np.ascontiguousarray(self.get_all()), engine.device)
            return engine.search_parallel(query, self._resident, k, metric)
        return self._search_streaming(np.asarray(query, dtype=np.float32).reshape(1, -1), k, metric, engine)

    def _search_streaming(self, q: np.ndarray, k: int, metric: str, engine: ParallelSearchEngine):
        """Stores larger than HBM (parallel_search.py:702-722 walks the file in 100k-row chunks): the file is streamed
        through TWO pinned staging buffers and two device buffers, so the host read of chunk c+1 (page cache / disk ->
        pinned memory), the host->device copy of chunk c and the fused scan + local top-k of chunk c-1 overlap; the per
        chunk lists are merged by the merge kernel with global row ids.  Bound by the slowest of the three (normally
        PCIe); ``last_stream_stats`` reports the achieved rate."""
        import time
        n, d, dev = self._n_vectors, self._dimensions, engine.device
        chunk = min(self.CHUNK_ROWS, n)
        kk = min(k, n)
        pins = [torch.empty((chunk, d), dtype=torch.float32).pin_memory() for _ in range(2)]
        devs = [torch.empty((chunk, d), dtype=torch.float32, device=dev) for _ in range(2)]
        copy_stream = torch.cuda.Stream(dev)
        compute = torch.cuda.current_stream(dev)
        h2d_done = [None, None]
        scan_done = [None, None]
        qd = torch.from_numpy(np.ascontiguousarray(q)).to(dev)
        parts = []
        t0 = time.perf_counter()
        for c, start in enumerate(range(0, n, chunk)):
            b = c & 1
            rows = min(chunk, n - start)
            if h2d_done[b] is not None:
                h2d_done[b].synchronize()                               # the copy that last read this pinned buffer is done
            pins[b][:rows].numpy()[...] = self._mmap[start:start + rows]   # host read, overlaps the GPU work in flight
            if scan_done[b] is not None:
                copy_stream.wait_event(scan_done[b])                    # the scan that last read this device buffer is done
            with torch.cuda.stream(copy_stream):
                devs[b][:rows].copy_(pins[b][:rows], non_blocking=True)
                ev = torch.cuda.Event()
                ev.record(copy_stream)
            h2d_done[b] = ev
            compute.wait_event(ev)
            dd, ii, _c = ops.scan_f32_topk(qd, devs[b][:rows], min(kk, rows), metric, None, None, start)
            ev2 = torch.cuda.Event()
            ev2.record(compute)
            scan_done[b] = ev2
            parts.append(pack_candidates(dd, ii, kk))
        dd, ii = unpack_candidates(torch.stack(parts))
        md, mi, mc = ops.merge_topk(dd, ii, kk)
        valid = int(mc[0].item())                                       # synchronises
        dt = time.perf_counter() - t0
        self.last_stream_stats = {"bytes": int(n) * d * 4, "seconds": dt, "gb_per_s": n * d * 4 / dt / 1e9, "chunks": len(parts),
                                  "chunk_rows": chunk}
        return [ParallelSearchResult(index=int(a), distance=float(b))
                for a, b in zip(mi[0, :valid].tolist(), md[0, :valid].tolist())]


class ParallelCollection:
    """Collection wrapper of parallel_search.py:757-952 on the GPU engine (see module docstring)."""

    def __init__(self, name: str, dimensions: int, metric: str = "cosine", n_workers: int = None, device=None, **_ignored):
        self.name, self.dimensions, self.metric = name, dimensions, metric
        self._engine = ParallelSearchEngine(n_workers=n_workers, device=device)
        self._vectors = np.zeros((0, dimensions), np.float32)
        self._ids: List[str] = []
        self._metadata: Dict[str, dict] = {}
        self._bq: Optional[BinaryQuantizer] = None
        self._codes = None

    def insert_batch(self, vectors: np.ndarray, ids: List[str] = None, metadata_list: List[dict] = None) -> List[str]:
        vectors = np.asarray(vectors, dtype=np.float32)
        n = len(vectors)
        ids = list(ids) if ids is not None else [str(uuid.uuid4()) for _ in range(n)]
        self._vectors = np.concatenate([self._vectors, vectors]) if len(self._vectors) else np.ascontiguousarray(vectors)
        self._ids.extend(ids)
        for i, m in zip(ids, metadata_list or [None] * n):
            if m:
                self._metadata[i] = m
        self._bq, self._codes = None, None
        return ids

    def _wrap(self, results: List[ParallelSearchResult]) -> List[ParallelSearchResult]:
        for r in results:
            r.id = self._ids[r.index]
            r.metadata = self._metadata.get(r.id, {})
        return results

    def search_hnsw(self, query: np.ndarray, k: int = 10):
        raise NotImplementedError("the HNSW index is outside the scope of this build; use search_parallel (exact) or "
                                  "search_hybrid (quantized candidates + exact re-rank)")

    def search_parallel(self, query: np.ndarray, k: int = 10, filter_fn: Callable[[dict], bool] = None
                        ) -> List[ParallelSearchResult]:
        """Exact search; ``filter_fn(metadata) -> bool`` becomes the row bitmask (parallel_search.py:857-893)."""
        if not self._ids:
            return []
        mask = None
        if filter_fn is not None:
            mask = np.fromiter((bool(filter_fn(self._metadata.get(i, {}))) for i in self._ids), bool, len(self._ids))
        return self._wrap(self._engine.search_parallel(query, self._vectors, k, self.metric, mask))

    def search_hybrid(self, query: np.ndarray, k: int = 10, hnsw_candidates: int = 100) -> List[ParallelSearchResult]:
        """Approximate candidates, then exact cosine re-rank of them (parallel_search.py:895-947).  ``hnsw_candidates``
        keeps its name; the candidates come from the Hamming scan of binary codes."""
        if not self._ids:
            return []
        if self._bq is None:
            self._bq = BinaryQuantizer(device=self._engine.device).train(self._vectors)
            self._codes = self._bq.to_device(self._bq.encode(self._vectors))
        cand, _ = self._bq.search(np.asarray(query, np.float32), self._codes, k=min(hnsw_candidates, len(self._ids)))
        cand = cand.cpu().numpy() if isinstance(cand, torch.Tensor) else cand
        idx, dist = self._engine.rerank(query, self._vectors, cand, k, "cosine")
        return self._wrap([ParallelSearchResult(index=int(a), distance=float(b)) for a, b in zip(idx[0], dist[0])])

    def count(self) -> int:
        return len(self._ids)
