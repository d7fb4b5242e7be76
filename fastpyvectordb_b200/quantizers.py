"""Drop-in ``ScalarQuantizer`` / ``BinaryQuantizer`` / ``ProductQuantizer`` served by the B200 kernels.

Mirrors the call surface of the reference's ``quantization.py`` (ScalarQuantizer :64-213, BinaryQuantizer
:282-407, ProductQuantizer :414-615): same constructor arguments, method names, public attributes
(``min_vals/max_vals/scale``, ``thresholds``, ``codebooks``, ``trained`` ...), return types and ValueErrors.
Encoders and every distance scan run on the GPU through ``ops.py``; NumPy arrays in give NumPy arrays out,
torch CUDA tensors in give torch CUDA tensors out (no host round trip).  There is no CPU scan path.

Additions on top of the reference API (all optional): ``search(..., filter_mask=)``, ``*_search_tensors``
and ``to_device(codes)`` which makes a code matrix resident in HBM so repeated scans do not re-upload it.
"""
from __future__ import annotations

from dataclasses import dataclass
from enum import Enum
from typing import Optional, Tuple

import numpy as np
import torch

from . import _native as N
from . import ops


class DistanceMetric(Enum):
    COSINE = "cosine"
    EUCLIDEAN = "l2"
    DOT_PRODUCT = "ip"


@dataclass
class ScalarQuantizerConfig:
    """Import compatibility with quantization.py:57-61 (the reference defines it and never reads it; 8 bits, min-max)."""
    bits: int = 8
    symmetric: bool = False


def _code_cache():
    """host code matrix -> device tensor with the engine's residency rules (engine._ResidentCache): reused only when
    the host array is read-only or a checksum of every byte still matches; large writeable arrays are re-uploaded."""
    from .engine import _ResidentCache
    return _ResidentCache(build=lambda arr, device: torch.from_numpy(np.ascontiguousarray(arr)).to(device))


class _DeviceMixin:
    device = None

    def _dev(self):
        if self.device is None:
            self.device = N.require_cuda(None)
        return self.device

    def _codes(self, codes, cache: bool = True) -> torch.Tensor:
        if isinstance(codes, torch.Tensor):
            t = codes.to(self._dev())
            return (t if t.dtype == torch.uint8 else t.to(torch.uint8)).contiguous()
        arr = np.asarray(codes)
        if arr.dtype != np.uint8:
            arr = arr.astype(np.uint8)
        if arr.ndim == 1:
            arr = arr.reshape(1, -1)
        if not cache:
            return torch.from_numpy(np.ascontiguousarray(arr)).to(self._dev())
        if not hasattr(self, "_code_cache"):
            self._code_cache = _code_cache()
        return self._code_cache.get(arr, self._dev())

    def to_device(self, codes) -> torch.Tensor:
        """Make a code matrix resident on the GPU; pass the returned tensor wherever codes are expected."""
        return self._codes(codes)

    def _f32(self, x, two_d=True) -> Tuple[torch.Tensor, bool]:
        """-> (device float32 tensor, input_was_torch)"""
        if isinstance(x, torch.Tensor):
            t = x.to(device=self._dev(), dtype=torch.float32)
            if two_d and t.ndim == 1:
                t = t.reshape(1, -1)
            return t.contiguous(), True
        a = np.asarray(x, dtype=np.float32)
        if two_d and a.ndim == 1:
            a = a.reshape(1, -1)
        return torch.from_numpy(np.ascontiguousarray(a)).to(self._dev()), False

    def _param(self, name: str, arr) -> torch.Tensor:
        """device copy of a small parameter array, refreshed when the public NumPy attribute is reassigned"""
        cache = self.__dict__.setdefault("_param_cache", {})
        a = np.ascontiguousarray(arr, dtype=np.float32)
        hit = cache.get(name)
        from .engine import _checksum
        sig = (a.shape, _checksum(a))                      # parameter arrays are small: every byte is checked
        if hit is not None and hit[0] is arr and hit[2] == sig:
            return hit[1]
        t = torch.from_numpy(a).to(self._dev())
        cache[name] = (arr, t, sig)
        return t

    def _mask(self, filter_mask, n):
        if filter_mask is None:
            return None
        if isinstance(filter_mask, torch.Tensor):
            if filter_mask.numel() != n:
                raise ValueError(f"filter_mask has {filter_mask.numel()} entries for {n} rows")
            return ops.pack_mask(filter_mask.to(self._dev()).reshape(-1) != 0)
        m = np.asarray(filter_mask).reshape(-1).astype(bool)
        if m.size != n:
            raise ValueError(f"filter_mask has {m.size} entries for {n} rows")
        return ops.pack_mask_host(m).to(self._dev())


def _large_k_search(self, all_dist: torch.Tensor, kk: int, filter_mask, as_torch):
    """k beyond the fused selector (MAX_K = 1024; cold path, the reference server caps k at 1000, server.py:72): one full
    distance row from the scan kernel, then a stable device sort (stable == lowest index first among equal distances).
    ``filter_mask`` has the np.where(mask, d, inf) semantics of the fused path: rejected rows are never returned."""
    n = all_dist.numel()
    n_ok = n
    if filter_mask is not None:
        m = filter_mask if isinstance(filter_mask, torch.Tensor) else torch.from_numpy(np.asarray(filter_mask).reshape(-1).astype(bool))
        m = m.to(all_dist.device).reshape(-1) != 0
        if m.numel() != n:
            raise ValueError(f"filter_mask has {m.numel()} entries for {n} rows")
        all_dist = torch.where(m, all_dist, torch.full_like(all_dist, float("inf")))
        n_ok = int(m.sum().item())
    d, order = torch.sort(all_dist, stable=True)
    kk = min(kk, n_ok)
    d, order = d[:kk], order[:kk]
    return (order, d) if as_torch else (order.cpu().numpy(), d.cpu().numpy())


def _finish_search(dist, idx, cnt, as_torch):
    """(device tensors) -> the reference's (indices, distances) pair trimmed to the valid count."""
    if as_torch:
        valid = int(cnt[0].item()) if cnt.numel() else 0
        return idx[0, :valid], dist[0, :valid]
    valid = int(cnt[0].item()) if cnt.numel() else 0
    return idx[0, :valid].cpu().numpy(), dist[0, :valid].cpu().numpy()


# ======================================================================================================
# Scalar quantization (quantization.py:64-276)
# ======================================================================================================
class ScalarQuantizer(_DeviceMixin):
    """f32 -> uint8 min-max quantizer (4x compression); API of quantization.py:64-213."""

    def __init__(self, dimensions: int = None, device=None):
        self.dimensions = dimensions
        self.trained = False
        self.min_vals: Optional[np.ndarray] = None
        self.max_vals: Optional[np.ndarray] = None
        self.scale: Optional[np.ndarray] = None
        self.device = N.require_cuda(device) if device is not None else None

    def train(self, vectors) -> "ScalarQuantizer":
        """Per-dimension min / max / scale (quantization.py:85-106); a zero range becomes scale 1.0."""
        if isinstance(vectors, torch.Tensor):
            v = vectors.to(torch.float32)
            v = v.reshape(1, -1) if v.ndim == 1 else v
            lo = v.min(dim=0).values.cpu().numpy()
            hi = v.max(dim=0).values.cpu().numpy()
        else:
            v = np.asarray(vectors, dtype=np.float32)
            v = v.reshape(1, -1) if v.ndim == 1 else v
            lo, hi = v.min(axis=0), v.max(axis=0)
        self.dimensions = v.shape[1]
        self.min_vals, self.max_vals = lo, hi
        scale = hi - lo
        self.scale = np.where(scale == 0, 1.0, scale).astype(np.float32)
        self.trained = True
        return self

    def _need_trained(self):
        if not self.trained:
            raise ValueError("Quantizer not trained. Call train() first.")

    def encode(self, vectors):
        """float32 -> uint8 codes, clip((v - min) / scale * 255, 0, 255) truncated (quantization.py:108-126)."""
        self._need_trained()
        v, was_torch = self._f32(vectors)
        codes = ops.sq_encode(v, self._param("min", self.min_vals), self._param("scale", self.scale))
        return codes if was_torch else codes.cpu().numpy()

    def decode(self, quantized):
        """uint8 -> approximate float32 (quantization.py:128-139); elementwise, evaluated where the codes live."""
        self._need_trained()
        if isinstance(quantized, torch.Tensor):
            return quantized.to(torch.float32) / 255.0 * self._param("scale", self.scale) + self._param("min", self.min_vals)
        return np.asarray(quantized).astype(np.float32) / 255.0 * self.scale + self.min_vals

    def encode_query(self, query):
        """Encode a query vector (quantization.py:141-143)."""
        return self.encode(query.reshape(1, -1))[0]

    def _scan(self, kind, query, quantized_db, k=0, filter_mask=None, want_all=True):
        self._need_trained()
        q, was_torch = self._f32(query)
        codes = self._codes(quantized_db)
        mn, sc = self._param("min", self.min_vals), self._param("scale", self.scale)
        qcodes = ops.sq_encode(q, mn, sc)                                     # the query is re-quantised first (:151)
        return ops.sq_scan(kind, qcodes, codes, mn, sc, k, self._mask(filter_mask, codes.shape[0]), 0, want_all), \
            was_torch or isinstance(quantized_db, torch.Tensor)

    def distances_l2(self, query, quantized_db):
        """Approximate L2 distances to every row (quantization.py:145-152)."""
        (_, _, _, out), t = self._scan(N.SQ_L2, query, quantized_db)
        return out[0] if t else out[0].cpu().numpy()

    def distances_cosine(self, query, quantized_db, norms=None):
        """Approximate cosine distances (quantization.py:154-174); ``norms`` is ignored there as well."""
        (_, _, _, out), t = self._scan(N.SQ_COSINE, query, quantized_db)
        return out[0] if t else out[0].cpu().numpy()

    def distances_dot(self, query, quantized_db):
        """Approximate negative dot products (quantization.py:176-181)."""
        (_, _, _, out), t = self._scan(N.SQ_DOT, query, quantized_db)
        return out[0] if t else out[0].cpu().numpy()

    def search(self, query, quantized_db, k: int = 10, metric: str = "l2", filter_mask=None):
        """Fused scan + top-k (what callers of distances_* do next: rag_demo.py:486-492).  -> (indices, distances)"""
        kind = {"l2": N.SQ_L2, "cosine": N.SQ_COSINE}.get(metric, N.SQ_DOT)
        n = len(quantized_db)
        kk = max(1, min(int(k), n, N.MAX_K)) if n else 1
        if self.tensor_core_scan and n and self.trained:
            codes = self._codes(quantized_db)
            # one query fills 3 of the 48 limb columns, but the scan is bound by reading the codes once either way and the
            # tensor-core pass needs no u8 -> float conversion (the SIMT scan is instruction bound at ~70 % of HBM)
            if ops.sq_mma_supported(n, codes.shape[1], kk):
                t = isinstance(query, torch.Tensor) or isinstance(quantized_db, torch.Tensor)
                dist, idx, cnt = self.search_batch_tensors(query, codes, kk, filter_mask, metric)
                return _finish_search(dist, idx, cnt, t)
        (dist, idx, cnt, _), t = self._scan(kind, query, quantized_db, kk, filter_mask, want_all=False)
        return _finish_search(dist, idx, cnt, t)

    #: batches go to the int8 tensor-core scan (csrc/fpv_sq_mma.cu) when the shape allows; False forces the SIMT scan
    tensor_core_scan = True

    def _row_term(self, dcodes: torch.Tensor, sc: torch.Tensor):
        """(C_row [N], max) of a device code matrix for the current ``scale``: built once per matrix (index build)."""
        from .engine import _checksum
        cache = self.__dict__.setdefault("_row_term_cache", {})
        key = (dcodes.data_ptr(), tuple(dcodes.shape), dcodes._version, _checksum(np.ascontiguousarray(self.scale, np.float32)))
        hit = cache.get(key)
        if hit is None:
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            hit = (dcodes,) + ops.sq_row_term(dcodes, sc)               # keep the source alive: the pointer is the key
            cache[key] = hit
        return hit[1], hit[2]

    def _row_terms_dc(self, dcodes: torch.Tensor, mn: torch.Tensor, sc: torch.Tensor):
        """(R_row [N], invn_row [N], maxima [2]) of a device code matrix for the current ``min_vals`` / ``scale``."""
        from .engine import _checksum
        cache = self.__dict__.setdefault("_row_terms_dc_cache", {})
        key = (dcodes.data_ptr(), tuple(dcodes.shape), dcodes._version, _checksum(np.ascontiguousarray(self.scale, np.float32)),
               _checksum(np.ascontiguousarray(self.min_vals, np.float32)))
        hit = cache.get(key)
        if hit is None:
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            hit = (dcodes,) + ops.sq_row_terms_dc(dcodes, mn, sc)        # keep the source alive: the pointer is the key
            cache[key] = hit
        return hit[1], hit[2], hit[3]

    def search_batch_tensors(self, queries, quantized_db, k: int = 10, filter_mask=None, metric: str = "l2"):
        """Top-k of a BATCH of queries over the uint8 codes -> device (dist [Q,k], idx [Q,k], count [Q]); ``metric`` as
        in :meth:`search` ("l2", "cosine", anything else = dot).

        One pass over the codes serves 16 queries: the weighted distance of quantization.py:217-236 is expanded so that
        its cross term is three exact int8 tensor-core dot products per (row, query) (``tcgen05.mma kind::i8``), a
        certified window is re-scored with the scan's own arithmetic, and the results are identical to
        ``search(q, ...)`` called per query.  Small or odd shapes take the SIMT scan, one pass per query."""
        self._need_trained()
        q, _ = self._f32(queries)
        codes = self._codes(quantized_db)
        n = codes.shape[0]
        mn, sc = self._param("min", self.min_vals), self._param("scale", self.scale)
        qcodes = ops.sq_encode(q, mn, sc)                                     # queries are re-quantised first (:151)
        kk = max(1, min(int(k), n, N.MAX_K)) if n else 1
        words = self._mask(filter_mask, n)
        kind = {"l2": N.SQ_L2, "cosine": N.SQ_COSINE}.get(metric, N.SQ_DOT)
        if self.tensor_core_scan and n and ops.sq_mma_supported(n, codes.shape[1], kk):
            if kind == N.SQ_L2:
                term, tmax = self._row_term(codes, sc)
                return ops.sq_l2_mma(qcodes, codes, mn, sc, term, tmax, kk, words, 0)
            rsum, rinv, maxima = self._row_terms_dc(codes, mn, sc)
            return ops.sq_dc_mma(kind, qcodes, codes, mn, sc, rsum, rinv, maxima, kk, words, 0)
        dist, idx, cnt, _ = ops.sq_scan(kind, qcodes, codes, mn, sc, kk, words, 0, False)
        return dist, idx, cnt

    def search_batch(self, queries, quantized_db, k: int = 10, filter_mask=None, metric: str = "l2"):
        """-> (indices [Q, k'], distances [Q, k']) NumPy arrays (torch tensors for torch inputs), k' = valid results."""
        t = isinstance(queries, torch.Tensor) or isinstance(quantized_db, torch.Tensor)
        dist, idx, cnt = self.search_batch_tensors(queries, quantized_db, k, filter_mask, metric)
        valid = int(cnt.min().item()) if cnt.numel() else 0
        idx, dist = idx[:, :valid], dist[:, :valid]
        return (idx, dist) if t else (idx.cpu().numpy(), dist.cpu().numpy())

    def memory_usage(self, n_vectors: int) -> dict:
        """Same accounting as quantization.py:183-194."""
        float32_bytes = n_vectors * self.dimensions * 4
        uint8_bytes = n_vectors * self.dimensions
        overhead = self.dimensions * 4 * 3
        return {"original_bytes": float32_bytes, "quantized_bytes": uint8_bytes + overhead,
                "compression_ratio": float32_bytes / (uint8_bytes + overhead),
                "savings_percent": (1 - (uint8_bytes + overhead) / float32_bytes) * 100}

    def save(self, path: str):
        """np.savez with the reference's field names (quantization.py:196-202)."""
        np.savez(path, min_vals=self.min_vals, max_vals=self.max_vals, scale=self.scale, dimensions=self.dimensions)

    @classmethod
    def load(cls, path: str) -> "ScalarQuantizer":
        """Inverse of :meth:`save` (quantization.py:204-213); files are interchangeable with the reference's."""
        data = np.load(path)
        sq = cls(dimensions=int(data["dimensions"]))
        sq.min_vals, sq.max_vals, sq.scale = data["min_vals"], data["max_vals"], data["scale"]
        sq.trained = True
        return sq


# ======================================================================================================
# Binary quantization (quantization.py:282-407)
# ======================================================================================================
class BinaryQuantizer(_DeviceMixin):
    """f32 -> 1 bit per dimension, Hamming distance (32x compression); API of quantization.py:282-407."""

    def __init__(self, dimensions: int = None, threshold: float = 0.0, device=None):
        self.dimensions = dimensions
        self.threshold = threshold
        self.trained = False
        self.thresholds: Optional[np.ndarray] = None
        self.device = N.require_cuda(device) if device is not None else None

    def train(self, vectors, use_median: bool = True) -> "BinaryQuantizer":
        """Per-dimension median (or the constant threshold) (quantization.py:307-327)."""
        if isinstance(vectors, torch.Tensor):
            v = vectors.to(torch.float32)
            v = v.reshape(1, -1) if v.ndim == 1 else v
            self.dimensions = v.shape[1]
            if use_median:
                # np.median semantics: mean of the two middle order statistics for an even count
                s, _ = torch.sort(v, dim=0)
                n = s.shape[0]
                med = s[n // 2] if n % 2 else (s[n // 2 - 1] + s[n // 2]) / 2
                self.thresholds = med.cpu().numpy()
            else:
                self.thresholds = np.full(self.dimensions, self.threshold)
        else:
            v = np.asarray(vectors, dtype=np.float32)
            v = v.reshape(1, -1) if v.ndim == 1 else v
            self.dimensions = v.shape[1]
            self.thresholds = np.median(v, axis=0) if use_median else np.full(self.dimensions, self.threshold)
        self.trained = True
        return self

    def encode(self, vectors):
        """float32 -> packed bits, MSB first, ceil(D/8) bytes per row (quantization.py:329-350)."""
        v, was_torch = self._f32(vectors)
        if self.trained:
            thr = self._param("thr", self.thresholds)
        else:
            thr = torch.full((v.shape[1],), float(self.threshold), dtype=torch.float32, device=self._dev())
        codes = ops.bq_encode(v, thr)
        return codes if was_torch else codes.cpu().numpy()

    def encode_query(self, query):
        """Encode a single query vector (quantization.py:352-354)."""
        return self.encode(query.reshape(1, -1))[0]

    def hamming_distances(self, query_bits, db_bits):
        """popcount(q XOR d) over the first ``dimensions`` bits as float32 (quantization.py:356-374)."""
        t = isinstance(db_bits, torch.Tensor) or isinstance(query_bits, torch.Tensor)
        qb = self._codes(query_bits if isinstance(query_bits, torch.Tensor) else np.asarray(query_bits).reshape(1, -1),
                         cache=False).reshape(1, -1)
        codes = self._codes(db_bits)
        _, _, _, out = ops.hamming(qb, codes, 0, self.dimensions or 0, None, 0, want_all=True)
        return out[0] if t else out[0].cpu().numpy()

    def search(self, query, db_bits, k: int = 10, filter_mask=None):
        """k nearest rows by Hamming distance (quantization.py:376-394) -> (indices, distances)."""
        q, was_torch = self._f32(query)
        qbits = self.encode(q)
        codes = self._codes(db_bits)
        n = codes.shape[0]
        if n == 0:
            e = (np.zeros(0, np.int64), np.zeros(0, np.float32))
            return (torch.from_numpy(e[0]), torch.from_numpy(e[1])) if was_torch else e
        kk = min(int(k), n)
        if kk > N.MAX_K:
            _, _, _, out = ops.hamming(qbits, codes, 0, self.dimensions or 0, None, 0, want_all=True)
            return _large_k_search(self, out[0], kk, filter_mask, was_torch or isinstance(db_bits, torch.Tensor))
        dist, idx, cnt, _ = ops.hamming(qbits, codes, kk, self.dimensions or 0, self._mask(filter_mask, n), 0)
        return _finish_search(dist, idx, cnt, was_torch or isinstance(db_bits, torch.Tensor))

    def search_batch_tensors(self, queries, db_bits, k: int = 10, filter_mask=None):
        """Hamming top-k of a BATCH of queries -> device (dist [Q,k], idx [Q,k], count [Q]).  Batches of 4+ queries over
        1024- or 2048-bit codes run on the int8 tensor cores (csrc/fpv_hamming_mma.cu: one pass over the packed codes
        serves 31 queries); the results are identical to ``search`` called per query."""
        q, _ = self._f32(queries)
        qbits = self.encode(q)
        codes = self._codes(db_bits)
        n = codes.shape[0]
        kk = max(1, min(int(k), n, N.MAX_K)) if n else 1
        dist, idx, cnt, _ = ops.hamming(qbits, codes, kk, self.dimensions or 0, self._mask(filter_mask, n), 0)
        return dist, idx, cnt

    def search_batch(self, queries, db_bits, k: int = 10, filter_mask=None):
        """-> (indices [Q, k'], distances [Q, k']); NumPy in, NumPy out."""
        t = isinstance(queries, torch.Tensor) or isinstance(db_bits, torch.Tensor)
        dist, idx, cnt = self.search_batch_tensors(queries, db_bits, k, filter_mask)
        valid = int(cnt.min().item()) if cnt.numel() else 0
        idx, dist = idx[:, :valid], dist[:, :valid]
        return (idx, dist) if t else (idx.cpu().numpy(), dist.cpu().numpy())

    def memory_usage(self, n_vectors: int) -> dict:
        """Same accounting as quantization.py:396-407."""
        float32_bytes = n_vectors * self.dimensions * 4
        binary_bytes = n_vectors * ((self.dimensions + 7) // 8)
        overhead = self.dimensions * 4 if self.thresholds is not None else 0
        return {"original_bytes": float32_bytes, "quantized_bytes": binary_bytes + overhead,
                "compression_ratio": float32_bytes / (binary_bytes + overhead),
                "savings_percent": (1 - (binary_bytes + overhead) / float32_bytes) * 100}


# ======================================================================================================
# Product quantization (quantization.py:414-615)
# ======================================================================================================
class ProductQuantizer(_DeviceMixin):
    """Product quantizer with ADC lookup tables; API of quantization.py:414-615."""

    def __init__(self, dimensions: int, num_subspaces: int = 8, num_centroids: int = 256, device=None):
        if dimensions % num_subspaces != 0:
            raise ValueError(f"Dimensions {dimensions} not divisible by {num_subspaces}")
        self.dimensions = dimensions
        self.num_subspaces = num_subspaces
        self.subspace_dim = dimensions // num_subspaces
        self.num_centroids = num_centroids
        self.codebooks: Optional[np.ndarray] = None
        self.trained = False
        self.device = N.require_cuda(device) if device is not None else None

    def _need_trained(self):
        if not self.trained:
            raise ValueError("PQ not trained. Call train() first.")

    def _cb(self) -> torch.Tensor:
        return self._param("cb", self.codebooks).reshape(self.num_subspaces, self.num_centroids, self.subspace_dim)

    def train(self, vectors, n_iter: int = 20, sample_size: int = None) -> "ProductQuantizer":
        """k-means++ seeding + Lloyd per subspace (quantization.py:444-508), run on the GPU.

        Random draws come from the global ``np.random`` state like the reference (:458, :486, :494) so
        ``np.random.seed`` controls it; the arithmetic is batched (one distance matrix per Lloyd step, running
        minimum for the seeding instead of the reference's O(K^2 n) recomputation), so centroids agree with
        the reference in quality, not bit for bit (k-means is chaotic in its rounding)."""
        from .pq_train import train_codebooks
        v, _ = self._f32(vectors)
        if sample_size and v.shape[0] > sample_size:
            pick = np.random.choice(v.shape[0], sample_size, replace=False)
            v = v[torch.from_numpy(pick).to(v.device)]
        self.codebooks = train_codebooks(v, self.num_subspaces, self.num_centroids, n_iter).cpu().numpy()
        self.trained = True
        return self

    def encode(self, vectors):
        """float32 -> (N, M) uint8 codes, first-min argmin per subspace (quantization.py:510-539)."""
        self._need_trained()
        v, was_torch = self._f32(vectors)
        if v.shape[1] != self.dimensions:
            raise ValueError(f"expected {self.dimensions} dimensions, got {v.shape[1]}")
        codes = ops.pq_encode(v, self._cb())
        return codes if was_torch else codes.cpu().numpy()

    def build_lookup_table(self, query):
        """(M, K) float32 squared distances from the query sub-vectors to all centroids (quantization.py:541-562)."""
        self._need_trained()
        q, was_torch = self._f32(np.asarray(query, dtype=np.float32).flatten() if not isinstance(query, torch.Tensor)
                                 else query.flatten())
        lut = ops.pq_build_lut(self._cb(), q)[0]
        return lut if was_torch else lut.cpu().numpy()

    def distances_with_table(self, lookup_table, codes):
        """sqrt(sum_m table[m, codes[:, m]]) for every row (quantization.py:564-578)."""
        t = isinstance(lookup_table, torch.Tensor) or isinstance(codes, torch.Tensor)
        lut, _ = self._f32(lookup_table, two_d=False)
        lut = lut.reshape(1, self.num_subspaces, -1).contiguous()
        _, _, _, out = ops.pq_adc(lut, self._codes(codes), 0, None, 0, want_all=True)
        return out[0] if t else out[0].cpu().numpy()

    def search(self, query, codes, k: int = 10, filter_mask=None):
        """k nearest rows by asymmetric distance (quantization.py:580-597) -> (indices, distances).
        ``filter_mask`` (bool per row) is applied inside the kernel (np.where(mask, d, inf) semantics,
        vectordb_optimized.py:692)."""
        self._need_trained()
        q, was_torch = self._f32(query.flatten() if isinstance(query, torch.Tensor) else np.asarray(query).flatten())
        t = was_torch or isinstance(codes, torch.Tensor)
        dcodes = self._codes(codes)
        n = dcodes.shape[0]
        if n == 0:
            e = (np.zeros(0, np.int64), np.zeros(0, np.float32))
            return (torch.from_numpy(e[0]), torch.from_numpy(e[1])) if t else e
        lut = ops.pq_build_lut(self._cb(), q)
        kk = min(int(k), n)
        if kk > N.MAX_K:
            _, _, _, out = ops.pq_adc(lut, dcodes, 0, None, 0, want_all=True)
            return _large_k_search(self, out[0], kk, filter_mask, t)
        words = self._mask(filter_mask, n)
        if self.fast_search and ops.pq_adc_packed_supported(1, n, self.num_subspaces, self.num_centroids, kk):
            dist, idx, cnt = ops.pq_adc_packed(lut, self._packed(dcodes), kk, words, 0)
        else:
            dist, idx, cnt, _ = ops.pq_adc(lut, dcodes, kk, words, 0)
        return _finish_search(dist, idx, cnt, t)

    def search_batch_tensors(self, queries, codes, k: int = 10, filter_mask=None):
        """ADC top-k of a BATCH of queries -> device (dist [Q,k], idx [Q,k], count [Q]).  Over >= 2^20 codes ONE pass
        serves four queries (csrc/fpv_pq.cu: fixed-point u16 x 4 table entries, one 64-bit lookup per code byte, the
        survivors re-scored in fp32); the results are identical to :meth:`search` called per query."""
        self._need_trained()
        q, _ = self._f32(queries)
        dcodes = self._codes(codes)
        n = dcodes.shape[0]
        kk = max(1, min(int(k), n, N.MAX_K)) if n else 1
        lut = ops.pq_build_lut(self._cb(), q)
        words = self._mask(filter_mask, n)
        if n and self.fast_search and ops.pq_adc_packed_supported(q.shape[0], n, self.num_subspaces, self.num_centroids, kk):
            return ops.pq_adc_packed(lut, self._packed(dcodes), kk, words, 0)
        dist, idx, cnt, _ = ops.pq_adc(lut, dcodes, kk, words, 0)
        return dist, idx, cnt

    def search_batch(self, queries, codes, k: int = 10, filter_mask=None):
        """-> (indices [Q, k'], distances [Q, k']) NumPy arrays (torch tensors for torch inputs), k' = valid results."""
        t = isinstance(queries, torch.Tensor) or isinstance(codes, torch.Tensor)
        dist, idx, cnt = self.search_batch_tensors(queries, codes, k, filter_mask)
        valid = int(cnt.min().item()) if cnt.numel() else 0
        idx, dist = idx[:, :valid], dist[:, :valid]
        return (idx, dist) if t else (idx.cpu().numpy(), dist.cpu().numpy())

    #: search() uses the bank-conflict-free rotated-subspace scan (distances equal the reference's to fp32 rounding);
    #: set False to force the exact-order kernel (bit-identical to distances_with_table / the reference).
    fast_search = True

    def _packed(self, dcodes: torch.Tensor) -> torch.Tensor:
        """lane-rotated copy of a device code matrix (built once per matrix; index-build work)"""
        cache = self.__dict__.setdefault("_packed_cache", {})
        key = (dcodes.data_ptr(), tuple(dcodes.shape), dcodes._version)
        hit = cache.get(key)
        if hit is None:
            if len(cache) >= 4:
                cache.pop(next(iter(cache)))
            hit = (dcodes, ops.pq_pack(dcodes))           # keep the source alive so the pointer key stays valid
            cache[key] = hit
        return hit[1]

    def memory_usage(self, n_vectors: int) -> dict:
        """Same accounting as quantization.py:599-615."""
        float32_bytes = n_vectors * self.dimensions * 4
        code_bytes = n_vectors * self.num_subspaces
        codebook_bytes = self.num_subspaces * self.num_centroids * self.subspace_dim * 4
        total = code_bytes + codebook_bytes
        return {"original_bytes": float32_bytes, "quantized_bytes": total, "code_bytes": code_bytes,
                "codebook_bytes": codebook_bytes, "compression_ratio": float32_bytes / total,
                "savings_percent": (1 - total / float32_bytes) * 100}
