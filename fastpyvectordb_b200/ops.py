"""Tensor-level wrappers over the C-ABI: torch CUDA tensors in, torch CUDA tensors out, no host copies.

These are the array-returning fast paths underneath the drop-in classes in ``engine.py`` / ``quantizers.py``.
Every function enqueues on the current torch CUDA stream and returns without synchronising.
"""
from __future__ import annotations

from typing import Optional, Tuple

import torch

from . import _native as N

_METRICS = {"cosine": N.METRIC_COSINE, "l2": N.METRIC_L2, "l2_diff": N.METRIC_L2_DIFF}


def metric_code(metric: str) -> int:
    """"cosine", "l2", anything else is inner product — the reference's own dispatch
    (parallel_search.py:85-98, 119-134)."""
    return _METRICS.get(metric, N.METRIC_IP)


def _f32c(t: torch.Tensor, name: str) -> torch.Tensor:
    if t.dtype != torch.float32 or not t.is_contiguous():
        raise ValueError(f"{name} must be a contiguous float32 CUDA tensor")
    return t


def _outs(q: int, k: int, device):
    dist = torch.empty((q, k), dtype=torch.float32, device=device)
    idx = torch.empty((q, k), dtype=torch.int64, device=device)
    cnt = torch.empty((q,), dtype=torch.int32, device=device)
    return dist, idx, cnt


def pack_mask(mask: torch.Tensor) -> torch.Tensor:
    """bool[N] (device) -> little-endian uint32 words, bit i <-> row i (include/fpv_b200.h)."""
    n = mask.numel()
    pad = (-n) % 32
    m = mask.reshape(-1).to(torch.int64)
    if pad:
        m = torch.cat([m, torch.zeros(pad, dtype=torch.int64, device=mask.device)])
    weights = (torch.ones(32, dtype=torch.int64, device=mask.device) << torch.arange(32, device=mask.device))
    words = (m.reshape(-1, 32) * weights).sum(dim=1)
    words = torch.where(words >= 2 ** 31, words - 2 ** 32, words)      # same 32 bits, as int32
    return words.to(torch.int32).contiguous()


def pack_mask_host(mask) -> torch.Tensor:
    """bool[N] NumPy array -> int32 word tensor on the host (same bit layout as :func:`pack_mask`)."""
    import numpy as np
    bits = np.packbits(np.asarray(mask, dtype=bool).reshape(-1), bitorder="little")
    pad = (-len(bits)) % 4
    if pad:
        bits = np.concatenate([bits, np.zeros(pad, np.uint8)])
    return torch.from_numpy(bits.view(np.int32).copy())


def row_sqnorm(db: torch.Tensor) -> torch.Tensor:
    _f32c(db, "db")
    n, d = db.shape
    out = torch.empty((n,), dtype=torch.float32, device=db.device)
    with N.guard(db.device):
        N.check(N.lib().fpv_row_sqnorm_f32(N.ptr(db), n, d, d, N.ptr(out), N.stream_ptr()), "fpv_row_sqnorm_f32")
    return out


def scan_f32_topk(queries: torch.Tensor, db: torch.Tensor, k: int, metric: str, mask_words=None, row_sq=None,
                  id_base: int = 0) -> Tuple[torch.Tensor, torch.Tensor, torch.Tensor]:
    _f32c(queries, "queries"), _f32c(db, "db")
    q, d = queries.shape
    n = db.shape[0]
    if db.shape[1] != d:
        raise ValueError(f"dimension mismatch: queries {d}, database {db.shape[1]}")
    dist, idx, cnt = _outs(q, k, db.device)
    with N.guard(db.device):
        L = N.lib()
        need = L.fpv_scan_f32_workspace(q, n, d, k)
        ws = N.workspace.get(db.device, need)
        N.check(L.fpv_scan_f32_topk(N.ptr(queries), q, N.ptr(db), n, d, d, metric_code(metric), k, N.ptr(mask_words),
                                    N.ptr(row_sq), id_base, N.ptr(dist), N.ptr(idx), N.ptr(cnt), N.ptr(ws), ws.numel(),
                                    N.stream_ptr()), "fpv_scan_f32_topk")
    return dist, idx, cnt


def to_bf16(src: torch.Tensor) -> torch.Tensor:
    _f32c(src, "src")
    out = torch.empty(src.shape, dtype=torch.bfloat16, device=src.device)
    with N.guard(src.device):
        N.check(N.lib().fpv_to_bf16(N.ptr(src), N.ptr(out), src.numel(), N.stream_ptr()), "fpv_to_bf16")
    return out


def gemm_topk(queries: torch.Tensor, db: torch.Tensor, k: int, metric: str, row_sq: torch.Tensor, aux, vmax: float,
              db_lowp=None, id_base: int = 0, mask_words=None, lowp_err=(0.0, 0.0)):
    """Tensor-core filter pass + certified exact re-rank (csrc/fpv_gemm_topk.cu).  kind is TF32 unless a bf16
    shadow copy is passed."""
    _f32c(queries, "queries"), _f32c(db, "db")
    q, d = queries.shape
    n = db.shape[0]
    kind = 0 if db_lowp is None else 1
    dist, idx, cnt = _outs(q, k, db.device)
    with N.guard(db.device):
        L = N.lib()
        ws = N.workspace.get(db.device, L.fpv_gemm_topk_workspace(q, n, d, k, kind))
        N.check(L.fpv_gemm_topk_f32(N.ptr(queries), q, N.ptr(db), N.ptr(db_lowp), n, d, metric_code(metric), k, kind,
                                    N.ptr(row_sq), N.ptr(aux), float(vmax), float(lowp_err[0]) if kind else 0.0,
                                    float(lowp_err[1]) if kind else 0.0, N.ptr(mask_words), id_base, N.ptr(dist), N.ptr(idx), N.ptr(cnt),
                                    N.ptr(ws), ws.numel(), N.stream_ptr()), "fpv_gemm_topk_f32")
    return dist, idx, cnt


def gemm_workspace(q: int, n: int, d: int, k: int, kind: int, device) -> torch.Tensor:
    """A private workspace for one in-flight two-phase (sharded) search: phase 2 reads what phase 1 left in it."""
    nbytes = N.lib().fpv_gemm_topk_workspace(q, n, d, k, kind)
    return torch.empty(int(nbytes) + 256, dtype=torch.uint8, device=device)


def gemm_filter_sharded(queries: torch.Tensor, db: torch.Tensor, k: int, metric: str, row_sq, aux, vmax: float, db_lowp=None,
                        mask_words=None, lowp_err=(0.0, 0.0), ws=None) -> torch.Tensor:
    """Phase 1 of the row-sharded tensor-core search: filter this shard, return its k best approximate values
    [Q, k] (int32 holding order-preserving uint32).  The candidate lists stay in this stream's workspace for
    :func:`gemm_finish_sharded`, which must be the next GEMM call on this device / stream."""
    _f32c(queries, "queries"), _f32c(db, "db")
    q, d = queries.shape
    n = db.shape[0]
    kind = 0 if db_lowp is None else 1
    approx = torch.empty((q, k), dtype=torch.int32, device=db.device)
    with N.guard(db.device):
        L = N.lib()
        if ws is None:
            ws = N.workspace.get(db.device, L.fpv_gemm_topk_workspace(q, n, d, k, kind))
        N.check(L.fpv_gemm_filter_sharded_f32(N.ptr(queries), q, N.ptr(db), N.ptr(db_lowp), n, d, metric_code(metric), k, kind,
                                              N.ptr(row_sq), N.ptr(aux), float(vmax), float(lowp_err[0]) if kind else 0.0,
                                              float(lowp_err[1]) if kind else 0.0, N.ptr(mask_words), N.ptr(approx), N.ptr(ws),
                                              ws.numel(), N.stream_ptr()), "fpv_gemm_filter_sharded_f32")
    return approx


def gemm_sample_sharded(queries: torch.Tensor, db: torch.Tensor, k: int, metric: str, row_sq, aux, vmax: float, db_lowp=None,
                        mask_words=None, lowp_err=(0.0, 0.0), ws=None) -> torch.Tensor:
    """First half of phase 1 when the shards exchange their samples: sampling slab only, returns this shard's k best
    group values [Q, k] (int32 holding order-preserving uint32).  :func:`gemm_slabs_sharded` must be the next GEMM call
    on this device / stream."""
    _f32c(queries, "queries"), _f32c(db, "db")
    q, d = queries.shape
    n = db.shape[0]
    kind = 0 if db_lowp is None else 1
    sample = torch.empty((q, k), dtype=torch.int32, device=db.device)
    with N.guard(db.device):
        L = N.lib()
        if ws is None:
            ws = N.workspace.get(db.device, L.fpv_gemm_topk_workspace(q, n, d, k, kind))
        N.check(L.fpv_gemm_sample_sharded_f32(N.ptr(queries), q, N.ptr(db), N.ptr(db_lowp), n, d, metric_code(metric), k, kind,
                                              N.ptr(row_sq), N.ptr(aux), float(vmax), float(lowp_err[0]) if kind else 0.0,
                                              float(lowp_err[1]) if kind else 0.0, N.ptr(mask_words), N.ptr(sample), N.ptr(ws),
                                              ws.numel(), N.stream_ptr()), "fpv_gemm_sample_sharded_f32")
    return sample


def gemm_slabs_sharded(queries: torch.Tensor, db: torch.Tensor, k: int, metric: str, row_sq, aux, vmax: float, sample_all,
                       shards: int, db_lowp=None, mask_words=None, lowp_err=(0.0, 0.0), flags_ptr: int = 0, epoch: int = 0,
                       ws=None) -> torch.Tensor:
    """Second half: ``sample_all`` = the gathered samples, an int32 [shards, Q, k] tensor or the raw device address of this
    rank's peer-memory gather area (then ``flags_ptr`` / ``epoch`` say what to wait for).  Returns approx [Q, k] as
    :func:`gemm_filter_sharded`."""
    import ctypes as C
    q, d = queries.shape
    n = db.shape[0]
    kind = 0 if db_lowp is None else 1
    if isinstance(sample_all, torch.Tensor):
        if tuple(sample_all.shape) != (shards, q, k) or sample_all.dtype != torch.int32 or not sample_all.is_contiguous():
            raise ValueError("sample_all must be a contiguous int32 [shards, Q, k] tensor")
        sptr = N.ptr(sample_all)
    else:
        sptr = C.c_void_p(int(sample_all))
    approx = torch.empty((q, k), dtype=torch.int32, device=db.device)
    with N.guard(db.device):
        L = N.lib()
        if ws is None:
            ws = N.workspace.get(db.device, L.fpv_gemm_topk_workspace(q, n, d, k, kind))
        N.check(L.fpv_gemm_slabs_sharded_f32(N.ptr(queries), q, N.ptr(db), N.ptr(db_lowp), n, d, metric_code(metric), k, kind,
                                             N.ptr(row_sq), N.ptr(aux), float(vmax), float(lowp_err[0]) if kind else 0.0,
                                             float(lowp_err[1]) if kind else 0.0, N.ptr(mask_words), sptr, shards,
                                             C.c_void_p(flags_ptr) if flags_ptr else None, epoch & 0xFFFFFFFF, N.ptr(approx),
                                             N.ptr(ws), ws.numel(), N.stream_ptr()), "fpv_gemm_slabs_sharded_f32")
    return approx


def gemm_finish_sharded(queries: torch.Tensor, db: torch.Tensor, k: int, metric: str, row_sq, approx_all: torch.Tensor,
                        db_lowp=None, id_base: int = 0, mask_words=None, ws=None):
    """Phase 2: ``approx_all`` [shards, Q, k] (the all-gathered phase-1 outputs) -> this shard's exact
    (dist, global idx, count) lists under the global limit."""
    q, d = queries.shape
    n = db.shape[0]
    kind = 0 if db_lowp is None else 1
    shards = approx_all.shape[0]
    if tuple(approx_all.shape[1:]) != (q, k) or approx_all.dtype != torch.int32 or not approx_all.is_contiguous():
        raise ValueError("approx_all must be a contiguous int32 [shards, Q, k] tensor")
    dist, idx, cnt = _outs(q, k, db.device)
    with N.guard(db.device):
        L = N.lib()
        if ws is None:
            ws = N.workspace.get(db.device, L.fpv_gemm_topk_workspace(q, n, d, k, kind))
        N.check(L.fpv_gemm_finish_sharded_f32(N.ptr(queries), q, N.ptr(db), N.ptr(db_lowp), n, d, metric_code(metric), k, kind,
                                              N.ptr(row_sq), N.ptr(mask_words), id_base, N.ptr(approx_all), shards, N.ptr(dist),
                                              N.ptr(idx), N.ptr(cnt), N.ptr(ws), ws.numel(), N.stream_ptr()),
                "fpv_gemm_finish_sharded_f32")
    return dist, idx, cnt


def gemm_finish_sharded_peer(queries: torch.Tensor, db: torch.Tensor, k: int, metric: str, row_sq, gather_ptr: int, flags_ptr: int,
                             epoch: int, shards: int, db_lowp=None, id_base: int = 0, mask_words=None):
    """Phase 2 reading the gathered phase-1 values from this rank's peer-memory gather area (raw device addresses);
    the kernel waits for the arrival flags first."""
    import ctypes as C
    q, d = queries.shape
    n = db.shape[0]
    kind = 0 if db_lowp is None else 1
    dist, idx, cnt = _outs(q, k, db.device)
    with N.guard(db.device):
        L = N.lib()
        ws = N.workspace.get(db.device, L.fpv_gemm_topk_workspace(q, n, d, k, kind))
        N.check(L.fpv_gemm_finish_sharded_peer_f32(N.ptr(queries), q, N.ptr(db), N.ptr(db_lowp), n, d, metric_code(metric), k, kind,
                                                   N.ptr(row_sq), N.ptr(mask_words), id_base, C.c_void_p(gather_ptr), shards,
                                                   C.c_void_p(flags_ptr), epoch & 0xFFFFFFFF, N.ptr(dist), N.ptr(idx), N.ptr(cnt),
                                                   N.ptr(ws), ws.numel(), N.stream_ptr()), "fpv_gemm_finish_sharded_peer_f32")
    return dist, idx, cnt


def gemm_last_flags(q: int, n: int, d: int, k: int, kind: int, device) -> torch.Tensor:
    """uint32-as-int32 [q]: 1 where the last gemm_topk call with this shape fell back to the exact scan."""
    off = N.lib().fpv_gemm_topk_flags_offset(q, n, d, k, kind)
    ws = N.workspace.get(device, off + 4 * q)
    return ws[off:off + 4 * q].view(torch.int32).clone()


def distances_f32(queries: torch.Tensor, db: torch.Tensor, metric: str, row_sq=None) -> torch.Tensor:
    _f32c(queries, "queries"), _f32c(db, "db")
    q, d = queries.shape
    n = db.shape[0]
    out = torch.empty((q, n), dtype=torch.float32, device=db.device)
    with N.guard(db.device):
        L = N.lib()
        ws = N.workspace.get(db.device, L.fpv_scan_f32_workspace(q, n, d, 0))
        N.check(L.fpv_distances_f32(N.ptr(queries), q, N.ptr(db), n, d, d, metric_code(metric), N.ptr(row_sq), N.ptr(out),
                                    N.ptr(ws), ws.numel(), N.stream_ptr()), "fpv_distances_f32")
    return out


def rerank_f32(queries: torch.Tensor, db: torch.Tensor, cand_idx: torch.Tensor, k: int, metric: str, row_sq=None,
               id_base: int = 0):
    _f32c(queries, "queries"), _f32c(db, "db")
    q, d = queries.shape
    n = db.shape[0]
    cand_idx = cand_idx.to(torch.int64).contiguous()
    c = cand_idx.shape[1]
    k = min(k, c)
    dist, idx, cnt = _outs(q, k, db.device)
    with N.guard(db.device):
        N.check(N.lib().fpv_rerank_f32(N.ptr(queries), q, N.ptr(db), n, d, d, metric_code(metric), N.ptr(cand_idx), c, k,
                                       N.ptr(row_sq), id_base, N.ptr(dist), N.ptr(idx), N.ptr(cnt), N.stream_ptr()),
                "fpv_rerank_f32")
    return dist, idx, cnt


def merge_topk(dist: torch.Tensor, idx: torch.Tensor, k_out: int):
    """dist/idx: [shards][Q][k_in] -> ([Q][k_out], [Q][k_out], [Q])."""
    s, q, k_in = dist.shape
    dist = dist.contiguous()
    idx = idx.contiguous()
    od, oi, oc = _outs(q, k_out, dist.device)
    with N.guard(dist.device):
        N.check(N.lib().fpv_merge_topk(N.ptr(dist), N.ptr(idx), s, q, k_in, k_out, N.ptr(od), N.ptr(oi), N.ptr(oc),
                                       N.stream_ptr()), "fpv_merge_topk")
    return od, oi, oc


def pack_topk(dist: torch.Tensor, idx: torch.Tensor, k_pad: int, id_base: int) -> torch.Tensor:
    """(dist [Q,kl] f32, idx [Q,kl] global i64) -> [Q,k_pad] int64 wire keys (see fpv_pack_topk)."""
    q, kl = dist.shape
    out = torch.empty((q, k_pad), dtype=torch.int64, device=dist.device)
    with N.guard(dist.device):
        N.check(N.lib().fpv_pack_topk(N.ptr(dist.contiguous()) if kl else None, N.ptr(idx.contiguous()) if kl else None, q, kl,
                                      k_pad, id_base, N.ptr(out), N.stream_ptr()), "fpv_pack_topk")
    return out


def merge_packed(packed: torch.Tensor, shard_bases: torch.Tensor, k_out: int):
    """packed [S,Q,k_in] int64 wire keys, shard_bases [S] int64 (device) -> merged (dist, idx, count)."""
    s, q, k_in = packed.shape
    od, oi, oc = _outs(q, k_out, packed.device)
    with N.guard(packed.device):
        N.check(N.lib().fpv_merge_packed(N.ptr(packed), N.ptr(shard_bases), s, q, k_in, k_out, N.ptr(od), N.ptr(oi), N.ptr(oc),
                                         N.stream_ptr()), "fpv_merge_packed")
    return od, oi, oc


def merge_packed_peer(gather_ptr: int, flags_ptr: int, epoch: int, shards: int, q: int, k_in: int, shard_bases: torch.Tensor, k_out: int):
    """The same merge over this rank's peer-memory gather area (raw device addresses): the kernel waits for the
    ``shards`` arrival flags to reach ``epoch`` first (csrc/fpv_peer.cu)."""
    import ctypes as C
    od, oi, oc = _outs(q, k_out, shard_bases.device)
    with N.guard(shard_bases.device):
        N.check(N.lib().fpv_merge_packed_peer(C.c_void_p(gather_ptr), N.ptr(shard_bases), shards, q, k_in, k_out, C.c_void_p(flags_ptr),
                                              epoch & 0xFFFFFFFF, N.ptr(od), N.ptr(oi), N.ptr(oc), N.stream_ptr()),
                "fpv_merge_packed_peer")
    return od, oi, oc


# ---------------------------------------------------------------------------------------------- binary
def bq_encode(vectors: torch.Tensor, thresholds: torch.Tensor) -> torch.Tensor:
    _f32c(vectors, "vectors")
    n, d = vectors.shape
    out = torch.empty((n, (d + 7) // 8), dtype=torch.uint8, device=vectors.device)
    with N.guard(vectors.device):
        N.check(N.lib().fpv_bq_encode(N.ptr(vectors), n, d, d, N.ptr(thresholds), N.ptr(out), N.stream_ptr()), "fpv_bq_encode")
    return out


#: batches of at least this many queries take the tensor-core Hamming scan when the shape allows (False: never)
HAMMING_TENSOR_CORES = True


def hamming_mma_supported(q: int, n: int, nbytes: int, k: int) -> bool:
    return bool(HAMMING_TENSOR_CORES and N.lib().fpv_hamming_mma_supported(q, n, nbytes, k))


def hamming_mma(qbits: torch.Tensor, codes: torch.Tensor, k: int, dims: int = 0, mask_words=None, id_base: int = 0):
    """Batched Hamming top-k on the int8 tensor cores (csrc/fpv_hamming_mma.cu); same results as :func:`hamming`."""
    q, nbytes = qbits.shape
    n = codes.shape[0]
    dist, idx, cnt = _outs(q, k, codes.device)
    with N.guard(codes.device):
        L = N.lib()
        ws = N.workspace.get(codes.device, L.fpv_hamming_mma_workspace(q, n, nbytes, k))
        N.check(L.fpv_hamming_mma_topk(N.ptr(qbits.contiguous()), q, N.ptr(codes), n, nbytes, dims, k, N.ptr(mask_words), id_base,
                                       N.ptr(dist), N.ptr(idx), N.ptr(cnt), N.ptr(ws), ws.numel(), N.stream_ptr()),
                "fpv_hamming_mma_topk")
    return dist, idx, cnt


def hamming_mma_dots(qbits: torch.Tensor, codes: torch.Tensor, dims: int = 0) -> torch.Tensor:
    """Test hook -> int32 [32, N]: row i < Q = -popc(x & q_i & dimmask), row 31 = -popc(x & dimmask)."""
    q, nbytes = qbits.shape
    n = codes.shape[0]
    out = torch.zeros((32, n), dtype=torch.int32, device=codes.device)
    with N.guard(codes.device):
        L = N.lib()
        ws = N.workspace.get(codes.device, L.fpv_hamming_mma_workspace(q, n, nbytes, 1))
        N.check(L.fpv_hamming_mma_dots(N.ptr(qbits.contiguous()), q, N.ptr(codes), n, nbytes, dims, N.ptr(out), N.ptr(ws), ws.numel(),
                                       N.stream_ptr()), "fpv_hamming_mma_dots")
    return out


def hamming(qbits: torch.Tensor, codes: torch.Tensor, k: int, dims: int = 0, mask_words=None, id_base: int = 0,
            want_all: bool = False):
    q, nbytes = qbits.shape
    if (not want_all and k > 0 and codes.dtype == torch.uint8 and codes.is_contiguous() and codes.shape[1] == nbytes
            and codes.data_ptr() % 16 == 0 and hamming_mma_supported(q, codes.shape[0], nbytes, k)):
        d, i, c = hamming_mma(qbits, codes, k, dims, mask_words, id_base)
        return d, i, c, None
    n = codes.shape[0]
    if codes.dtype != torch.uint8 or qbits.dtype != torch.uint8 or not codes.is_contiguous() or not qbits.is_contiguous():
        raise ValueError("codes/qbits must be contiguous uint8 CUDA tensors")
    if n and codes.shape[1] != nbytes:
        raise ValueError(f"code width mismatch: query {nbytes} bytes, database {codes.shape[1]} bytes")
    dist = idx = cnt = None
    if k > 0:
        dist, idx, cnt = _outs(q, k, codes.device)
    out_all = torch.empty((q, n), dtype=torch.float32, device=codes.device) if want_all else None
    with N.guard(codes.device):
        L = N.lib()
        ws = N.workspace.get(codes.device, L.fpv_hamming_workspace(q, n, nbytes, k))
        N.check(L.fpv_hamming_topk(N.ptr(qbits), q, N.ptr(codes), n, nbytes, int(dims or 0), k, N.ptr(mask_words), id_base,
                                   N.ptr(dist), N.ptr(idx), N.ptr(cnt), N.ptr(out_all), N.ptr(ws), ws.numel(),
                                   N.stream_ptr()), "fpv_hamming_topk")
    return dist, idx, cnt, out_all


# ---------------------------------------------------------------------------------------------- product
def pq_encode(vectors: torch.Tensor, codebooks: torch.Tensor) -> torch.Tensor:
    _f32c(vectors, "vectors"), _f32c(codebooks, "codebooks")
    n, d = vectors.shape
    m, kc, dsub = codebooks.shape
    out = torch.empty((n, m), dtype=torch.uint8, device=vectors.device)
    with N.guard(vectors.device):
        N.check(N.lib().fpv_pq_encode(N.ptr(vectors), n, d, d, N.ptr(codebooks), m, kc, N.ptr(out), N.stream_ptr()),
                "fpv_pq_encode")
    return out


def pq_build_lut(codebooks: torch.Tensor, queries: torch.Tensor) -> torch.Tensor:
    _f32c(queries, "queries"), _f32c(codebooks, "codebooks")
    m, kc, dsub = codebooks.shape
    q = queries.shape[0]
    lut = torch.empty((q, m, kc), dtype=torch.float32, device=queries.device)
    with N.guard(queries.device):
        N.check(N.lib().fpv_pq_build_lut(N.ptr(codebooks), m, kc, dsub, N.ptr(queries), q, N.ptr(lut), N.stream_ptr()),
                "fpv_pq_build_lut")
    return lut


def pq_adc(lut: torch.Tensor, codes: torch.Tensor, k: int, mask_words=None, id_base: int = 0, want_all: bool = False):
    _f32c(lut, "lut")
    q, m, kc = lut.shape
    n = codes.shape[0]
    if codes.dtype != torch.uint8 or not codes.is_contiguous() or (n and codes.shape[1] != m):
        raise ValueError("codes must be a contiguous uint8 [N, M] CUDA tensor")
    dist = idx = cnt = None
    if k > 0:
        dist, idx, cnt = _outs(q, k, codes.device)
    out_all = torch.empty((q, n), dtype=torch.float32, device=codes.device) if want_all else None
    with N.guard(codes.device):
        L = N.lib()
        ws = N.workspace.get(codes.device, L.fpv_pq_adc_workspace(q, n, m, kc, k))
        N.check(L.fpv_pq_adc_topk(N.ptr(lut), q, N.ptr(codes), n, m, kc, k, N.ptr(mask_words), id_base, N.ptr(dist),
                                  N.ptr(idx), N.ptr(cnt), N.ptr(out_all), N.ptr(ws), ws.numel(), N.stream_ptr()),
                "fpv_pq_adc_topk")
    return dist, idx, cnt, out_all


def pq_pack(codes: torch.Tensor) -> torch.Tensor:
    """[N, M] uint8 -> lane-rotated copy for pq_adc_packed (M % 16 == 0)."""
    n, m = codes.shape
    out = torch.empty_like(codes)
    with N.guard(codes.device):
        N.check(N.lib().fpv_pq_pack(N.ptr(codes), n, m, N.ptr(out), N.stream_ptr()), "fpv_pq_pack")
    return out


def pq_adc_packed_supported(q: int, n: int, m: int, kc: int, k: int) -> bool:
    return k >= 1 and N.lib().fpv_pq_adc_packed_workspace(q, n, m, kc, k) > 0


def pq_adc_packed(lut: torch.Tensor, packed: torch.Tensor, k: int, mask_words=None, id_base: int = 0):
    _f32c(lut, "lut")
    q, m, kc = lut.shape
    n = packed.shape[0]
    dist, idx, cnt = _outs(q, k, packed.device)
    with N.guard(packed.device):
        L = N.lib()
        ws = N.workspace.get(packed.device, L.fpv_pq_adc_packed_workspace(q, n, m, kc, k))
        N.check(L.fpv_pq_adc_packed_topk(N.ptr(lut), q, N.ptr(packed), n, m, kc, k, N.ptr(mask_words), id_base,
                                         N.ptr(dist), N.ptr(idx), N.ptr(cnt), N.ptr(ws), ws.numel(), N.stream_ptr()),
                "fpv_pq_adc_packed_topk")
    return dist, idx, cnt


# ---------------------------------------------------------------------------------------------- scalar
def sq_encode(vectors: torch.Tensor, min_vals: torch.Tensor, scale: torch.Tensor) -> torch.Tensor:
    _f32c(vectors, "vectors")
    n, d = vectors.shape
    out = torch.empty((n, d), dtype=torch.uint8, device=vectors.device)
    with N.guard(vectors.device):
        N.check(N.lib().fpv_sq_encode(N.ptr(vectors), n, d, d, N.ptr(min_vals), N.ptr(scale), N.ptr(out), N.stream_ptr()),
                "fpv_sq_encode")
    return out


def sq_scan(kind: int, qcodes: torch.Tensor, codes: torch.Tensor, min_vals: torch.Tensor, scale: torch.Tensor, k: int,
            mask_words=None, id_base: int = 0, want_all: bool = False):
    q, d = qcodes.shape
    n = codes.shape[0]
    if codes.dtype != torch.uint8 or not codes.is_contiguous() or (n and codes.shape[1] != d):
        raise ValueError("codes must be a contiguous uint8 [N, D] CUDA tensor")
    dist = idx = cnt = None
    if k > 0:
        dist, idx, cnt = _outs(q, k, codes.device)
    out_all = torch.empty((q, n), dtype=torch.float32, device=codes.device) if want_all else None
    with N.guard(codes.device):
        L = N.lib()
        ws = N.workspace.get(codes.device, L.fpv_sq_workspace(q, n, d, k))
        N.check(L.fpv_sq_topk(kind, N.ptr(qcodes), q, N.ptr(codes), n, d, N.ptr(min_vals), N.ptr(scale), k,
                              N.ptr(mask_words), id_base, N.ptr(dist), N.ptr(idx), N.ptr(cnt), N.ptr(out_all), N.ptr(ws),
                              ws.numel(), N.stream_ptr()), "fpv_sq_topk")
    return dist, idx, cnt, out_all


# ---- uint8 scalar L2 on the int8 tensor cores (csrc/fpv_sq_mma.cu) -------------------------------------------------------
def sq_mma_supported(n: int, d: int, k: int) -> bool:
    return bool(N.lib().fpv_sq_mma_supported(n, d, k))


def sq_row_term(codes: torch.Tensor, scale: torch.Tensor):
    """Per-row term C_row = sum_j (scale_j/255)^2 code_j^2 of the expanded distance, and its maximum (a 1-element device
    tensor): computed once per code matrix (index build)."""
    n, d = codes.shape
    if codes.dtype != torch.uint8 or not codes.is_contiguous():
        raise ValueError("codes must be a contiguous uint8 [N, D] CUDA tensor")
    term = torch.empty((n,), dtype=torch.float32, device=codes.device)
    tmax = torch.zeros((1,), dtype=torch.float32, device=codes.device)
    with N.guard(codes.device):
        N.check(N.lib().fpv_sq_row_term(N.ptr(codes), n, d, N.ptr(scale), N.ptr(term), N.ptr(tmax), N.stream_ptr()), "fpv_sq_row_term")
    return term, tmax


def sq_l2_mma(qcodes: torch.Tensor, codes: torch.Tensor, min_vals: torch.Tensor, scale: torch.Tensor, row_term: torch.Tensor,
              row_term_max: torch.Tensor, k: int, mask_words=None, id_base: int = 0):
    """Batched L2 top-k over uint8 codes on the int8 tensor cores; same results as ``sq_scan(SQ_L2, ...)``."""
    q, d = qcodes.shape
    n = codes.shape[0]
    if codes.dtype != torch.uint8 or not codes.is_contiguous() or codes.shape[1] != d:
        raise ValueError("codes must be a contiguous uint8 [N, D] CUDA tensor")
    dist, idx, cnt = _outs(q, k, codes.device)
    with N.guard(codes.device):
        L = N.lib()
        ws = N.workspace.get(codes.device, L.fpv_sq_mma_workspace(q, n, d, k))
        N.check(L.fpv_sq_l2_mma_topk(N.ptr(qcodes), q, N.ptr(codes), n, d, N.ptr(min_vals), N.ptr(scale), N.ptr(row_term),
                                     N.ptr(row_term_max), k, N.ptr(mask_words), id_base, N.ptr(dist), N.ptr(idx), N.ptr(cnt),
                                     N.ptr(ws), ws.numel(), N.stream_ptr()), "fpv_sq_l2_mma_topk")
    return dist, idx, cnt


def sq_row_terms_dc(codes: torch.Tensor, min_vals: torch.Tensor, scale: torch.Tensor):
    """Per-row terms of the dot / cosine tensor-core scans: R_row = sum_j code_j, invn_row = 1 / (|decode(row)| + 1e-8),
    and their maxima (a 2-element device tensor): computed once per code matrix (index build)."""
    n, d = codes.shape
    if codes.dtype != torch.uint8 or not codes.is_contiguous():
        raise ValueError("codes must be a contiguous uint8 [N, D] CUDA tensor")
    rsum = torch.empty((n,), dtype=torch.float32, device=codes.device)
    rinv = torch.empty((n,), dtype=torch.float32, device=codes.device)
    maxima = torch.zeros((2,), dtype=torch.float32, device=codes.device)
    with N.guard(codes.device):
        N.check(N.lib().fpv_sq_row_terms_dc(N.ptr(codes), n, d, N.ptr(min_vals), N.ptr(scale), N.ptr(rsum), N.ptr(rinv),
                                            N.ptr(maxima), N.stream_ptr()), "fpv_sq_row_terms_dc")
    return rsum, rinv, maxima


def sq_dc_mma(kind: int, qcodes: torch.Tensor, codes: torch.Tensor, min_vals: torch.Tensor, scale: torch.Tensor,
              row_sum: torch.Tensor, row_invn: torch.Tensor, maxima: torch.Tensor, k: int, mask_words=None, id_base: int = 0):
    """Batched dot / cosine top-k over uint8 codes on the int8 tensor cores; same results as ``sq_scan(kind, ...)``."""
    q, d = qcodes.shape
    n = codes.shape[0]
    if codes.dtype != torch.uint8 or not codes.is_contiguous() or codes.shape[1] != d:
        raise ValueError("codes must be a contiguous uint8 [N, D] CUDA tensor")
    dist, idx, cnt = _outs(q, k, codes.device)
    with N.guard(codes.device):
        L = N.lib()
        ws = N.workspace.get(codes.device, L.fpv_sq_mma_workspace(q, n, d, k))
        N.check(L.fpv_sq_dc_mma_topk(kind, N.ptr(qcodes), q, N.ptr(codes), n, d, N.ptr(min_vals), N.ptr(scale), N.ptr(row_sum),
                                     N.ptr(row_invn), N.ptr(maxima), k, N.ptr(mask_words), id_base, N.ptr(dist), N.ptr(idx),
                                     N.ptr(cnt), N.ptr(ws), ws.numel(), N.stream_ptr()), "fpv_sq_dc_mma_topk")
    return dist, idx, cnt


def sq_mma_last_flags(q: int, n: int, d: int, k: int, device) -> torch.Tensor:
    off = N.lib().fpv_sq_mma_flags_offset(q, n, d, k)
    ws = N.workspace.get(device, off + 4 * q)
    return ws[off:off + 4 * q].view(torch.int32).clone()


def sq_mma_limb_dots(qcode: torch.Tensor, codes: torch.Tensor, scale: torch.Tensor):
    """Test hook -> (limbs [3, Dp] uint8, tensor-core dots [3, N] int32, CUDA-core dots [3, N] int32) for ONE query."""
    n, d = codes.shape
    dp = (d + 127) // 128 * 128
    limbs = torch.empty((3, dp), dtype=torch.uint8, device=codes.device)
    a = torch.empty((3, n), dtype=torch.int32, device=codes.device)
    b = torch.empty((3, n), dtype=torch.int32, device=codes.device)
    with N.guard(codes.device):
        L = N.lib()
        ws = N.workspace.get(codes.device, L.fpv_sq_mma_workspace(1, n, d, 1))
        N.check(L.fpv_sq_mma_limb_dots(N.ptr(qcode.contiguous()), N.ptr(codes), n, d, N.ptr(scale), N.ptr(limbs), N.ptr(a), N.ptr(b),
                                       N.ptr(ws), ws.numel(), N.stream_ptr()), "fpv_sq_mma_limb_dots")
    return limbs, a, b
