"""Dispatch of query batches to the tensor-core path (csrc/fpv_gemm_topk.cu).

The path is exact (certified re-rank + device-side exact fallback), so the dispatch is purely a performance
decision: batches of at least ``ParallelSearchEngine.GEMM_MIN_BATCH`` queries (any batch once the bf16 shadow copy
exists), k <= 256, rows a multiple of 4 floats and at least 4096 of them go to tcgen05, with or without a row
filter (applied in the epilogue); everything else stays on the HBM-bound fp32 scan.
"""
from __future__ import annotations

import os

import torch

from . import ops

MAX_K = 256
# "tf32": tensor-core pass straight from the fp32 rows (no extra memory); "bf16": pass over a bf16 shadow copy
# (half the operand traffic and twice the MMA rate, +50% memory).  Both give the same exact results.
MODE = os.environ.get("FPV_GEMM_MODE", "auto")


def available(index, n_queries: int, k: int) -> bool:
    if k > MAX_K or index.n < 4096 or index.n >= 2 ** 31:
        return False
    if index.d % 4 != 0 or index.d < 16:
        return False
    return True


def _aux(index, metric: str):
    """per-row epilogue operand, built once per index and metric (index-build plumbing, torch ops)"""
    cache = index.__dict__.setdefault("_gemm_aux", {})
    if "vmax" not in cache:
        cache["vmax"] = float(torch.sqrt(index.row_sq.max()).item()) if index.n else 0.0
    if metric == "cosine":
        if "rinv" not in cache:
            cache["rinv"] = (1.0 / (torch.sqrt(index.row_sq) + 1e-10)).contiguous()
        return cache["rinv"], cache["vmax"]
    if metric == "l2":
        return index.row_sq, cache["vmax"]
    return None, cache["vmax"]


# "auto": the BF16 pass whenever its shadow copy fits.  Large batches are tensor bound (2x MMA rate); small ones are
# bound by reading the rows once, and the shadow is half the bytes (measured, 1M x 768: Q = 64 0.61 -> 0.43 ms).
BF16_MIN_BATCH = 1


def _effective_mode(mode, index, k: int, n_queries: int = 0) -> str:
    mode = mode or MODE
    if mode == "auto":
        mode = "bf16" if n_queries >= BF16_MIN_BATCH else "tf32"
        if mode == "bf16" and index._lowp is None:
            free, _total = torch.cuda.mem_get_info(index.device)
            if index.n * index.d * 2 > 0.5 * free:
                mode = "tf32"                               # no room for the shadow copy
    if mode == "bf16" and (k > 128 or index.d % 8 != 0):
        return "tf32"           # the coarser pass keeps 4k candidates per query; beyond k=128 TF32 is the better filter
    return mode


def has_shadow(index, k: int) -> bool:
    """True when the index already holds the bf16 shadow copy and a search with this k would use it."""
    return index._lowp is not None and _effective_mode(None, index, k, 1) == "bf16"


def last_fallback_fraction(index, n_queries: int, k: int, mode: str = None) -> float:
    """Fraction of the queries of the most recent search() of this shape that failed the certificate and were
    recomputed by the exact scan (diagnostics; forces a device sync)."""
    kind = 0 if _effective_mode(mode, index, k, n_queries) == "tf32" else 1
    flags = ops.gemm_last_flags(n_queries, index.n, index.d, k, kind, index.device)
    return float(flags.float().mean().item())


def search(q: torch.Tensor, index, k: int, metric: str, mode: str = None, mask_words=None):
    mode = _effective_mode(mode, index, k, q.shape[0])
    aux, vmax = _aux(index, metric)
    lowp, err = None, (0.0, 0.0)
    if mode == "bf16":
        if index._lowp is None:
            index._lowp = ops.to_bf16(index.rows)
        lowp = index._lowp
        err = _shadow_error(index)
    return ops.gemm_topk(q, index.rows, k, metric, index.row_sq, aux, vmax, lowp, index.id_base, mask_words, err)


def _shadow_error(index):
    """(max_i |v_i - bf16(v_i)|, max_i |v_i - bf16(v_i)| / |v_i|) of the shadow copy, measured once per index (index
    build plumbing, torch ops in row chunks).  Inflated by 1e-4 for the fp32 rounding of the norms themselves."""
    cache = index.__dict__.setdefault("_gemm_aux", {})
    if "lowp_err" not in cache:
        abs_max, rel_max = 0.0, 0.0
        step = max(1, (64 << 20) // max(index.d, 1))
        for lo in range(0, index.n, step):
            rows = index.rows[lo:lo + step]
            r = torch.linalg.vector_norm(rows - index._lowp[lo:lo + step].float(), dim=1)
            nrm = torch.sqrt(index.row_sq[lo:lo + step])
            abs_max = max(abs_max, float(r.max().item()))
            rel_max = max(rel_max, float(torch.where(nrm > 0, r / nrm, torch.zeros_like(r)).max().item()))
        cache["lowp_err"] = (abs_max * 1.0001 + 1e-30, rel_max * 1.0001 + 1e-30)
    return cache["lowp_err"]


# ---- row-sharded search: the same path split around the one exchange it needs (sharded.py drives the collectives) ----
def sharded_bounds(index, mode: str):
    """(vmax, (db_err_abs, db_err_rel)) of THIS shard for ``mode``; the caller replaces them by the maxima over all
    shards (set_sharded_bounds) so that one error bound E holds on every shard."""
    _aux(index, "ip")
    cache = index.__dict__["_gemm_aux"]
    err = (0.0, 0.0)
    if mode == "bf16":
        if index._lowp is None:
            index._lowp = ops.to_bf16(index.rows)
        err = _shadow_error(index)
    return cache["vmax"], err


def set_sharded_bounds(index, vmax: float, err):
    cache = index.__dict__.setdefault("_gemm_aux", {})
    cache["vmax"] = float(vmax)
    if err[0] > 0.0:
        cache["lowp_err"] = (float(err[0]), float(err[1]))


def filter_sharded(q: torch.Tensor, index, k: int, metric: str, mode: str, mask_words=None, ws=None) -> torch.Tensor:
    aux, vmax = _aux(index, metric)
    lowp, err = (index._lowp, _shadow_error(index)) if mode == "bf16" else (None, (0.0, 0.0))
    return ops.gemm_filter_sharded(q, index.rows, k, metric, index.row_sq, aux, vmax, lowp, mask_words, err, ws)


def sample_sharded(q: torch.Tensor, index, k: int, metric: str, mode: str, mask_words=None, ws=None) -> torch.Tensor:
    aux, vmax = _aux(index, metric)
    lowp, err = (index._lowp, _shadow_error(index)) if mode == "bf16" else (None, (0.0, 0.0))
    return ops.gemm_sample_sharded(q, index.rows, k, metric, index.row_sq, aux, vmax, lowp, mask_words, err, ws)


def slabs_sharded(q: torch.Tensor, index, k: int, metric: str, mode: str, sample_all, shards: int, mask_words=None,
                  flags_ptr: int = 0, epoch: int = 0, ws=None) -> torch.Tensor:
    aux, vmax = _aux(index, metric)
    lowp, err = (index._lowp, _shadow_error(index)) if mode == "bf16" else (None, (0.0, 0.0))
    return ops.gemm_slabs_sharded(q, index.rows, k, metric, index.row_sq, aux, vmax, sample_all, shards, lowp, mask_words, err,
                                  flags_ptr, epoch, ws)


def finish_sharded(q: torch.Tensor, index, k: int, metric: str, mode: str, approx_all: torch.Tensor, mask_words=None, ws=None):
    lowp = index._lowp if mode == "bf16" else None
    return ops.gemm_finish_sharded(q, index.rows, k, metric, index.row_sq, approx_all, lowp, index.id_base, mask_words, ws)
