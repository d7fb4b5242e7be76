"""Tensor-core (tcgen05) batched search path — dispatch shim.  See csrc/fpv_gemm_topk.cu."""
from __future__ import annotations


def available(index, n_queries: int, k: int) -> bool:
    return False


def search(q, index, k, metric):  # pragma: no cover
    raise NotImplementedError
