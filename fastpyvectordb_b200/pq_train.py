"""GPU k-means for ProductQuantizer.train (quantization.py:444-508).

Same algorithm as the reference's ``_kmeans`` (k-means++ seeding, then ``n_iter`` Lloyd steps that leave an
empty cluster's centroid unchanged), restructured for the device and run for ALL subspaces at once:

* seeding keeps a running minimum distance per subspace (O(K n) instead of the reference's O(K^2 n) list rebuild,
  :488-491); every seeding step is a handful of batched [M, n] tensor operations;
* the assignment of a Lloyd step is exactly ``ProductQuantizer.encode`` against the current centroids, so it runs on
  the library's own argmin kernel (``fpv_pq_encode``: the reference's ``np.sum((data - c) ** 2, axis=1)`` in NumPy's
  pairwise order, first minimum wins) -- one launch for all subspaces -- followed by one ``index_add_`` + ``bincount``.

Random draws use the global ``np.random`` state in the reference's order (per subspace: one ``randint`` then K-1
``choice(n, p=...)``, each of which consumes one uniform double) so ``np.random.seed`` governs the result.
100k x 768, M = 48, K = 256, 20 iterations: 6.3 s with one subspace at a time and chunked torch distance matrices,
see ``profiles/r02_index_build.json`` for the current figure (the reference: 485 s for 5k rows).
torch ops are used here on purpose: training is index-build plumbing, not the scan hot path.
"""
from __future__ import annotations

import numpy as np
import torch


def _draw(n: int, k: int, m: int):
    """The reference's random stream: subspace by subspace, one randint and K-1 uniforms (np.random.choice with p= is
    inverse-CDF sampling on ONE uniform double: cdf.searchsorted(u, side="right"))."""
    firsts = np.empty(m, np.int64)
    us = np.empty((m, max(k - 1, 0)), np.float64)
    for j in range(m):
        firsts[j] = int(np.random.randint(n))
        us[j] = np.random.random_sample(max(k - 1, 0))
    return firsts, us


def _train(vectors: torch.Tensor, m: int, k: int, n_iter: int, assign=None) -> torch.Tensor:
    """vectors [n, m * dsub] contiguous fp32 on the device -> codebooks [m, k, dsub].

    ``assign(vectors, centroids [m, k, dsub]) -> [n, m]`` replaces the assignment step (default: the library's argmin
    kernel, which needs CUDA tensors); the host-logic test of the random stream passes a plain-torch one to run without a GPU."""
    from . import ops
    if assign is None and not vectors.is_cuda:
        raise RuntimeError("ProductQuantizer training runs on the GPU (fpv_pq_encode is the assignment step): pass CUDA tensors")
    n, dim = vectors.shape
    dsub = dim // m
    dev = vectors.device
    data3 = vectors.view(n, m, dsub).permute(1, 0, 2).contiguous()              # [m, n, dsub]
    firsts, us = _draw(n, k, m)
    firsts = torch.from_numpy(firsts).to(dev)
    us = torch.from_numpy(us).to(dev)
    ar = torch.arange(m, device=dev)
    cent = torch.zeros((m, k, dsub), dtype=torch.float32, device=dev)
    cent[:, 0] = data3[ar, firsts]
    mind = ((data3 - cent[:, 0, None, :]) ** 2).sum(dim=2)                      # [m, n]
    uniform_cdf = torch.arange(1, n + 1, device=dev, dtype=torch.float64)[None, :]
    for i in range(1, k):
        cdf = torch.cumsum(mind.double(), dim=1)
        total = cdf[:, -1]
        ok = torch.isfinite(total) & (total > 0)
        # degenerate (all points already chosen): uniform, like the reference's fallback probabilities
        target = torch.where(ok, us[:, i - 1] * total, us[:, i - 1] * n)
        cdf = torch.where(ok[:, None], cdf, uniform_cdf)
        pick = torch.searchsorted(cdf, target[:, None], right=True).clamp_(max=n - 1)[:, 0]
        cent[:, i] = data3[ar, pick]
        mind = torch.minimum(mind, ((data3 - cent[:, i, None, :]) ** 2).sum(dim=2))
    del data3, mind
    base = (ar * k)[None, :]
    flat = vectors.view(n * m, dsub)
    for _ in range(n_iter):
        codes = assign(vectors, cent) if assign is not None else ops.pq_encode(vectors, cent.contiguous())   # [n, m]: the assignment step
        idx = (codes.to(torch.int64) + base).view(-1)
        sums = torch.zeros((m * k, dsub), dtype=torch.float32, device=dev).index_add_(0, idx, flat)
        counts = torch.bincount(idx, minlength=m * k).to(torch.float32)
        alive = (counts > 0)[:, None]
        cent = torch.where(alive, sums / counts.clamp(min=1.0)[:, None], cent.view(m * k, dsub)).view(m, k, dsub)
    return cent


def _kmeans(data: torch.Tensor, k: int, n_iter: int, assign=None) -> torch.Tensor:
    """One subspace: [n, d] -> [k, d] (the reference's ``_kmeans`` signature)."""
    return _train(data.contiguous(), 1, k, n_iter, assign)[0]


def train_codebooks(vectors: torch.Tensor, m: int, k: int, n_iter: int) -> torch.Tensor:
    if k > 256:
        raise ValueError("ProductQuantizer codes are uint8: at most 256 centroids per subspace")
    return _train(vectors.contiguous(), m, k, n_iter)
