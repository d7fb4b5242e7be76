"""GPU k-means for ProductQuantizer.train (quantization.py:444-508).

Same algorithm as the reference's ``_kmeans`` (k-means++ seeding, then ``n_iter`` Lloyd steps that leave an
empty cluster's centroid unchanged), restructured for the device: the seeding keeps a running minimum
distance (O(K n) instead of the reference's O(K^2 n) list rebuild, :488-491) and a Lloyd step is one
distance matrix + ``index_add_``.  Random draws use the global ``np.random`` state in the reference's order
(one ``randint`` then K-1 ``choice(n, p=...)`` per subspace) so ``np.random.seed`` governs it.
torch ops are used here on purpose: training is index-build plumbing, not the scan hot path.
"""
from __future__ import annotations

import numpy as np
import torch


def _kmeans(data: torch.Tensor, k: int, n_iter: int) -> torch.Tensor:
    n, d = data.shape
    cent = torch.zeros((k, d), dtype=torch.float32, device=data.device)
    cent[0] = data[np.random.randint(n)]
    mind = ((data - cent[0]) ** 2).sum(dim=1)
    for i in range(1, k):
        total = mind.sum()
        probs = (mind / total).double().cpu().numpy()
        s = probs.sum()
        if not np.isfinite(s) or s <= 0:
            probs = np.full(n, 1.0 / n)
        else:
            probs = probs / s
        cent[i] = data[np.random.choice(n, p=probs)]
        mind = torch.minimum(mind, ((data - cent[i]) ** 2).sum(dim=1))
    for _ in range(n_iter):
        assign = torch.empty(n, dtype=torch.int64, device=data.device)
        step = max(1, (64 << 20) // (4 * k * max(d, 1)))
        for s0 in range(0, n, step):
            blk = data[s0:s0 + step]
            dist = ((blk[:, None, :] - cent[None, :, :]) ** 2).sum(dim=2)      # exact form, like the reference
            assign[s0:s0 + step] = dist.argmin(dim=1)
        sums = torch.zeros_like(cent).index_add_(0, assign, data)
        counts = torch.bincount(assign, minlength=k).to(torch.float32)
        alive = counts > 0
        cent[alive] = sums[alive] / counts[alive, None]
    return cent


def train_codebooks(vectors: torch.Tensor, m: int, k: int, n_iter: int) -> torch.Tensor:
    n, dim = vectors.shape
    dsub = dim // m
    out = torch.zeros((m, k, dsub), dtype=torch.float32, device=vectors.device)
    for j in range(m):
        out[j] = _kmeans(vectors[:, j * dsub:(j + 1) * dsub].contiguous(), k, n_iter)
    return out
