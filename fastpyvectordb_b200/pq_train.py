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
    # The reference draws one randint and then K-1 ``np.random.choice(n, p=...)`` (quantization.py:486-494); choice() is
    # inverse-CDF sampling on ONE uniform double (cdf.searchsorted(u, side="right")).  Drawing the same numbers up front
    # and doing the cumsum + searchsorted on the device keeps np.random.seed in charge of the seeding without a
    # host round trip per centroid (the first version copied n probabilities to the host K times per subspace).
    first = int(np.random.randint(n))
    us = torch.from_numpy(np.random.random_sample(max(k - 1, 0))).to(data.device)        # float64
    cent[0] = data[first]
    mind = ((data - cent[0]) ** 2).sum(dim=1)
    for i in range(1, k):
        cdf = torch.cumsum(mind.double(), dim=0)
        total = cdf[-1]
        ok = torch.isfinite(total) & (total > 0)
        # degenerate (all points already chosen): uniform, like the reference's fallback probabilities
        target = torch.where(ok, us[i - 1] * total, us[i - 1] * n)
        cdf = torch.where(ok, cdf, torch.arange(1, n + 1, device=data.device, dtype=torch.float64))
        pick = torch.searchsorted(cdf, target.reshape(1), right=True).clamp_(max=n - 1)
        cent[i] = data[pick[0]]
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
        alive = (counts > 0)[:, None]
        cent = torch.where(alive, sums / counts.clamp(min=1.0)[:, None], cent)   # an empty cluster keeps its centroid
    return cent


def train_codebooks(vectors: torch.Tensor, m: int, k: int, n_iter: int) -> torch.Tensor:
    n, dim = vectors.shape
    dsub = dim // m
    out = torch.zeros((m, k, dsub), dtype=torch.float32, device=vectors.device)
    for j in range(m):
        out[j] = _kmeans(vectors[:, j * dsub:(j + 1) * dsub].contiguous(), k, n_iter)
    return out
