#!/usr/bin/env python
"""Small-batch search (Q queries, 1M x 768, L2 top-100) a few times: run under
`ncu --metrics gpu__time_duration.sum --clock-control none --csv` to get the per-launch breakdown, or alone to
print the CUDA-event time per search.  usage: profile_small_batch.py [Q] [mode] [rows] [dim] [k] [metric]
(BASELINE configs[0]: profile_small_batch.py 1000 auto 100000 384 10 cosine)"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastpyvectordb_b200 as fpv  # noqa: E402
from fastpyvectordb_b200 import engine_gemm  # noqa: E402


def main():
    q = int(sys.argv[1]) if len(sys.argv) > 1 else 64
    mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    rows = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
    dim = int(sys.argv[4]) if len(sys.argv) > 4 else 768
    k = int(sys.argv[5]) if len(sys.argv) > 5 else 100
    metric = sys.argv[6] if len(sys.argv) > 6 else "l2"
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    db = torch.randn((rows, dim), generator=g, device=dev)
    db /= db.norm(dim=1, keepdim=True)
    qs = torch.randn((q, dim), generator=g, device=dev)
    qs /= qs.norm(dim=1, keepdim=True)
    index = fpv.GpuIndex(db)
    for _ in range(3):
        engine_gemm.search(qs, index, k, metric, mode=mode)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        engine_gemm.search(qs, index, k, metric, mode=mode)
    e1.record()
    torch.cuda.synchronize()
    print(f"Q={q} mode={mode} {rows}x{dim} k={k} {metric}: {e0.elapsed_time(e1) / 10:.4f} ms per search")


if __name__ == "__main__":
    main()
