#!/usr/bin/env python
"""One rank's share of the row-sharded search, on one GPU, for `ncu` launch lists (ncu cannot follow several ranks).

    python tools/profile_sharded_local.py [--rows-total 1000000] [--shards 8] [--queries 4096] [--k 100]

Runs exactly what rank 0 of `--shards` ranks runs (phase-1 filter, phase-2 re-rank under the global limit, pack, merge);
the two all-gathers are replaced by copies of this rank's own buffers (statistically the same global k-th value for
i.i.d. rows), so every kernel sees realistic inputs.  Prints CUDA-event times per phase."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastpyvectordb_b200 as fpv  # noqa: E402
from fastpyvectordb_b200 import engine_gemm as eg, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows-total", type=int, default=1_000_000)
    ap.add_argument("--shards", type=int, default=8)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--metric", default="l2")
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--pooled", type=int, default=1, help="1: sample exchange (default path for shards >= 70K rows), 0: local samples")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    n = a.rows_total // a.shards
    g = torch.Generator(device=dev); g.manual_seed(42)
    rows = torch.randn((n, a.dim), generator=g, device=dev)
    rows /= rows.norm(dim=1, keepdim=True)
    index = fpv.GpuIndex(rows, dev)
    q = torch.randn((a.queries, a.dim), generator=g, device=dev)
    q /= q.norm(dim=1, keepdim=True)
    mode = eg._effective_mode(None, index, a.k, a.queries)
    eg.sharded_bounds(index, mode)
    bases = torch.arange(a.shards, dtype=torch.int64, device=dev) * n

    def step(ev=None):
        def mark(i):
            if ev is not None:
                ev[i].record()
        mark(0)
        if a.pooled:     # the shards pool their samples (eight copies of this shard's sample stand in for the gather)
            sample = eg.sample_sharded(q, index, a.k, a.metric, mode)
            pooled = sample.unsqueeze(0).expand(a.shards, -1, -1).contiguous()
            approx = eg.slabs_sharded(q, index, a.k, a.metric, mode, pooled, a.shards)
        else:
            approx = eg.filter_sharded(q, index, a.k, a.metric, mode)
        mark(1)
        gathered = approx.unsqueeze(0).expand(a.shards, -1, -1).contiguous()
        mark(2)
        d, i, c = eg.finish_sharded(q, index, a.k, a.metric, mode, gathered)
        mark(3)
        keys = ops.pack_topk(d, i, a.k, 0)
        wire = keys.unsqueeze(0).expand(a.shards, -1, -1).contiguous()
        mark(4)
        out = ops.merge_packed(wire, bases, a.k)
        mark(5)
        return out

    for _ in range(3):
        step()
    torch.cuda.synchronize()
    names = ["phase1 filter+tighten", "(copy for gather)", "phase2 select+rerank(+flag scan)", "pack (+copy)", "merge"]
    tot = np.zeros(5)
    for _ in range(a.steps):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
        step(ev)
        torch.cuda.synchronize()
        tot += [ev[i].elapsed_time(ev[i + 1]) for i in range(5)]
    for nm, t in zip(names, tot / a.steps):
        print(f"{nm:36s} {t * 1e3:9.1f} us")
    print(f"{'sum':36s} {tot.sum() / a.steps * 1e3:9.1f} us   ({n} rows per shard, {a.shards} shards, Q={a.queries}, k={a.k}, {mode})")


if __name__ == "__main__":
    main()
