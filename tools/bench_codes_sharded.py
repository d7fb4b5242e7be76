#!/usr/bin/env python
"""Row-sharded quantized scans (BASELINE configs[3], [4]) on N GPUs: one rank per GPU (torchrun), random codes of the
named shapes generated per shard, ShardedCodeSearch (local fused scan -> NCCL all-gather of 8-byte keys -> merge).

    torchrun --nproc-per-node N tools/bench_codes_sharded.py [--only pq,hamming,sq] [--steps 10]

Prints one JSON object (rank 0): per config ms per step (CUDA events, barrier on both sides, max over ranks), queries/s
and the per-GPU HBM roofline fraction on ALGORITHMIC bytes (the shard's codes read once per pass)."""
import argparse
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from fastpyvectordb_b200 import ops  # noqa: E402
from fastpyvectordb_b200.sharded import ShardedCodeSearch, shard_bounds  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default="pq,hamming,sq")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--pq-rows", type=int, default=200_000_000)
    ap.add_argument("--code-rows", type=int, default=20_000_000)
    a = ap.parse_args()
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps):
        for _ in range(3):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        t = torch.tensor([e0.elapsed_time(e1) / steps], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    out = {"n_gpus": world, "peak_hbm_gbs": hbm}
    g = torch.Generator(device=dev)
    g.manual_seed(1234 + rank)
    want = a.only.split(",")
    if "pq" in want:
        n = a.pq_rows
        lo, hi = shard_bounds(n, world, rank)
        if (hi - lo) * 48 * 2.2 < torch.cuda.mem_get_info(dev)[0]:
            codes = torch.randint(0, 256, (hi - lo, 48), dtype=torch.uint8, device=dev, generator=g)
            sh = ShardedCodeSearch("pq", codes, n)
            del codes
            torch.cuda.empty_cache()
            cb = (torch.randn((48, 256, 16), device=dev, generator=torch.Generator(device=dev).manual_seed(7)) / np.sqrt(768)).contiguous()
            mask = ops.pack_mask(torch.rand(hi - lo, device=dev, generator=g) < 0.25)
            for qn in (1, 4):
                qs = torch.randn((qn, 768), device=dev, generator=torch.Generator(device=dev).manual_seed(999))
                lut = ops.pq_build_lut(cb, qs)
                for tag, m in (("mask25", mask), ("nomask", None)):
                    ms = timed(lambda: sh.search_tensors(lut, 100, m), a.steps)
                    nbytes = (hi - lo) * 48.0 * qn          # the reference API is one query per scan: each query reads the codes
                    out[f"pq_adc_{n}x48B_q{qn}_{tag}_top100"] = {
                        "ms_per_step": ms, "qps": qn / (ms * 1e-3), "rows_per_gpu": hi - lo,
                        "per_gpu_hbm_frac_algorithmic": nbytes / (ms * 1e-3) / 1e9 / hbm}
            del sh, mask
            torch.cuda.empty_cache()
        else:
            out["pq"] = "skipped: shard does not fit"
    if "hamming" in want:
        n = a.code_rows
        lo, hi = shard_bounds(n, world, rank)
        codes = torch.randint(0, 256, (hi - lo, 128), dtype=torch.uint8, device=dev, generator=g)
        sh = ShardedCodeSearch("hamming", codes, n, dims=1024)
        for qn in (1, 16):
            qb = torch.randint(0, 256, (qn, 128), dtype=torch.uint8, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
            ms = timed(lambda: sh.search_tensors(qb, 100), a.steps)
            out[f"hamming_{n}x1024b_q{qn}_top100"] = {"ms_per_step": ms, "qps": qn / (ms * 1e-3), "rows_per_gpu": hi - lo,
                                                     "per_gpu_hbm_frac_algorithmic": (hi - lo) * 128.0 / (ms * 1e-3) / 1e9 / hbm}
        del sh, codes
        torch.cuda.empty_cache()
    if "sq" in want:
        n = a.code_rows
        lo, hi = shard_bounds(n, world, rank)
        codes = torch.randint(0, 256, (hi - lo, 1024), dtype=torch.uint8, device=dev, generator=g)
        mn = torch.full((1024,), -0.1, device=dev)
        sc = torch.full((1024,), 0.2, device=dev)
        sh = ShardedCodeSearch("sq", codes, n, sq_params=(mn, sc))
        for qn in (1, 16):
            qc = torch.randint(0, 256, (qn, 1024), dtype=torch.uint8, device=dev, generator=torch.Generator(device=dev).manual_seed(6))
            ms = timed(lambda: sh.search_tensors(qc, 100), a.steps)
            out[f"sq_u8_l2_{n}x1024_q{qn}_top100"] = {"ms_per_step": ms, "qps": qn / (ms * 1e-3), "rows_per_gpu": hi - lo,
                                                    "per_gpu_hbm_frac_algorithmic": (hi - lo) * 1024.0 / (ms * 1e-3) / 1e9 / hbm}
        del sh, codes
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
