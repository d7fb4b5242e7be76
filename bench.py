#!/usr/bin/env python
"""Benchmark of the B200 search hot path (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one batch of Q queries searched exactly (top-k) against the resident database.
Default workload = BASELINE.json configs[1]: 1M x 768 fp32 database, Q = 4096 queries, L2, top-100.
For N > 1 the SAME job is split over the ranks (strong scaling), --shard auto|rows|queries:
  rows    - database row-sharded: local fused top-k per rank, NCCL all-gather of the packed (distance, id)
            candidates, k-way merge kernel on every rank (the only option when the database exceeds one GPU);
  queries - database replicated, the query batch split, no data-path collective (auto picks this when the
            database takes < 1/4 of one GPU's HBM; the row-sharded time of the same job is reported beside it
            as `row_sharded`).

Prints ONE JSON line (rank 0).  `value` = queries/s with inputs already in HBM; `e2e` = same through the public
array API with pinned host queries in and host results out every step; `roofline` = algorithmic flops (or bytes)
of the dominant kernel / CUDA-event time against MEASURED_PEAKS.json; `cpu_baseline` = the oracle port of the
reference's NumPy/BLAS path timed on this box's host cores on a bounded sample.
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

ROW_BLOCK = 65536          # rows generated per RNG stream: database content is independent of the shard count


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000)
    ap.add_argument("--dim", type=int, default=768)
    ap.add_argument("--queries", type=int, default=4096)
    ap.add_argument("--k", type=int, default=100)
    ap.add_argument("--metric", default="l2")
    ap.add_argument("--cpu-queries", type=int, default=256, help="bounded query sample for the CPU baseline")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-regimes", action="store_true", help="skip the extra small-batch / quantized regime lines")
    ap.add_argument("--shard", default="auto", choices=["auto", "rows", "queries"],
                    help="N > 1: 'rows' = database row-sharded, local top-k + NCCL all-gather + merge kernel; 'queries' = database "
                         "replicated, the query batch split over the ranks, no collective; 'auto' = queries when the database "
                         "(fp32 rows + bf16 shadow) takes less than a quarter of one GPU's HBM, rows otherwise")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------ data
def gen_rows(lo: int, hi: int, dim: int, device):
    """Rows [lo, hi) of the synthetic database: standard normal, unit-normalised (the generator of
    examples/benchmark_parallel.py:212-217), one RNG stream per ROW_BLOCK so any sharding sees the same rows."""
    import torch
    out = torch.empty((hi - lo, dim), dtype=torch.float32, device=device)
    b0, b1 = lo // ROW_BLOCK, (hi - 1) // ROW_BLOCK if hi > lo else -1
    for b in range(b0, b1 + 1):
        g = torch.Generator(device=device)
        g.manual_seed(42 + b)
        blk = torch.randn((ROW_BLOCK, dim), generator=g, device=device, dtype=torch.float32)
        blk /= blk.norm(dim=1, keepdim=True)
        s, e = max(lo, b * ROW_BLOCK), min(hi, (b + 1) * ROW_BLOCK)
        out[s - lo:e - lo] = blk[s - b * ROW_BLOCK:e - b * ROW_BLOCK]
    return out


def gen_queries(q: int, dim: int) -> np.ndarray:
    x = np.random.default_rng(999).standard_normal((q, dim)).astype(np.float32)   # benchmark_parallel.py:329-330
    return x / np.linalg.norm(x, axis=1, keepdims=True)


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(hbm=d["hbm_gbs"], tensor_burst=d["bf16_tflops"], tensor_sust=d.get("bf16_tflops_sustained", d["bf16_tflops"]),
                    src="measured (MEASURED_PEAKS.json)")
    return dict(hbm=6650.0, tensor_burst=1590.0, tensor_sust=1400.0, src="fallback (B200_PROFILING.md)")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
            except ValueError:
                continue
            for name, val in zip(names, parts[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------ CPU arm
def cpu_threads():
    try:
        from threadpoolctl import threadpool_info
        n = max([i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"] or [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def cpu_reference_qps(db_host: np.ndarray, qs: np.ndarray, k: int, metric: str, reps: int, warm: int):
    """The reference's CPU path (oracle port of search_batch_parallel, parallel_search.py:259-311: one sgemm +
    per-row argpartition/argsort) on a bounded query sample; returns (median QPS, seconds per rep list)."""
    from oracle import oracle as O
    times = []
    for i in range(warm + reps):
        t0 = time.perf_counter()
        O.search_batch_parallel(qs, db_host, k, metric)
        dt = time.perf_counter() - t0
        if i >= warm:
            times.append(dt)
    return len(qs) / statistics.median(times), times


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rows = args.rows
    # host-side generation of the same distribution (the CPU arm has no GPU dependency); bounded row count keeps the
    # run within minutes: the full database when it is 1M rows, never more than 1M.
    n = min(rows, 1_000_000)
    rng = np.random.default_rng(42)
    db = rng.standard_normal((n, args.dim), dtype=np.float32)
    db /= np.linalg.norm(db, axis=1, keepdims=True)
    qs = gen_queries(args.cpu_queries, args.dim)
    qps, times = cpu_reference_qps(db, qs, args.k, args.metric, max(1, args.steps), max(0, args.warmup))
    scale = n / rows                                    # linear extrapolation if the sample has fewer rows
    value = qps * scale
    line = {
        "impl": "reference", "metric": "exact top-k QPS", "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"exact {args.metric} top-{args.k}, {rows}x{args.dim} fp32 DB, query batch {args.queries}" +
                               (" (BASELINE configs[1])" if (rows, args.dim, args.k) == (1_000_000, 768, 100) else ""),
                   "rows_total": rows, "dim": args.dim, "queries_per_step": args.queries, "k": args.k, "metric": args.metric,
                   "sample_queries_per_step": args.cpu_queries, "rows_in_sample": n},
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cpu_threads(), "kind": "port",
                         "sample": f"{args.cpu_queries} queries x {n} rows per step, oracle port of "
                                   "ParallelSearchEngine.search_batch_parallel (NumPy/OpenBLAS sgemm + argpartition)"},
        "e2e": {"value": value, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------ GPU arm
def main():
    args = parse()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist
    import fastpyvectordb_b200 as fpv
    from fastpyvectordb_b200 import _native, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _native.lib()
    P = peaks()

    n_total, dim, Q, k, metric = args.rows, args.dim, args.queries, args.k, args.metric
    # How the job is split over the ranks.  Both splits leave the answer unchanged (tests/test_gpu_sharded.py):
    #   rows    - the reference's chunk -> local top-k -> merge structure with GPUs as chunks (needed when the
    #             database does not fit one GPU: configs[2], configs[4]); per-query costs (threshold warm-up slabs,
    #             exact re-rank, merge) are paid on EVERY rank, so a small database scales poorly;
    #   queries - database replicated, the batch's queries split: no data-path collective, every cost divides by N.
    shard = args.shard
    if world == 1:
        shard = "none"
    elif shard == "auto":
        hbm = torch.cuda.get_device_properties(dev).total_memory
        shard = "queries" if n_total * dim * 6 <= 0.25 * hbm and Q >= 128 * world else "rows"
    per = (n_total + world - 1) // world
    lo, hi = min(rank * per, n_total), min((rank + 1) * per, n_total)
    eng = fpv.ParallelSearchEngine(device=dev)
    q_host = gen_queries(Q, dim)
    if shard == "queries":
        index = fpv.GpuIndex(gen_rows(0, n_total, dim, dev), dev, id_base=0)
        q_per = (Q + world - 1) // world
        q_lo, q_hi = min(rank * q_per, Q), min((rank + 1) * q_per, Q)
    else:
        index = fpv.GpuIndex(gen_rows(lo, hi, dim, dev), dev, id_base=lo)
        q_lo, q_hi = 0, Q
    Q_local = q_hi - q_lo
    q_pin = torch.from_numpy(np.ascontiguousarray(q_host[q_lo:q_hi])).pin_memory()
    q_dev = q_pin.to(dev)
    k_local = min(k, index.n)

    sharded = None
    if shard == "rows":
        from fastpyvectordb_b200.sharded import ShardedSearchEngine
        sharded = ShardedSearchEngine(index, n_total, engine=eng)

    def step_device_full(qd):
        if sharded is not None:       # local fused top-k -> one packed NCCL all-gather -> merge kernel on every rank
            return sharded.search_tensors(qd, k, metric)
        return eng.search_tensors(qd, index, k_local, metric)

    def step_device(qd):
        d, i, _c = step_device_full(qd)
        return d, i

    out_d = torch.empty((Q_local, min(k, n_total)), dtype=torch.float32).pin_memory()
    out_i = torch.empty((Q_local, min(k, n_total)), dtype=torch.int64).pin_memory()

    def step_e2e():
        qd = q_pin.to(dev, non_blocking=True)
        d, i = step_device(qd)
        out_d.copy_(d, non_blocking=True)
        out_i.copy_(i, non_blocking=True)
        torch.cuda.current_stream().synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    # ---- device-resident timing ------------------------------------------------------------------
    for _ in range(max(3, args.warmup)):
        step_device(q_dev)
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = lib.fpv_launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for _ in range(args.steps):
        step_device(q_dev)
    e1.record()
    barrier()
    launches = lib.fpv_launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop() if rank == 0 else None
    ms_step = ms_total / args.steps
    value = Q / (ms_step * 1e-3)

    # ---- end to end (pinned host queries in, host results out, every step) ---------------------------
    for _ in range(2):
        step_e2e()
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_e2e()
    barrier()
    e2e_serial_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    # the serving form of the same public API: SearchPipeline double-buffers, so the H2D of batch i+1 and the D2H of
    # batch i-1 overlap the kernels of batch i.  Every step still copies its queries in from pinned host memory and
    # its (distance, id) result out to host memory, and every result is fetched inside the timed region.
    from fastpyvectordb_b200 import SearchPipeline
    pipe = SearchPipeline(eng, k=k, metric=metric, search_fn=step_device_full)
    for _ in range(2):
        pipe.result(pipe.submit(q_pin))
    barrier()
    t0 = time.perf_counter()
    prev = None
    for _ in range(args.steps):
        t = pipe.submit(q_pin)
        if prev is not None:
            pipe.result(prev)
        prev = t
    pipe.result(prev)
    barrier()
    e2e_s = max_over_ranks(time.perf_counter() - t0) / args.steps
    copies = world if shard == "queries" else 1        # whole-job bytes: every rank copies its own slice
    e2e = {"value": Q / e2e_s, "unit": "queries/s", "h2d_bytes_per_step": int(q_pin.numel() * 4) * copies,
           "d2h_bytes_per_step": int(out_d.numel() * 4 + out_i.numel() * 8) * copies,
           "serial_value": Q / e2e_serial_s,
           "note": "database resident in HBM (uploaded once at index build); queries H2D + top-k D2H inside the timed region, "
                   "every step; value = SearchPipeline (copies of neighbouring batches overlap the kernels), serial_value = one "
                   "synchronous copy-in / search / copy-out per step"}

    # ---- roofline of the dominant kernel ---------------------------------------------------------
    n_local = index.n
    if Q_local >= eng.GEMM_MIN_BATCH:
        # dominant kernel = gemm_filter_kernel, launched once per slab.  Its launches are timed live with CUDA events
        # recorded by the library on the launching stream (fpv_gemm_profile), over args.steps more steps.
        flops = 2.0 * Q_local * n_local * dim               # per rank (the roofline is per GPU)
        peak = P["tensor_sust"] if ms_total > 1000 else P["tensor_burst"]
        lib.fpv_gemm_profile(1)
        k_ms, k_launches = 0.0, 0
        fm, fl = ctypes.c_float(0), ctypes.c_int(0)
        for _ in range(args.steps):
            step_device(q_dev)
            lib.fpv_gemm_profile_read(ctypes.byref(fm), ctypes.byref(fl))
            k_ms += fm.value
            k_launches += fl.value
        lib.fpv_gemm_profile(0)
        k_ms = max_over_ranks(k_ms) / args.steps
        if k_launches > 0 and k_ms > 0:
            ach = flops / (k_ms * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "kernel": "gemm_filter_kernel", "launches_per_step": k_launches // args.steps, "kernel_ms_per_step": k_ms,
                    "whole_step_frac": flops / (ms_step * 1e-3) / 1e12 / peak,
                    "note": "achieved = 2*Q*N_local*D flops of one step / summed CUDA-event duration of the step's "
                            "gemm_filter_kernel launches (one per slab); whole_step_frac divides by the full step time "
                            "(filter + threshold tightening + exact re-rank) instead"}
        else:
            ach = flops / (ms_step * 1e-3) / 1e12
            roof = {"bound": "tensor", "achieved": ach, "peak": peak, "unit": "TFLOP/s", "frac": ach / peak,
                    "note": "whole step time (the filter kernel was not used for this shape)"}
    else:
        nbytes = float(n_local) * dim * 4
        ach = nbytes / (ms_step * 1e-3) / 1e9
        roof = {"bound": "hbm", "achieved": ach, "peak": P["hbm"], "unit": "GB/s", "frac": ach / P["hbm"],
                "note": "achieved = N_local*D*4 bytes / CUDA-event time of the whole step (scan + finalize)"}
    roof["traffic"] = None
    roof["peak_source"] = P["src"]
    tpath = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tpath):
        try:
            roof["traffic"] = json.load(open(tpath)).get(f"q{Q}_n{n_total}_d{dim}")
        except Exception:
            pass

    line = {
        "metric": "exact top-k QPS", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"exact {metric} top-{k}, {n_total}x{dim} fp32 DB, query batch {Q}" +
                               (" (BASELINE configs[1])" if (n_total, dim, k) == (1_000_000, 768, 100) else
                                " (BASELINE configs[2])" if (n_total, dim, k, metric) == (50_000_000, 768, 10, "cosine") else ""),
                   "rows_total": n_total, "rows_per_gpu": n_local, "dim": dim, "queries_per_step": Q, "k": k, "metric": metric,
                   "queries_per_gpu": Q_local,
                   "sharding": {"none": "none", "rows": f"rows/{world} + NCCL all-gather + merge kernel",
                                "queries": f"database replicated, queries/{world}, no collective"}[shard],
                   "l2_policy": "database (%.2f GB per GPU) is larger than the 126 MB L2" % (n_local * dim * 4 / 1e9)},
        "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }

    if Q_local >= eng.GEMM_MIN_BATCH:
        from fastpyvectordb_b200 import engine_gemm
        if engine_gemm.available(index, Q_local, k_local):
            line["config"]["tensor_core_pass"] = engine_gemm._effective_mode(None, index, k_local, Q_local)
            line["config"]["exact_fallback_fraction"] = engine_gemm.last_fallback_fraction(index, Q_local, k_local)

    # ---- the row-sharded split of the same job, measured beside the query split (N > 1, database small enough) ----
    if shard == "queries" and not args.no_regimes:
        from fastpyvectordb_b200.sharded import ShardedSearchEngine
        sub = fpv.GpuIndex(index.rows[lo:hi].clone(), dev, id_base=lo)
        rs = ShardedSearchEngine(sub, n_total, engine=eng)
        q_all = torch.from_numpy(q_host).to(dev)
        for _ in range(3):
            rs.search_tensors(q_all, k, metric)
        barrier()
        r0, r1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        r0.record()
        for _ in range(args.steps):
            rs.search_tensors(q_all, k, metric)
        r1.record()
        barrier()
        rs_ms = max_over_ranks(r0.elapsed_time(r1)) / args.steps
        line["row_sharded"] = {"value": Q / (rs_ms * 1e-3), "unit": "queries/s", "ms_per_step": rs_ms,
                               "sharding": f"rows/{world} + NCCL all-gather + merge kernel", "rows_per_gpu": hi - lo}

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        db_host = index.rows.cpu().numpy()
        cq = min(args.cpu_queries, Q)
        qps, times = cpu_reference_qps(db_host, q_host[:cq], k, metric, reps=3, warm=1)
        line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cpu_threads(), "kind": "port",
                                "sample": f"{cq} of the {Q} queries x all {n_local} rows, median of 3, oracle port of "
                                          "search_batch_parallel (NumPy/OpenBLAS sgemm + per-row argpartition)",
                                "host_cpus": os.cpu_count()}
        del db_host

    # ---- other regimes of the north-star (short, N = 1 only; reported, not the headline) ----------------
    if rank == 0 and world == 1 and not args.no_regimes:
        try:
            from bench_regimes import run_regimes
            line["regimes"] = run_regimes(eng, index, q_host, P, k, metric)
        except Exception as exc:  # never lose the headline line
            line["regimes"] = {"error": repr(exc)}

    if rank == 0:
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
