#!/usr/bin/env python
"""Benchmark of the B200 search hot path (contract: see the task brief / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

A *step* is one batch of Q queries searched exactly (top-k) against the resident database.
Default workload = BASELINE.json configs[1]: 1M x 768 fp32 database, Q = 4096 queries, L2, top-100.
For N > 1 the SAME job is split over the ranks (strong scaling).  The headline split is the north-star's:
  rows    - database row-sharded: sampling slab per rank, all-gather of the k best group values (first threshold of the
            whole job), tensor-core filter, all-gather of the k best approximate values, phase-2 exact re-rank under
            the global limit, all-gather of the packed (distance, id) lists, merge kernel on every rank
            (fastpyvectordb_b200/sharded.py; the exchanges run over NVLink peer memory, NCCL as the fallback).
Beside it the line carries `replicated` (database replicated, the query batch split, no collective: --shard queries
makes it the headline instead) and `rows_8m` (the row-sharded search of an 8M x 768 database, the size where the
shards are large enough for the per-query costs to amortise; its N=1 value is the single-GPU anchor).

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
    ap.add_argument("--shard", default="rows", choices=["rows", "queries"],
                    help="N > 1 headline split: 'rows' (default, the north-star's) = database row-sharded, two-phase search + "
                         "NCCL all-gathers + merge kernel; 'queries' = database replicated, the query batch split, no collective")
    ap.add_argument("--no-extras", action="store_true", help="skip the `replicated` and `rows_8m` blocks")
    ap.add_argument("--big-rows", type=int, default=8_000_000, help="rows of the `rows_8m` block")
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
    """BLAS threads in use right now."""
    try:
        from threadpoolctl import threadpool_info
        n = max([i.get("num_threads", 1) for i in threadpool_info() if i.get("user_api") == "blas"] or [1])
        return int(n)
    except Exception:
        return os.cpu_count() or 1


def blas_all_cores():
    """Context manager: give NumPy's BLAS every host core (capped at the wheel's MAX_THREADS = 64), whatever the
    launcher exported -- torchrun sets OMP_NUM_THREADS=1, which made the round-1 reference arm 2x slower at N > 1
    than at N = 1.  The thread count actually used is what cpu_threads() then reports."""
    try:
        from threadpoolctl import threadpool_limits
        return threadpool_limits(limits=min(os.cpu_count() or 1, 64), user_api="blas")
    except Exception:
        import contextlib
        return contextlib.nullcontext()


def base_config(args):
    """The workload description shared verbatim by both arms (so the driver can see they ran the same job)."""
    n, d, k, m = args.rows, args.dim, args.k, args.metric
    tag = (" (BASELINE configs[1])" if (n, d, k) == (1_000_000, 768, 100) else
           " (BASELINE configs[2])" if (n, d, k, m) == (50_000_000, 768, 10, "cosine") else
           " (BASELINE configs[0])" if (n, d, k, m, args.queries) == (100_000, 384, 10, "cosine", 1000) else "")
    return {"workload": f"exact {m} top-{k}, {n}x{d} fp32 DB, query batch {args.queries}{tag}",
            "rows_total": n, "dim": d, "queries_per_step": args.queries, "k": k, "metric": m,
            "sampled": True,
            "sampled_note": "GPU arm: every step searches all queries; its cpu_baseline and the reference arm time a bounded "
                            f"sample of {min(args.cpu_queries, args.queries)} queries per step against all rows (QPS of that "
                            "path is flat in the batch size: its time is the per-query argpartition over N distances)",
            "l2_policy": "database (%.2f GB) is larger than the 126 MB L2; nothing is flushed between steps" % (n * d * 4 / 1e9)}


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
    cq = min(args.cpu_queries, args.queries)
    qs = gen_queries(cq, args.dim)
    with blas_all_cores():
        cores = cpu_threads()
        qps, times = cpu_reference_qps(db, qs, args.k, args.metric, max(1, args.steps), max(0, args.warmup))
    scale = n / rows                                    # linear extrapolation if the sample has fewer rows
    value = qps * scale
    line = {
        "impl": "reference", "metric": "exact top-k QPS", "value": value, "unit": "queries/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.median(times),
        "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(args),
        "cpu_baseline": {"value": value, "unit": "queries/s", "cores": cores, "kind": "port", "host_cpus": os.cpu_count(),
                         "sample": f"{cq} queries x {n} rows per step, oracle port of "
                                   "ParallelSearchEngine.search_batch_parallel (NumPy/OpenBLAS sgemm + argpartition), BLAS "
                                   "thread count set explicitly (independent of the launcher's OMP_NUM_THREADS)"},
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
    from fastpyvectordb_b200 import _native, engine_gemm
    from fastpyvectordb_b200.sharded import ShardedSearchEngine, shard_bounds

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
    shard = "none" if world == 1 else args.shard
    eng = fpv.ParallelSearchEngine(device=dev)
    q_host = gen_queries(Q, dim)

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

    def time_steps(fn, steps, warm):
        """device-timed: barrier + synchronize on both sides, CUDA events, max over ranks -> ms per step"""
        for _ in range(warm):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        return max_over_ranks(e0.elapsed_time(e1)) / steps

    # ---- the two splits of the job ----------------------------------------------------------------------------
    #   rows    - the reference's chunk -> local top-k -> merge structure with GPUs as chunks (the north-star split; the
    #             only one possible when the database exceeds one GPU: configs[2], configs[4]);
    #   queries - database replicated, the batch's queries split: no data-path collective.
    lo, hi = shard_bounds(n_total, world, rank)
    q_per = (Q + world - 1) // world
    q_lo, q_hi = min(rank * q_per, Q), min((rank + 1) * q_per, Q)

    def make_rows_split():
        index = fpv.GpuIndex(gen_rows(lo, hi, dim, dev), dev, id_base=lo)
        sharded = ShardedSearchEngine(index, n_total, engine=eng) if world > 1 else None
        return index, sharded

    def make_query_split():
        return fpv.GpuIndex(gen_rows(0, n_total, dim, dev), dev, id_base=0)

    sharded = None
    if shard == "queries":
        index = make_query_split()
        my_q = slice(q_lo, q_hi)
    else:
        index, sharded = make_rows_split()
        my_q = slice(0, Q)
    Q_local = my_q.stop - my_q.start
    q_pin = torch.from_numpy(np.ascontiguousarray(q_host[my_q])).pin_memory()
    q_dev = q_pin.to(dev)
    k_local = min(k, index.n)

    def step_device_full(qd):
        if sharded is not None:
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
    # whole-job bytes = the sum over the ranks: with the query split every rank copies its own slice of the batch, with
    # the row split every rank copies all the queries in and the merged answer out
    h2d = int(q_pin.numel() * 4) * world
    d2h = int(out_d.numel() * 4 + out_i.numel() * 8) * world
    # two forms of the same end-to-end measurement are taken (both copy every step's queries in and results out inside
    # the timed region); the headline is the faster one and `mode` says which -- on a box whose host is busy the
    # double-buffered form can lose to the plain one (seen once in this pool: 343K vs 827K on the same GPU)
    e2e_best = min(e2e_s, e2e_serial_s)
    e2e = {"value": Q / e2e_best, "unit": "queries/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
           "mode": "pipelined (SearchPipeline)" if e2e_s <= e2e_serial_s else "serial (copy in, search, copy out, wait)",
           "pipelined_value": Q / e2e_s, "serial_value": Q / e2e_serial_s,
           "note": "database resident in HBM (uploaded once at index build); queries H2D + top-k D2H inside the timed region, "
                   "every step, on every rank; pipelined_value = SearchPipeline (copies of neighbouring batches overlap the kernels), "
                   "serial_value = one synchronous copy-in / search / copy-out per step; value = the better of the two"}

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
                    "note": "per GPU: achieved = 2*Q*N_local*D flops of one step / summed CUDA-event duration of the step's "
                            "gemm_filter_kernel launches (one per slab); whole_step_frac divides by the full step time "
                            "(filter + threshold tightening + exchange + exact re-rank + merge) instead"}
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
    if os.path.exists(tpath) and world == 1:
        try:
            roof["traffic"] = json.load(open(tpath)).get(f"q{Q}_n{n_total}_d{dim}")
        except Exception:
            pass

    cfg = base_config(args)
    cfg.update({"rows_per_gpu": n_local, "queries_per_gpu": Q_local,
                "sharding": {"none": "none",
                             "rows": f"rows/{world}: two-phase search, 3 small all-gathers (k best sample group values -> first "
                                     "threshold of the whole job; k best approximate values; packed exact lists) + merge kernel",
                             "queries": f"database replicated, queries/{world}, no collective"}[shard]})
    line = {
        "metric": "exact top-k QPS", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
        "warmup": max(3, args.warmup), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
        "roofline": roof, "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks,
    }
    if sharded is not None:
        px = sharded.topk.peer
        line["config"]["pooled_sample"] = bool(sharded._sample_exchange_ok())
        line["config"]["exchange"] = ("NVLink peer-memory stores + in-kernel arrival flags (csrc/fpv_peer.cu)" if px is not None and px.ok
                                      else "NCCL all_gather_into_tensor")
    if Q_local >= eng.GEMM_MIN_BATCH and engine_gemm.available(index, Q_local, k_local):
        line["config"]["tensor_core_pass"] = engine_gemm._effective_mode(None, index, k_local, Q_local)
        line["config"]["exact_fallback_fraction"] = engine_gemm.last_fallback_fraction(index, Q_local, k_local)

    # ---- parity of the N-rank answer against a single-rank recomputation (outside every timed region) -------------
    if world > 1:
        sample = min(64, Q)
        d_n, i_n = step_device(q_dev)
        rows_here = index.rows if shard == "queries" else None
        if rows_here is None:                       # row split: rebuild the whole database on this GPU if it fits
            free, _tot = torch.cuda.mem_get_info(dev)
            if n_total * dim * 6 < 0.6 * free:
                rows_here = gen_rows(0, n_total, dim, dev)
        ok = None
        if rows_here is not None:
            whole = fpv.GpuIndex(rows_here, dev, id_base=0)
            qs_s = torch.from_numpy(np.ascontiguousarray(q_host[my_q][:sample])).to(dev)
            d_1, i_1, _ = eng.search_tensors(qs_s, whole, min(k, n_total), metric)
            ok = bool(torch.equal(i_1, i_n[:sample]) and torch.equal(d_1, d_n[:sample]))
            t = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MIN)
            ok = bool(t.item())
            del whole
            if shard != "queries":
                del rows_here
        line["parity_checked"] = ok
        line["parity_note"] = (f"first {sample} queries of every rank's answer compared bit for bit (ids and fp32 distances) with "
                               "a single-GPU search of the whole database on the same GPU, outside the timed region"
                               if ok is not None else "database does not fit one GPU: not checked here (tests/test_gpu_multirank.py)")

    # ---- the other split of the same job, and the 8M-row row-sharded job (reported beside the headline) ------------
    if not args.no_extras and Q >= 128 * world:
        index = sharded = pipe = None                # release the headline's database before building the others
        torch.cuda.empty_cache()

        def other_split():
            if shard == "rows":
                idx2 = make_query_split()
                qd2 = torch.from_numpy(np.ascontiguousarray(q_host[q_lo:q_hi])).to(dev)
                ms2 = time_steps(lambda: eng.search_tensors(qd2, idx2, min(k, idx2.n), metric), args.steps, 3)
                return "replicated", {"value": Q / (ms2 * 1e-3), "unit": "queries/s", "ms_per_step": ms2,
                                      "sharding": f"database replicated, queries/{world}, no collective"}
            idx2, sh2 = make_rows_split()
            qd2 = torch.from_numpy(q_host).to(dev)
            ms2 = time_steps(lambda: sh2.search_tensors(qd2, k, metric), args.steps, 3)
            return "row_sharded", {"value": Q / (ms2 * 1e-3), "unit": "queries/s", "ms_per_step": ms2,
                                   "sharding": f"rows/{world}: two-phase search + 2 all-gathers + merge kernel"}

        def big_rows():
            nb = args.big_rows
            blo, bhi = shard_bounds(nb, world, rank)
            free, _tot = torch.cuda.mem_get_info(dev)
            fits = torch.tensor([1 if (nb > n_total and (bhi - blo) * dim * 6 * 1.15 < free) else 0], dtype=torch.int32, device=dev)
            if world > 1:
                dist.all_reduce(fits, op=dist.ReduceOp.MIN)
            if not fits.item():
                return None
            big = fpv.GpuIndex(gen_rows(blo, bhi, dim, dev), dev, id_base=blo)
            shb = ShardedSearchEngine(big, nb, engine=eng) if world > 1 else None
            qd2 = torch.from_numpy(q_host).to(dev)
            fn = (lambda: shb.search_tensors(qd2, k, metric)) if shb is not None else (lambda: eng.search_tensors(qd2, big, k, metric))
            steps_b = max(2, min(args.steps, 5))
            msb = time_steps(fn, steps_b, 3)
            fl = 2.0 * Q * (bhi - blo) * dim
            return {"value": Q / (msb * 1e-3), "unit": "queries/s", "ms_per_step": msb, "steps": steps_b,
                    "rows_total": nb, "rows_per_gpu": bhi - blo,
                    "whole_step_frac_of_tensor_burst": fl / (msb * 1e-3) / 1e12 / P["tensor_burst"],
                    "sharding": "none" if world == 1 else f"rows/{world}: two-phase search + 2 all-gathers + merge kernel",
                    "note": f"exact {metric} top-{k}, {nb}x{dim} fp32 DB, query batch {Q}: the row-sharded job at a shard size "
                            "where the per-query costs amortise; the N=1 line is its single-GPU anchor"}

        if world > 1:
            name, blk = other_split()
            line[name] = blk
            torch.cuda.empty_cache()
        blk = big_rows()
        if blk is not None:
            line["rows_8m"] = blk
        torch.cuda.empty_cache()

    # ---- CPU baseline (rank 0, N = 1 only) -----------------------------------------------------------
    if rank == 0 and world == 1 and (not args.no_cpu_baseline or not args.no_regimes) and index is None:
        index = fpv.GpuIndex(gen_rows(0, n_total, dim, dev), dev, id_base=0)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        db_host = index.rows.cpu().numpy()
        cq = min(args.cpu_queries, Q)
        with blas_all_cores():
            cores = cpu_threads()
            qps, times = cpu_reference_qps(db_host, q_host[:cq], k, metric, reps=3, warm=1)
        line["cpu_baseline"] = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                                "sample": f"{cq} of the {Q} queries x all {n_total} rows, median of 3, oracle port of "
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
