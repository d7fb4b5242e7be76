"""Extra bench lines for the other regimes BASELINE.json's north_star names (small-batch fp32 scan, uint8 scalar,
binary/Hamming and PQ-ADC scans).  Called by bench.py on rank 0 at N=1; each regime is timed with CUDA events
over inputs that are larger than L2, after a warm-up on its own work (see _time), and reported against the measured HBM peak with
ALGORITHMIC bytes (codes or rows read once per scan).  Random codes are used for throughput;
parity is the job of tests/."""
from __future__ import annotations

import time

import numpy as np
import torch

from fastpyvectordb_b200 import _native, ops


def _time(fn, iters=10, warm=3, min_warm_s=0.25):
    # The regimes run back to back in one process.  A tensor-bound regime leaves the GPU at its power cap (SM clock
    # ~1.3 GHz) for a moment, which slows an instruction-heavy HBM-bound scan measured right after it (Hamming:
    # 0.46 -> 0.53 ms), while idling between regimes drops the clocks the other way (fp32 scan 0.52 -> 0.71 ms).
    # So every HBM-bound regime warms up on its own work for a quarter of a second before it is timed; the tensor-bound
    # lines keep the headline's short warm-up (min_warm_s=0: they are compared with the burst tensor peak like it).
    t0 = time.perf_counter()
    n_warm = 0
    while n_warm < warm or time.perf_counter() - t0 < min_warm_s:
        fn()
        n_warm += 1
        if n_warm % 8 == 0:
            torch.cuda.synchronize()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def _hbm(name, nbytes, ms, q, P, extra=None):
    ach = nbytes / (ms * 1e-3) / 1e9
    out = {"ms_per_scan": ms, "queries_per_scan": q, "qps": q / (ms * 1e-3), "algorithmic_bytes": nbytes,
           "roofline": {"bound": "hbm", "achieved": ach, "peak": P["hbm"], "unit": "GB/s", "frac": ach / P["hbm"]}}
    if extra:
        out.update(extra)
    return name, out


def run_regimes(eng, index, q_host, P, k, metric):
    dev = index.device
    res = {}
    n, d = index.n, index.d
    # ---- fp32 small batches (HBM-bound: every row read once per batch) --------------------------------
    # Q = 1 runs the exact fp32 scan kernel; Q >= GEMM_MIN_BATCH the tensor-core filter + exact fp32 re-rank, which
    # for these batch sizes is bound by reading the rows once, not by the tensor pipe.  The filter reads the bf16
    # shadow copy when the index has one (half the bytes; `pass` says which), so its algorithmic bytes are N*D*2.
    from fastpyvectordb_b200 import engine_gemm
    for qn in (1, 8, 64):
        if q_host.shape[0] < qn:
            continue
        qd = torch.from_numpy(q_host[:qn]).to(dev)
        if qn < eng.GEMM_MIN_BATCH:
            # the exact fp32 scan kernel itself (what an index without a shadow copy, or a filtered search, runs)
            ms = _time(lambda: ops.scan_f32_topk(qd, index.rows, k, metric, None, index.row_sq, index.id_base))
            name, r = _hbm(f"f32_scan_q{qn}_{n}x{d}_{metric}_top{k}", float(n) * d * 4, ms, qn, P)
            res[name] = r
            if not (engine_gemm.has_shadow(index, k) and engine_gemm.available(index, qn, k)):
                continue
        ms = _time(lambda: eng.search_tensors(qd, index, k, metric))
        mode = engine_gemm._effective_mode(None, index, k, qn)
        # the numerator is the ALGORITHMIC N*D*4 bytes of the fp32 rows, whatever is actually read; the
        # bf16 filter reads a half-size shadow copy, so `frac` can exceed 1 -- `bytes_read` / `frac_of_bytes_read` say
        # how close the kernel is to the HBM roofline on the bytes it really moves.
        name, r = _hbm(f"f32_tc_q{qn}_{n}x{d}_{metric}_top{k}", float(n) * d * 4, ms, qn, P, {"pass": mode})
        read = float(n) * d * (2 if mode == "bf16" else 4)
        r["roofline"]["bytes_read"] = read
        r["roofline"]["frac_of_bytes_read"] = read / (ms * 1e-3) / 1e9 / P["hbm"]
        res[name] = r

    # ---- the headline batch again with a 25 % row filter (filter_mask / Collection.query(where=...)) -------------
    if q_host.shape[0] >= 256 and engine_gemm.available(index, q_host.shape[0], k):
        qd = torch.from_numpy(q_host).to(dev)
        keep = torch.rand(n, device=dev) < 0.25
        ms = _time(lambda: eng.search_tensors(qd, index, k, metric, filter_mask=keep), iters=5, min_warm_s=0.0)
        qn = q_host.shape[0]
        flops = 2.0 * qn * n * d
        res[f"f32_tc_q{qn}_{n}x{d}_{metric}_top{k}_mask25"] = {
            "ms_per_scan": ms, "queries_per_scan": qn, "qps": qn / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": P["tensor_burst"],
                         "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / P["tensor_burst"]},
            "note": "whole step incl. packing the boolean mask; the filter is one word per 32-row chunk in the epilogue"}

    # ---- the other metric configs[1] names (inner product), full batch ------------------------------------------
    if q_host.shape[0] >= 256 and engine_gemm.available(index, q_host.shape[0], k):
        other = "ip" if metric != "ip" else "l2"
        qd = torch.from_numpy(q_host).to(dev)
        ms = _time(lambda: eng.search_tensors(qd, index, k, other), iters=5, min_warm_s=0.0)
        qn = q_host.shape[0]
        flops = 2.0 * qn * n * d
        res[f"f32_tc_q{qn}_{n}x{d}_{other}_top{k}"] = {
            "ms_per_scan": ms, "queries_per_scan": qn, "qps": qn / (ms * 1e-3),
            "roofline": {"bound": "tensor", "achieved": flops / (ms * 1e-3) / 1e12, "peak": P["tensor_burst"],
                         "unit": "TFLOP/s", "frac": flops / (ms * 1e-3) / 1e12 / P["tensor_burst"]},
            "note": "whole step (filter + tighten + exact re-rank)"}

    # ---- binary / Hamming: 20M x 1024 bits (BASELINE configs[3]) ---------------------------------------
    nb = 20_000_000
    codes = torch.randint(0, 256, (nb, 128), dtype=torch.uint8, device=dev)
    qb = torch.randint(0, 256, (1, 128), dtype=torch.uint8, device=dev)
    ms = _time(lambda: ops.hamming(qb, codes, 100, 1024))
    name, r = _hbm("hamming_q1_20Mx1024b_top100", float(nb) * 128, ms, 1, P)
    res[name] = r
    # batches: one pass over the packed codes on the int8 tensor cores (csrc/fpv_hamming_mma.cu) serves 31 queries
    for qn in (16, 31):
        qbn = torch.randint(0, 256, (qn, 128), dtype=torch.uint8, device=dev)
        ms = _time(lambda: ops.hamming(qbn, codes, 100, 1024), iters=5)
        name, r = _hbm(f"hamming_q{qn}_20Mx1024b_top100", float(nb) * 128, ms, qn, P,
                       {"kernel": "ham_mma_kernel (bits expanded to int8 in tensor memory, tcgen05.mma kind::i8 with A from TMEM)"
                                  if ops.hamming_mma_supported(qn, nb, 128, 100) else "hamming_fast_kernel (CUDA cores)"})
        res[name] = r
    ops.HAMMING_TENSOR_CORES = False
    try:
        ms = _time(lambda: ops.hamming(qbn[:16].contiguous(), codes, 100, 1024), iters=3)
    finally:
        ops.HAMMING_TENSOR_CORES = True
    name, r = _hbm("hamming_simt_q16_20Mx1024b_top100", float(nb) * 128, ms, 16, P,
                   {"kernel": "hamming_fast_kernel (CUDA cores): 4 queries share each pass; POPC-issue bound beyond ~3 queries per pass"})
    res[name] = r
    del codes

    # ---- uint8 scalar quantizer: 20M x 1024 codes (BASELINE configs[3]) ----------------------------------
    # The L2 scan runs on the int8 tensor cores (csrc/fpv_sq_mma.cu): one pass over the codes serves 16 queries.
    # `sq_u8_l2_simt_q1` is the CUDA-core scan it replaced (instruction bound: 3.25 instructions per code byte).
    ns = 20_000_000
    try:
        codes = torch.randint(0, 256, (ns, 1024), dtype=torch.uint8, device=dev)
        mn = torch.full((1024,), -0.1, device=dev)
        sc = torch.full((1024,), 0.2, device=dev)
        t0 = time.perf_counter()
        term, tmax = ops.sq_row_term(codes, sc)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        for qn in (1, 16):
            qc = torch.randint(0, 256, (qn, 1024), dtype=torch.uint8, device=dev)
            ms = _time(lambda: ops.sq_l2_mma(qc, codes, mn, sc, term, tmax, 100), iters=5)
            name, r = _hbm(f"sq_u8_l2_q{qn}_20Mx1024_top100", float(ns) * 1024, ms, qn, P,
                           {"kernel": "sq_mma_kernel (tcgen05.mma kind::i8, three exact limb dots per row and query) + certified "
                                      "exact re-score", "exact_fallback_queries": int(ops.sq_mma_last_flags(qn, ns, 1024, 100, dev).sum().item()),
                            "row_term_build_ms_once_per_index": build_ms})
            res[name] = r
        qc = torch.randint(0, 256, (1, 1024), dtype=torch.uint8, device=dev)
        ms = _time(lambda: ops.sq_scan(_native.SQ_L2, qc, codes, mn, sc, 100), iters=5)
        name, r = _hbm("sq_u8_l2_simt_q1_20Mx1024_top100", float(ns) * 1024, ms, 1, P, {"kernel": "sq_l2_tma_kernel (CUDA cores)"})
        res[name] = r
        del term
        # distances_dot / distances_cosine on the same tensor-core machinery (signed weights shifted, per-row sums and
        # inverse norms from the index build); `..._simt_q1` = the CUDA-core scan of the same metric
        t0 = time.perf_counter()
        rsum, rinv, maxima = ops.sq_row_terms_dc(codes, mn, sc)
        torch.cuda.synchronize()
        build_ms = (time.perf_counter() - t0) * 1e3
        for metric, kind in (("dot", _native.SQ_DOT), ("cosine", _native.SQ_COSINE)):
            for qn in (1, 16):
                qc = torch.randint(0, 256, (qn, 1024), dtype=torch.uint8, device=dev)
                ms = _time(lambda: ops.sq_dc_mma(kind, qc, codes, mn, sc, rsum, rinv, maxima, 100), iters=5)
                name, r = _hbm(f"sq_u8_{metric}_q{qn}_20Mx1024_top100", float(ns) * 1024, ms, qn, P,
                               {"kernel": "sq_mma_kernel<kind> (tcgen05.mma kind::i8) + certified exact re-score",
                                "exact_fallback_queries": int(ops.sq_mma_last_flags(qn, ns, 1024, 100, dev).sum().item()),
                                "row_terms_build_ms_once_per_index": build_ms})
                res[name] = r
            qc = torch.randint(0, 256, (1, 1024), dtype=torch.uint8, device=dev)
            ms = _time(lambda: ops.sq_scan(kind, qc, codes, mn, sc, 100), iters=3)
            name, r = _hbm(f"sq_u8_{metric}_simt_q1_20Mx1024_top100", float(ns) * 1024, ms, 1, P, {"kernel": "sq_scan_kernel (CUDA cores)"})
            res[name] = r
        del codes, rsum, rinv
    except torch.cuda.OutOfMemoryError as exc:   # bounded: never take the box down for an extra line
        res["sq_u8_l2_q1_20Mx1024_top100"] = {"skipped": repr(exc)}

    # ---- PQ ADC: 25M codes = one GPU's share of 200M x (48 x 8 bit), 25% bitmask in-kernel (configs[4]) --------
    npq = 25_000_000
    codes = torch.randint(0, 256, (npq, 48), dtype=torch.uint8, device=dev)
    cb = (torch.randn((48, 256, 16), device=dev) / np.sqrt(768)).contiguous()
    qd = torch.from_numpy(q_host[:1]).to(dev)
    lut = ops.pq_build_lut(cb, qd)
    mask = ops.pack_mask(torch.rand(npq, device=dev) < 0.25)
    packed = ops.pq_pack(codes)          # lane-rotated copy, built once per index (what ProductQuantizer.search uses)
    for tag, m in (("mask25", mask), ("nomask", None)):
        ms = _time(lambda: ops.pq_adc_packed(lut, packed, 100, m))
        name, r = _hbm(f"pq_adc_q1_25Mx48B_{tag}_top100", float(npq) * 48 + (npq / 8 if m is not None else 0), ms, 1, P)
        res[name] = r
    # query batches: one pass per four queries over fixed-point u16 tables + exact re-score of the survivors
    lut4 = ops.pq_build_lut(cb, torch.from_numpy(q_host[:4]).to(dev))
    for tag, m in (("mask25", mask), ("nomask", None)):
        ms = _time(lambda: ops.pq_adc_packed(lut4, packed, 100, m))
        name, r = _hbm(f"pq_adc_q4_25Mx48B_{tag}_top100", float(npq) * 48 + (npq / 8 if m is not None else 0), ms, 4, P,
                       {"kernel": "pq_adc_quad_kernel (four queries per 64-bit lookup) + pq_quad_rescore_kernel"})
        res[name] = r
    del codes
    torch.cuda.empty_cache()
    return res
