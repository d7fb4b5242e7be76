#!/usr/bin/env python
"""Launch EVERY kernel of libfpv_b200.so once at a representative size, for one `ncu --set full` capture
(BASELINE north_star: "every kernel has committed ncu counters").

    ncu --set full --clock-control none --import-source on -k regex:fpv -o all python tools/profile_all.py
    python tools/ncu_summary.py all.ncu-rep > profiles/r02_all_kernels_ncu_summary.txt

Shapes are the BASELINE shapes scaled to fit one short run (--scale); the big scans have their own captures."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastpyvectordb_b200 as fpv  # noqa: E402
from fastpyvectordb_b200 import _native, engine_gemm as eg, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=0.25)
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)
    s = a.scale
    # ---- float path: row norms, bf16 shadow, prep / filter (sample + slabs) / tighten / finish, scan, distances, rerank
    n, d = int(1_000_000 * s), 768
    rows = torch.randn((n, d), device=dev)
    rows /= rows.norm(dim=1, keepdim=True)
    index = fpv.GpuIndex(rows, dev)                                  # row_sqnorm_kernel
    eng = fpv.ParallelSearchEngine(device=dev)
    q = torch.randn((1024, d), device=dev)
    eng.search_tensors(q, index, 100, "l2")                          # to_bf16, gemm_prep, gemm_filter (pair), tighten_warp, gemm_finish2
    eng.search_tensors(q[:64], index, 100, "cosine")                 # one-CTA filter kernel
    mask = torch.rand(n, device=dev) < 0.25
    eng.search_tensors(q[:256], index, 10, "ip", filter_mask=mask)
    ops.scan_f32_topk(q[:1].contiguous(), index.rows, 100, "l2", None, index.row_sq, 0)      # scan_f32 + finalize
    ops.distances_f32(q[:2].contiguous(), index.rows[:100000].contiguous(), "cosine", index.row_sq[:100000].contiguous())
    ops.distances_f32(q[:2].contiguous(), index.rows[:100000].contiguous(), "l2_diff")       # l2_diff_rows_kernel
    cand = torch.randint(0, n, (256, 100), device=dev)
    ops.rerank_f32(q[:256].contiguous(), index.rows, cand, 10, "cosine", index.row_sq, 0)    # rerank_kernel
    # ---- sharded phases + exchange kernels (one rank's share, 8 emulated shards)
    mode = eg._effective_mode(None, index, 100, 1024)
    eg.sharded_bounds(index, mode)
    approx = eg.filter_sharded(q, index, 100, "l2", mode)
    dd, ii, cc = eg.finish_sharded(q, index, 100, "l2", mode, approx.unsqueeze(0).expand(8, -1, -1).contiguous())
    keys = ops.pack_topk(dd, ii, 100, 0)                             # pack_topk_kernel
    bases = torch.arange(8, dtype=torch.int64, device=dev) * n
    ops.merge_packed(keys.unsqueeze(0).expand(8, -1, -1).contiguous(), bases, 100)          # merge_rank_kernel
    ops.merge_topk(dd.unsqueeze(0).expand(4, -1, -1).contiguous(), ii.unsqueeze(0).expand(4, -1, -1).contiguous(), 100)  # merge_kernel
    del index, rows
    torch.cuda.empty_cache()
    # ---- encoders and quantized scans
    nv = int(2_000_000 * s)
    x = torch.randn((nv, 1024), device=dev) * 0.1
    mn = x.min(0).values.contiguous(); sc = (x.max(0).values - mn).contiguous()
    codes = ops.sq_encode(x, mn, sc)                                 # sq_encode_kernel
    thr = torch.zeros(1024, device=dev)
    bits = ops.bq_encode(x, thr)                                     # bq_encode_kernel
    qc = ops.sq_encode(x[:16].contiguous(), mn, sc)
    term, tmax = ops.sq_row_term(codes, sc)                          # sq_row_term_kernel
    ops.sq_l2_mma(qc, codes, mn, sc, term, tmax, 100)                # sq_mma_prep / sq_mma_kernel / tighten / sq_mma_finish
    rsum, rinv, maxima = ops.sq_row_terms_dc(codes, mn, sc)         # sq_row_terms_dc_kernel
    ops.sq_dc_mma(_native.SQ_DOT, qc, codes, mn, sc, rsum, rinv, maxima, 100)     # sq_mma_prep_dc / sq_mma_kernel<DOT> / sq_mma_finish_dc
    ops.sq_dc_mma(_native.SQ_COSINE, qc, codes, mn, sc, rsum, rinv, maxima, 100)  # sq_mma_kernel<COSINE>
    ops.sq_scan(_native.SQ_L2, qc[:1].contiguous(), codes, mn, sc, 100)      # sq_prep + sq_l2_tma_kernel
    ops.sq_scan(_native.SQ_DOT, qc[:1].contiguous(), codes, mn, sc, 100)     # sq_scan_kernel
    ops.sq_scan(_native.SQ_COSINE, qc[:1].contiguous(), codes[:200000].contiguous(), mn, sc, 100)
    qb = ops.bq_encode(x[:16].contiguous(), thr)
    ops.hamming(qb[:1].contiguous(), bits, 100, 1024)                # hamming_fast_kernel<8,1>
    ops.hamming(qb, bits, 100, 1024)                                 # hamming_fast_kernel<8,4>
    del x, codes, bits
    torch.cuda.empty_cache()
    npq = max(int(8_000_000 * s), 1_200_000)                        # >= 2^20 rows: the bound + filter form of the PQ scan
    xv = torch.randn((min(npq, 500_000), 768), device=dev)
    cb = (torch.randn((48, 256, 16), device=dev) / np.sqrt(768)).contiguous()
    ops.pq_encode(xv, cb)                                            # pq_encode_kernel
    lut = ops.pq_build_lut(cb, torch.randn((1, 768), device=dev))   # pq_lut_kernel
    pcodes = torch.randint(0, 256, (npq, 48), dtype=torch.uint8, device=dev)
    packed = ops.pq_pack(pcodes)                                     # pq_pack_kernel
    words = ops.pack_mask(torch.rand(npq, device=dev) < 0.25)
    ops.pq_adc_packed(lut, packed, 100, words)                       # rot table + pq_sample_min + pq_tau + pq_adc_filter + pq_filter_finish
    lut4 = ops.pq_build_lut(cb, torch.randn((4, 768), device=dev))
    ops.pq_adc_packed(lut4, packed, 100, words)                      # pq_quad_table + pq_sample_min_quad + pq_adc_quad + pq_quad_rescore
    ops.pq_adc(lut, pcodes, 100, words)                              # pq_adc_kernel
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
