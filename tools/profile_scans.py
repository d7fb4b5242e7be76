#!/usr/bin/env python
"""Launch each scan kernel of the bench regimes a few times (for `ncu --set full -k regex:... `).

    python tools/profile_scans.py [--scale 1.0]

Shapes follow bench_regimes.py (BASELINE configs[1], [3], [4]); --scale shrinks the row counts."""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from fastpyvectordb_b200 import _native, ops  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--scale", type=float, default=1.0)
    ap.add_argument("--reps", type=int, default=2)
    ap.add_argument("--only", default="")
    a = ap.parse_args()
    dev = torch.device("cuda", 0)
    torch.manual_seed(0)

    def want(name):
        return not a.only or name in a.only.split(",")

    if want("f32"):
        n = int(1_000_000 * a.scale)
        db = torch.randn((n, 768), device=dev)
        rsq = ops.row_sqnorm(db)
        for qn in (1, 8):
            q = torch.randn((qn, 768), device=dev)
            for _ in range(a.reps):
                ops.scan_f32_topk(q, db, 100, "l2", None, rsq, 0)
        del db
    if want("hamming"):
        n = int(20_000_000 * a.scale)
        codes = torch.randint(0, 256, (n, 128), dtype=torch.uint8, device=dev)
        qb = torch.randint(0, 256, (1, 128), dtype=torch.uint8, device=dev)
        for _ in range(a.reps):
            ops.hamming(qb, codes, 100, 1024)
        del codes
    if want("sq"):
        n = int(20_000_000 * a.scale)
        codes = torch.randint(0, 256, (n, 1024), dtype=torch.uint8, device=dev)
        qc = torch.randint(0, 256, (1, 1024), dtype=torch.uint8, device=dev)
        mn = torch.full((1024,), -0.1, device=dev)
        sc = torch.full((1024,), 0.2, device=dev)
        for _ in range(a.reps):
            ops.sq_scan(_native.SQ_L2, qc, codes, mn, sc, 100)
        del codes
    if want("pq"):
        n = int(25_000_000 * a.scale)
        codes = torch.randint(0, 256, (n, 48), dtype=torch.uint8, device=dev)
        cb = (torch.randn((48, 256, 16), device=dev) / np.sqrt(768)).contiguous()
        lut = ops.pq_build_lut(cb, torch.randn((1, 768), device=dev))
        mask = ops.pack_mask(torch.rand(n, device=dev) < 0.25)
        packed = ops.pq_pack(codes)
        for _ in range(a.reps):
            ops.pq_adc_packed(lut, packed, 100, mask)
            ops.pq_adc_packed(lut, packed, 100, None)
            ops.pq_adc(lut, codes, 100, None)
    torch.cuda.synchronize()
    print("done")


if __name__ == "__main__":
    main()
