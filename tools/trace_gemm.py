#!/usr/bin/env python
"""Experiment: per-slab phase timers of gemm_filter_kernel (needs a library built with
FPV_EXTRA_NVCC_FLAGS=-DFPV_GEMM_TRACE python -m fastpyvectordb_b200.build --force).

Prints, per slab, the per-tile averages (cycles) over all CTAs of:
  epilogue warp 4: aux staging + named barrier | wait for the accumulator | tile processing
  MMA thread: wait for a free accumulator | wait for operands
"""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fastpyvectordb_b200 as fpv  # noqa: E402
from fastpyvectordb_b200 import _native, engine_gemm  # noqa: E402


def main():
    rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    mode = sys.argv[2] if len(sys.argv) > 2 else "bf16"
    q = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(42)
    db = torch.randn((rows, 768), generator=g, device=dev)
    db /= db.norm(dim=1, keepdim=True)
    qs = torch.randn((q, 768), generator=g, device=dev)
    qs /= qs.norm(dim=1, keepdim=True)
    index = fpv.GpuIndex(db)
    for _ in range(2):
        engine_gemm.search(qs, index, 100, "l2", mode=mode)
    torch.cuda.synchronize()
    lib = ctypes.CDLL(_native.LIB_PATH)
    buf = (ctypes.c_ulonglong * (8 * 148 * 8))()
    rc = lib.fpv_debug_trace(buf)
    assert rc == 0, rc
    t = np.array(buf, dtype=np.float64).reshape(8, 148, 8)
    print("slab  tiles/CTA | epi: aux+bar  wait_acc   work | mma: wait_tempty  wait_full  (cycles per tile)")
    for s in range(8):
        tiles = t[s, :, 5]
        if tiles.sum() == 0:
            continue
        n = np.maximum(tiles, 1)
        lead = t[s, :, 4] > 0                      # CTAs whose MMA thread ran (the leaders of CTA pairs)
        if not lead.any():
            lead[:] = True
        print(f"{s:4d} {tiles.mean():10.1f} | {np.mean(t[s,:,0]/n):10.0f} {np.mean(t[s,:,1]/n):9.0f} {np.mean(t[s,:,2]/n):6.0f} |"
              f" {np.mean((t[s,:,3]/n)[lead]):14.0f} {np.mean((t[s,:,4]/n)[lead]):10.0f}")


if __name__ == "__main__":
    main()
