#!/usr/bin/env python
"""Index-build throughput (encoders + trainer), one B200: CUDA-event times, rows/s and the HBM roofline
fraction on algorithmic bytes (fp32 rows read once + codes written once).

    python tools/bench_index_build.py > profiles/r02_index_build.json"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import fastpyvectordb_b200 as fpv  # noqa: E402
from fastpyvectordb_b200 import ops  # noqa: E402
from bench_regimes import _time  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {"hbm_gbs": 6650.0}
    hbm = peaks["hbm_gbs"]
    out = {"peak_hbm_gbs": hbm}

    def line(name, ms, rows, nbytes, extra=None):
        r = {"ms": ms, "rows_per_s": rows / (ms * 1e-3), "algorithmic_bytes": nbytes,
             "hbm_frac": nbytes / (ms * 1e-3) / 1e9 / hbm}
        if extra:
            r.update(extra)
        out[name] = r

    n, d = 4_000_000, 1024
    x = torch.randn((n, d), device=dev) * 0.1
    mn = x.min(0).values.contiguous(); sc = (x.max(0).values - mn).contiguous()
    ms = _time(lambda: ops.sq_encode(x, mn, sc), iters=5)
    line("sq_encode_4Mx1024", ms, n, n * d * 5.0)
    thr = torch.zeros(d, device=dev)
    ms = _time(lambda: ops.bq_encode(x, thr), iters=5)
    line("bq_encode_4Mx1024", ms, n, n * d * 4.0 + n * d / 8)
    codes = ops.sq_encode(x, mn, sc)
    ms = _time(lambda: ops.sq_row_term(codes, sc), iters=5)
    line("sq_row_term_4Mx1024", ms, n, n * d * 1.0 + n * 4)
    del x, codes
    torch.cuda.empty_cache()
    n, d = 2_000_000, 768
    x = torch.randn((n, d), device=dev)
    cb = (torch.randn((48, 256, 16), device=dev) / np.sqrt(768)).contiguous()
    ms = _time(lambda: ops.pq_encode(x, cb), iters=3)
    flops = 2.0 * n * 48 * 256 * 16 * 1.5          # sub, mul, add per (row, centroid, dim)
    line("pq_encode_2Mx768_m48", ms, n, n * d * 4.0 + n * 48, {"fp32_tflops": flops / (ms * 1e-3) / 1e12,
                                                               "note": "compute bound: 3 * 256 fp32 operations per input element"})
    pcodes = ops.pq_encode(x[:1_000_000].contiguous(), cb)
    ms = _time(lambda: ops.pq_pack(pcodes), iters=5)
    line("pq_pack_1Mx48", ms, 1_000_000, 1_000_000 * 96.0)
    # trainer: 100k x 768 sample, M = 48, K = 256, 20 Lloyd iterations (the reference: 485 s for 5k rows)
    np.random.seed(0)
    pq = fpv.ProductQuantizer(768, 48, 256, device=dev)
    sample = x[:100_000].contiguous()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    pq.train(sample, n_iter=20)
    torch.cuda.synchronize()
    dt = time.perf_counter() - t0
    out["pq_train_100kx768_m48_k256_20iter"] = {"seconds": dt, "rows_per_s": 100_000 / dt}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
