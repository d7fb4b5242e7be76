"""uint8 scalar-quantizer scans at BASELINE configs[3] (20M x 1024): tensor-core L2 / dot / cosine (Q = 1, 16) and the
CUDA-core scans of the same metrics (Q = 1)."""
import sys
import torch
sys.path.insert(0, ".")
from fastpyvectordb_b200 import ops, _native
from bench_regimes import _time
dev = torch.device("cuda", 0)
ns = int(sys.argv[1]) if len(sys.argv) > 1 else 20_000_000
codes = torch.randint(0, 256, (ns, 1024), dtype=torch.uint8, device=dev)
mn = torch.full((1024,), -0.1, device=dev)
sc = torch.full((1024,), 0.2, device=dev)
term, tmax = ops.sq_row_term(codes, sc)
rsum, rinv, maxima = ops.sq_row_terms_dc(codes, mn, sc)
for qn in (1, 16):
    qc = torch.randint(0, 256, (qn, 1024), dtype=torch.uint8, device=dev)
    ms = _time(lambda: ops.sq_l2_mma(qc, codes, mn, sc, term, tmax, 100), iters=5)
    print("l2 mma q", qn, round(ms, 3), "ms", round(ns * 1024 / ms / 1e6), "GB/s", "fallbacks", int(ops.sq_mma_last_flags(qn, ns, 1024, 100, dev).sum()))
    for name, kind in (("dot", _native.SQ_DOT), ("cosine", _native.SQ_COSINE)):
        ms = _time(lambda: ops.sq_dc_mma(kind, qc, codes, mn, sc, rsum, rinv, maxima, 100), iters=5)
        print(name, "mma q", qn, round(ms, 3), "ms", round(ns * 1024 / ms / 1e6), "GB/s", "fallbacks", int(ops.sq_mma_last_flags(qn, ns, 1024, 100, dev).sum()))
qc = torch.randint(0, 256, (1, 1024), dtype=torch.uint8, device=dev)
for name, kind in (("l2", _native.SQ_L2), ("dot", _native.SQ_DOT), ("cosine", _native.SQ_COSINE)):
    ms = _time(lambda: ops.sq_scan(kind, qc, codes, mn, sc, 100), iters=3)
    print(name, "simt q 1", round(ms, 3), "ms", round(ns * 1024 / ms / 1e6), "GB/s")
