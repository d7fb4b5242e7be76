import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from fastpyvectordb_b200 import ops
from bench_regimes import _time
dev = torch.device("cuda", 0)
npq = 25_000_000
codes = torch.randint(0, 256, (npq, 48), dtype=torch.uint8, device=dev)
cb = (torch.randn((48, 256, 16), device=dev) / np.sqrt(768)).contiguous()
mask = ops.pack_mask(torch.rand(npq, device=dev) < 0.25)
packed = ops.pq_pack(codes)
del codes
for qn in (1, 4):
    lut = ops.pq_build_lut(cb, torch.randn((qn, 768), device=dev))
    for tag, m in (("mask25", mask), ("nomask", None)):
        ms = _time(lambda: ops.pq_adc_packed(lut, packed, 100, m))
        print("pq q", qn, tag, round(ms, 4), "ms", round(npq * 48 * qn / ms / 1e6), "GB/s")
