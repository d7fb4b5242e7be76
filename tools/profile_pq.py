"""PQ ADC scans at the BASELINE configs[4] per-GPU shape (25M x 48 B codes) -- run it under ncu:
   ncu --set full -k regex:'pq_adc_(filter|quad|rot)' -c 6 -o gpurun_out/pq python tools/profile_pq.py"""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from fastpyvectordb_b200 import ops
dev = torch.device("cuda", 0)
npq = 25_000_000
codes = torch.randint(0, 256, (npq, 48), dtype=torch.uint8, device=dev)
cb = (torch.randn((48, 256, 16), device=dev) / np.sqrt(768)).contiguous()
mask = ops.pack_mask(torch.rand(npq, device=dev) < 0.25)
packed = ops.pq_pack(codes)
del codes
for qn in (1, 4):
    lut = ops.pq_build_lut(cb, torch.randn((qn, 768), device=dev))
    for m in (mask, None):
        ops.pq_adc_packed(lut, packed, 100, m)
torch.cuda.synchronize()
