import torch, sys
sys.path.insert(0, ".")
from fastpyvectordb_b200 import ops
from bench_regimes import _time
dev = torch.device("cuda", 0)
nb = 20_000_000
codes = torch.randint(0, 256, (nb, 128), dtype=torch.uint8, device=dev)
for qn in (4, 16, 31):
    qb = torch.randint(0, 256, (qn, 128), dtype=torch.uint8, device=dev)
    ms = _time(lambda: ops.hamming(qb, codes, 100, 1024), iters=5)
    print("hamming_mma q", qn, round(ms, 4), "ms")
