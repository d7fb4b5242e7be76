"""Batched Hamming scan on the int8 tensor cores (csrc/fpv_hamming_mma.cu) against
 (a) exact integer arithmetic: the raw accumulators equal -popcount(x & q) computed with NumPy, bit for bit, and
 (b) the CUDA-core scan (same distances and ids, ties by lowest row), incl. the row filter, two passes of queries and a
     tie group too large for the sorter (device-side fallback)."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu

POP8 = np.array([bin(i).count("1") for i in range(256)], dtype=np.int32)


@pytest.mark.parametrize("n,nbytes,dims", [(70000, 128, 0), (66000, 256, 0), (65536, 128, 1000)])
def test_tensor_core_popcount_dots_are_exact(n, nbytes, dims):
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(3)
    codes = rng.integers(0, 256, (n, nbytes), dtype=np.uint8)
    qb = rng.integers(0, 256, (7, nbytes), dtype=np.uint8)
    out = ops.hamming_mma_dots(torch.from_numpy(qb).cuda(), torch.from_numpy(codes).cuda(), dims).cpu().numpy()
    dm = np.full(nbytes, 0xFF, np.uint8)
    if dims:
        bits = np.zeros(nbytes * 8, np.uint8)
        bits[:dims] = 1
        dm = np.packbits(bits)
    assert np.array_equal(out[31], -POP8[codes & dm].sum(axis=1))
    for qi in range(7):
        assert np.array_equal(out[qi], -POP8[codes & qb[qi] & dm].sum(axis=1)), f"query {qi}"
    assert (out[7:31] == 0).all()


@pytest.mark.parametrize("n,nbytes,dims,q,k,mask", [(70000, 128, 1024, 16, 100, False), (131072, 128, 1000, 40, 10, True),
                                                    (66000, 256, 2048, 5, 100, True), (100000, 128, 1024, 31, 1000, False)])
def test_tensor_core_hamming_equals_cuda_core_scan(n, nbytes, dims, q, k, mask):
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(9)
    codes = torch.from_numpy(rng.integers(0, 256, (n, nbytes), dtype=np.uint8)).cuda()
    codes[4321] = codes[99]                                             # exact duplicate rows: ties broken by the lower row
    qb = torch.from_numpy(rng.integers(0, 256, (q, nbytes), dtype=np.uint8)).cuda()
    qb[1] = codes[99]                                                   # distance 0 twice
    words = ops.pack_mask(torch.from_numpy(rng.random(n) < 0.3).cuda()) if mask else None
    assert ops.hamming_mma_supported(q, n, nbytes, k)
    d1, i1, c1, _ = ops.hamming(qb, codes, k, dims, words)              # tensor-core route (q >= 4)
    for qi in range(q):
        d0, i0, c0, _ = ops.hamming(qb[qi:qi + 1].contiguous(), codes, k, dims, words)     # one query: CUDA-core scan
        assert torch.equal(i1[qi], i0[0]) and torch.equal(d1[qi], d0[0]) and int(c1[qi]) == int(c0[0]), f"query {qi}"
    if not mask:
        host = codes.cpu().numpy()
        ref = O.bq_hamming(qb[1].cpu().numpy(), host, dims)
        O.check_topk(ref, i1[1].cpu().numpy(), d1[1].cpu().numpy(), k, integer=True)
        assert d1[1, 0].item() == 0.0 and i1[1, :2].tolist() == [99, 4321]


def test_huge_tie_group_falls_back_on_the_device():
    from fastpyvectordb_b200 import ops
    rng = np.random.default_rng(1)
    n = 70000
    codes = torch.zeros((n, 128), dtype=torch.uint8, device="cuda")
    codes[:, 0] = torch.from_numpy(rng.integers(0, 2, n).astype(np.uint8)).cuda()          # two distinct rows: 35000-row tie groups
    qb = torch.zeros((6, 128), dtype=torch.uint8, device="cuda")
    d1, i1, c1, _ = ops.hamming(qb, codes, 50, 1024)
    ops.HAMMING_TENSOR_CORES = False
    try:
        d0, i0, c0, _ = ops.hamming(qb, codes, 50, 1024)
    finally:
        ops.HAMMING_TENSOR_CORES = True
    assert torch.equal(i1, i0) and torch.equal(d1, d0) and torch.equal(c1, c0)


def test_binary_quantizer_batch_search_matches_single_query_search():
    import fastpyvectordb_b200 as fpv
    rng = np.random.default_rng(4)
    x = rng.standard_normal((80000, 1024)).astype(np.float32)
    bq = fpv.BinaryQuantizer(1024).train(x[:5000])
    codes = bq.to_device(bq.encode(x))
    qs = rng.standard_normal((9, 1024)).astype(np.float32)
    idx, dist = bq.search_batch(qs, codes, k=20)
    idx, dist = (idx.cpu().numpy(), dist.cpu().numpy()) if isinstance(idx, torch.Tensor) else (idx, dist)
    for qi in range(9):
        i0, d0 = bq.search(qs[qi], codes, k=20)
        i0, d0 = (i0.cpu().numpy(), d0.cpu().numpy()) if isinstance(i0, torch.Tensor) else (i0, d0)
        assert np.array_equal(idx[qi], i0) and np.array_equal(dist[qi], d0)
