"""Parity at BASELINE.json's FULL sizes (one B200), where the NumPy oracle cannot hold a Q x N matrix.

Each test checks the CUDA path through size-independent properties plus an independent recomputation of the
distances of ALL rows for a few queries (plain fp32 NumPy for the float path; plain torch integer / gather ops on
the device for the quantized scans — no kernel of this repo is involved in the reference side), judged by the same
tie-aware checker as the small cases (`oracle.check_topk`).

    configs[1]  1M x 768 fp32, Q = 4096, L2 top-100           (tensor-core path, BF16 filter + exact re-rank)
    configs[3]  20M x 1024: binary codes (Hamming) and uint8 scalar codes (L2)
    configs[4]  25M x 48-byte PQ codes (one GPU's shard of 200M), 25 % row bitmask
"""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


def _need(gb):
    free, _total = torch.cuda.mem_get_info()
    if free < gb * 2**30:
        pytest.skip(f"needs {gb} GiB of free device memory")


@pytest.fixture(scope="module")
def fpv():
    import fastpyvectordb_b200 as m
    return m


def test_c1_full_size_batch_search(fpv):
    from fastpyvectordb_b200 import engine_gemm
    _need(12)
    n, d, q, k = 1_000_000, 768, 4096, 100
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(7)
    db = torch.randn((n, d), generator=g, device=dev)
    db /= db.norm(dim=1, keepdim=True)
    qs = torch.randn((q, d), generator=g, device=dev)
    qs /= qs.norm(dim=1, keepdim=True)
    planted = torch.arange(64, device=dev) * 15013 + 11          # query i (< 64) is row planted[i] exactly
    db[planted] = qs[:64]
    db[123456] = db[planted[0]]                                  # and an exact duplicate: tie broken by index
    index = fpv.GpuIndex(db)
    eng = fpv.ParallelSearchEngine()
    dist, idx, cnt = eng.search_tensors(qs, index, k, "l2")
    torch.cuda.synchronize()
    assert engine_gemm.last_fallback_fraction(index, q, k) <= 0.01
    assert (cnt == k).all()
    # ordered by (distance, index), ids unique and in range
    assert (dist[:, 1:] >= dist[:, :-1]).all()
    tie = dist[:, 1:] == dist[:, :-1]
    assert (idx[:, 1:][tie] > idx[:, :-1][tie]).all()
    assert idx.min() >= 0 and idx.max() < n
    srt = idx.sort(dim=1).values
    assert (srt[:, 1:] != srt[:, :-1]).all()
    # planted rows come first with distance ~0 (the reference's sqrt(max(0, q.q + v.v - 2 q.v)) form is noisy at 0)
    assert torch.equal(idx[1:64, 0], planted[1:64]) and (dist[:64, 0] < 2e-3).all()
    lo, hi = sorted((int(planted[0]), 123456))
    assert idx[0, 0] == lo and idx[0, 1] == hi and dist[0, 0] == dist[0, 1]
    # batch-split and path invariance: bit-identical answers from a small batch (fp32 scan kernel) and a sub-batch
    from fastpyvectordb_b200 import ops
    sd, si, _ = ops.scan_f32_topk(qs[100:103].contiguous(), index.rows, k, "l2", None, index.row_sq, 0)
    assert torch.equal(si, idx[100:103]) and torch.equal(sd, dist[100:103])
    sd, si, _ = eng.search_tensors(qs[100:103], index, k, "l2")      # 1-3 queries reuse the bf16 shadow once it exists
    assert torch.equal(si, idx[100:103]) and torch.equal(sd, dist[100:103])
    bd, bi, _ = eng.search_tensors(qs[1024:1536], index, k, "l2")
    assert torch.equal(bi, idx[1024:1536]) and torch.equal(bd, dist[1024:1536])
    # independent fp32 recomputation of all 1M distances for a sample of queries (oracle formula, NumPy on the host)
    db_h = db.cpu().numpy()
    sample = [0, 1, 63, 64, 2047, 4095]
    ref = O.distances_batch(qs[sample].cpu().numpy(), db_h, "l2")
    for j, qi in enumerate(sample):
        O.check_topk(ref[j], idx[qi].cpu().numpy(), dist[qi].cpu().numpy(), k, squared_near_zero=True)
    # inner product on the same index (configs[1] names both metrics)
    dist_ip, idx_ip, _ = eng.search_tensors(qs[:512], index, k, "ip")
    ref = O.distances_batch(qs[[5, 300]].cpu().numpy(), db_h, "ip")
    for j, qi in enumerate([5, 300]):
        O.check_topk(ref[j], idx_ip[qi].cpu().numpy(), dist_ip[qi].cpu().numpy(), k)


def test_c3_full_size_hamming(fpv):
    from fastpyvectordb_b200 import ops
    _need(12)
    n, nbytes, k = 20_000_000, 128, 100
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(11)
    codes = torch.randint(0, 256, (n, nbytes), generator=g, device=dev, dtype=torch.uint8)
    qb = torch.randint(0, 256, (2, nbytes), generator=g, device=dev, dtype=torch.uint8)
    # planted rows at known Hamming distances 0, 1, 2, ..., 9 from query 0 (far below the ~430 of random rows)
    for t in range(10):
        row = qb[0].clone()
        for b in range(t):
            row[b] ^= 1
        codes[1_999_999 * (t + 1)] = row
    dist, idx, cnt, _ = ops.hamming(qb, codes, k, 1024)
    torch.cuda.synchronize()
    assert (cnt == k).all()
    assert idx[0, :10].tolist() == [1_999_999 * (t + 1) for t in range(10)]
    assert dist[0, :10].tolist() == [float(t) for t in range(10)]
    # independent recomputation: XOR + 256-entry popcount table, in row chunks, plain torch
    table = torch.tensor([bin(i).count("1") for i in range(256)], dtype=torch.int32, device=dev)
    for qi in range(2):
        ref = torch.empty(n, dtype=torch.int32, device=dev)
        for lo in range(0, n, 2_000_000):
            x = torch.bitwise_xor(codes[lo:lo + 2_000_000], qb[qi])
            ref[lo:lo + 2_000_000] = table[x.long()].sum(dim=1)
        O.check_topk(ref.cpu().numpy(), idx[qi].cpu().numpy(), dist[qi].cpu().numpy(), k, integer=True)
        d_i, i_i = dist[qi], idx[qi]
        assert ((d_i[1:] > d_i[:-1]) | ((d_i[1:] == d_i[:-1]) & (i_i[1:] > i_i[:-1]))).all()   # ties by lowest index


def test_c3_full_size_scalar_l2(fpv):
    from fastpyvectordb_b200 import ops
    _need(30)
    n, d, k = 20_000_000, 1024, 100
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(13)
    codes = torch.randint(0, 256, (n, d), generator=g, device=dev, dtype=torch.uint8)
    qc = torch.randint(0, 256, (1, d), generator=g, device=dev, dtype=torch.uint8)
    mn = torch.rand(d, generator=g, device=dev) - 0.5
    sc = torch.rand(d, generator=g, device=dev) * 0.2 + 0.01
    codes[7_777_777] = qc[0]                                      # distance 0
    near = qc[0].clone()
    near[5] = near[5] ^ 1                                         # one code off by one
    codes[19_999_999] = near
    dist, idx, cnt, _ = ops.sq_scan(0, qc, codes, mn, sc, k)      # kind 0 = L2 (quantization.py:145-152)
    torch.cuda.synchronize()
    assert cnt[0] == k and idx[0, 0] == 7_777_777 and dist[0, 0] == 0.0 and idx[0, 1] == 19_999_999
    # the reference's arithmetic (quantization.py:217-236): (q - b) int16 -> fp32 * (scale / 255) -> square -> sum -> sqrt
    s255 = sc / 255.0
    ref = torch.empty(n, dtype=torch.float32, device=dev)
    qf = qc[0].to(torch.int16)
    for lo in range(0, n, 500_000):
        diff = (qf - codes[lo:lo + 500_000].to(torch.int16)).to(torch.float32) * s255
        ref[lo:lo + 500_000] = torch.sqrt((diff * diff).sum(dim=1))
    O.check_topk(ref.cpu().numpy(), idx[0].cpu().numpy(), dist[0].cpu().numpy(), k)
    # the int8 tensor-core scan (what ScalarQuantizer.search runs at this size): same bits as the CUDA-core scan
    term, tmax = ops.sq_row_term(codes, sc)
    q3 = torch.cat([qc, torch.randint(0, 256, (2, d), generator=g, device=dev, dtype=torch.uint8)])
    md, mi, mc = ops.sq_l2_mma(q3, codes, mn, sc, term, tmax, k)
    assert int(ops.sq_mma_last_flags(3, n, d, k, dev).sum()) == 0
    assert torch.equal(mi[0], idx[0]) and torch.equal(md[0], dist[0]) and bool((mc == k).all())


@pytest.mark.parametrize("metric", ["dot", "cosine"])
def test_c3_full_size_scalar_dot_cosine_on_the_tensor_cores(fpv, metric):
    """distances_dot / distances_cosine + top-100 over 20M x 1024 codes on the int8 tensor cores, against the reference's
    arithmetic (quantization.py:154-181, 239-251) recomputed with plain torch in row chunks."""
    from fastpyvectordb_b200 import ops, _native
    _need(30)
    n, d, k = 20_000_000, 1024, 100
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(19)
    codes = torch.randint(0, 256, (n, d), generator=g, device=dev, dtype=torch.uint8)
    qc = torch.randint(0, 256, (2, d), generator=g, device=dev, dtype=torch.uint8)
    mn = torch.rand(d, generator=g, device=dev) - 0.5
    sc = torch.rand(d, generator=g, device=dev) * 0.2 + 0.01
    codes[3_333_333] = qc[0]                                      # the query's own codes: cosine distance ~0
    kind = _native.SQ_DOT if metric == "dot" else _native.SQ_COSINE
    rsum, rinv, maxima = ops.sq_row_terms_dc(codes, mn, sc)
    dist, idx, cnt = ops.sq_dc_mma(kind, qc, codes, mn, sc, rsum, rinv, maxima, k)
    torch.cuda.synchronize()
    assert bool((cnt == k).all()) and int(ops.sq_mma_last_flags(2, n, d, k, dev).sum()) == 0
    if metric == "cosine":
        assert idx[0, 0] == 3_333_333 and abs(float(dist[0, 0])) < 1e-5
    for qi in range(2):
        qdec = qc[qi].to(torch.float32) / 255.0 * sc + mn
        if metric == "cosine":
            qdec = qdec / (torch.linalg.norm(qdec) + 1e-8)
        ref = torch.empty(n, dtype=torch.float32, device=dev)
        for lo in range(0, n, 500_000):
            dec = codes[lo:lo + 500_000].to(torch.float32) / 255.0 * sc + mn
            if metric == "cosine":
                dec = dec / (torch.linalg.norm(dec, dim=1, keepdim=True) + 1e-8)
                ref[lo:lo + 500_000] = 1.0 - dec @ qdec
            else:
                ref[lo:lo + 500_000] = -(dec @ qdec)
        O.check_topk(ref.cpu().numpy(), idx[qi].cpu().numpy(), dist[qi].cpu().numpy(), k, rtol=1e-5)
    # and the CUDA-core scan of the same metric returns the same bits
    sd, si, scn, _ = ops.sq_scan(kind, qc[:1].contiguous(), codes, mn, sc, k)
    assert torch.equal(si[0], idx[0]) and torch.equal(sd[0], dist[0])


def test_c4_full_shard_pq_adc_with_bitmask(fpv):
    from fastpyvectordb_b200 import ops
    _need(8)
    n, m, kc, k = 25_000_000, 48, 256, 100
    dev = torch.device("cuda", 0)
    g = torch.Generator(device=dev)
    g.manual_seed(17)
    codes = torch.randint(0, 256, (n, m), generator=g, device=dev, dtype=torch.uint8)
    lut = torch.rand((1, m, kc), generator=g, device=dev) * 0.05
    mask = torch.rand(n, generator=g, device=dev) < 0.25
    words = ops.pack_mask(mask)
    # reference: sequential-in-m fp32 accumulation of table[m, codes[:, m]] (quantization.py:571-578), then sqrt
    ref = torch.zeros(n, dtype=torch.float32, device=dev)
    for j in range(m):
        ref += lut[0, j][codes[:, j].long()]
    ref = torch.sqrt(ref)
    ref_h, mask_h = ref.cpu().numpy(), mask.cpu().numpy()
    # exact-order kernel on the reference code layout: bit-identical distances
    dist, idx, cnt, _ = ops.pq_adc(lut, codes, k, words)
    torch.cuda.synchronize()
    assert cnt[0] == k and mask[idx[0]].all()
    assert torch.equal(dist[0], ref[idx[0]])
    O.check_topk(ref_h, idx[0].cpu().numpy(), dist[0].cpu().numpy(), k, valid=mask_h)
    # packed (bank-conflict-free) layout: same rows, distances to fp32 rounding
    if ops.pq_adc_packed_supported(1, n, m, kc, k):
        packed = ops.pq_pack(codes)
        pd, pi, pc = ops.pq_adc_packed(lut, packed, k, words)[:3]
        O.check_topk(ref_h, pi[0].cpu().numpy(), pd[0].cpu().numpy(), k, valid=mask_h)
        # a batch of four queries shares one pass (fixed-point tables + exact re-score): query 0 must come back bit for bit
        lut4 = torch.cat([lut, torch.rand((3, m, kc), generator=g, device=dev) * 0.05])
        qd, qi4, qcn = ops.pq_adc_packed(lut4, packed, k, words)[:3]
        assert torch.equal(qi4[0], pi[0]) and torch.equal(qd[0], pd[0]) and bool((qcn == k).all())
        ref3 = torch.zeros(n, dtype=torch.float32, device=dev)
        for j in range(m):
            ref3 += lut4[3, j][codes[:, j].long()]
        O.check_topk(torch.sqrt(ref3).cpu().numpy(), qi4[3].cpu().numpy(), qd[3].cpu().numpy(), k, valid=mask_h)
