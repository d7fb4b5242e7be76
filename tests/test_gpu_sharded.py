"""Shard-count invariance on one GPU: the ranks of a sharded search are emulated as sequential local searches
over row slices (global ids via id_base), merged by the CUDA merge kernel, and compared with the unsharded
search.  Every exact kernel uses the same summation order and the same (distance, id) rule, so the answers
must be bit-identical for any shard count — which is what makes the multi-GPU result well defined."""
import numpy as np
import pytest
import torch

from oracle import oracle as O

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shards", [2, 3, 8])
@pytest.mark.parametrize("q,k,metric", [(4, 10, "cosine"), (40, 100, "l2"), (1, 100, "ip")])
def test_float_search_is_shard_count_invariant(shards, q, k, metric):
    import fastpyvectordb_b200 as fpv
    from fastpyvectordb_b200 import ops
    from fastpyvectordb_b200.sharded import pack_candidates, shard_bounds, unpack_candidates
    n, d = 30000, 128
    rng = np.random.default_rng(42)
    db = rng.standard_normal((n, d)).astype(np.float32)
    db[17] = db[n - 5]                                           # exact tie across shards
    qs = np.random.default_rng(999).standard_normal((q, d)).astype(np.float32)
    eng = fpv.ParallelSearchEngine()
    whole_d, whole_i, _ = eng.search_tensors(qs, fpv.GpuIndex(db), k, metric)
    packed, wire = [], []
    for r in range(shards):
        lo, hi = shard_bounds(n, shards, r)
        idx = fpv.GpuIndex(db[lo:hi], id_base=lo)
        dl, il, _ = eng.search_tensors(qs, idx, min(k, hi - lo), metric)
        packed.append(pack_candidates(dl, il, k))
        wire.append(ops.pack_topk(dl, il, k, lo))                  # the 8-byte wire format of the NCCL exchange
    dd, ii = unpack_candidates(torch.stack(packed))
    md, mi, mc = ops.merge_topk(dd, ii, k)
    assert torch.equal(mi, whole_i) and torch.equal(md, whole_d) and (mc == k).all()
    bases = torch.tensor([shard_bounds(n, shards, r)[0] for r in range(shards)], dtype=torch.int64, device="cuda")
    wd, wi, wc = ops.merge_packed(torch.stack(wire), bases, k)
    assert torch.equal(wi, whole_i) and torch.equal(wd, whole_d) and (wc == k).all()
    ref = O.distances_batch(qs, db, metric)
    for qi in range(q):
        O.check_topk(ref[qi], mi[qi].cpu().numpy(), md[qi].cpu().numpy(), k, squared_near_zero=(metric == "l2"))


def test_sharded_code_search_single_rank_equals_direct_scan():
    """ShardedCodeSearch with a world of one: same answer as the scan kernels called directly (the multi-rank gather
    and merge are covered by test_sharded_cpu.py and by the shard-count invariance tests in this file)."""
    from fastpyvectordb_b200 import ops
    from fastpyvectordb_b200.sharded import ShardedCodeSearch
    rng = np.random.default_rng(9)
    n = 40000
    codes = torch.from_numpy(rng.integers(0, 256, (n, 64), dtype=np.uint8)).cuda()
    qb = torch.from_numpy(rng.integers(0, 256, (3, 64), dtype=np.uint8)).cuda()
    d0, i0, _, _ = ops.hamming(qb, codes, 50, 512)
    d1, i1, c1 = ShardedCodeSearch("hamming", codes, n, dims=512).search_tensors(qb, 50)
    assert torch.equal(i0, i1) and torch.equal(d0, d1) and (c1 == 50).all()
    pcodes = torch.from_numpy(rng.integers(0, 256, (n, 48), dtype=np.uint8)).cuda()
    lut = torch.from_numpy(rng.random((2, 48, 256)).astype(np.float32)).cuda()
    words = ops.pack_mask(torch.from_numpy(rng.random(n) < 0.25).cuda())
    s = ShardedCodeSearch("pq", pcodes, n)
    d1, i1, c1 = s.search_tensors(lut, 20, words)
    if ops.pq_adc_packed_supported(2, n, 48, 256, 20):
        d0, i0, _ = ops.pq_adc_packed(lut, ops.pq_pack(pcodes), 20, words)
    else:
        d0, i0, _, _ = ops.pq_adc(lut, pcodes, 20, words)
    assert torch.equal(i0, i1) and torch.equal(d0, d1) and (c1 == 20).all()


def test_quantized_scans_are_shard_count_invariant():
    import fastpyvectordb_b200 as fpv
    from fastpyvectordb_b200 import ops
    from fastpyvectordb_b200.sharded import pack_candidates, shard_bounds, unpack_candidates
    rng = np.random.default_rng(5)
    n = 50000
    codes = torch.from_numpy(rng.integers(0, 256, (n, 128), dtype=np.uint8)).cuda()
    qb = torch.from_numpy(rng.integers(0, 256, (2, 128), dtype=np.uint8)).cuda()
    wd, wi, _, _ = ops.hamming(qb, codes, 100, 1024)
    packed = []
    for r in range(4):
        lo, hi = shard_bounds(n, 4, r)
        dl, il, _, _ = ops.hamming(qb, codes[lo:hi].contiguous(), 100, 1024, None, lo)
        packed.append(pack_candidates(dl, il, 100))
    dd, ii = unpack_candidates(torch.stack(packed))
    md, mi, _ = ops.merge_topk(dd, ii, 100)
    assert torch.equal(mi, wi) and torch.equal(md, wd)
    # PQ ADC with the row bitmask sharded with the rows
    pcodes = torch.from_numpy(rng.integers(0, 256, (n, 48), dtype=np.uint8)).cuda()
    cb = torch.from_numpy((rng.standard_normal((48, 256, 16)) / 27.7).astype(np.float32)).cuda()
    q = torch.from_numpy(rng.standard_normal((1, 768)).astype(np.float32)).cuda()
    lut = ops.pq_build_lut(cb, q)
    mask = torch.from_numpy(rng.random(n) < 0.25).cuda()
    wd, wi, _, _ = ops.pq_adc(lut, pcodes, 100, ops.pack_mask(mask))
    packed = []
    for r in range(5):
        lo, hi = shard_bounds(n, 5, r)
        dl, il, _, _ = ops.pq_adc(lut, pcodes[lo:hi].contiguous(), 100, ops.pack_mask(mask[lo:hi]), lo)
        packed.append(pack_candidates(dl, il, 100))
    dd, ii = unpack_candidates(torch.stack(packed))
    md, mi, _ = ops.merge_topk(dd, ii, 100)
    assert torch.equal(mi, wi) and torch.equal(md, wd)
