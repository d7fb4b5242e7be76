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


@pytest.mark.parametrize("shards", [2, 5])
@pytest.mark.parametrize("q,k,metric,mode", [(40, 100, "l2", "bf16"), (300, 10, "cosine", "bf16"), (130, 64, "ip", "tf32")])
def test_two_phase_sharded_search_equals_unsharded(shards, q, k, metric, mode):
    """The row-sharded tensor-core path (phase 1 filter -> exchange of the k best approximate values -> phase 2
    re-rank under the GLOBAL limit -> merge) with the ranks emulated one after the other on one GPU, each with its own
    workspace.  The answer must be bit-identical to the unsharded search, and every shard must have re-ranked fewer
    rows than a local window would hold."""
    import fastpyvectordb_b200 as fpv
    from fastpyvectordb_b200 import engine_gemm as eg, ops
    from fastpyvectordb_b200.sharded import shard_bounds
    n, d = 60000, 128
    rng = np.random.default_rng(42)
    db = rng.standard_normal((n, d)).astype(np.float32)
    db[17] = db[n - 5]                                           # exact tie across shards
    db[40000:40100] *= 3.0                                       # the row-norm maximum lives in one shard only
    qs = torch.from_numpy(np.random.default_rng(999).standard_normal((q, d)).astype(np.float32)).cuda()
    eng = fpv.ParallelSearchEngine()
    whole_d, whole_i, _ = eng.search_tensors(qs, fpv.GpuIndex(db), k, metric)
    idxs, wss = [], []
    for r in range(shards):
        lo, hi = shard_bounds(n, shards, r)
        idxs.append(fpv.GpuIndex(db[lo:hi], id_base=lo))
    bounds = [eg.sharded_bounds(ix, mode) for ix in idxs]        # what the MAX all-reduce does
    vmax = max(b[0] for b in bounds)
    err = (max(b[1][0] for b in bounds), max(b[1][1] for b in bounds))
    approx = []
    for ix in idxs:
        eg.set_sharded_bounds(ix, vmax, err)
        ws = ops.gemm_workspace(q, ix.n, d, k, 0 if mode == "tf32" else 1, "cuda")
        wss.append(ws)
        approx.append(eg.filter_sharded(qs, ix, k, metric, mode, ws=ws))
    gathered = torch.stack(approx).contiguous()
    wire, counts = [], []
    for ix, ws in zip(idxs, wss):
        dl, il, cl = eg.finish_sharded(qs, ix, k, metric, mode, gathered, ws=ws)
        assert (il[dl.isinf()] == -1).all()
        counts.append(cl)
        wire.append(ops.pack_topk(dl, il, k, ix.id_base))
    bases = torch.tensor([ix.id_base for ix in idxs], dtype=torch.int64, device="cuda")
    md, mi, mc = ops.merge_packed(torch.stack(wire), bases, k)
    assert torch.equal(mi, whole_i) and torch.equal(md, whole_d) and (mc == k).all()
    # the point of the exchange: the shards together re-rank about one window, not one window each
    total = torch.stack(counts).sum(0).float().mean().item()
    assert total <= min(k * shards, 3.0 * k + 64), total
    ref = O.distances_batch(qs.cpu().numpy()[:8], db, metric)
    for qi in range(8):
        O.check_topk(ref[qi], mi[qi].cpu().numpy(), md[qi].cpu().numpy(), k, squared_near_zero=(metric == "l2"))


@pytest.mark.parametrize("shards", [2, 3])
@pytest.mark.parametrize("q,k,metric,mode", [(300, 100, "l2", "bf16"), (40, 10, "cosine", "bf16"), (130, 64, "ip", "tf32")])
def test_pooled_sample_sharded_search_equals_unsharded(shards, q, k, metric, mode):
    """Shards of >= 70K rows pool their samples: sampling slab per shard -> exchange of the k best group values -> the k-th
    best of the WHOLE job is every shard's first threshold -> filtering slabs -> the two exchanges of the two-phase route.
    Ranks emulated one after the other on one GPU, each with its own workspace; bit-identical to the unsharded search."""
    import fastpyvectordb_b200 as fpv
    from fastpyvectordb_b200 import engine_gemm as eg, ops
    from fastpyvectordb_b200.sharded import shard_bounds
    n, d = 75000 * shards, 64
    rng = np.random.default_rng(4)
    db = rng.standard_normal((n, d)).astype(np.float32)
    db[17] = db[n - 5]                                           # exact tie across shards
    db[n - 300:n - 200] *= 3.0                                   # the row-norm maximum lives in the last shard only
    qs = torch.from_numpy(np.random.default_rng(99).standard_normal((q, d)).astype(np.float32)).cuda()
    eng = fpv.ParallelSearchEngine()
    whole_d, whole_i, _ = eng.search_tensors(qs, fpv.GpuIndex(db), k, metric)
    idxs = [fpv.GpuIndex(db[lo:hi], id_base=lo) for lo, hi in (shard_bounds(n, shards, r) for r in range(shards))]
    bounds = [eg.sharded_bounds(ix, mode) for ix in idxs]
    vmax = max(b[0] for b in bounds)
    err = (max(b[1][0] for b in bounds), max(b[1][1] for b in bounds))
    wss, samples = [], []
    for ix in idxs:
        eg.set_sharded_bounds(ix, vmax, err)
        wss.append(ops.gemm_workspace(q, ix.n, d, k, 0 if mode == "tf32" else 1, "cuda"))
        samples.append(eg.sample_sharded(qs, ix, k, metric, mode, ws=wss[-1]))
    pooled = torch.stack(samples).contiguous()
    approx = [eg.slabs_sharded(qs, ix, k, metric, mode, pooled, shards, ws=ws) for ix, ws in zip(idxs, wss)]
    gathered = torch.stack(approx).contiguous()
    wire = []
    for ix, ws in zip(idxs, wss):
        dl, il, cl = eg.finish_sharded(qs, ix, k, metric, mode, gathered, ws=ws)
        wire.append(ops.pack_topk(dl, il, k, ix.id_base))
    bases = torch.tensor([ix.id_base for ix in idxs], dtype=torch.int64, device="cuda")
    md, mi, mc = ops.merge_packed(torch.stack(wire), bases, k)
    assert torch.equal(mi, whole_i) and torch.equal(md, whole_d) and (mc == k).all()


def test_merge_packed_rank_merge_edge_cases():
    """fpv_merge_packed ranks entries by binary search over the sorted lists: ties across shards go to the lower
    global id, empty slots are skipped, short and empty lists are fine."""
    from fastpyvectordb_b200 import ops
    inf = float("inf")
    d = torch.tensor([[[0.5, 0.5, 2.0, inf]], [[0.5, 1.0, inf, inf]], [[inf, inf, inf, inf]]], dtype=torch.float32).cuda()   # [3 shards][1][4]
    i = torch.tensor([[[3, 9, 4, -1]], [[100, 101, -1, -1]], [[-1, -1, -1, -1]]], dtype=torch.int64).cuda()
    bases = torch.tensor([0, 100, 200], dtype=torch.int64).cuda()
    wire = torch.stack([ops.pack_topk(d[s], i[s], 4, int(bases[s])) for s in range(3)])
    md, mi, mc = ops.merge_packed(wire, bases, 6)
    assert mi[0].tolist() == [3, 9, 100, 101, 4, -1] and int(mc[0]) == 5
    assert md[0, :5].tolist() == [0.5, 0.5, 0.5, 1.0, 2.0] and md[0, 5].item() == inf
    md, mi, mc = ops.merge_packed(wire, bases, 2)
    assert mi[0].tolist() == [3, 9] and int(mc[0]) == 2
