"""world_size-2 gloo tests (CPU) of the row-sharding host logic: bounds, global ids, packing, gather layout, merge.
The local scan and the merge kernel are replaced by oracle-based callables here (test infrastructure); the GPU
versions of the same path are covered by tests/test_gpu_sharded.py and bench.py --gpus N."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fastpyvectordb_b200.sharded import ShardedTopK, pack_candidates, shard_bounds, unpack_candidates
from oracle import oracle as O


def test_shard_bounds_cover_rows_exactly_once():
    for n in (0, 1, 7, 8, 9, 1000, 1_000_003):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert a <= b == c <= d


def test_pack_roundtrip_keeps_bits_and_pads():
    d = torch.tensor([[0.5, -1.25, float("inf")], [1e-30, -0.0, 3.0]], dtype=torch.float32)
    i = torch.tensor([[7, 2**40, 5], [1, 0, 9]], dtype=torch.int64)
    p = pack_candidates(d, i, 5)
    assert p.shape == (2, 5, 2)
    dd, ii = unpack_candidates(p.unsqueeze(0))
    assert torch.equal(ii[0, :, :3], i) and torch.equal(dd[0, :, :3].view(torch.int32), d.view(torch.int32))
    assert (ii[0, :, 3:] == -1).all() and torch.isinf(dd[0, :, 3:]).all()


def _oracle_merge(d, i, k_out):
    """stand-in for fpv_merge_topk with the same contract: order by (distance, id), ids < 0 are empty slots"""
    s, q, k = d.shape
    od = torch.full((q, k_out), float("inf"))
    oi = torch.full((q, k_out), -1, dtype=torch.int64)
    oc = torch.zeros(q, dtype=torch.int32)
    for qi in range(q):
        dd = d[:, qi].reshape(-1).numpy()
        ii = i[:, qi].reshape(-1).numpy()
        keep = ii >= 0
        order = np.lexsort((ii[keep], dd[keep]))[:k_out]
        od[qi, :len(order)] = torch.from_numpy(dd[keep][order])
        oi[qi, :len(order)] = torch.from_numpy(ii[keep][order])
        oc[qi] = len(order)
    return od, oi, oc


def _worker(rank, world, port, n, d, k, metric, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        rng = np.random.default_rng(42)
        db = rng.standard_normal((n, d)).astype(np.float32)
        if n > 5:
            db[5] = db[n - 3]                               # a cross-shard exact tie
        qs = np.random.default_rng(999).standard_normal((3, d)).astype(np.float32)
        topk = ShardedTopK(n, merge_fn=_oracle_merge)
        assert topk.peer_setup("cpu", 1024, 1024) is False      # peer memory needs NCCL + CUDA IPC: gloo keeps the collective route
        lo, hi = topk.lo, topk.hi
        assert (lo, hi) == shard_bounds(n, world, rank)
        rows = db[lo:hi]
        kl = min(k, hi - lo)
        dl = np.zeros((len(qs), kl), np.float32)
        il = np.zeros((len(qs), kl), np.int64)
        for qi, q in enumerate(qs):
            if kl:
                dist_local = O.distances_single(q, rows, metric)
                idx, dd = O.canonical_topk(dist_local, kl)
                dl[qi], il[qi] = dd, idx + lo               # global ids: local + shard start (parallel_search.py:353)
        md, mi, mc = topk.merge(torch.from_numpy(dl), torch.from_numpy(il), k)
        if rank == 0:
            out.put((md.numpy(), mi.numpy(), mc.numpy()))
        gathered = [None] * world
        dist.all_gather_object(gathered, (mi.numpy().tobytes(), md.numpy().tobytes()))
        assert all(g == gathered[0] for g in gathered)      # every rank holds the same merged answer
    finally:
        dist.destroy_process_group()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("n,k,metric", [(501, 10, "cosine"), (64, 100, "l2"), (3, 5, "ip"), (1, 4, "l2")])
def test_two_rank_sharded_search_matches_unsharded_oracle(n, k, metric):
    world, d = 2, 16
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, d, k, metric, out)) for r in range(world)]
    for p in procs:
        p.start()
    md, mi, mc = out.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    rng = np.random.default_rng(42)
    db = rng.standard_normal((n, d)).astype(np.float32)
    if n > 5:
        db[5] = db[n - 3]
    qs = np.random.default_rng(999).standard_normal((3, d)).astype(np.float32)
    kk = min(k, n)
    assert md.shape == (3, kk) and (mc == kk).all()
    for qi, q in enumerate(qs):
        ref = O.distances_single(q, db, metric)
        # per-shard slices give the same per-row arithmetic as the unsharded call up to BLAS blocking: tolerance
        O.check_topk(ref, mi[qi], md[qi], k, rtol=1e-5, squared_near_zero=(metric == "l2"))
