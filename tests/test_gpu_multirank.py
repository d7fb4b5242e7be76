"""Real multi-rank check (needs >= 2 GPUs; skipped otherwise): two NCCL ranks, one GPU each, run the row-sharded float
search (two-phase tensor-core route and the one-phase route), the sharded Hamming scan and the sharded PQ ADC scan
with a row bitmask; every rank must hold an answer that is bit-identical to the single-GPU answer over the whole
database.  This is the device branch of ShardedTopK.merge (pack_topk -> all_gather_into_tensor -> merge_packed) and
the approx exchange of ShardedSearchEngine, exercised with real collectives (missing #4)."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import fastpyvectordb_b200 as fpv
        from fastpyvectordb_b200 import ops
        from fastpyvectordb_b200.sharded import ShardedCodeSearch, ShardedSearchEngine, shard_bounds
        problems = []
        n, d = 40000, 128
        rng = np.random.default_rng(42)
        db = rng.standard_normal((n, d)).astype(np.float32)
        db[17] = db[n - 5]                                            # exact tie across the shards
        db[30000:30050] *= 2.5                                        # largest row norms only in the second shard
        lo, hi = shard_bounds(n, world, rank)
        eng = fpv.ParallelSearchEngine(device=dev)
        sh = ShardedSearchEngine(fpv.GpuIndex(db[lo:hi], dev, id_base=lo), n, engine=eng)
        whole = fpv.GpuIndex(db, dev)
        mask = torch.from_numpy(rng.random(n) < 0.3).to(dev)
        for q, k, metric in [(200, 100, "l2"), (64, 10, "cosine"), (5, 20, "ip"), (2, 7, "l2")]:
            qs = torch.from_numpy(np.random.default_rng(999 + q).standard_normal((q, d)).astype(np.float32)).to(dev)
            wd, wi, _ = eng.search_tensors(qs, whole, k, metric)
            for two_phase in (True, False):
                sh.two_phase = two_phase
                md, mi, mc = sh.search_tensors(qs, k, metric)
                if not (torch.equal(mi, wi) and torch.equal(md, wd) and bool((mc == k).all())):
                    problems.append(("float", q, k, metric, two_phase))
            sh.two_phase = True
            if q >= eng.GEMM_MIN_BATCH:                               # row filter, sharded with the rows
                wd, wi, _ = eng.search_tensors(qs, whole, k, metric, filter_mask=mask)
                md, mi, _ = sh.search_tensors(qs, k, metric, local_mask_words=ops.pack_mask(mask[lo:hi]))
                if not (torch.equal(mi, wi) and torch.equal(md, wd)):
                    problems.append(("float+mask", q, k, metric))
        # shards of >= 70K rows pool their SAMPLES (a third small exchange: the k-th best group value of the whole job is
        # every shard's first threshold); over peer memory and over NCCL, with and without it: same bits
        nb = 150_000
        dbb = np.random.default_rng(7).standard_normal((nb, 64)).astype(np.float32)
        dbb[5] = dbb[nb - 9]
        lob, hib = shard_bounds(nb, world, rank)
        wholeb = fpv.GpuIndex(dbb, dev)
        maskb = torch.from_numpy(np.random.default_rng(8).random(nb) < 0.3).to(dev)
        for route in ("peer", "nccl"):
            os.environ["FPV_PEER_EXCHANGE"] = "1" if route == "peer" else "0"
            shb = ShardedSearchEngine(fpv.GpuIndex(dbb[lob:hib], dev, id_base=lob), nb, engine=eng)
            for q, k, metric in [(300, 100, "l2"), (64, 10, "cosine")]:
                qs = torch.from_numpy(np.random.default_rng(77 + q).standard_normal((q, 64)).astype(np.float32)).to(dev)
                wd, wi, _ = eng.search_tensors(qs, wholeb, k, metric)
                for pooled in (True, False):
                    shb.sample_exchange = pooled
                    if pooled and not shb._sample_exchange_ok():
                        problems.append(("sample exchange not taken", route))
                    md, mi, mc = shb.search_tensors(qs, k, metric)
                    if not (torch.equal(mi, wi) and torch.equal(md, wd) and bool((mc == k).all())):
                        problems.append(("float/pooled sample", route, q, k, metric, pooled))
                shb.sample_exchange = True
                wd, wi, _ = eng.search_tensors(qs, wholeb, k, metric, filter_mask=maskb)
                md, mi, _ = shb.search_tensors(qs, k, metric, local_mask_words=ops.pack_mask(maskb[lob:hib]))
                if not (torch.equal(mi, wi) and torch.equal(md, wd)):
                    problems.append(("float/pooled sample + mask", route, q, k, metric))
        os.environ["FPV_PEER_EXCHANGE"] = "1"
        del wholeb, shb
        # Hamming codes
        codes = torch.from_numpy(rng.integers(0, 256, (n, 64), dtype=np.uint8)).to(dev)
        qb = torch.from_numpy(rng.integers(0, 256, (3, 64), dtype=np.uint8)).to(dev)
        wd, wi, _, _ = ops.hamming(qb, codes, 50, 512)
        md, mi, mc = ShardedCodeSearch("hamming", codes[lo:hi].contiguous(), n, dims=512).search_tensors(qb, 50)
        if not (torch.equal(mi, wi) and torch.equal(md, wd) and bool((mc == 50).all())):
            problems.append(("hamming",))
        # PQ ADC with the bitmask sharded with the rows
        pcodes = torch.from_numpy(rng.integers(0, 256, (n, 48), dtype=np.uint8)).to(dev)
        lut = torch.from_numpy(rng.random((2, 48, 256)).astype(np.float32)).to(dev)
        wd, wi, _, _ = ops.pq_adc(lut, pcodes, 20, ops.pack_mask(mask))
        md, mi, mc = ShardedCodeSearch("pq", pcodes[lo:hi].contiguous(), n).search_tensors(lut, 20, ops.pack_mask(mask[lo:hi]))
        same_ids = torch.equal(mi, wi)
        close = torch.allclose(md, wd, rtol=1e-6, atol=0)            # the packed layout sums the subspaces in another order
        if not (same_ids and close and bool((mc == 20).all())):
            problems.append(("pq", same_ids, close))
        # every rank holds the same merged answer
        chk = torch.stack([mi.sum(), md.double().sum().long()])
        both = [torch.empty_like(chk) for _ in range(world)]
        dist.all_gather(both, chk)
        if not all(torch.equal(b, both[0]) for b in both):
            problems.append(("ranks disagree",))
        torch.cuda.synchronize()
        out.put((rank, problems))
    except Exception as exc:                                          # pragma: no cover
        import traceback
        out.put((rank, [("exception", repr(exc), traceback.format_exc()[-1500:])]))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs two GPUs")
def test_two_nccl_ranks_match_the_single_gpu_answer():
    import torch.multiprocessing as mp
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    results = [out.get(timeout=600) for _ in range(world)]
    for p in procs:
        p.join(timeout=120)
    for rank, problems in results:
        assert not problems, (rank, problems)
    assert all(p.exitcode == 0 for p in procs)
