"""Row-sharded search over the GPUs of one box: one process per GPU, ``torch.distributed`` (NCCL over NVLink).

The reference's only "sharded" code is ``search_chunked_parallel`` (parallel_search.py:313-368): row chunks ->
local top-k with ``global_idx = local + start`` (:353) -> ``_merge_top_k`` (:137-156).  Here chunks are GPUs:

    shard g owns rows [g*ceil(N/G), ...)            (contiguous, ids are global: id_base = first row)
    every rank runs the same fused local top-k        (no data-path collective during the scan)
    ONE all-gather of the packed (distance, id) lists  (Q*k*16 bytes per rank: latency bound, not bandwidth bound)
    every rank merges the G sorted lists               (fpv_merge_topk: same (distance, id) order -> the answer is
                                                        independent of the shard count)

The collective and the merge are injectable so that the host-side logic (bounds, padding, packing, gather
layout) is covered by world_size-2 ``gloo`` tests on CPU; the defaults are NCCL + the CUDA merge kernel.
"""
from __future__ import annotations

from typing import Callable, Optional, Tuple

import os

import torch
import torch.distributed as dist


def shard_bounds(n_total: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous row range of ``rank``: ceil(N/G) rows per shard, the last one may be short/empty."""
    per = (n_total + world - 1) // world
    lo = min(rank * per, n_total)
    return lo, min(lo + per, n_total)


def pack_candidates(d: torch.Tensor, i: torch.Tensor, k: int) -> torch.Tensor:
    """(dist [Q,kl] f32, idx [Q,kl] i64) -> [Q,k,2] int64 (id, distance bits), padded with (-1, +inf) to k columns."""
    q, kl = d.shape
    out = torch.empty((q, k, 2), dtype=torch.int64, device=d.device)
    out[:, :, 0] = -1
    out[:, :, 1] = 0x7F800000                                   # +inf
    if kl:
        out[:, :kl, 0] = i
        out[:, :kl, 1] = d.contiguous().view(torch.int32).to(torch.int64)
    return out


def unpack_candidates(packed: torch.Tensor) -> Tuple[torch.Tensor, torch.Tensor]:
    """[S,Q,k,2] int64 -> (dist [S,Q,k] f32, idx [S,Q,k] i64)."""
    idx = packed[..., 0].contiguous()
    d = packed[..., 1].to(torch.int32).contiguous().view(torch.float32)
    return d, idx


def _default_merge(d: torch.Tensor, i: torch.Tensor, k_out: int):
    from . import ops
    return ops.merge_topk(d, i, k_out)


def _peer_exchange_enabled() -> bool:
    import os
    return os.environ.get("FPV_PEER_EXCHANGE", "1") != "0"


class PeerExchange:
    """All-gathers of the row-sharded search over NVLink peer memory (csrc/fpv_peer.cu) instead of NCCL calls.

    Every rank owns one IPC-exported device region: [4 flag arrays | counters | approx area x 2 | key area x 2].  ``put``
    stores a local tensor into slot ``rank`` of the chosen area of EVERY rank's region with P2P stores and publishes the
    epoch in the peers' flag words; the consumer kernels (phase-2 re-rank, merge) wait on the local flags themselves, so
    there is no collective launch and no host synchronisation on the data path.  The two areas of a kind alternate with
    the epoch parity (see the memory-model note in fpv_peer.cu).  Setup (handle exchange) is one all-gather; if IPC is
    unavailable on any rank every rank falls back to NCCL.
    """
    FLAG_SLOTS = 256

    def __init__(self, group, device, world: int, rank: int):
        self.group, self.device, self.world, self.rank = group, device, world, rank
        self.cap_approx = self.cap_keys = 0
        self.base = None
        self.peers = []          # mapped addresses of every rank's region (ints)
        self.peers_dev = None    # the same as a device int64 tensor for the kernels
        self.epoch = {"a": 0, "k": 0, "s": 0}     # one counter per kind: a kind's two areas alternate with ITS epoch parity
        self.ok = False

    _KIND_INDEX = {"a": 0, "k": 1, "s": 2}

    # layout -----------------------------------------------------------------------------------------------------
    def _offsets(self):
        flags = 6 * self.FLAG_SLOTS * 4                     # six flag arrays: (approx, keys, sample) x parity
        counters = 256
        a = _round_up(self.world * self.cap_approx, 256)
        k = _round_up(self.world * self.cap_keys, 256)
        off_a0 = flags + counters
        off_s0 = off_a0 + 2 * a + 2 * k                     # the sample areas have the size of the approx areas
        return dict(flags=0, counter=flags, a=(off_a0, off_a0 + a), k=(off_a0 + 2 * a, off_a0 + 2 * a + k),
                    s=(off_s0, off_s0 + a), total=off_s0 + 2 * a)

    def ensure(self, approx_bytes: int, key_bytes: int) -> bool:
        """Collective: (re)allocate the regions when a larger slot is needed.  Returns False (on every rank) when peer
        memory cannot be used."""
        import ctypes as C
        from . import _native as N
        if self.base is not None and approx_bytes <= self.cap_approx and key_bytes <= self.cap_keys:
            return self.ok
        self.close()
        self.cap_approx = _round_up(max(approx_bytes, 1 << 16), 256)
        self.cap_keys = _round_up(max(key_bytes, 1 << 17), 256)
        off = self._offsets()
        L = N.lib()
        handle = (C.c_ubyte * 64)()
        ptr = C.c_void_p()
        ok = 1
        with torch.cuda.device(self.device):
            if L.fpv_peer_alloc(off["total"], C.byref(ptr), handle) != 0:
                ok = 0
            hbytes = torch.tensor(list(handle) + [ok], dtype=torch.uint8, device=self.device)
            allh = torch.empty((self.world, 65), dtype=torch.uint8, device=self.device)
            dist.all_gather_into_tensor(allh, hbytes, group=self.group)
            allh = allh.cpu().numpy()
            ok = int(allh[:, 64].min())
            self.base = ptr.value if ptr.value else None
            self.peers = []
            if ok:
                for r in range(self.world):
                    if r == self.rank:
                        self.peers.append(self.base)
                        continue
                    hp = (C.c_ubyte * 64)(*[int(x) for x in allh[r, :64]])
                    pp = C.c_void_p()
                    if L.fpv_peer_open(hp, C.byref(pp)) != 0:
                        ok = 0
                        break
                    self.peers.append(pp.value)
            flag = torch.tensor([ok], dtype=torch.int32, device=self.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.group)       # every rank takes the same route
            self.ok = bool(flag.item())
            if self.ok:
                self.peers_dev = torch.tensor(self.peers, dtype=torch.int64, device=self.device)
                self.epoch = {"a": 0, "k": 0, "s": 0}
            dist.barrier(group=self.group)
        return self.ok

    def close(self):
        import ctypes as C
        from . import _native as N
        if self.base is None:
            return
        L = N.lib()
        torch.cuda.synchronize(self.device)
        if dist.is_initialized():
            try:
                dist.barrier(group=self.group)          # nobody unmaps while a peer may still store into the region
            except Exception:
                pass
        for r, p in enumerate(self.peers):
            if r != self.rank and p:
                L.fpv_peer_close(C.c_void_p(p))
        L.fpv_peer_free(C.c_void_p(self.base))
        self.base, self.peers, self.peers_dev, self.ok = None, [], None, False

    # data path --------------------------------------------------------------------------------------------------
    def next_epoch(self, kind: str) -> int:
        self.epoch[kind] += 1
        return self.epoch[kind]

    def put(self, kind: str, src: torch.Tensor, epoch: int):
        """kind "a" (approx values), "k" (packed keys) or "s" (sample values): store ``src`` into slot ``rank`` of every
        rank's area."""
        import ctypes as C
        from . import _native as N
        off = self._offsets()
        par = epoch & 1
        nbytes = src.numel() * src.element_size()
        ki = self._KIND_INDEX[kind]
        flag_off = off["flags"] + (2 * ki + par) * self.FLAG_SLOTS * 4
        with N.guard(self.device):
            N.check(N.lib().fpv_peer_put(N.ptr(src), nbytes, N.ptr(self.peers_dev), self.world, self.rank, off[kind][par], nbytes, flag_off,
                                         epoch & 0xFFFFFFFF, C.c_void_p(self.base + off["counter"] + 64 * ki),
                                         N.stream_ptr()), "fpv_peer_put")

    def area(self, kind: str, epoch: int):
        """(address of this rank's gather area, address of its arrival flags) for ``kind`` at ``epoch``."""
        off = self._offsets()
        par = epoch & 1
        return self.base + off[kind][par], self.base + off["flags"] + (2 * self._KIND_INDEX[kind] + par) * self.FLAG_SLOTS * 4


def _round_up(v: int, a: int) -> int:
    return (v + a - 1) // a * a


class ShardedTopK:
    """Gather + merge of per-rank top-k lists.  ``local`` results must carry GLOBAL ids (id_base = shard start)."""

    def __init__(self, n_total: int, group=None, merge_fn: Optional[Callable] = None):
        self.n_total = int(n_total)
        self.group = group
        self.world = dist.get_world_size(group) if dist.is_initialized() else 1
        self.rank = dist.get_rank(group) if dist.is_initialized() else 0
        self.lo, self.hi = shard_bounds(self.n_total, self.world, self.rank)
        self._custom_merge = merge_fn
        self.merge_fn = merge_fn or _default_merge
        self._gather_buf = None
        self._bases = None
        self.peer: Optional[PeerExchange] = None       # set by peer_setup(): exchanges by NVLink peer stores

    def gather(self, packed: torch.Tensor) -> torch.Tensor:
        """[Q,k,2] per rank -> [world,Q,k,2] on every rank."""
        if self.world == 1:
            return packed.unsqueeze(0)
        shape = (self.world,) + tuple(packed.shape)
        buf = self._gather_buf
        if buf is None or buf.shape != shape or buf.device != packed.device:
            buf = torch.empty(shape, dtype=packed.dtype, device=packed.device)
            self._gather_buf = buf
        try:
            dist.all_gather_into_tensor(buf, packed.contiguous(), group=self.group)
        except (RuntimeError, NotImplementedError):          # backends without the flat variant
            parts = [buf[r] for r in range(self.world)]
            dist.all_gather(parts, packed.contiguous(), group=self.group)
        return buf

    def peer_setup(self, device, approx_bytes: int, key_bytes: int) -> bool:
        """Collective, call it on every rank: prepare (or grow) the peer-memory exchange for slots of these sizes.  Returns
        whether peer memory is in use; False (NCCL all-gathers) when disabled by FPV_PEER_EXCHANGE=0, on a non-NCCL
        group or when CUDA IPC is unavailable on any rank."""
        if self.world == 1 or not _peer_exchange_enabled() or not dist.is_initialized() or dist.get_backend(self.group) != "nccl":
            return False
        if self.peer is None:
            self.peer = PeerExchange(self.group, torch.device(device), self.world, self.rank)
        return self.peer.ensure(approx_bytes, key_bytes)

    def merge(self, d_local: torch.Tensor, i_local: torch.Tensor, k: int):
        """Local (dist, global idx) lists [Q, <=k] -> merged (dist [Q,k'], idx [Q,k'], count [Q]), k' = min(k, N)."""
        k_out = min(int(k), self.n_total)
        if self._custom_merge is None and d_local.is_cuda:
            # device path: one pack kernel (8-byte wire keys), ONE all-gather, one merge kernel
            from . import ops
            if self._bases is None or self._bases.device != d_local.device:
                self._bases = torch.tensor([shard_bounds(self.n_total, self.world, r)[0] for r in range(self.world)],
                                           dtype=torch.int64, device=d_local.device)
            keys = ops.pack_topk(d_local, i_local, k_out, self.lo)
            nbytes = keys.numel() * 8
            if self.world > 1 and (self.peer is None or (self.peer.ok and nbytes > self.peer.cap_keys)):
                # collective (every rank merges the same shapes): size the peer-memory key areas once per growth
                self.peer_setup(d_local.device, self.peer.cap_approx if self.peer is not None else 0, nbytes)
            if self.peer is not None and self.peer.ok and nbytes % 16 == 0 and nbytes <= self.peer.cap_keys and self.world <= 256:
                # the packed lists go straight into every peer's gather area; the merge kernel waits for their arrival
                ep = self.peer.next_epoch("k")
                self.peer.put("k", keys, ep)
                addr, flags = self.peer.area("k", ep)
                return ops.merge_packed_peer(addr, flags, ep, self.world, keys.shape[0], keys.shape[1], self._bases, k_out)
            return ops.merge_packed(self.gather(keys), self._bases, k_out)
        gathered = self.gather(pack_candidates(d_local, i_local, k_out))
        d, i = unpack_candidates(gathered)
        return self.merge_fn(d, i, k_out)


class ShardedCodeSearch:
    """Quantized scans over row-sharded codes (BASELINE configs[3], [4]): every rank holds the codes of rows [lo, hi)
    (and, for PQ, the slice of the row bitmask), runs the fused scan + top-k kernel with ``id_base = lo`` and the
    lists are merged exactly like the float path (one all-gather of 8-byte keys + merge kernel).

    ``kind``: "hamming" (codes [n_local, nbytes] uint8, queries = packed bits [Q, nbytes]),
              "pq" (codes [n_local, M] uint8, queries = ADC tables [Q, M, Kc] fp32 from ``ops.pq_build_lut``),
              "sq" (uint8 scalar codes [n_local, D], queries = re-quantised query codes [Q, D] uint8; L2;
                    ``sq_params = (min_vals, scale)`` device tensors; shards of >= 65536 rows run on the int8 tensor cores).
    """

    def __init__(self, kind: str, local_codes: torch.Tensor, n_total: int, group=None, dims: int = 0, sq_params=None):
        from . import ops
        if kind not in ("hamming", "pq", "sq"):
            raise ValueError(f"unknown kind {kind!r}")
        if kind == "sq" and sq_params is None:
            raise ValueError("kind 'sq' needs sq_params = (min_vals, scale)")
        self.kind, self.dims = kind, int(dims)
        self.sq_params = sq_params
        self._row_term = None
        self.topk = ShardedTopK(n_total, group)
        if local_codes.shape[0] != self.topk.hi - self.topk.lo:
            raise ValueError(f"rank {self.topk.rank} holds {local_codes.shape[0]} rows, "
                             f"expected {self.topk.hi - self.topk.lo}")
        self.codes = local_codes.contiguous()
        self._packed = None
        if kind == "pq" and self.codes.shape[0]:
            self._packed = ops.pq_pack(self.codes)      # bank-conflict-free layout, built once per shard

    def search_tensors(self, queries: torch.Tensor, k: int = 10, local_mask_words: Optional[torch.Tensor] = None):
        """Every rank passes the same queries (and its own slice of the mask); returns merged (dist, idx, count)."""
        from . import ops
        n_local = self.codes.shape[0]
        k_local = min(int(k), n_local)
        nq = queries.shape[0]
        if k_local == 0:
            d = torch.empty((nq, 0), dtype=torch.float32, device=self.codes.device)
            i = torch.empty((nq, 0), dtype=torch.int64, device=self.codes.device)
        elif self.kind == "hamming":
            d, i, _c, _ = ops.hamming(queries, self.codes, k_local, self.dims, local_mask_words, self.topk.lo)
        elif self.kind == "sq":
            from . import _native as N
            mn, sc = self.sq_params
            if ops.sq_mma_supported(n_local, self.codes.shape[1], k_local):
                if self._row_term is None:
                    self._row_term = ops.sq_row_term(self.codes, sc)           # once per shard (index build)
                d, i, _c = ops.sq_l2_mma(queries, self.codes, mn, sc, self._row_term[0], self._row_term[1], k_local,
                                         local_mask_words, self.topk.lo)
            else:
                d, i, _c, _ = ops.sq_scan(N.SQ_L2, queries, self.codes, mn, sc, k_local, local_mask_words, self.topk.lo)
        else:
            m, kc = queries.shape[1], queries.shape[2]
            if self._packed is not None and ops.pq_adc_packed_supported(nq, n_local, m, kc, k_local):
                d, i, _c = ops.pq_adc_packed(queries, self._packed, k_local, local_mask_words, self.topk.lo)
            else:
                d, i, _c, _ = ops.pq_adc(queries, self.codes, k_local, local_mask_words, self.topk.lo)
        return self.topk.merge(d, i, k)


class ShardedSearchEngine:
    """Exact float search over a row-sharded database (BASELINE configs[2]): each rank holds rows [lo, hi).

    Batches take the TWO-PHASE tensor-core path (csrc/fpv_gemm_topk.cu, fpv_gemm_*_sharded_f32):

        phase 1   every rank filters its own rows on the tensor cores and keeps its candidates
        exchange  all-gather of the k best APPROXIMATE values per query and shard      (Q*k*4 bytes per rank)
        phase 2   every rank selects the k-th best approximate value of the WHOLE job, and re-ranks in exact fp32
                  only its own rows below (that + 2E): the row gather of the job is paid once, split over the ranks
        exchange  all-gather of the packed exact (distance, row) lists                 (Q*k*8 bytes per rank)
        merge     fpv_merge_packed on every rank

    The error bound E must hold on every shard, so the row-norm maximum and the measured bf16 residuals are
    all-reduced (MAX) once.  Without the first exchange every rank would re-rank a full local window (measured
    r1: 0.39 ms per 4096 queries on EVERY rank, the same as one GPU pays for the whole database).
    Everything else (fewer than GEMM_MIN_BATCH queries, k > 256, a shard below 4096 rows) takes the one-phase route:
    local fused top-k, one all-gather, merge.
    """

    def __init__(self, local_rows, n_total: int, group=None, device=None, engine=None):
        from .engine import GpuIndex, ParallelSearchEngine
        self.topk = ShardedTopK(n_total, group)
        self.engine = engine or ParallelSearchEngine(device=device)
        self.index = local_rows if isinstance(local_rows, GpuIndex) else GpuIndex(local_rows, self.engine.device,
                                                                                   id_base=self.topk.lo)
        if self.index.n != self.topk.hi - self.topk.lo:
            raise ValueError(f"rank {self.topk.rank} holds {self.index.n} rows, expected {self.topk.hi - self.topk.lo}")
        self.index.id_base = self.topk.lo
        self._modes = {}                 # mode -> True once the shard bounds have been made global
        self._bf16_everywhere = None
        self._approx_buf = None
        self._sample_buf = None
        self._peer_sized = set()
        self.sample_exchange = os.environ.get("FPV_SAMPLE_EXCHANGE", "1") != "0"   # False: every shard keeps its own sample
        self.two_phase = True            # set False to force the one-phase route (A/B measurements)

    # ---- decisions every rank must take identically (functions of global quantities only)
    def _min_shard_rows(self) -> int:
        t = self.topk
        return min(hi - lo for lo, hi in (shard_bounds(t.n_total, t.world, r) for r in range(t.world)))

    def _two_phase_ok(self, nq: int, k: int) -> bool:
        from . import engine_gemm
        t = self.topk
        return (self.two_phase and t.world > 1 and nq >= self.engine.GEMM_MIN_BATCH and k <= engine_gemm.MAX_K
                and t.world * k <= 4096 and self._min_shard_rows() >= max(4096, k) and self.index.d % 4 == 0
                and self.index.d >= 16)

    def _sample_exchange_ok(self) -> bool:
        """Pool the shards' samples (one more small exchange)?  Needs a sampling slab on EVERY shard: more than 2 x 128
        tiles of 256 rows (csrc/fpv_gemm_topk.cu: `sampling`)."""
        return self.sample_exchange and self.topk.world > 1 and self._min_shard_rows() >= 70000

    def _mode(self, k: int, nq: int) -> str:
        """tensor-core operand format, identical on every rank (the local choice depends on free memory)"""
        from . import engine_gemm
        local = engine_gemm._effective_mode(None, self.index, k, nq)
        if self._bf16_everywhere is None:
            local_any = engine_gemm._effective_mode(None, self.index, 1, max(nq, engine_gemm.BF16_MIN_BATCH))
            flag = torch.tensor([1 if local_any == "bf16" else 0], dtype=torch.int32, device=self.index.device)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self.topk.group)
            self._bf16_everywhere = bool(flag.item())
        mode = local if self._bf16_everywhere else "tf32"
        if mode == "bf16" and (k > 128 or self.index.d % 8 != 0):
            mode = "tf32"
        if mode not in self._modes:
            from . import engine_gemm as eg
            vmax, err = eg.sharded_bounds(self.index, mode)
            b = torch.tensor([vmax, err[0], err[1]], dtype=torch.float64, device=self.index.device)
            dist.all_reduce(b, op=dist.ReduceOp.MAX, group=self.topk.group)
            vmax, e0, e1 = (float(x) for x in b.tolist())
            eg.set_sharded_bounds(self.index, vmax, (e0, e1))
            self._modes[mode] = True
        return mode

    def search_tensors(self, queries, k: int = 10, metric: str = "cosine", local_mask_words: Optional[torch.Tensor] = None):
        """Every rank passes the same queries (and its own slice of the row filter, packed by ``ops.pack_mask``);
        every rank gets the same merged (dist, idx, count)."""
        k = int(k)
        if isinstance(queries, torch.Tensor) and queries.is_cuda:
            nq = 1 if queries.ndim == 1 else queries.shape[0]
        else:
            nq = 1 if getattr(queries, "ndim", 2) == 1 else len(queries)
        k_out = min(k, self.topk.n_total)
        if k_out >= 1 and self._two_phase_ok(nq, k_out):
            return self._search_two_phase(queries, k_out, metric, local_mask_words)
        k_local = min(k, self.index.n)
        if k_local > 0:
            if local_mask_words is not None:
                raise ValueError("a row filter needs the two-phase (batched) route: pass at least GEMM_MIN_BATCH queries")
            d, i, _c = self.engine.search_tensors(queries, self.index, k_local, metric)
        else:
            d = torch.empty((nq, 0), dtype=torch.float32, device=self.index.device)
            i = torch.empty((nq, 0), dtype=torch.int64, device=self.index.device)
        return self.topk.merge(d, i, k)

    def _search_two_phase(self, queries, k: int, metric: str, mask_words):
        from . import _native as N
        from . import engine_gemm as eg
        from . import ops
        index, t = self.index, self.topk
        q = self.engine._queries_to_device(queries, index.d)
        if q.shape[1] != index.d:
            raise ValueError(f"dimension mismatch: query has {q.shape[1]}, database has {index.d}")
        mode = self._mode(k, q.shape[0])
        nq = q.shape[0]
        a_bytes, k_bytes = nq * k * 4, nq * min(k, t.n_total) * 8
        if (nq, k) not in self._peer_sized:                     # collective, once per shape: size the peer-memory areas
            self._peer_sized.add((nq, k))
            t.peer_setup(index.device, a_bytes, k_bytes)
        use_peer = t.peer is not None and t.peer.ok and a_bytes % 16 == 0 and a_bytes <= t.peer.cap_approx
        with N.guard(index.device):          # nothing else may touch this stream's workspace between the phases
            if self._sample_exchange_ok():
                # the shards pool their samples: the k-th best group value of the WHOLE job is everyone's first threshold
                sample = eg.sample_sharded(q, index, k, metric, mode, mask_words)
                if use_peer:
                    ep = t.peer.next_epoch("s")
                    t.peer.put("s", sample, ep)
                    addr, flags = t.peer.area("s", ep)
                    approx = eg.slabs_sharded(q, index, k, metric, mode, addr, t.world, mask_words, flags, ep)
                else:
                    shape = (t.world,) + tuple(sample.shape)
                    sbuf = self._sample_buf
                    if sbuf is None or sbuf.shape != shape:
                        sbuf = self._sample_buf = torch.empty(shape, dtype=torch.int32, device=index.device)
                    dist.all_gather_into_tensor(sbuf, sample, group=t.group)
                    approx = eg.slabs_sharded(q, index, k, metric, mode, sbuf, t.world, mask_words)
            else:
                approx = eg.filter_sharded(q, index, k, metric, mode, mask_words)
            if use_peer:
                # exchange 1 over NVLink peer stores; the phase-2 kernel itself waits for the peers' values
                ep = t.peer.next_epoch("a")
                t.peer.put("a", approx, ep)
                addr, flags = t.peer.area("a", ep)
                lowp = index._lowp if mode == "bf16" else None
                d, i, _c = ops.gemm_finish_sharded_peer(q, index.rows, k, metric, index.row_sq, addr, flags, ep, t.world, lowp,
                                                        index.id_base, mask_words)
                return t.merge(d, i, k)
            shape = (t.world,) + tuple(approx.shape)
            buf = self._approx_buf
            if buf is None or buf.shape != shape:
                buf = self._approx_buf = torch.empty(shape, dtype=torch.int32, device=index.device)
            dist.all_gather_into_tensor(buf, approx, group=t.group)
            d, i, _c = eg.finish_sharded(q, index, k, metric, mode, buf, mask_words)
        return t.merge(d, i, k)
