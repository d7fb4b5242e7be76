"""ctypes binding of ``libfpv_b200.so`` (the C-ABI declared in ``include/fpv_b200.h``).

There is no CPU fallback: if the library is missing or a call fails this module raises.  Device memory is
owned by torch; this layer only passes raw device pointers, sizes and the current CUDA stream.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libfpv_b200.so")

METRIC_COSINE, METRIC_L2, METRIC_IP, METRIC_L2_DIFF = 0, 1, 2, 3
SQ_L2, SQ_DOT, SQ_COSINE = 0, 1, 2
MAX_K = 1024

_p = C.c_void_p
_i64 = C.c_int64
_i = C.c_int
_sz = C.c_size_t

# name -> (restype, argtypes); must list every symbol of include/fpv_b200.h (tests/test_abi.py checks)
SIGNATURES = {
    "fpv_abi_version": (_i, []),
    "fpv_last_error": (C.c_char_p, []),
    "fpv_launch_count": (C.c_longlong, []),
    "fpv_row_sqnorm_f32": (_i, [_p, _i64, _i, _i64, _p, _p]),
    "fpv_scan_f32_workspace": (_sz, [_i64, _i64, _i, _i]),
    "fpv_scan_f32_topk": (_i, [_p, _i64, _p, _i64, _i, _i64, _i, _i, _p, _p, _i64, _p, _p, _p, _p, _sz, _p]),
    "fpv_gemm_topk_workspace": (_sz, [_i64, _i64, _i, _i, _i]),
    "fpv_gemm_topk_f32": (_i, [_p, _i64, _p, _p, _i64, _i, _i, _i, _i, _p, _p, C.c_float, C.c_float, C.c_float, _p, _i64, _p, _p, _p,
                          _p, _sz, _p]),
    "fpv_gemm_filter_sharded_f32": (_i, [_p, _i64, _p, _p, _i64, _i, _i, _i, _i, _p, _p, C.c_float, C.c_float, C.c_float, _p, _p,
                                    _p, _sz, _p]),
    "fpv_gemm_sample_sharded_f32": (_i, [_p, _i64, _p, _p, _i64, _i, _i, _i, _i, _p, _p, C.c_float, C.c_float, C.c_float, _p, _p,
                                    _p, _sz, _p]),
    "fpv_gemm_slabs_sharded_f32": (_i, [_p, _i64, _p, _p, _i64, _i, _i, _i, _i, _p, _p, C.c_float, C.c_float, C.c_float, _p, _p,
                                   _i, _p, C.c_uint32, _p, _p, _sz, _p]),
    "fpv_gemm_finish_sharded_f32": (_i, [_p, _i64, _p, _p, _i64, _i, _i, _i, _i, _p, _p, _i64, _p, _i, _p, _p, _p, _p, _sz, _p]),
    "fpv_to_bf16": (_i, [_p, _p, _i64, _p]),
    "fpv_gemm_topk_flags_offset": (_sz, [_i64, _i64, _i, _i, _i]),
    "fpv_gemm_profile": (_i, [_i]),
    "fpv_gemm_profile_read": (_i, [_p, _p]),
    "fpv_distances_f32": (_i, [_p, _i64, _p, _i64, _i, _i64, _i, _p, _p, _p, _sz, _p]),
    "fpv_rerank_f32": (_i, [_p, _i64, _p, _i64, _i, _i64, _i, _p, _i, _i, _p, _i64, _p, _p, _p, _p]),
    "fpv_merge_topk": (_i, [_p, _p, _i, _i64, _i, _i, _p, _p, _p, _p]),
    "fpv_pack_topk": (_i, [_p, _p, _i64, _i, _i, _i64, _p, _p]),
    "fpv_merge_packed": (_i, [_p, _p, _i, _i64, _i, _i, _p, _p, _p, _p]),
    "fpv_peer_alloc": (_i, [_sz, _p, _p]),
    "fpv_peer_open": (_i, [_p, _p]),
    "fpv_peer_close": (_i, [_p]),
    "fpv_peer_free": (_i, [_p]),
    "fpv_peer_put": (_i, [_p, _sz, _p, _i, _i, _sz, _sz, _sz, C.c_uint32, _p, _p]),
    "fpv_merge_packed_peer": (_i, [_p, _p, _i, _i64, _i, _i, _p, C.c_uint32, _p, _p, _p, _p]),
    "fpv_gemm_finish_sharded_peer_f32": (_i, [_p, _i64, _p, _p, _i64, _i, _i, _i, _i, _p, _p, _i64, _p, _i, _p, C.c_uint32, _p, _p, _p,
                                         _p, _sz, _p]),
    "fpv_bq_encode": (_i, [_p, _i64, _i, _i64, _p, _p, _p]),
    "fpv_hamming_workspace": (_sz, [_i64, _i64, _i, _i]),
    "fpv_hamming_topk": (_i, [_p, _i64, _p, _i64, _i, _i, _i, _p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "fpv_hamming_mma_supported": (_i, [_i64, _i64, _i, _i]),
    "fpv_hamming_mma_workspace": (_sz, [_i64, _i64, _i, _i]),
    "fpv_hamming_mma_topk": (_i, [_p, _i64, _p, _i64, _i, _i, _i, _p, _i64, _p, _p, _p, _p, _sz, _p]),
    "fpv_hamming_mma_dots": (_i, [_p, _i64, _p, _i64, _i, _i, _p, _p, _sz, _p]),
    "fpv_pq_encode": (_i, [_p, _i64, _i, _i64, _p, _i, _i, _p, _p]),
    "fpv_pq_build_lut": (_i, [_p, _i, _i, _i, _p, _i64, _p, _p]),
    "fpv_pq_adc_workspace": (_sz, [_i64, _i64, _i, _i, _i]),
    "fpv_pq_adc_topk": (_i, [_p, _i64, _p, _i64, _i, _i, _i, _p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "fpv_pq_pack": (_i, [_p, _i64, _i, _p, _p]),
    "fpv_pq_adc_packed_workspace": (_sz, [_i64, _i64, _i, _i, _i]),
    "fpv_pq_adc_packed_topk": (_i, [_p, _i64, _p, _i64, _i, _i, _i, _p, _i64, _p, _p, _p, _p, _sz, _p]),
    "fpv_sq_encode": (_i, [_p, _i64, _i, _i64, _p, _p, _p, _p]),
    "fpv_sq_workspace": (_sz, [_i64, _i64, _i, _i]),
    "fpv_sq_topk": (_i, [_i, _p, _i64, _p, _i64, _i, _p, _p, _i, _p, _i64, _p, _p, _p, _p, _p, _sz, _p]),
    "fpv_sq_row_term": (_i, [_p, _i64, _i, _p, _p, _p, _p]),
    "fpv_sq_mma_supported": (_i, [_i64, _i, _i]),
    "fpv_sq_mma_workspace": (_sz, [_i64, _i64, _i, _i]),
    "fpv_sq_l2_mma_topk": (_i, [_p, _i64, _p, _i64, _i, _p, _p, _p, _p, _i, _p, _i64, _p, _p, _p, _p, _sz, _p]),
    "fpv_sq_row_terms_dc": (_i, [_p, _i64, _i, _p, _p, _p, _p, _p, _p]),
    "fpv_sq_dc_mma_topk": (_i, [_i, _p, _i64, _p, _i64, _i, _p, _p, _p, _p, _p, _i, _p, _i64, _p, _p, _p, _p, _sz, _p]),
    "fpv_sq_mma_flags_offset": (_sz, [_i64, _i64, _i, _i]),
    "fpv_sq_mma_limb_dots": (_i, [_p, _p, _i64, _i, _p, _p, _p, _p, _p, _sz, _p]),
}

_lib = None
_lock = threading.Lock()


class NativeError(RuntimeError):
    pass


def lib():
    """Load the CUDA library (once).  Raises if it has not been built — there is no other code path."""
    global _lib
    if _lib is None:
        with _lock:
            if _lib is None:
                if not os.path.exists(LIB_PATH):
                    raise NativeError(
                        f"{LIB_PATH} is missing: build it with `python -m fastpyvectordb_b200.build` "
                        "(nvcc, sm_100a). fastpyvectordb_b200 has no CPU fallback.")
                handle = C.CDLL(LIB_PATH)
                for name, (res, args) in SIGNATURES.items():
                    fn = getattr(handle, name)
                    fn.restype, fn.argtypes = res, args
                if handle.fpv_abi_version() != 1:
                    raise NativeError("libfpv_b200.so ABI version mismatch; rebuild")
                _lib = handle
    return _lib


def check(rc: int, what: str):
    if rc != 0:
        msg = lib().fpv_last_error().decode(errors="replace")
        if rc == 1:
            raise ValueError(f"{what}: {msg}")
        raise NativeError(f"{what} failed (code {rc}): {msg}")


def ptr(t):
    """Device pointer of a torch CUDA tensor (None -> NULL)."""
    if t is None:
        return None
    if not t.is_cuda:
        raise NativeError("internal error: host tensor passed to a device entry point")
    return C.c_void_p(t.data_ptr())


def stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise NativeError("fastpyvectordb_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    lib()
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise NativeError(f"device {dev} is not a CUDA device")
    if dev.index is None:
        dev = torch.device("cuda", torch.cuda.current_device())
    return dev


class Workspace:
    """Grow-only scratch buffer per (device, stream) — the C-ABI never allocates.  Kernels of one stream run in order,
    so consecutive calls on a stream may share one buffer; calls on different streams get different buffers."""

    def __init__(self):
        self._buf = {}

    def get(self, device: torch.device, nbytes: int) -> torch.Tensor:
        nbytes = max(int(nbytes), 256)
        key = (device, torch.cuda.current_stream(device).cuda_stream)
        buf = self._buf.get(key)
        if buf is None or buf.numel() < nbytes:
            buf = torch.empty(int(nbytes * 1.25) + 4096, dtype=torch.uint8, device=device)
            self._buf[key] = buf
        return buf


workspace = Workspace()

_device_locks = {}


def device_lock(device) -> "threading.RLock":
    """One re-entrant lock per device.  An operation is several kernel launches that share the scratch buffer; the
    launches of two host threads must not interleave (ctypes releases the GIL), so every ops.* call enqueues its whole
    launch sequence under this lock.  The GPU work itself is asynchronous: the lock is held for microseconds."""
    dev = torch.device(device)
    lk = _device_locks.get(dev)
    if lk is None:
        with _lock:
            lk = _device_locks.setdefault(dev, threading.RLock())
    return lk


class guard:
    """``with guard(device):`` = torch.cuda.device(device) + the device's enqueue lock."""

    def __init__(self, device):
        self._dev = torch.device(device)
        self._ctx = torch.cuda.device(self._dev)
        self._lk = device_lock(self._dev)

    def __enter__(self):
        self._lk.acquire()
        try:
            self._ctx.__enter__()
        except BaseException:
            self._lk.release()
            raise
        return self

    def __exit__(self, *exc):
        try:
            return self._ctx.__exit__(*exc)
        finally:
            self._lk.release()
