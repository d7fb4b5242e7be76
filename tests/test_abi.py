"""CPU-side checks of the C-ABI boundary: the library loads, exports exactly what include/fpv_b200.h declares,
the ctypes table covers it, and argument validation works without touching a GPU."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "fpv_b200.h")


@pytest.fixture(scope="module")
def built_lib():
    from fastpyvectordb_b200 import build
    return build.build()


def _declared():
    text = open(HEADER).read()
    return sorted(set(re.findall(r"FPV_API[^;]*?\b(fpv_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_symbols():
    names = _declared()
    assert "fpv_scan_f32_topk" in names and "fpv_hamming_topk" in names and "fpv_pq_adc_topk" in names
    assert len(names) >= 18


def test_library_exports_every_declared_symbol(built_lib):
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\bT (fpv_[a-z0-9_]+)", out))
    declared = set(_declared())
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert exported <= declared, f"exported but not declared in the header: {sorted(exported - declared)}"


def test_ctypes_table_matches_header(built_lib):
    from fastpyvectordb_b200 import _native
    assert set(_native.SIGNATURES) == set(_declared())
    lib = _native.lib()
    assert lib.fpv_abi_version() == 1
    # parameter counts in the table match the header prototypes
    text = open(HEADER).read()
    for name, (_res, args) in _native.SIGNATURES.items():
        proto = re.search(r"FPV_API[^;(]*\b%s\s*\(([^;]*?)\)\s*;" % name, text, re.S).group(1).strip()
        n_params = 0 if proto in ("void", "") else proto.count(",") + 1
        assert n_params == len(args), f"{name}: header has {n_params} parameters, ctypes table {len(args)}"


def test_argument_validation_needs_no_gpu(built_lib):
    from fastpyvectordb_b200 import _native
    lib = _native.lib()
    null = ctypes.c_void_p(None)
    # k out of range -> FPV_ERR_INVALID before anything is launched
    rc = lib.fpv_scan_f32_topk(null, 1, null, 10, 8, 8, 0, 5000, null, null, 0, null, null, null, null, 0, null)
    assert rc == 1 and b"k=" in lib.fpv_last_error()
    rc = lib.fpv_scan_f32_topk(null, 1, null, 10, 8, 8, 7, 5, null, null, 0, null, null, null, null, 0, null)
    assert rc == 1 and b"metric" in lib.fpv_last_error()
    rc = lib.fpv_pq_adc_topk(null, 1, null, 10, 8, 300, 5, null, 0, null, null, null, null, null, 0, null)
    assert rc == 1
    rc = lib.fpv_merge_topk(null, null, 0, 1, 1, 1, null, null, null, null)
    assert rc == 1
    assert lib.fpv_scan_f32_workspace(4, 1000, 64, 10) > 0
    assert lib.fpv_hamming_workspace(1, 1000, 16, 10) > 0


def test_product_path_fails_loudly_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    import numpy as np
    import fastpyvectordb_b200 as fpv
    from fastpyvectordb_b200._native import NativeError
    with pytest.raises(NativeError):
        fpv.ParallelSearchEngine()
    bq = fpv.BinaryQuantizer(dimensions=16)
    with pytest.raises(NativeError):
        bq.encode(np.zeros((2, 16), np.float32))


def test_product_code_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "fastpyvectordb_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "/root/reference" not in text, f"{f} reads the reference tree"
