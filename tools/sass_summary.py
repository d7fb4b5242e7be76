#!/usr/bin/env python
"""Instruction-mix summary of the key kernels from `cuobjdump -sass` of the in-tree objects (no GPU needed):
the tcgen05 / TMEM / TMA mnemonics (UTCHMMA = f16/bf16/tf32 MMA, UTCIMMA = int8 MMA, LDTM / STTM = tensor-memory
load / store, UTMALDG = TMA tensor load, UTCBAR = tcgen05.commit) prove which hardware path a kernel uses.
    python tools/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OBJS = {"fpv_gemm_topk.o": ["gemm_filter_kernelILi1ELi1ELi2E", "gemm_filter_kernelILi0ELi0ELi1E", "gemm_finish2_kernel"],
        "fpv_sq_mma.o": ["sq_mma_kernelILi0E", "sq_mma_kernelILi2E", "sq_mma_finish_kernel", "sq_mma_finish_dc_kernelILi2E"],
        "fpv_hamming_mma.o": ["ham_mma_kernel"],
        "fpv_pq.o": ["pq_adc_filter_kernelILi3ELb0E", "pq_adc_quad_kernelILi3ELb0E", "pq_sample_min_kernelILi3ELb0E", "pq_adc_rot_kernelILi3ELb0E"],
        "fpv_sq.o": ["sq_l2_tma_kernelILi2ELb0E", "sq_scan_kernelILi1ELb1E"],
        "fpv_hamming.o": ["hamming_fast_kernelILi8ELi1E"], "fpv_scan_f32.o": ["scan_f32_kernelILi1ELb1E"]}
MNEM = ["UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCBAR", "UTMALDG", "UBLKCP", "LDTM", "STTM", "SYNCS", "LDS", "STS", "PRMT", "POPC",
        "VABSDIFF4", "FADD", "FFMA", "IADD3", "LOP3", "SHF", "ATOMG", "ATOMS", "LDG", "STG"]

print("# SASS instruction mix (cuobjdump -sass of fastpyvectordb_b200/build/*.o, sm_100a); regenerate: python tools/sass_summary.py\n")
for obj, kernels in OBJS.items():
    txt = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "fastpyvectordb_b200", "build", obj)], capture_output=True, text=True).stdout
    for f in re.split(r"\n\s*Function : ", txt)[1:]:
        name = f.split("\n", 1)[0].strip()
        if not any(k in name for k in kernels):
            continue
        c, n = collections.Counter(), 0
        for ln in f.splitlines():
            m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", ln)
            if m:
                n += 1
                c[m.group(1).split(".")[0]] += 1
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        print(f"## {obj}: {dem[:150]}")
        print(f"instructions: {n}   " + "  ".join(f"{k}={c[k]}" for k in MNEM if c[k]) + "\n")
