"""Numerical check of the error bounds that certify the tensor-core dot / cosine scans of the uint8 scalar quantizer
(csrc/fpv_sq_mma.cu: sq_mma_prep_dc_kernel) and of the superset property of the four-query PQ filter
(csrc/fpv_pq.cu: pq_quad_table_kernel / pq_adc_quad_kernel).  The device arithmetic is restated here in NumPy (fp32 where
the device uses fp32, FMAs emulated through float64) and compared with float64 ground truth: the bound must cover the
observed |approximation - scan value| with room to spare, on shapes with benign and with hostile value ranges.
No GPU needed: this is the host-side proof obligation behind "the tensor cores only filter"."""
import numpy as np
import pytest

F = np.float32
U = 2.0 ** -24


def _fma32(a, b, c):
    return (np.asarray(a, np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(F)


def _consts(qcode, mn, sc, cosine):
    s255 = (sc / F(255.0)).astype(F)                                       # c0
    qr = ((qcode.astype(F) / F(255.0)).astype(F) * sc).astype(F) + mn     # decode, fp32 op by op
    qr = qr.astype(F)
    if cosine:
        qr = (qr * F(1.0 / (np.sqrt(np.sum(qr.astype(np.float64) ** 2)) + 1e-8))).astype(F)
    return s255, mn.astype(F), qr


@pytest.mark.parametrize("cosine", [False, True])
@pytest.mark.parametrize("case", ["benign", "wide_ranges", "offset_minima", "tiny_scales"])
def test_sq_dot_cosine_error_bound_covers_the_approximation(case, cosine):
    rng = np.random.default_rng(hash((case, cosine)) % 2 ** 32)
    n, d = 4000, 1024
    mn = (-0.2 - 0.3 * rng.random(d)).astype(F)
    sc = (0.4 + 0.5 * rng.random(d)).astype(F)
    if case == "wide_ranges":
        sc *= (10.0 ** rng.uniform(-2, 2, d)).astype(F)
    elif case == "offset_minima":
        mn = (50.0 + rng.random(d)).astype(F)                              # large C_q: heavy cancellation against c R_row
    elif case == "tiny_scales":
        sc = (1e-3 * rng.random(d) + 1e-5).astype(F)
    codes = rng.integers(0, 256, (n, d), dtype=np.uint8)
    codes[0] = 255
    codes[1] = 0 if not cosine else 1
    qcode = rng.integers(0, 256, d, dtype=np.uint8)
    s, m, w = _consts(qcode, mn, sc, cosine)
    s64, m64, w64 = s.astype(np.float64), m.astype(np.float64), w.astype(np.float64)
    b = codes.astype(np.float64)
    # ---- what sq_mma_prep_dc_kernel computes
    a = s64 * w64
    amax = np.abs(a).max()
    c_f = F(amax) * F(1.000001)
    alpha_f = F(2.0 * float(c_f) / 16777215.0) * F(1.0000002)
    A = np.clip(np.rint((a + float(c_f)) / float(alpha_f)), 0, 16777215).astype(np.int64)
    C = float(np.sum(m64 * w64))
    G = float(np.sum(np.maximum(np.abs(m64), np.abs(m64 + 255.0 * s64)) * np.abs(w64)))
    R = codes.astype(np.int64).sum(axis=1)
    nrm = np.sqrt(((b * s64 + m64) ** 2).sum(axis=1))
    invn = (1.0 / (nrm + 1e-8)).astype(F)
    rmax, imax = float(R.max()), float(invn.max())
    inner = rmax * float(alpha_f) * 0.5 + U * (10.1 * float(c_f) * rmax + 3.0 * abs(C) * 1.000001)
    if cosine:
        E = 1.25 * (inner * imax + ((3.0 * d / 64.0 + 36.0) * 1.0001 + 5.0) * U)
    else:
        E = 1.25 * (inner + 1.01 * (d / 32.0 + 22.0) * U * G * 1.000001)
    # ---- the epilogue in fp32: three exact limb dots, recombination, row terms
    L0 = codes.astype(np.int64) @ (A & 255)
    L1 = codes.astype(np.int64) @ ((A >> 8) & 255)
    L2 = codes.astype(np.int64) @ (A >> 16)
    tsum = _fma32(L2.astype(F), F(65536.0), _fma32(L1.astype(F), F(256.0), L0.astype(F)))
    approx = _fma32(-alpha_f, tsum, _fma32(c_f, R.astype(F), F(-C)))
    if cosine:
        approx = _fma32(approx, invn, F(1.0))
    # ---- the scan's value: fp32 accumulation (sequential per lane of 32, then a tree) and float64 ground truth
    dec32 = _fma32(codes.astype(F), s, m)
    lanes = np.zeros((n, 32), F)
    nl = np.zeros((n, 32), F)
    for j in range(d):                                                    # 16-code chunks, chunk c -> lane c % 32
        lane = (j // 16) % 32
        lanes[:, lane] = _fma32(dec32[:, j], w[j], lanes[:, lane])
        nl[:, lane] = _fma32(dec32[:, j], dec32[:, j], nl[:, lane])
    acc = lanes
    nr = nl
    while acc.shape[1] > 1:
        acc = (acc[:, ::2] + acc[:, 1::2]).astype(F)
        nr = (nr[:, ::2] + nr[:, 1::2]).astype(F)
    acc, nr = acc[:, 0], nr[:, 0]
    if cosine:
        scan = (F(1.0) - (acc / (np.sqrt(nr).astype(F) + F(1e-8))).astype(F)).astype(F)
        truth = 1.0 - ((b * s64 + m64) @ w64) / (nrm + 1e-8)
    else:
        scan = -acc
        truth = -((b * s64 + m64) @ w64)
    worst = float(np.abs(approx.astype(np.float64) - scan.astype(np.float64)).max())
    assert worst <= E, (case, cosine, worst, E)
    assert float(np.abs(approx.astype(np.float64) - truth).max()) <= E
    # the bound is not vacuous either: within two orders of magnitude of what is observed on benign data
    if case == "benign":
        assert E <= 300 * max(worst, 1e-12), (worst, E)


@pytest.mark.parametrize("case", ["random", "one_dominant_subspace", "constant_subspaces"])
def test_pq_fixed_point_filter_is_a_superset(case):
    rng = np.random.default_rng(3)
    n, m, kc = 200_000, 48, 256
    lut = (rng.random((m, kc)) * 0.05).astype(F)
    if case == "one_dominant_subspace":
        lut[7] *= F(1000.0)
    elif case == "constant_subspaces":
        lut[::2] = F(0.3)
    codes = rng.integers(0, kc, (n, m), dtype=np.uint8)
    mn = lut.min(axis=1).astype(np.float64)
    inv = 65535.0 / float((lut.max(axis=1).astype(np.float64) - lut.min(axis=1).astype(np.float64)).sum())
    ent = np.minimum(65535, np.floor((lut.astype(np.float64) - mn[:, None]) * inv)).astype(np.int64)
    B = float(mn.sum())
    E = ent[np.arange(m)[None, :], codes].sum(axis=1)
    assert E.max() < 65536                                                 # no carry between the packed u16 fields
    sums32 = np.zeros(n, F)                                                # fp32 sums in some order (two chains, like the device)
    a0 = np.zeros(n, F)
    a1 = np.zeros(n, F)
    for j in range(m):
        v = lut[j][codes[:, j]]
        if j & 1:
            a1 = (a1 + v).astype(F)
        else:
            a0 = (a0 + v).astype(F)
    sums32 = (a0 + a1).astype(F)
    for quant in (1e-5, 1e-3, 0.05):
        thr2 = F(np.quantile(sums32, quant))
        x = (float(thr2) * (1.0 + 2e-5) - B) * inv + 2.0
        T = 65535 if x >= 65535 else (-1 if x < 0 else int(x))
        passing = sums32 <= thr2
        assert (E[passing] <= T).all(), (case, quant)
        # and the filter is selective: it lets through little more than the rows that pass
        assert (E <= T).sum() <= passing.sum() * 1.5 + 200, (case, quant, int((E <= T).sum()), int(passing.sum()))
