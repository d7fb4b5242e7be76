// Product quantizer: encode, lookup-table build and the ADC scan with an in-kernel row bitmask and fused top-k.
//
// Replaces ProductQuantizer.encode (quantization.py:520-539), build_lookup_table (:551-562),
// distances_with_table (:571-578) and search (:588-597).
//
// Bit-exactness: the reference's arithmetic on this path is plain fp32 with a fixed order, so it is
// reproduced exactly: (a) np.sum over a contiguous last axis uses NumPy's pairwise scheme (8 strided
// accumulators up to 128 elements, recursive halves above) — np_pairwise() below is that scheme;
// (b) distances_with_table adds table[m, code] sequentially for m = 0..M-1; (c) sqrtf is IEEE-correct.
// __fadd_rn/__fmul_rn/__fsub_rn stop the compiler from contracting into FMAs.
//
// Layout: codes [N][M] uint8 row major as the reference stores them; one lane per row, the row's M bytes are
// fetched with 128-bit loads when M % 16 == 0 (a warp's 32 rows are one contiguous 32*M byte span, every
// fetched sector is fully used); the query's M x Kc fp32 table lives in shared memory.
#include <algorithm>
#include <cstdlib>

#include "fpv_common.cuh"
#include "fpv_select.cuh"

namespace fpv {

// NumPy's pairwise summation of term(0..n) in fp32 (numpy/_core/src/umath/loops_utils.h.src, pairwise_sum).
template <typename F>
__device__ float np_pairwise(F term, int off, int n) {
    if (n < 8) {
        float r = 0.f;
        for (int i = 0; i < n; ++i) r = __fadd_rn(r, term(off + i));
        return r;
    }
    if (n <= 128) {
        float r0 = term(off), r1 = term(off + 1), r2 = term(off + 2), r3 = term(off + 3);
        float r4 = term(off + 4), r5 = term(off + 5), r6 = term(off + 6), r7 = term(off + 7);
        int i = 8;
        for (; i < n - (n % 8); i += 8) {
            r0 = __fadd_rn(r0, term(off + i)); r1 = __fadd_rn(r1, term(off + i + 1));
            r2 = __fadd_rn(r2, term(off + i + 2)); r3 = __fadd_rn(r3, term(off + i + 3));
            r4 = __fadd_rn(r4, term(off + i + 4)); r5 = __fadd_rn(r5, term(off + i + 5));
            r6 = __fadd_rn(r6, term(off + i + 6)); r7 = __fadd_rn(r7, term(off + i + 7));
        }
        float res = __fadd_rn(__fadd_rn(__fadd_rn(r0, r1), __fadd_rn(r2, r3)),
                              __fadd_rn(__fadd_rn(r4, r5), __fadd_rn(r6, r7)));
        for (; i < n; ++i) res = __fadd_rn(res, term(off + i));
        return res;
    }
    int n2 = n / 2;
    n2 -= n2 % 8;
    float a = np_pairwise(term, off, n2);
    float b = np_pairwise(term, off + n2, n - n2);
    return __fadd_rn(a, b);
}

// ---------------------------------------------------------------------------------------------------- LUT
// grid (M, Q), block = 256: table[q][m][k] = sum_d (cb[m][k][d] - q[m*dsub + d])^2     (quantization.py:560)
__global__ void pq_lut_kernel(const float* __restrict__ cb, int M, int Kc, int dsub,
                              const float* __restrict__ queries, float* __restrict__ lut) {
    extern __shared__ float qsub[];
    const int m = blockIdx.x;
    const int64_t q = blockIdx.y;
    const int D = M * dsub;
    for (int d = threadIdx.x; d < dsub; d += blockDim.x) qsub[d] = queries[q * D + m * dsub + d];
    __syncthreads();
    for (int k = threadIdx.x; k < Kc; k += blockDim.x) {
        const float* c = cb + ((size_t)m * Kc + k) * dsub;
        auto term = [&](int d) { float t = __fsub_rn(__ldg(c + d), qsub[d]); return __fmul_rn(t, t); };
        lut[((size_t)q * M + m) * Kc + k] = np_pairwise(term, 0, dsub);
    }
}

// ---------------------------------------------------------------------------------------------------- encode
// grid (row tiles, M), block = 128 rows: argmin_k sum_d (sub[d] - cb[m][k][d])^2, first minimum wins (np.argmin).
__global__ void __launch_bounds__(128) pq_encode_kernel(const float* __restrict__ v, int64_t N, int D, int64_t ld,
                                                        const float* __restrict__ cb, int M, int Kc, int dsub,
                                                        uint8_t* __restrict__ codes) {
    extern __shared__ float sm[];
    float* cbs = sm;                               // [Kc][dsub]
    float* subs = sm + (size_t)Kc * dsub;          // [128][dsub + 1]
    const int m = blockIdx.y;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = threadIdx.x; i < Kc * dsub; i += blockDim.x) cbs[i] = cb[(size_t)m * Kc * dsub + i];
    float* mine = subs + (size_t)threadIdx.x * (dsub + 1);
    if (row < N)
        for (int d = 0; d < dsub; ++d) mine[d] = v[row * ld + m * dsub + d];
    __syncthreads();
    if (row >= N) return;
    float best = INFINITY;
    int arg = 0;
    for (int k = 0; k < Kc; ++k) {
        const float* c = cbs + (size_t)k * dsub;
        auto term = [&](int d) { float t = __fsub_rn(mine[d], c[d]); return __fmul_rn(t, t); };
        float dist = np_pairwise(term, 0, dsub);
        if (dist < best) { best = dist; arg = k; }          // strict: first minimum, NaN never wins (np.argmin
    }                                                       // would return the first NaN; inputs are finite)
    codes[row * M + m] = (uint8_t)arg;
}

// The same argmin for the common sub-vector sizes, with the row's sub-vector in REGISTERS and the centroids read from
// shared memory as broadcast float4: the generic kernel above reads both operands of every term from shared memory
// (8192 32-bit loads per row and subspace for 12288 flops; ncu: 0.9 G bank conflicts on 500K x 768 rows, 4.0 M rows/s =
// 3 % of the fp32 peak).  Same terms, same np_pairwise order, same strict `<`: bit-identical codes.
template <int DSUB>
__global__ void __launch_bounds__(128) pq_encode_reg_kernel(const float* __restrict__ v, int64_t N, int D, int64_t ld,
                                                            const float* __restrict__ cb, int M, int Kc,
                                                            uint8_t* __restrict__ codes) {
    extern __shared__ __align__(16) float cbs_r[];                   // [Kc][DSUB]
    const int m = blockIdx.y;
    const int64_t row = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (int i = threadIdx.x; i < Kc * DSUB; i += blockDim.x) cbs_r[i] = cb[(size_t)m * Kc * DSUB + i];
    float x[DSUB];
    if (row < N) {
        const float* src = v + row * ld + (size_t)m * DSUB;
        if (DSUB % 4 == 0 && ((reinterpret_cast<uintptr_t>(src) & 15) == 0)) {
#pragma unroll
            for (int d4 = 0; d4 < DSUB / 4; ++d4) {
                const float4 t = __ldg(reinterpret_cast<const float4*>(src) + d4);
                x[4 * d4] = t.x; x[4 * d4 + 1] = t.y; x[4 * d4 + 2] = t.z; x[4 * d4 + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int d = 0; d < DSUB; ++d) x[d] = __ldg(src + d);
        }
    } else {
#pragma unroll
        for (int d = 0; d < DSUB; ++d) x[d] = 0.f;
    }
    __syncthreads();
    if (row >= N) return;
    float best = INFINITY;
    int arg = 0;
#pragma unroll 2
    for (int k = 0; k < Kc; ++k) {
        float c[DSUB];
        if (DSUB % 4 == 0) {
#pragma unroll
            for (int d4 = 0; d4 < DSUB / 4; ++d4) {
                const float4 t = reinterpret_cast<const float4*>(cbs_r + (size_t)k * DSUB)[d4];      // warp-wide broadcast
                c[4 * d4] = t.x; c[4 * d4 + 1] = t.y; c[4 * d4 + 2] = t.z; c[4 * d4 + 3] = t.w;
            }
        } else {
#pragma unroll
            for (int d = 0; d < DSUB; ++d) c[d] = cbs_r[(size_t)k * DSUB + d];
        }
        auto term = [&](int d) { float t = __fsub_rn(x[d], c[d]); return __fmul_rn(t, t); };
        const float dist = np_pairwise(term, 0, DSUB);
        if (dist < best) { best = dist; arg = k; }
    }
    codes[row * M + m] = (uint8_t)arg;
}

// ---------------------------------------------------------------------------------------------------- ADC scan
struct PqParams {
    const float* lut;          // [Q][M][Kc]
    const uint8_t* codes;      // [N][M]
    const uint32_t* mask;
    uint64_t* partials;
    float* out_all;
    int64_t Q, N;
    int M, Kc, K, CAP, parts;
    const uint32_t* only_flagged;   // optional [Q]: queries with a zero entry already have their answer (filter path)
    const float* rot_tab;           // [Q][nblk * Kc * 64] the lane-rotated tables, built once per call (pq_rot_table_kernel)
    int sample_spacing;             // sample passes: every sample_spacing-th group of 32 rows (the sample spans the whole database)
};

__device__ __forceinline__ float adc4(const float* lut_m, int Kc, uint32_t w, float acc, int kmax) {
#pragma unroll
    for (int b = 0; b < 4; ++b) {
        int code = min((int)((w >> (8 * b)) & 0xFFu), kmax);
        acc = __fadd_rn(acc, lut_m[b * Kc + code]);
    }
    return acc;
}

// MODE 2: M % 16 == 0 (uint4 loads), 1: M % 4 == 0 (u32 loads), 0: bytes.  grid = (parts, Q), block = 256.
template <int MODE>
__global__ void __launch_bounds__(256) pq_adc_kernel(PqParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* lut = reinterpret_cast<float*>(smem_raw);
    const size_t lut_bytes = align_up((size_t)p.M * p.Kc * 4, 16);
    uint64_t* sel_base = reinterpret_cast<uint64_t*>(smem_raw + lut_bytes);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q = blockIdx.y;
    const float* src = p.lut + (size_t)q * p.M * p.Kc;
    for (int i = threadIdx.x; i < p.M * p.Kc; i += blockDim.x) lut[i] = src[i];
    WarpSelect<1> sel;
    const bool select = p.K > 0;
    if (select) sel.init(sel_base + (size_t)warp * (p.K + p.CAP), p.K, p.CAP, lane);
    __syncthreads();
    const int kmax = p.Kc - 1;
    const int64_t ngroups = (p.N + 31) / 32;
    for (int64_t g = (int64_t)blockIdx.x * W + warp; g < ngroups; g += (int64_t)gridDim.x * W) {
        const int64_t row = g * 32 + lane;
        const bool valid = row < p.N && (!p.mask || mask_bit(p.mask, row));
        float acc = 0.f;
        if (row < p.N && (valid || p.out_all)) {
            const uint8_t* r = p.codes + row * p.M;
            if (MODE == 2) {
                const uint4* r4 = reinterpret_cast<const uint4*>(r);
                for (int c = 0; c < p.M / 16; ++c) {
                    uint4 w = ldg_nc_u4(r4 + c);
                    const float* l = lut + (size_t)c * 16 * p.Kc;
                    acc = adc4(l, p.Kc, w.x, acc, kmax);
                    acc = adc4(l + 4 * p.Kc, p.Kc, w.y, acc, kmax);
                    acc = adc4(l + 8 * p.Kc, p.Kc, w.z, acc, kmax);
                    acc = adc4(l + 12 * p.Kc, p.Kc, w.w, acc, kmax);
                }
            } else if (MODE == 1) {
                const uint32_t* r1 = reinterpret_cast<const uint32_t*>(r);
                for (int c = 0; c < p.M / 4; ++c) acc = adc4(lut + (size_t)c * 4 * p.Kc, p.Kc, __ldg(r1 + c), acc, kmax);
            } else {
                for (int m = 0; m < p.M; ++m) acc = __fadd_rn(acc, lut[(size_t)m * p.Kc + min((int)__ldg(r + m), kmax)]);
            }
        }
        const float d = sqrtf(acc);
        if (p.out_all && row < p.N) p.out_all[q * p.N + row] = d;
        if (select) sel.add_lanes(0, make_key(d, (uint32_t)row), valid, lane);
    }
    if (select) {
        sel.flush_all(lane);
        block_merge_store<1>(sel_base, p.K, p.CAP, 1, p.partials + ((size_t)q * p.parts + blockIdx.x) * p.K, 0);
    }
}

// Byte offset of position `pos` of row `row` in the packed copy.  Full groups of 32 rows are stored chunk-major:
// the 16-byte chunk v of lane l = row % 32 sits at (v * 32 + l) * 16 inside the group's 32 * M bytes, so one 128-bit
// load per lane reads 512 CONTIGUOUS bytes per warp.  (Row-major, the lanes of such a load are 48 bytes apart and
// touch 12 cache lines: ncu showed the L1TEX pipe 84-90 % busy, as many tag wavefronts for the code loads as data
// wavefronts for the 48 table lookups.)  The rows of a last partial group stay row-major (the copy is exactly N * M bytes).
__host__ __device__ __forceinline__ int64_t pq_packed_offset(int64_t row, int pos, int64_t N, int M) {
    const int64_t g = row >> 5;
    if (g * 32 + 32 > N) return row * M + pos;
    return g * 32 * M + (((int64_t)(pos >> 4) * 32 + (row & 31)) << 4) + (pos & 15);
}

// ---------------------------------------------------------------------------------------------------- rotated ADC
// Conflict-free lookups.  With the plain layout all 32 lanes look up the SAME subspace at the same time, the bank is
// the (random) code -> ~3.5-way conflicts and the scan runs at 21% of HBM (ncu: 42-75% smem wavefronts).  Here the
// subspaces are split into blocks of 32 (plus one of 16); inside a block lane l at step s handles subspace
// base + (s + l) % size, and the table is stored [code][64] with column c = s + l  (value of subspace c % size), so
// the bank is (s + l) % 32: distinct for the 32 lanes, no wrap arithmetic.  The code bytes are stored pre-rotated
// (pq_pack_kernel, once per index): position s of row r holds the code of subspace base + (s + r % 32) % size, so a
// lane still reads its row with plain 128-bit loads and static byte indices.  The sum order differs per lane, so
// these distances agree with the reference to fp32 rounding (1e-6), not bit for bit — the exact-order kernel above
// stays the one behind distances_with_table().
__global__ void pq_pack_kernel(const uint8_t* __restrict__ codes, int64_t N, int M, uint8_t* __restrict__ out) {
    const int64_t total = N * M;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t row = i / M;
        const int pos = (int)(i - row * M);
        const int l = (int)(row & 31);
        const int blk = pos >> 5, base = blk << 5;
        const int size = (M - base) >= 32 ? 32 : 16;
        const int s = pos - base;
        out[pq_packed_offset(row, pos, N, M)] = codes[row * M + base + ((s + l) % size)];
    }
}

// The rotated tables of every query, built once per call: tab[q][blk][code][c] = lut[q][base + c % size][code].  Each
// scanning CTA then copies its query's 128 KB table with coalesced 128-bit loads; building it inside every CTA read the
// LUT with a 1 KB stride (one sector per element) and cost ~20 us per CTA -- as much as the sample pass itself.
__global__ void __launch_bounds__(256) pq_rot_table_kernel(const float* __restrict__ lut_all, int M, int Kc, float* __restrict__ tab_all) {
    const int nblk = (M + 31) >> 5;
    const int per_q = nblk * Kc * 64;
    const float* lut = lut_all + (size_t)blockIdx.y * M * Kc;
    float* tab = tab_all + (size_t)blockIdx.y * per_q;
    // consecutive threads walk the codes of one (block, column): coalesced reads, 256-byte-strided writes (small)
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_q; i += gridDim.x * blockDim.x) {
        const int code = i % Kc, c = (i / Kc) & 63, blk = i / (64 * Kc);
        const int base = blk << 5, size = (M - base) >= 32 ? 32 : 16;
        tab[((size_t)blk * Kc + code) * 64 + c] = lut[(size_t)(base + c % size) * Kc + code];
    }
}

__device__ __forceinline__ void pq_load_table(float* tab, const float* __restrict__ src, int n) {
    const uint4* s4 = reinterpret_cast<const uint4*>(src);
    uint4* d4 = reinterpret_cast<uint4*>(tab);
    for (int i = threadIdx.x; i < n / 4; i += blockDim.x) d4[i] = __ldg(s4 + i);
}

// Row walker of the filter passes: warp `w0` of the grid takes the 32-row groups g0, g0 + gstep, ...  One group's code
// vectors (and its word of the row bitmask) are fetched while the previous group is looked up; the two register
// buffers alternate by unrolling (no copies), and the indices are 32-bit (N < 2^32): the first version spent ~105 of
// its 253 instructions per group on 64-bit index arithmetic, buffer moves and a bitmask load issued right before its use.
template <int NV, class Body>
__device__ __forceinline__ void pq_walk_rows(const PqParams& p, uint32_t g, const uint32_t gstep, const int lane, Body body) {
    const uint32_t N = (uint32_t)p.N;
    const uint32_t ngroups = (uint32_t)((p.N + 31) / 32);
    auto fetch = [&](uint4 (&buf)[NV], uint32_t& mw, uint32_t gg) {
        if (gg * 32u + 32u <= N) {                               // full group: chunk-major, 512 contiguous bytes per load
            const uint4* src = reinterpret_cast<const uint4*>(p.codes + (size_t)gg * (32 * NV * 16)) + lane;
#pragma unroll
            for (int v = 0; v < NV; ++v) buf[v] = ldg_nc_u4(src + v * 32);
        } else {                                                 // last partial group: row-major (pq_packed_offset)
            const uint32_t row = min(gg * 32u + (uint32_t)lane, N - 1u);
            const uint4* src = reinterpret_cast<const uint4*>(p.codes + (size_t)row * (NV * 16));
#pragma unroll
            for (int v = 0; v < NV; ++v) buf[v] = ldg_nc_u4(src + v);
        }
        mw = p.mask ? __ldg(p.mask + gg) : 0xFFFFFFFFu;
    };
    if (g >= ngroups) return;
    uint4 a[NV], b[NV];
    uint32_t ma, mb = 0;
    fetch(a, ma, g);
    while (true) {
        if (g + gstep < ngroups) fetch(b, mb, g + gstep);
        { const uint32_t row = g * 32u + (uint32_t)lane; body(a, row, row < N && ((ma >> lane) & 1u)); }
        g += gstep;
        if (g >= ngroups) break;
        if (g + gstep < ngroups) fetch(a, ma, g + gstep);
        { const uint32_t row = g * 32u + (uint32_t)lane; body(b, row, row < N && ((mb >> lane) & 1u)); }
        g += gstep;
        if (g >= ngroups) break;
    }
}

// The 16 * NV lookups of one row (lane = row % 32) in the rotated fp32 table: a lookup is
//   PRMT (code << 8 | lane*4)  +  LDS [R + UR + imm]  +  FADD;  two alternating chains halve the dependent-add latency.
// Every kernel that needs the fast-path sum of a row calls this, so the sum of a row is the same bits everywhere.
template <int NV, bool CLAMP>
__device__ __forceinline__ float pq_row_sum(const uint4 (&cur)[NV], const uint32_t tbase, const uint32_t blk_bytes,
                                            const uint32_t lane4, const uint32_t kmax8) {
    float acc = 0.f, acc2 = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        // warp-uniform part of the address (table block, s offset 0 / 16 columns)
        const uint32_t tv = tbase + (uint32_t)(v >> 1) * blk_bytes + (v & 1) * 64;
        const uint32_t ws[4] = {cur[v].x, cur[v].y, cur[v].z, cur[v].w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                // one PRMT builds (code << 8) | lane*4: byte 1 = code, byte 0 = lane*4, bytes 2-3 = 0
                uint32_t off = __byte_perm(ws[u], lane4, 0x6504u | (b << 4));
                if (CLAMP) off = min(off, kmax8 | lane4);
                float val;
                asm("ld.shared.f32 %0, [%1];" : "=f"(val) : "r"(tv + off + (uint32_t)((u * 4 + b) * 4)));
                if (b & 1) acc2 += val; else acc += val;
            }
    }
    return acc + acc2;
}

// One CTA per SM, 512 or 1024 threads.  smem: nblk tables of [Kc][64] floats, then one selector per warp.
// NV = M / 16 (16-byte vectors per row); CLAMP guards codes >= Kc when Kc < 256.
template <int NV, bool CLAMP>
__global__ void __launch_bounds__(1024, 1) pq_adc_rot_kernel(PqParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tab = reinterpret_cast<float*>(smem_raw);
    const int nblk = (p.M + 31) >> 5;
    uint64_t* sel_base = reinterpret_cast<uint64_t*>(smem_raw + (size_t)nblk * p.Kc * 64 * 4);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q = blockIdx.y;
    if (p.only_flagged && p.only_flagged[q] == 0) return;
    pq_load_table(tab, p.rot_tab + (size_t)q * nblk * p.Kc * 64, nblk * p.Kc * 64);
    WarpSelect<1> sel;
    const bool select = p.K > 0;
    if (select) sel.init(sel_base + (size_t)warp * (p.K + p.CAP), p.K, p.CAP, lane);
    __syncthreads();
    // 32-bit shared address of column `lane` of table 0; a lookup is  PRMT (code << 8)  +  IADD  +  LDS [.. + imm]  + FADD
    const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t lane4 = (uint32_t)lane * 4u;
    const uint32_t blk_bytes = (uint32_t)p.Kc * 256u;
    const uint32_t kmax8 = (uint32_t)(p.Kc - 1) << 8;
    pq_walk_rows<NV>(p, blockIdx.x * W + warp, gridDim.x * W, lane,
                     [&](const uint4 (&cur)[NV], const uint32_t row, const bool valid) {
        const float sum = valid ? pq_row_sum<NV, CLAMP>(cur, tbase, blk_bytes, lane4, kmax8) : 0.f;
        const float d = sqrtf(sum);
        if (select) sel.add_lanes(0, make_key(d, row), valid, lane);
    });
    if (select) {
        sel.flush_all(lane);
        block_merge_store<1>(sel_base, p.K, p.CAP, 1, p.partials + ((size_t)q * p.parts + blockIdx.x) * p.K, 0);
    }
}

// ---------------------------------------------------------------------------------------------------- filtered ADC
// Large scans (>= 1M rows) run as bound + filter.  The selector of pq_adc_rot_kernel costs as much as the lookups:
// every one of the 4736 warps keeps its own sorted top-K list and re-sorts it ~6 times (ncu r1: 31 M of the kernel's
// 68 M shared-memory wavefronts and a third of its instructions were the selectors').  So:
//   1. pq_sample_min_kernel looks up a SAMPLE of S rows (every (N/S)-th group of 32 rows, so that it is representative
//      whatever the order of the rows) and keeps only the minimum sum per warp --
//      G = 148 x 32 group minima, each a different row;
//   2. pq_tau_kernel takes the k-th smallest group minimum: at least k rows have a sum <= that value, so it bounds the
//      final k-th sum, and with G >> k it is as tight as the (k + k^2 / 2G)-th smallest sum of the whole sample
//      (the round-2 version ran the selector kernel on the sample: 35-70 us + two merge launches for the same bound);
//   3. the filter pass over ALL rows is the same conflict-free lookups, ONE compare per row and a rare warp-aggregated
//      append to the query's candidate list (expected hits ~ N k / S, a few thousand);
//   4. pq_filter_finish_kernel selects and sorts the top-k of the candidates.
// A query whose list overflows (adversarial order, a bitmask that leaves fewer than k sample groups) is recomputed by
// pq_adc_rot_kernel, gated on a device-side flag.
constexpr int PQF_CAP = 16384;          // candidate slots per query
constexpr int PQF_SEL = 4096;           // keys at or below the k-th value that the finish kernel can sort

struct PqFilter {
    const float* thr2;          // [Q] bound on the squared distance (pq_tau_kernel); +inf: no bound, the query overflows
    uint32_t* cnt;              // [Q]
    uint64_t* cand;             // [Q][PQF_CAP]  ordered(squared sum) << 32 | row
    int k;
};

// grid (parts, Q), 1024 threads; every p.sample_spacing-th group of 32 rows.  gmin[q][blockIdx.x * 32 + warp] = smallest sum
// the warp saw.
template <int NV, bool CLAMP>
__global__ void __launch_bounds__(1024, 1) pq_sample_min_kernel(PqParams p, float* __restrict__ gmin) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tab = reinterpret_cast<float*>(smem_raw);
    const int nblk = (p.M + 31) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q = blockIdx.y;
    pq_load_table(tab, p.rot_tab + (size_t)q * nblk * p.Kc * 64, nblk * p.Kc * 64);
    __syncthreads();
    const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t lane4 = (uint32_t)lane * 4u;
    const uint32_t blk_bytes = (uint32_t)p.Kc * 256u;
    const uint32_t kmax8 = (uint32_t)(p.Kc - 1) << 8;
    float best = INFINITY;
    pq_walk_rows<NV>(p, (blockIdx.x * W + warp) * (uint32_t)p.sample_spacing, gridDim.x * W * (uint32_t)p.sample_spacing, lane,
                     [&](const uint4 (&cur)[NV], const uint32_t row, const bool valid) {
        if (valid) best = fminf(best, pq_row_sum<NV, CLAMP>(cur, tbase, blk_bytes, lane4, kmax8));
    });
#pragma unroll
    for (int o = 16; o; o >>= 1) best = fminf(best, __shfl_xor_sync(FPV_FULL_MASK, best, o));
    if (lane == 0) gmin[((size_t)q * gridDim.x + blockIdx.x) * W + warp] = best;
}

// one CTA per query: thr2[q] = k-th smallest finite group minimum (+inf when there are fewer than k)
__global__ void __launch_bounds__(1024) pq_tau_kernel(const float* __restrict__ gmin, int groups, int k, float* __restrict__ thr2) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);            // [groups]
    __shared__ uint32_t hist[256];
    __shared__ int s_bin, s_need, s_n;
    const int q = blockIdx.x;
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    int nv = 0;
    for (int i = threadIdx.x; i < groups; i += blockDim.x) {
        const float v = gmin[(size_t)q * groups + i];
        const bool ok = v < INFINITY;                                // NaN sums (NaN tables) are no bound either
        keys[i] = ok ? make_key(v, (uint32_t)i) : FPV_KEY_MAX;
        nv += ok;
    }
    if (nv) atomicAdd(&s_n, nv);
    __syncthreads();
    if (s_n < k) {                                                   // uniform
        if (threadIdx.x == 0) thr2[q] = INFINITY;
        return;
    }
    const uint64_t kth = block_radix_select(keys, groups, k, hist, &s_bin, &s_need);
    if (threadIdx.x == 0) thr2[q] = ordered_to_f32((uint32_t)(kth >> 32));
}

template <int NV, bool CLAMP>
__global__ void __launch_bounds__(1024, 1) pq_adc_filter_kernel(PqParams p, PqFilter f) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* tab = reinterpret_cast<float*>(smem_raw);
    const int nblk = (p.M + 31) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q = blockIdx.y;
    pq_load_table(tab, p.rot_tab + (size_t)q * nblk * p.Kc * 64, nblk * p.Kc * 64);
    __syncthreads();
    const float thr2 = f.thr2[q];                                    // the sum of a sample row computed by the same code
    const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(tab);
    const uint32_t lane4 = (uint32_t)lane * 4u;
    const uint32_t blk_bytes = (uint32_t)p.Kc * 256u;
    const uint32_t kmax8 = (uint32_t)(p.Kc - 1) << 8;
    uint64_t* cand = f.cand + (size_t)q * PQF_CAP;
    pq_walk_rows<NV>(p, blockIdx.x * W + warp, gridDim.x * W, lane,
                     [&](const uint4 (&cur)[NV], const uint32_t row, const bool valid) {
        const float sum = valid ? pq_row_sum<NV, CLAMP>(cur, tbase, blk_bytes, lane4, kmax8) : 0.f;
        const bool hit = valid && sum <= thr2;
        const uint32_t m = __ballot_sync(FPV_FULL_MASK, hit);
        if (m) {                                                 // rare: ~N k / S hits per query in the whole scan
            const int leader = __ffs(m) - 1;
            uint32_t pos = 0;
            if (lane == leader) pos = atomicAdd(f.cnt + q, (uint32_t)__popc(m));
            pos = __shfl_sync(FPV_FULL_MASK, pos, leader) + (uint32_t)__popc(m & ((1u << lane) - 1u));
            if (hit && pos < (uint32_t)PQF_CAP) cand[pos] = ((uint64_t)f32_to_ordered(sum) << 32) | (uint64_t)row;
        }
    });
}

// ---------------------------------------------------------------------------------------------------- four queries per pass
// Query batches (Q >= 2) share ONE pass over the codes.  The fp32 tables of four queries do not fit in shared memory
// (4 x 128 KB) and four 32-bit lookups per code byte would cost what four passes cost, so this pass is a FILTER on
// fixed-point tables and the survivors are re-scored with the fp32 arithmetic of pq_adc_filter_kernel:
//   entry(q, m, code) = floor((lut[q][m][code] - min_code lut[q][m][.]) * inv_q),   four queries
//   packed in one 8-byte table entry (u16 each) -> ONE 64-bit lookup + two packed 16-bit adds (the compiler fuses
//   pairs of them into IADD3) serve four queries; the sums stay below 2^16, so no carry crosses a field.
//   inv_q = 65535 / sum_m range(q, m).  With B_q = sum_m min(q, m):  B_q + E / inv_q <= exact sum, so every row
//   whose fp32 sum passes `sum <= thr2` has E <= T_q = floor((thr2 (1 + 2e-5) - B_q) inv_q) + 2   (2e-5 covers the 49
//   fp32 roundings of the exact sum, +2 the double roundings of entries and bound): the filter is a superset.
// The table has NO redundant columns (64 KB per block of 32 subspaces, 128 KB at M = 48): the wrapped column index
// ((lane + s) & 31) * 8 of every step s is precomputed per lane as one BYTE (11 registers hold the 32 of them, three
// per register plus a zero byte) and the same PRMT that extracts the code byte puts it in byte 0 of the address:
//   off = PRMT(codes, lx[s / 3]) = code << 8 | column * 8;  LDS.64 [table + off];  2 x packed add
// The codes are the lane-rotated copy of fpv_pq_pack (shared with the one-query kernels).  Survivors (rows only) are
// appended to the queries' candidate lists; pq_quad_rescore_kernel replaces each by the key of the EXACT fp32 sum in
// the lane order of pq_adc_filter_kernel (or drops it), so the answer is bit-identical to the one-query path.
struct PqQuad {
    const uint2* tab;        // [G][nblk][Kc][32] entries of 4 x u16
    const double* stats;     // [G * 4][2]: B_q, inv_q (0: degenerate table, every row passes -> fallback)
};

// grid (slices, G): every CTA recomputes the group's minima / ranges (4 M warp reductions over Kc values) and writes
// its slice of the table; consecutive threads walk the codes of one (block, column): coalesced LUT reads.
__global__ void __launch_bounds__(1024) pq_quad_table_kernel(const float* __restrict__ lut_all, int64_t Q, int M, int Kc,
                                                            uint2* __restrict__ tab_all, double* __restrict__ stats) {
    __shared__ float s_mn[4][96], s_mx[4][96];
    __shared__ double s_inv[4];
    const int nblk = (M + 31) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q0 = (int64_t)blockIdx.y * 4;
    for (int pidx = warp; pidx < 4 * M; pidx += W) {
        const int j = pidx / M, m = pidx - j * M;
        float mn = INFINITY, mx = -INFINITY;
        if (q0 + j < Q) {
            const float* l = lut_all + ((size_t)(q0 + j) * M + m) * Kc;
            for (int c = lane; c < Kc; c += 32) { const float v = __ldg(l + c); mn = fminf(mn, v); mx = fmaxf(mx, v); }
        }
        for (int o = 16; o; o >>= 1) {
            mn = fminf(mn, __shfl_xor_sync(FPV_FULL_MASK, mn, o));
            mx = fmaxf(mx, __shfl_xor_sync(FPV_FULL_MASK, mx, o));
        }
        if (lane == 0) { s_mn[j][m] = mn; s_mx[j][m] = mx; }
    }
    __syncthreads();
    if (threadIdx.x < 4) {
        const int j = threadIdx.x;
        double B = 0.0, inv = 0.0;
        if (q0 + j < Q) {
            double range = 0.0;                                      // sum over the subspaces of (max - min)
            bool finite = true;
            for (int m = 0; m < M; ++m) {
                B += (double)s_mn[j][m];
                const float r = s_mx[j][m] - s_mn[j][m];
                finite = finite && (r >= 0.f) && (r < INFINITY);
                range += (double)r;
            }
            // the 16-bit budget is shared in proportion to the ranges: sum_m floor(range_m inv) <= 65535 (one scale for
            // all subspaces, so the fixed-point sums add up; a subspace with a large range no longer costs the others
            // their resolution, as 65535 / (M max_m range_m) did)
            if (finite && range > 0.0 && range < 1e300) inv = 65535.0 / range;
            if (blockIdx.x == 0) { stats[(q0 + j) * 2] = B; stats[(q0 + j) * 2 + 1] = inv; }
        }
        s_inv[j] = inv;
    }
    __syncthreads();
    const int cap = 65535;
    const int per_g = nblk * Kc * 32;
    uint2* tab = tab_all + (size_t)blockIdx.y * per_g;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < per_g; i += gridDim.x * blockDim.x) {
        const int code = i % Kc, c = (i / Kc) & 31, blk = i / (32 * Kc);
        const int base = blk << 5, size = (M - base) >= 32 ? 32 : 16;
        const int m = base + c % size;
        uint32_t e[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            e[j] = 0;
            if (q0 + j < Q && s_inv[j] > 0.0) {
                const double x = ((double)__ldg(lut_all + ((size_t)(q0 + j) * M + m) * Kc + code) - (double)s_mn[j][m]) * s_inv[j];
                e[j] = x >= 0.0 ? (uint32_t)min((double)cap, floor(x)) : 0u;         // NaN -> 0 (a lower bound all the same)
            }
        }
        tab[((size_t)blk * Kc + code) * 32 + c] = make_uint2(e[0] | (e[1] << 16), e[2] | (e[3] << 16));
    }
}

// The 16 * NV lookups of one row in the four-query table: {sums of queries 0 | 1 << 16, sums of queries 2 | 3 << 16}.
template <int NV, bool CLAMP>
__device__ __forceinline__ uint2 pq_row_quad(const uint4 (&cur)[NV], const uint32_t tbase, const uint32_t blk_bytes,
                                             const uint32_t (&lx)[11], const uint32_t kmax8) {
    uint32_t a01 = 0, a23 = 0, b01 = 0, b23 = 0;                                // two chains per pair of queries
#pragma unroll
    for (int v = 0; v < NV; ++v) {
        const uint32_t tv = tbase + (uint32_t)(v >> 1) * blk_bytes;
        const uint32_t ws[4] = {cur[v].x, cur[v].y, cur[v].z, cur[v].w};
#pragma unroll
        for (int u = 0; u < 4; ++u)
#pragma unroll
            for (int b = 0; b < 4; ++b) {
                const int s = (v & 1) * 16 + u * 4 + b;                          // step inside the block of 32 subspaces
                // byte 0 = wrapped column * 8 (byte s % 3 of lx[s / 3]), byte 1 = code, bytes 2-3 = 0 (byte 3 of lx)
                uint32_t off = __byte_perm(ws[u], lx[s / 3], 0x7700u | (b << 4) | (4 + s % 3));
                if (CLAMP) off = off > (kmax8 | 0xFFu) ? (kmax8 | (off & 0xFFu)) : off;
                uint32_t e01, e23;
                asm("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(e01), "=r"(e23) : "r"(tv + off));
                if (b & 1) { b01 += e01; b23 += e23; } else { a01 += e01; a23 += e23; }
            }
    }
    return make_uint2(a01 + b01, a23 + b23);
}

__device__ __forceinline__ void pq_quad_load_table(unsigned char* smem_raw, const PqQuad& qd, int group, int nblk, int Kc) {
    const int n16 = nblk * Kc * 16;                                             // 16-byte pieces: 2 entries each
    const uint4* s4 = reinterpret_cast<const uint4*>(qd.tab + (size_t)group * nblk * Kc * 32);
    uint4* d4 = reinterpret_cast<uint4*>(smem_raw);
    for (int i = threadIdx.x; i < n16; i += blockDim.x) d4[i] = __ldg(s4 + i);
}

__device__ __forceinline__ void pq_quad_columns(uint32_t (&lx)[11], int lane) {  // byte t of lx[j]: ((lane + 3j + t) & 31) * 8
#pragma unroll
    for (int j = 0; j < 11; ++j)
        lx[j] = (((lane + 3 * j) & 31) << 3) | (((lane + 3 * j + 1) & 31) << 11) | (((lane + 3 * j + 2) & 31) << 19);
}

// grid (parts, G), 1024 threads; every p.sample_spacing-th group of 32 rows.  Per warp and query the smallest fixed-point sum E, written as
// the UPPER bound of that row's fp32 sum:  exact sum < B + (E + M) / inv  (every entry is a floor), times 1 + 1e-5 for
// the fp32 roundings of the sum, rounded up to fp32.
template <int NV, bool CLAMP>
__global__ void __launch_bounds__(1024, 1) pq_sample_min_quad_kernel(PqParams p, PqQuad qd, float* __restrict__ gmin) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int nblk = (p.M + 31) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q0 = (int64_t)blockIdx.y * 4;
    pq_quad_load_table(smem_raw, qd, blockIdx.y, nblk, p.Kc);
    __syncthreads();
    const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t blk_bytes = (uint32_t)p.Kc * 256u;
    const uint32_t kmax8 = (uint32_t)(p.Kc - 1) << 8;
    uint32_t lx[11];
    pq_quad_columns(lx, lane);
    uint32_t best[4] = {0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu, 0xFFFFFFFFu};
    pq_walk_rows<NV>(p, (blockIdx.x * W + warp) * (uint32_t)p.sample_spacing, gridDim.x * W * (uint32_t)p.sample_spacing, lane,
                     [&](const uint4 (&cur)[NV], const uint32_t row, const bool valid) {
        if (valid) {
            const uint2 e = pq_row_quad<NV, CLAMP>(cur, tbase, blk_bytes, lx, kmax8);
            best[0] = min(best[0], e.x & 0xFFFFu); best[1] = min(best[1], e.x >> 16);
            best[2] = min(best[2], e.y & 0xFFFFu); best[3] = min(best[3], e.y >> 16);
        }
    });
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const uint32_t b = __reduce_min_sync(FPV_FULL_MASK, best[j]);
        if (lane == 0 && q0 + j < p.Q) {
            const double B = qd.stats[(q0 + j) * 2], inv = qd.stats[(q0 + j) * 2 + 1];
            float ub = INFINITY;
            if (b != 0xFFFFFFFFu && inv > 0.0) {
                const double x = (B + ((double)b + (double)p.M + 1.0) / inv) * (1.0 + 1e-5);
                ub = __double2float_ru(x);
            }
            gmin[((size_t)(q0 + j) * gridDim.x + blockIdx.x) * W + warp] = ub;
        }
    }
}

template <int NV, bool CLAMP>
__global__ void __launch_bounds__(1024, 1) pq_adc_quad_kernel(PqParams p, PqFilter f, PqQuad qd) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ int s_T[4];
    const int nblk = (p.M + 31) >> 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q0 = (int64_t)blockIdx.y * 4;
    pq_quad_load_table(smem_raw, qd, blockIdx.y, nblk, p.Kc);
    if (threadIdx.x < 4) {
        const int64_t q = q0 + threadIdx.x;
        int T = -1;                                                             // padding query: nothing passes
        if (q < p.Q) {
            const float thr2 = f.thr2[q];
            const double B = qd.stats[q * 2], inv = qd.stats[q * 2 + 1];
            const double x = ((double)thr2 * (1.0 + 2e-5) - B) * inv + 2.0;
            T = !(inv > 0.0) || !(x == x) || x >= 65535.0 ? 65535 : x < 0.0 ? -1 : (int)x;
        }
        s_T[threadIdx.x] = T;
    }
    __syncthreads();
    const int T0 = s_T[0], T1 = s_T[1], T2 = s_T[2], T3 = s_T[3];
    const uint32_t tbase = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t blk_bytes = (uint32_t)p.Kc * 256u;
    const uint32_t kmax8 = (uint32_t)(p.Kc - 1) << 8;
    uint32_t lx[11];
    pq_quad_columns(lx, lane);
    pq_walk_rows<NV>(p, blockIdx.x * W + warp, gridDim.x * W, lane,
                     [&](const uint4 (&cur)[NV], const uint32_t row, const bool valid) {
        uint32_t hits = 0;
        if (valid) {
            const uint2 e = pq_row_quad<NV, CLAMP>(cur, tbase, blk_bytes, lx, kmax8);
            hits = ((int)(e.x & 0xFFFFu) <= T0 ? 1u : 0u) | ((int)(e.x >> 16) <= T1 ? 2u : 0u) |
                   ((int)(e.y & 0xFFFFu) <= T2 ? 4u : 0u) | ((int)(e.y >> 16) <= T3 ? 8u : 0u);
        }
        if (__any_sync(FPV_FULL_MASK, hits != 0)) {                              // rare
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const bool hit = (hits >> j) & 1u;
                const uint32_t m = __ballot_sync(FPV_FULL_MASK, hit);
                if (m) {
                    const int leader = __ffs(m) - 1;
                    uint32_t pos = 0;
                    if (lane == leader) pos = atomicAdd(f.cnt + q0 + j, (uint32_t)__popc(m));
                    pos = __shfl_sync(FPV_FULL_MASK, pos, leader) + (uint32_t)__popc(m & ((1u << lane) - 1u));
                    if (hit && pos < (uint32_t)PQF_CAP) f.cand[(size_t)(q0 + j) * PQF_CAP + pos] = (uint64_t)row;
                }
            }
        }
    });
}

// grid (PQF_CAP / 256, Q): candidate rows of the four-query pass -> keys of their exact fp32 sums, accumulated in the
// order pq_adc_filter_kernel uses for that row (lane = row % 32, two alternating chains), or FPV_KEY_MAX when the
// exact sum fails the bound the one-query filter applies.
__global__ void __launch_bounds__(256) pq_quad_rescore_kernel(PqParams p, PqFilter f) {
    const int64_t q = blockIdx.y;
    const uint32_t c = f.cnt[q];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (c > (uint32_t)PQF_CAP || i >= c) return;                                 // overflow: the fallback answers this query
    const float thr2 = f.thr2[q];
    uint64_t* slot = f.cand + (size_t)q * PQF_CAP + i;
    const uint32_t row = (uint32_t)*slot;
    const int l = (int)(row & 31u), kmax = p.Kc - 1;
    const float* lut = p.lut + (size_t)q * p.M * p.Kc;
    float acc = 0.f, acc2 = 0.f;
    for (int pos = 0; pos < p.M; ++pos) {
        const int base = (pos >> 5) << 5, size = (p.M - base) >= 32 ? 32 : 16;
        const int m = base + (pos - base + l) % size;
        const float val = __ldg(lut + (size_t)m * p.Kc + min((int)__ldg(p.codes + pq_packed_offset(row, pos, p.N, p.M)), kmax));
        if (pos & 1) acc2 = __fadd_rn(acc2, val); else acc = __fadd_rn(acc, val);
    }
    const float sum = __fadd_rn(acc, acc2);
    *slot = sum <= thr2 ? (((uint64_t)f32_to_ordered(sum) << 32) | (uint64_t)row) : FPV_KEY_MAX;
}

// one CTA per query: candidates of the filter pass (squared sums) -> the final top-k
__global__ void __launch_bounds__(1024) pq_filter_finish_kernel(const uint64_t* __restrict__ cand_all, const uint32_t* __restrict__ cnt,
                                                               uint32_t* __restrict__ flags, int k, int64_t id_base,
                                                               float* __restrict__ out_dist, int64_t* __restrict__ out_idx,
                                                               int32_t* __restrict__ out_count) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);            // [PQF_CAP]
    uint64_t* sel = keys + PQF_CAP;                                  // [PQF_SEL]
    __shared__ uint32_t hist[256];
    __shared__ int s_bin, s_need, s_n;
    const int q = blockIdx.x;
    const uint32_t c_raw = cnt[q];
    if (c_raw > (uint32_t)PQF_CAP) {                                 // overflow: pq_adc_rot_kernel answers this query
        if (threadIdx.x == 0) flags[q] = 1;
        return;
    }
    const int c = (int)c_raw;
    const uint64_t* mine = cand_all + (size_t)q * PQF_CAP;
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const uint64_t key = mine[i];                                // FPV_KEY_MAX: dropped by pq_quad_rescore_kernel
        keys[i] = key == FPV_KEY_MAX ? key : make_key(sqrtf(ordered_to_f32((uint32_t)(key >> 32))), (uint32_t)key);   // the distance the scan returns
    }
    int ns = 0;                                                      // real entries
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    const int tot = c;
    for (int i = threadIdx.x; i < tot; i += blockDim.x) ns += keys[i] != FPV_KEY_MAX;
    if (ns) atomicAdd(&s_n, ns);
    __syncthreads();
    const int have = s_n;                                            // real entries
    const int kk = min(k, have);
    __syncthreads();
    if (threadIdx.x == 0) s_n = 0;
    __syncthreads();
    uint32_t kth_v = 0xFFFFFFFFu;
    if (kk > 0) kth_v = (uint32_t)(block_radix_select(keys, tot, kk, hist, &s_bin, &s_need) >> 32);
    for (int i = threadIdx.x; i < tot; i += blockDim.x) {
        const uint64_t key = keys[i];
        if (key != FPV_KEY_MAX && (uint32_t)(key >> 32) <= kth_v) {
            const int pos = atomicAdd(&s_n, 1);
            if (pos < PQF_SEL) sel[pos] = key;
        }
    }
    __syncthreads();
    const int R = s_n;
    if (R > PQF_SEL) {                                               // thousands of ties on the k-th distance
        if (threadIdx.x == 0) flags[q] = 1;
        return;
    }
    int P2 = 2; while (P2 < R) P2 <<= 1;
    for (int i = R + threadIdx.x; i < P2; i += blockDim.x) sel[i] = FPV_KEY_MAX;
    __syncthreads();
    block_bitonic_sort(sel, P2);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = i < kk;
        const uint64_t key = ok ? sel[i] : FPV_KEY_MAX;
        out_dist[(size_t)q * k + i] = ok ? ordered_to_f32((uint32_t)(key >> 32)) : INFINITY;
        out_idx[(size_t)q * k + i] = ok ? id_base + (int64_t)(uint32_t)key : -1;
    }
    if (out_count && threadIdx.x == 0) out_count[q] = kk;
}

struct PqPlan { int K, CAP, parts; size_t off_part, total, smem; };
static PqPlan plan_pq(int64_t Q, int64_t N, int M, int Kc, int k) {
    PqPlan pl{};
    pl.K = k > 0 ? sel_K(k) : 0;
    pl.CAP = k > 0 ? sel_CAP(pl.K) : 0;
    pl.smem = align_up((size_t)M * Kc * 4, 16) + (size_t)8 * (pl.K + pl.CAP) * 8;
    int per_sm = (int)((size_t)(220 * 1024) / (pl.smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    int64_t want = (int64_t)sm_count() * per_sm;
    int64_t parts = Q > 0 ? (want + Q - 1) / Q : want;
    int64_t max_parts = (N + 255) / 256;
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    pl.parts = (int)parts;
    pl.off_part = 0;
    pl.total = 256 + (size_t)(Q > 0 ? Q : 0) * pl.parts * pl.K * 8;
    return pl;
}

template <int MODE>
static int launch_adc(const PqParams& p, const PqPlan& pl, cudaStream_t st) {
    if (pl.smem > 48 * 1024)
        FPV_CUDA(cudaFuncSetAttribute(pq_adc_kernel<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    pq_adc_kernel<MODE><<<dim3(pl.parts, (unsigned)p.Q), 256, pl.smem, st>>>(p);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

}  // namespace fpv

using namespace fpv;

extern "C" int fpv_pq_build_lut(const float* codebooks, int m, int kc, int dsub, const float* queries, int64_t q,
                                float* lut, void* stream) {
    FPV_REQUIRE(m >= 1 && kc >= 1 && kc <= 256 && dsub >= 1 && q >= 0, "pq_build_lut: bad shape m=%d kc=%d dsub=%d q=%lld",
                m, kc, dsub, (long long)q);
    FPV_REQUIRE(q <= 65535 && m <= 65535, "pq_build_lut: grid too large");
    if (q == 0) return FPV_OK;
    FPV_REQUIRE(codebooks && queries && lut, "pq_build_lut: null pointer");
    pq_lut_kernel<<<dim3(m, (unsigned)q), 256, (size_t)dsub * 4, (cudaStream_t)stream>>>(codebooks, m, kc, dsub, queries, lut);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

extern "C" int fpv_pq_encode(const float* vectors, int64_t n, int d, int64_t ld, const float* codebooks, int m, int kc,
                             uint8_t* out_codes, void* stream) {
    FPV_REQUIRE(n >= 0 && d >= 1 && m >= 1 && d % m == 0 && ld >= d && kc >= 1 && kc <= 256,
                "pq_encode: bad shape n=%lld d=%d m=%d kc=%d", (long long)n, d, m, kc);
    FPV_REQUIRE(m <= 65535, "pq_encode: m too large");
    if (n == 0) return FPV_OK;
    FPV_REQUIRE(vectors && codebooks && out_codes, "pq_encode: null pointer");
    int dsub = d / m;
    if (dsub == 4 || dsub == 8 || dsub == 16 || dsub == 32) {        // register-resident sub-vectors (same codes)
        typedef void (*RegKernel)(const float*, int64_t, int, int64_t, const float*, int, int, uint8_t*);
        const RegKernel rk = dsub == 4 ? pq_encode_reg_kernel<4> : dsub == 8 ? pq_encode_reg_kernel<8>
                           : dsub == 16 ? pq_encode_reg_kernel<16> : pq_encode_reg_kernel<32>;
        const size_t rsmem = (size_t)kc * dsub * 4;                  // <= 32 KB
        rk<<<dim3((unsigned)((n + 127) / 128), m), 128, rsmem, (cudaStream_t)stream>>>(vectors, n, d, ld, codebooks, m, kc, out_codes);
        FPV_LAUNCH_CHECK();
        return FPV_OK;
    }
    size_t smem = ((size_t)kc * dsub + (size_t)128 * (dsub + 1)) * 4;
    FPV_REQUIRE(smem <= (size_t)max_smem_optin(), "pq_encode: kc=%d dsub=%d needs %zu B shared memory", kc, dsub, smem);
    if (smem > 48 * 1024)
        FPV_CUDA(cudaFuncSetAttribute(pq_encode_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    pq_encode_kernel<<<dim3((unsigned)((n + 127) / 128), m), 128, smem, (cudaStream_t)stream>>>(vectors, n, d, ld, codebooks,
                                                                                                  m, kc, dsub, out_codes);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

namespace fpv {
struct PqRotPlan { int K, CAP, parts, warps; size_t total, smem; bool ok;
                   bool filter; int64_t sample_rows; int sample_parts, sample_groups; size_t off_tab, off_gmin, off_thr2, off_cnt, off_flags, off_cand;
                   bool quad; int groups; size_t quad_smem, off_qtab, off_qstats; };
// FPV_PQ_FILTER=0 keeps the one-pass selector kernel for every size (A/B measurements)
// FPV_PQ_QUAD=0 scans once per query even for query batches (A/B measurements)
static bool pq_quad_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FPV_PQ_QUAD"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
static bool pq_filter_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FPV_PQ_FILTER"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}
static PqRotPlan plan_pq_rot(int64_t Q, int64_t N, int M, int Kc, int k) {
    PqRotPlan pl{};
    pl.K = sel_K(k > 0 ? k : 1);
    pl.CAP = sel_CAP(pl.K);
    const int nblk = (M + 31) / 32;
    const size_t tables = (size_t)nblk * Kc * 64 * 4;
    pl.warps = 32;
    while (pl.warps > 8 && tables + (size_t)pl.warps * (pl.K + pl.CAP) * 8 > (size_t)max_smem_optin()) pl.warps >>= 1;
    pl.smem = tables + (size_t)pl.warps * (pl.K + pl.CAP) * 8;
    const int nv = M / 16;
    pl.ok = (M % 16 == 0) && (nv == 1 || nv == 2 || nv == 3 || nv == 4 || nv == 6) && pl.smem <= (size_t)max_smem_optin() && k >= 1;
    int64_t parts = Q > 0 ? ((int64_t)sm_count() + Q - 1) / Q : sm_count();
    int64_t max_parts = (N + 511) / 512;
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    pl.parts = (int)parts;
    size_t o = align_up(256 + (size_t)(Q > 0 ? Q : 0) * std::max<int64_t>(pl.parts, 2 * (int64_t)sm_count()) * pl.K * 8, 256);
    pl.off_tab = o; o += align_up((size_t)(Q > 0 ? Q : 0) * tables, 256);
    // two-pass form for large scans: sample rows S with N k / S ~ 8192 expected hits, at most a quarter of the rows
    pl.filter = pl.ok && pq_filter_enabled() && N >= (1 << 20) && Q >= 1;
    if (pl.filter) {
        int64_t S = (int64_t)((double)N * k / 8192.0);       // ~8192 expected hits in a 16384-slot list
        S = std::max<int64_t>(S, 131072);
        S = std::min<int64_t>(S, N / 4);
        S = (S + 1023) / 1024 * 1024;
        pl.sample_rows = S;
        // one group minimum per warp of a 1024-thread CTA per SM (pq_sample_min_kernel)
        pl.sample_parts = (int)std::max<int64_t>(1, std::min<int64_t>(sm_count(), S / 1024));
        pl.sample_groups = pl.sample_parts * 32;
        const size_t Qz = (size_t)Q;
        pl.off_gmin = o; o += align_up(Qz * pl.sample_groups * 4, 256);
        pl.off_thr2 = o; o += align_up(Qz * 4, 256);
        pl.off_cnt = o;   o += align_up(Qz * 4, 256);
        pl.off_flags = o; o += align_up(Qz * 4, 256);
        pl.off_cand = o;  o += Qz * PQF_CAP * 8;
        // query batches: one pass per group of four queries over fixed-point tables (pq_adc_quad_kernel)
        pl.quad_smem = (size_t)nblk * Kc * 32 * 8;
        pl.groups = (int)((Q + 3) / 4);
        pl.quad = Q >= 2 && M <= 96 && pq_quad_enabled() && pl.quad_smem + 1024 <= (size_t)max_smem_optin();
        if (pl.quad) {
            pl.off_qtab = o;   o += align_up((size_t)pl.groups * pl.quad_smem, 256);
            pl.off_qstats = o; o += align_up((size_t)pl.groups * 4 * 16, 256);
        }
    }
    pl.total = o;
    return pl;
}
}  // namespace fpv

extern "C" int fpv_pq_pack(const uint8_t* codes, int64_t n, int m, uint8_t* out_packed, void* stream) {
    FPV_REQUIRE(n >= 0 && m >= 16 && m % 16 == 0, "pq_pack: needs m %% 16 == 0, got m=%d", m);
    if (n == 0) return FPV_OK;
    FPV_REQUIRE(codes && out_packed && codes != out_packed, "pq_pack: null or aliased pointer");
    int64_t blocks = (n * m + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    pq_pack_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(codes, n, m, out_packed);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

extern "C" size_t fpv_pq_adc_packed_workspace(int64_t q, int64_t n, int m, int kc, int k) {
    if (m <= 0 || kc <= 0 || k <= 0) return 0;
    PqRotPlan pl = plan_pq_rot(q, n, m, kc, k);
    return pl.ok ? pl.total : 0;          // 0: shape not supported by the rotated kernel, use fpv_pq_adc_topk
}

extern "C" int fpv_pq_adc_packed_topk(const float* lut, int64_t q, const uint8_t* packed, int64_t n, int m, int kc,
                                      int k, const uint32_t* mask_words, int64_t id_base,
                                      float* out_dist, int64_t* out_idx, int32_t* out_count,
                                      void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FPV_REQUIRE(q >= 0 && q <= 65535 && n >= 0 && n < (1ll << 32) && kc >= 1 && kc <= 256, "pq_adc_packed: bad shape");
    FPV_REQUIRE(k >= 1 && k <= FPV_MAX_K, "pq_adc_packed: k=%d outside [1,%d]", k, FPV_MAX_K);
    if (q == 0) return FPV_OK;
    PqRotPlan pl = plan_pq_rot(q, n, m, kc, k);
    if (!pl.ok) { set_error("pq_adc_packed: m=%d kc=%d k=%d not supported by the rotated kernel", m, kc, k); return FPV_ERR_UNSUPPORTED; }
    FPV_REQUIRE(lut && (packed || n == 0) && out_dist && out_idx, "pq_adc_packed: null pointer");
    FPV_REQUIRE((reinterpret_cast<uintptr_t>(packed) & 15) == 0, "pq_adc_packed: codes must be 16-byte aligned");
    if (!ws || ws_bytes < pl.total) { set_error("pq_adc_packed: workspace %zu < %zu", ws_bytes, pl.total); return FPV_ERR_WORKSPACE; }
    PqParams p{};
    p.lut = lut; p.codes = packed; p.mask = mask_words; p.partials = reinterpret_cast<uint64_t*>(ws); p.out_all = nullptr;
    p.Q = q; p.N = n; p.M = m; p.Kc = kc; p.K = pl.K; p.CAP = pl.CAP; p.parts = pl.parts;
    typedef void (*RotKernel)(PqParams);
    RotKernel kern = nullptr;
    const bool clamp = kc < 256;
    switch (m / 16) {
        case 1: kern = clamp ? pq_adc_rot_kernel<1, true> : pq_adc_rot_kernel<1, false>; break;
        case 2: kern = clamp ? pq_adc_rot_kernel<2, true> : pq_adc_rot_kernel<2, false>; break;
        case 3: kern = clamp ? pq_adc_rot_kernel<3, true> : pq_adc_rot_kernel<3, false>; break;
        case 4: kern = clamp ? pq_adc_rot_kernel<4, true> : pq_adc_rot_kernel<4, false>; break;
        default: kern = clamp ? pq_adc_rot_kernel<6, true> : pq_adc_rot_kernel<6, false>; break;
    }
    FPV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    {
        float* tabs = reinterpret_cast<float*>(static_cast<char*>(ws) + pl.off_tab);
        const int per_q = ((m + 31) / 32) * kc * 64;
        pq_rot_table_kernel<<<dim3((unsigned)std::min(64, (per_q + 255) / 256), (unsigned)q), 256, 0, st>>>(lut, m, kc, tabs);
        FPV_LAUNCH_CHECK();
        p.rot_tab = tabs;
    }
    if (!pl.filter) {
        kern<<<dim3(pl.parts, (unsigned)q), pl.warps * 32, pl.smem, st>>>(p);
        FPV_LAUNCH_CHECK();
        return launch_finalize(p.partials, q, pl.parts, pl.K, k, id_base, out_dist, out_idx, out_count, st);
    }
    // ---- bound + filter: group minima of a sample -> bound; pure filter over all rows; select (see pq_adc_filter_kernel)
    char* w = static_cast<char*>(ws);
    float* gmin = reinterpret_cast<float*>(w + pl.off_gmin);
    float* thr2 = reinterpret_cast<float*>(w + pl.off_thr2);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(w + pl.off_cnt);
    uint32_t* flags = reinterpret_cast<uint32_t*>(w + pl.off_flags);
    uint64_t* cand = reinterpret_cast<uint64_t*>(w + pl.off_cand);
    FPV_CUDA(cudaMemsetAsync(cnt, 0, (size_t)(pl.off_cand - pl.off_cnt), st));        // cnt and flags
    PqParams ps = p;                                               // the sample: every spacing-th group, over the whole database
    ps.sample_spacing = (int)std::max<int64_t>(1, ((n + 31) / 32) / std::max<int64_t>(1, pl.sample_rows / 32));
    PqFilter f{};
    f.thr2 = thr2; f.cnt = cnt; f.cand = cand; f.k = k;
    const size_t fin_smem = (size_t)(PQF_CAP + PQF_SEL) * 8;
    const size_t tau_smem = (size_t)pl.sample_groups * 8;
    const int nv = m / 16;
#define FPV_PQ_PICK(NAME) (nv == 1 ? (clamp ? NAME<1, true> : NAME<1, false>) : nv == 2 ? (clamp ? NAME<2, true> : NAME<2, false>) : \
                           nv == 3 ? (clamp ? NAME<3, true> : NAME<3, false>) : nv == 4 ? (clamp ? NAME<4, true> : NAME<4, false>) : \
                                     (clamp ? NAME<6, true> : NAME<6, false>))
    if (pl.quad) {
        auto sk = FPV_PQ_PICK(pq_sample_min_quad_kernel);
        auto qk = FPV_PQ_PICK(pq_adc_quad_kernel);
        PqQuad qd{};
        uint2* qtab = reinterpret_cast<uint2*>(w + pl.off_qtab);
        double* qstats = reinterpret_cast<double*>(w + pl.off_qstats);
        qd.tab = qtab; qd.stats = qstats;
        pq_quad_table_kernel<<<dim3(16, (unsigned)pl.groups), 1024, 0, st>>>(lut, q, m, kc, qtab, qstats);
        FPV_LAUNCH_CHECK();
        FPV_CUDA(cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.quad_smem));
        sk<<<dim3(pl.sample_parts, (unsigned)pl.groups), 1024, pl.quad_smem, st>>>(ps, qd, gmin);
        FPV_LAUNCH_CHECK();
        pq_tau_kernel<<<(unsigned)q, 1024, tau_smem, st>>>(gmin, pl.sample_groups, k, thr2);
        FPV_LAUNCH_CHECK();
        FPV_CUDA(cudaFuncSetAttribute(qk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.quad_smem));
        qk<<<dim3(sm_count(), (unsigned)pl.groups), 1024, pl.quad_smem, st>>>(p, f, qd);
        FPV_LAUNCH_CHECK();
        pq_quad_rescore_kernel<<<dim3(PQF_CAP / 256, (unsigned)q), 256, 0, st>>>(p, f);
        FPV_LAUNCH_CHECK();
    } else {
        auto sk = FPV_PQ_PICK(pq_sample_min_kernel);
        auto fk = FPV_PQ_PICK(pq_adc_filter_kernel);
        const size_t tab_smem = (size_t)((m + 31) / 32) * kc * 64 * 4;
        FPV_CUDA(cudaFuncSetAttribute(sk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_smem));
        sk<<<dim3(pl.sample_parts, (unsigned)q), 1024, tab_smem, st>>>(ps, gmin);
        FPV_LAUNCH_CHECK();
        pq_tau_kernel<<<(unsigned)q, 1024, tau_smem, st>>>(gmin, pl.sample_groups, k, thr2);
        FPV_LAUNCH_CHECK();
        FPV_CUDA(cudaFuncSetAttribute(fk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)tab_smem));
        fk<<<dim3(pl.parts, (unsigned)q), 1024, tab_smem, st>>>(p, f);
        FPV_LAUNCH_CHECK();
    }
#undef FPV_PQ_PICK
    FPV_CUDA(cudaFuncSetAttribute(pq_filter_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    pq_filter_finish_kernel<<<(unsigned)q, 1024, fin_smem, st>>>(cand, cnt, flags, k, id_base, out_dist, out_idx, out_count);
    FPV_LAUNCH_CHECK();
    // overflowed queries (normally none): the one-pass selector kernel over all rows, gated on the device-side flags
    p.only_flagged = flags;
    kern<<<dim3(pl.parts, (unsigned)q), pl.warps * 32, pl.smem, st>>>(p);
    FPV_LAUNCH_CHECK();
    return launch_finalize(p.partials, q, pl.parts, pl.K, k, id_base, out_dist, out_idx, out_count, st, flags);
}

extern "C" size_t fpv_pq_adc_workspace(int64_t q, int64_t n, int m, int kc, int k) {
    if (m <= 0 || kc <= 0 || k < 0) return 256;
    return plan_pq(q, n, m, kc, k).total;
}

extern "C" int fpv_pq_adc_topk(const float* lut, int64_t q, const uint8_t* codes, int64_t n, int m, int kc,
                               int k, const uint32_t* mask_words, int64_t id_base,
                               float* out_dist, int64_t* out_idx, int32_t* out_count, float* out_all,
                               void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FPV_REQUIRE(q >= 0 && n >= 0 && m >= 1 && kc >= 1 && kc <= 256, "pq_adc: bad shape q=%lld n=%lld m=%d kc=%d",
                (long long)q, (long long)n, m, kc);
    FPV_REQUIRE(k >= 0 && k <= FPV_MAX_K, "pq_adc: k=%d outside [0,%d]", k, FPV_MAX_K);
    FPV_REQUIRE(k > 0 || out_all, "pq_adc: nothing to do (k == 0 and out_all == NULL)");
    FPV_REQUIRE(n < (1ll << 32), "pq_adc: N=%lld rows per call exceeds 2^32-1 (shard the database)", (long long)n);
    FPV_REQUIRE(q <= 65535, "pq_adc: at most 65535 queries per call");
    if (q == 0) return FPV_OK;
    FPV_REQUIRE(lut && (codes || n == 0), "pq_adc: null pointer");
    FPV_REQUIRE(k == 0 || (out_dist && out_idx), "pq_adc: null output");
    PqPlan pl = plan_pq(q, n, m, kc, k);
    if (!ws || ws_bytes < pl.total) { set_error("pq_adc: workspace %zu < %zu", ws_bytes, pl.total); return FPV_ERR_WORKSPACE; }
    FPV_REQUIRE(pl.smem <= (size_t)max_smem_optin(), "pq_adc: m=%d kc=%d k=%d needs %zu B shared memory", m, kc, k, pl.smem);
    PqParams p{};
    p.lut = lut; p.codes = codes; p.mask = mask_words; p.partials = reinterpret_cast<uint64_t*>(ws); p.out_all = out_all;
    p.Q = q; p.N = n; p.M = m; p.Kc = kc; p.K = pl.K; p.CAP = pl.CAP; p.parts = pl.parts;
    const uintptr_t a = reinterpret_cast<uintptr_t>(codes);
    int rc;
    if (m % 16 == 0 && (a & 15) == 0) rc = launch_adc<2>(p, pl, st);
    else if (m % 4 == 0 && (a & 3) == 0) rc = launch_adc<1>(p, pl, st);
    else rc = launch_adc<0>(p, pl, st);
    if (rc != FPV_OK) return rc;
    if (k > 0) return launch_finalize(p.partials, q, pl.parts, pl.K, k, id_base, out_dist, out_idx, out_count, st);
    return FPV_OK;
}
