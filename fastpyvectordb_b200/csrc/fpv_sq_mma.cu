// uint8 scalar-quantizer L2 scan on the int8 tensor cores (tcgen05.mma kind::i8, u8 x u8 -> s32 in TMEM).
//
// Replaces ScalarQuantizer.distances_l2 -> _sq_distances_l2_vectorized (quantization.py:145-152, 217-236) followed by
// the caller's top-k, for up to 16 queries per pass over the codes.  The reference distance
//     d = sqrt( sum_j ((qc_j - b_j) * s_j)^2 ),   s_j = fp32(scale_j / 255),  qc = the re-quantised query (:151)
// is a WEIGHTED sum, not an integer dot product, so the tensor cores only FILTER (same certificate as the float path,
// fpv_gemm_topk.cu); the returned distances come from the reference's own arithmetic on the few surviving rows.
//
//   d^2 = A_q + C_row - 2 * sum_j a_j b_j        A_q = sum_j w_j qc_j^2,  C_row = sum_j w_j b_j^2,  a_j = w_j qc_j,  w_j = s_j^2
//
//   * C_row: one fp32 per row, computed once per code matrix in fp64 (fpv_sq_row_term) -- index build work;
//   * the cross term: a_j is fixed-pointed per query, a_j ~ alpha * A_j with A_j a 24-bit unsigned integer split
//     into three 8-bit limbs; the three limb dots  L_l = sum_j A_j^(l) b_j  are EXACT s32 integers from the MMA
//     (<= 1024 * 255 * 255 < 2^31), bit-identical to np.dot in int64 (tests/test_gpu_sq_mma.py), and
//     sum_j a_j b_j ~ alpha * (L_0 + 256 L_1 + 65536 L_2);
//   * the MMA runs with the CODES as the M operand (128 rows per tile, straight from TMA, no u8 -> float conversion:
//     the SIMT scan is bound by exactly that conversion, 3.25 instructions per code byte) and the 3 x 16 limb rows as
//     the N operand (N = 48), so one TMEM lane = one database row and an epilogue thread sees all 16 queries of its
//     row: recombine, compare with the query's threshold, append the rare hit to the query's candidate list.
//
// E bounds |approx d^2 - exactly-evaluated d^2| rigorously (limb truncation + every fp32 rounding of both
// evaluations, see sq_mma_prep_kernel); every row of the true top-k has approx <= a_k + 2E, so the finish kernel
// re-scores the rows inside that window with the arithmetic of the SIMT scan kernels (fpv_sq.cu, bit-identical
// distances and order) and a query whose window does not fit is recomputed by the SIMT scan on the device.
//
// Kernel shape: persistent, one CTA per SM, 256 threads: warp 0 TMA producer (8-stage ring of 128 rows x 128 B,
// SWIZZLE_128B; the limb matrix is loaded once and stays resident), warp 1 MMA issuer (4 x UTCIMMA per K block, two
// 64-column accumulators), warp 2 TMEM allocator, warps 4-7 epilogue.  HBM bound: one byte per code, read once for
// all 16 queries.
#include <cuda.h>

#include <cstdlib>
#include <mutex>

#include "fpv_common.cuh"
#include "fpv_select.cuh"
#include "fpv_sq_common.cuh"
#include "fpv_tc.cuh"

namespace fpv {

constexpr int SQM_BM = 128;                         // database rows per tile (UMMA M)
constexpr int SQM_QB = 16;                          // queries per pass
constexpr int SQM_BN = 3 * SQM_QB;                  // limb rows (UMMA N)
constexpr int SQM_COLS = 64;                        // TMEM columns per accumulator (48 used)
constexpr int SQM_KROW = 128;                       // bytes of K per shared-memory row (one swizzle span)
constexpr int SQM_STAGES = 8;
constexpr int SQM_A_BYTES = SQM_BM * SQM_KROW;      // 16 KB
constexpr int SQM_B_BYTES = SQM_BN * SQM_KROW;      // 6 KB per K block
constexpr int SQM_MAX_KB = 8;                       // D <= 1024
constexpr int SQM_CAP = 16384;                      // candidate slots per query
constexpr int SQM_THREADS = 256;
constexpr int SQM_RMAX = 4096;                      // rows re-scored exactly per query at most
constexpr size_t SQM_OFF_BAR = (size_t)SQM_STAGES * SQM_A_BYTES + (size_t)SQM_MAX_KB * SQM_B_BYTES;
constexpr size_t SQM_SMEM = SQM_OFF_BAR + 256 + SQM_QB * 4 * 4;

struct SqmParams {
    const float* row_term;      // [N]  L2: C_row;  DOT / COSINE: R_row = sum_j b_j
    const float* row_term2;     // [N]  COSINE: 1 / (|decoded row| + 1e-8)
    const uint32_t* mask;       // optional row filter
    const float* qconst;        // [QB][4]: L2: 2*alpha, A_q, -, 1;  DOT / COSINE: alpha, -C_q, -, c  (slot 2 = bound, set here)
    const float* thr;           // [QB]  -(bound): a row passes when approx d^2 <= bound
    uint32_t* cnt;              // [QB]
    uint64_t* cand;             // [QB][SQM_CAP]  ordered(approx d^2) << 32 | row
    int64_t N;
    int nq;                     // queries of this pass (<= QB)
    int tile0, ntiles;          // 128-row tiles of this slab, in SCAN order: scan position s is tile (s * perm_stride) % tiles_total
    int perm_stride, tiles_total;   // perm_stride coprime to tiles_total (a bijection): the first slab spans the whole code matrix
    int nkb;                    // K blocks of 128 bytes
    int32_t* dump;              // test hook: raw limb dots of query 0, [3][N]
};

__device__ __forceinline__ void tc_mma_i8(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// c_format S32 (2 << 4), a / b format unsigned 8 bit (0), both K-major, N >> 3 at bit 17, M >> 4 at bit 24
constexpr uint32_t SQM_IDESC = (2u << 4) | ((uint32_t)(SQM_BN >> 3) << 17) | ((uint32_t)(SQM_BM >> 4) << 24);

template <int KIND>
__global__ void __launch_bounds__(SQM_THREADS, 1)
sq_mma_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, SqmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) __trap();
    const uint32_t sA = base, sB = base + SQM_STAGES * SQM_A_BYTES;
    const uint32_t bars = base + (uint32_t)SQM_OFF_BAR;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * SQM_STAGES, bar_b = bars + 16 * SQM_STAGES;
    const uint32_t bar_tfull = bar_b + 8, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + SQM_OFF_BAR + 16 * SQM_STAGES + 8 + 32);
    float* qc = reinterpret_cast<float*>(smem_raw + SQM_OFF_BAR + 256);        // [QB][4]: 2 alpha, A_q, bound, -
    const int warp = __shfl_sync(FPV_FULL_MASK, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < SQM_STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        mbar_init(bar_b, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 4); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < SQM_QB * 4) {
        const int q = threadIdx.x >> 2, f = threadIdx.x & 3;
        float v = p.qconst[threadIdx.x];
        if (f == 2) v = q < p.nq ? -p.thr[q] : -INFINITY;                        // bound; padding queries never hit
        qc[threadIdx.x] = v;
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(2 * SQM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(FPV_FULL_MASK, *tmem_slot, 0);

    if (warp == 0) {                                    // ---------------- TMA producer
        if (elect_one()) {                              // the limb matrix: loaded once, resident for the whole launch
            mbar_expect_tx(bar_b, (uint32_t)(p.nkb * SQM_B_BYTES));
            for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(sB + kb * SQM_B_BYTES, &tmB, bar_b, kb * SQM_KROW, 0);
        }
        __syncwarp();
        int stage = 0; uint32_t phase = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const int row0 = (int)(((int64_t)(p.tile0 + t) * p.perm_stride) % p.tiles_total) * SQM_BM;
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                if (elect_one()) {
                    mbar_expect_tx(bar_full + 8 * stage, SQM_A_BYTES);
                    tma_load_2d(sA + stage * SQM_A_BYTES, &tmA, bar_full + 8 * stage, kb * SQM_KROW, row0);
                }
                __syncwarp();
                if (++stage == SQM_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {                             // ---------------- MMA issuer
        mbar_wait(bar_b, 0);
        tc_fence_after();
        int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * SQM_COLS;
            for (int kb = 0; kb < p.nkb; ++kb) {
                mbar_wait(bar_full + 8 * stage, phase);
                tc_fence_after();
                if (elect_one()) {
                    const uint64_t ad = make_smem_desc(sA + stage * SQM_A_BYTES), bd = make_smem_desc(sB + kb * SQM_B_BYTES);
#pragma unroll
                    for (int k = 0; k < SQM_KROW / 32; ++k)         // 32 bytes (= 32 u8 elements) of K per instruction
                        tc_mma_i8(d_tmem, ad + 2 * k, bd + 2 * k, SQM_IDESC, (kb | k) != 0);
                    tc_commit(bar_empty + 8 * stage);
                }
                __syncwarp();
                if (++stage == SQM_STAGES) { stage = 0; phase ^= 1; }
            }
            if (elect_one()) tc_commit(bar_tfull + 8 * as);
            __syncwarp();
            as ^= 1; if (as == 0) aphase ^= 1;
        }
    } else if (warp >= 4) {                             // ---------------- epilogue: one thread per database row
        const int quarter = warp & 3;
        int as = 0; uint32_t aphase = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const int64_t row = (((int64_t)(p.tile0 + t) * p.perm_stride) % p.tiles_total) * SQM_BM + quarter * 32 + lane;
            const bool valid = row < p.N && (!p.mask || mask_bit(p.mask, row));
            const float rt = row < p.N ? __ldg(p.row_term + row) : 0.f;
            const float rt2 = KIND == FPV_SQ_COSINE && row < p.N ? __ldg(p.row_term2 + row) : 0.f;
            mbar_wait(bar_tfull + 8 * as, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * SQM_COLS;
            uint32_t r0[16], r1[16], r2[16];
            TMEM_LD16(r0, taddr);
            TMEM_LD16(r1, taddr + 16);
            TMEM_LD16(r2, taddr + 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * as);          // the MMA warp may overwrite this accumulator now
            if (p.dump) {                                             // uniform; test hook only
                if (row < p.N) { p.dump[row] = (int)r0[0]; p.dump[p.N + row] = (int)r0[1]; p.dump[2 * p.N + row] = (int)r0[2]; }
                as ^= 1; if (as == 0) aphase ^= 1;
                continue;
            }
#pragma unroll
            for (int qi = 0; qi < SQM_QB; ++qi) {
                if (qi < p.nq) {                                      // uniform
                    const int c = 3 * qi;
                    const uint32_t w0 = c < 16 ? r0[c & 15] : (c < 32 ? r1[c & 15] : r2[c & 15]);
                    const uint32_t w1 = c + 1 < 16 ? r0[(c + 1) & 15] : (c + 1 < 32 ? r1[(c + 1) & 15] : r2[(c + 1) & 15]);
                    const uint32_t w2 = c + 2 < 16 ? r0[(c + 2) & 15] : (c + 2 < 32 ? r1[(c + 2) & 15] : r2[(c + 2) & 15]);
                    const float tsum = fmaf((float)(int)w2, 65536.0f, fmaf((float)(int)w1, 256.0f, (float)(int)w0));
                    // L2: d^2 = A_q + C_row - 2 alpha T;  DOT: -(alpha T - c R_row + C_q);  COSINE: 1 + DOT / (|row| + 1e-8)
                    float d2;
                    if (KIND == FPV_SQ_L2) d2 = fmaf(-qc[qi * 4 + 0], tsum, qc[qi * 4 + 1] + rt);
                    else {
                        d2 = fmaf(-qc[qi * 4 + 0], tsum, fmaf(qc[qi * 4 + 3], rt, qc[qi * 4 + 1]));
                        if (KIND == FPV_SQ_COSINE) d2 = fmaf(d2, rt2, 1.0f);
                    }
                    const bool hit = valid && d2 <= qc[qi * 4 + 2];
                    const uint32_t m = __ballot_sync(FPV_FULL_MASK, hit);
                    if (m) {
                        const int leader = __ffs(m) - 1;
                        uint32_t pos = 0;
                        if (lane == leader) pos = atomicAdd(p.cnt + qi, (uint32_t)__popc(m));
                        pos = __shfl_sync(FPV_FULL_MASK, pos, leader) + (uint32_t)__popc(m & ((1u << lane) - 1u));
                        if (hit && pos < (uint32_t)SQM_CAP)
                            p.cand[(size_t)qi * SQM_CAP + pos] = ((uint64_t)f32_to_ordered(d2) << 32) | (uint64_t)(uint32_t)row;
                    }
                }
            }
            as ^= 1; if (as == 0) aphase ^= 1;
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * SQM_COLS) : "memory");
    }
}

// ---- per-row term C_row = sum_j w_j b_j^2 (fp64 accumulate, one fp32 rounding), and its maximum ------------------
__global__ void __launch_bounds__(256) sq_row_term_kernel(const uint8_t* __restrict__ codes, int64_t N, int D,
                                                          const float* __restrict__ scale, float* __restrict__ row_term,
                                                          uint32_t* __restrict__ max_bits) {
    extern __shared__ double w_s[];                                   // [D]  w_j = s_j^2, s_j = fp32(scale_j / 255)
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        const double s = (double)__fdiv_rn(scale[j], 255.0f);
        w_s[j] = s * s;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    float local_max = 0.f;
    for (int64_t row = (int64_t)blockIdx.x * W + warp; row < N; row += (int64_t)gridDim.x * W) {
        const uint8_t* r = codes + row * D;
        double acc = 0.0;
        if ((D & 3) == 0 && ((reinterpret_cast<uintptr_t>(codes) & 3) == 0)) {
            for (int j = lane * 4; j < D; j += 128) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(r + j));
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const double x = (double)((w >> (8 * b)) & 0xFFu);
                    acc = fma(w_s[j + b] * x, x, acc);
                }
            }
        } else {
            for (int j = lane; j < D; j += 32) { const double x = (double)__ldg(r + j); acc = fma(w_s[j] * x, x, acc); }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(FPV_FULL_MASK, acc, o);
        const float c = (float)acc;
        if (lane == 0) row_term[row] = c;
        local_max = fmaxf(local_max, c);
    }
    if (lane == 0 && local_max > 0.f) atomicMax(max_bits, __float_as_uint(local_max));   // non-negative floats order like uints
}

// ---- per query: 24-bit fixed point of a_j = w_j qc_j in three byte limbs, A_q, alpha, the error bound ---------------
// bmat [QB * 3][Dp] u8 (row 3*q + l = limb l of query q, zero padded); qconst [QB][4]; ebound / thr / cnt / flags [QB].
__global__ void __launch_bounds__(256) sq_mma_prep_kernel(const uint8_t* __restrict__ qcodes, int nq, int D, int Dp,
                                                          const float* __restrict__ scale, const uint32_t* __restrict__ cmax_bits,
                                                          uint8_t* __restrict__ bmat, float* __restrict__ qconst,
                                                          float* __restrict__ ebound, float* __restrict__ thr,
                                                          uint32_t* __restrict__ cnt, uint32_t* __restrict__ flags) {
    const int q = blockIdx.x;
    __shared__ double red[32];
    __shared__ double s_amax, s_A;
    uint8_t* b0 = bmat + (size_t)(3 * q) * Dp;
    if (q >= nq) {                                                    // padding query of the pass: zero limbs, never hits
        for (int j = threadIdx.x; j < 3 * Dp; j += blockDim.x) b0[j] = 0;
        if (threadIdx.x < 4) qconst[q * 4 + threadIdx.x] = 0.f;
        if (threadIdx.x == 0) { ebound[q] = 0.f; thr[q] = INFINITY; cnt[q] = 0; flags[q] = 0; }
        return;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    double amax = 0.0, A = 0.0;
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        const double s = (double)__fdiv_rn(scale[j], 255.0f), w = s * s, c = (double)qcodes[(size_t)q * D + j];
        amax = fmax(amax, w * c);
        A = fma(w * c, c, A);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        amax = fmax(amax, __shfl_xor_sync(FPV_FULL_MASK, amax, o));
        A += __shfl_xor_sync(FPV_FULL_MASK, A, o);
    }
    if (lane == 0) { red[warp] = amax; red[16 + warp] = A; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0, a = 0.0;
        for (int w2 = 0; w2 < W; ++w2) { m = fmax(m, red[w2]); a += red[16 + w2]; }
        s_amax = m; s_A = a;
    }
    __syncthreads();
    // alpha as a float (the epilogue multiplies in fp32); limbs are computed against exactly this value
    const float alpha_f = s_amax > 0.0 ? (float)(s_amax / 16777215.0) * 1.0000002f : 1.0f;
    const double alpha = (double)alpha_f;
    for (int j = threadIdx.x; j < Dp; j += blockDim.x) {
        uint32_t Aj = 0;
        if (j < D) {
            const double s = (double)__fdiv_rn(scale[j], 255.0f);
            const double a = s * s * (double)qcodes[(size_t)q * D + j];
            long long v = __double2ll_rn(a / alpha);
            Aj = (uint32_t)(v < 0 ? 0 : (v > 16777215ll ? 16777215ll : v));
        }
        b0[j] = (uint8_t)(Aj & 0xFFu);
        b0[Dp + j] = (uint8_t)((Aj >> 8) & 0xFFu);
        b0[2 * Dp + j] = (uint8_t)(Aj >> 16);
    }
    if (threadIdx.x == 0) {
        const float Aq = (float)s_A;
        const float cmax = __uint_as_float(*cmax_bits);
        // |approx - exact evaluation| <=
        //   limb truncation   D * 255 * alpha            (|a_j - alpha A_j| <= alpha / 2, b_j <= 255, times 2)
        // + fp32 roundings of the recombination: I2F of three limb dots, two FMAs, the fp32 alpha, A_q + C_row, the
        //   final FMA, C_row itself: each <= 2^-24 of a quantity <= A_q + C_row            -> 12 * 2^-24 (A_q + C_max)
        // + the exact kernel's own rounding: every term round(|q-b| s_j)^2 carries 2 * 2^-24, a chain is D/32 + 1 FMAs
        //   + 5 butterfly adds + 1, all terms non-negative                          -> (D/32 + 10) * 2^-24 (A_q + C_max)
        // (d^2 <= A_q + C_row because the cross term is non-negative); 25 % slack on top.
        const float u = 5.9604645e-8f;
        const float e = 1.25f * ((float)D * 255.0f * alpha_f + (22.0f + (float)D / 32.0f) * u * (Aq + cmax));
        qconst[q * 4 + 0] = 2.0f * alpha_f;
        qconst[q * 4 + 1] = Aq;
        qconst[q * 4 + 2] = 0.f;
        qconst[q * 4 + 3] = 0.f;
        ebound[q] = e;
        thr[q] = -INFINITY;
        cnt[q] = 0;
        flags[q] = 0;
    }
}

// ---- DOT / COSINE -----------------------------------------------------------------------------------------------------
// distances_dot (quantization.py:176-181, 239-251) and distances_cosine (:154-174) on the same machinery.  With the
// scan's per-dimension constants s_j = fp32(scale_j / 255), m_j = min_j, w_j = decoded query (normalised for cosine):
//     X = sum_j (b_j s_j + m_j) w_j = sum_j a_j b_j + C_q,      a_j = s_j w_j (SIGNED),   C_q = sum_j m_j w_j
// The limbs must be unsigned, so a_j is shifted by c = max_j |a_j|:  a'_j = a_j + c in [0, 2c] is fixed-pointed to
// 24 bits against alpha = 2c / (2^24 - 1), and  sum_j a_j b_j = alpha T - c R_row + tau  with T = L_0 + 256 L_1 + 65536 L_2
// (exact integers from the MMA), R_row = sum_j b_j (one exact fp32 per row, index build) and |tau| <= R_row alpha / 2.
//     DOT     approx = -(alpha T - c R_row + C_q)
//     COSINE  approx = 1 + DOT * invn_row,   invn_row = 1 / (|decoded row| + 1e-8)    (second per-row term)
// Error bounds (u = 2^-24; derivation in DESIGN.md): every fp32 rounding of the recombination acts on a quantity
// <= alpha T + c R + |C_q| <= 3 c R_max + |C_q|:
//     E_inner = R_max alpha / 2 + u (10.1 c R_max + 3 |C_q|)
//     E_dot   = E_inner + 1.01 (D/32 + 22) u G,   G = sum_j max(|m_j|, |m_j + 255 s_j|) |w_j|   (the scan's own roundings)
//     E_cos   = E_inner invn_max + ((3D/64 + 36) 1.0001 + 5) u                                  (|w| <= 1)
// A row with a (near-)zero decoded vector makes invn_max ~ 1e8 and every window overflows: the queries then fall back to
// the SIMT scan on the device (correct, slow) -- cosine over data with zero vectors is ill-defined in the reference too.

// per row: R_row = sum_j b_j (fp32, exact: <= 255 * 1024) and invn_row; maxima[0] = max R_row, maxima[1] = max invn_row
__global__ void __launch_bounds__(256) sq_row_terms_dc_kernel(const uint8_t* __restrict__ codes, int64_t N, int D,
                                                              const float* __restrict__ mn, const float* __restrict__ scale,
                                                              float* __restrict__ row_sum, float* __restrict__ row_invn,
                                                              uint32_t* __restrict__ max_bits) {
    extern __shared__ double sm_d[];                                   // [2][D]: s_j, m_j
    double* s_s = sm_d;
    double* m_s = sm_d + D;
    for (int j = threadIdx.x; j < D; j += blockDim.x) { s_s[j] = (double)__fdiv_rn(scale[j], 255.0f); m_s[j] = (double)mn[j]; }
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    float max_r = 0.f, max_i = 0.f;
    const bool vec = (D & 3) == 0 && ((reinterpret_cast<uintptr_t>(codes) & 3) == 0);
    for (int64_t row = (int64_t)blockIdx.x * W + warp; row < N; row += (int64_t)gridDim.x * W) {
        const uint8_t* r = codes + row * D;
        double nrm = 0.0;
        uint32_t sum = 0;
        if (vec) {
            for (int j = lane * 4; j < D; j += 128) {
                const uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(r + j));
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const uint32_t x = (w >> (8 * b)) & 0xFFu;
                    sum += x;
                    const double dec = fma((double)x, s_s[j + b], m_s[j + b]);
                    nrm = fma(dec, dec, nrm);
                }
            }
        } else {
            for (int j = lane; j < D; j += 32) {
                const uint32_t x = __ldg(r + j);
                sum += x;
                const double dec = fma((double)x, s_s[j], m_s[j]);
                nrm = fma(dec, dec, nrm);
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) { nrm += __shfl_xor_sync(FPV_FULL_MASK, nrm, o); sum += __shfl_xor_sync(FPV_FULL_MASK, sum, o); }
        const float rs = (float)sum;
        const float iv = (float)(1.0 / (sqrt(nrm) + 1e-8));
        if (lane == 0) { row_sum[row] = rs; row_invn[row] = iv; }
        max_r = fmaxf(max_r, rs);
        max_i = fmaxf(max_i, iv);
    }
    if (lane == 0) {                                                   // non-negative floats order like uints
        if (max_r > 0.f) atomicMax(max_bits, __float_as_uint(max_r));
        if (max_i > 0.f) atomicMax(max_bits + 1, __float_as_uint(max_i));
    }
}

// per query slot: limbs of a'_j, qconst = {alpha, -C_q, -, c}, the error bound.  consts [nq][3][Dc] from sq_prep_kernel
// (c0 = s_j, c1 = m_j, c2 = w_j), Dc = D rounded up to 16; bmat rows of Dp = D rounded up to 128 bytes.
__global__ void __launch_bounds__(256) sq_mma_prep_dc_kernel(int kind, const float* __restrict__ consts, int nq, int D, int Dc, int Dp,
                                                             const uint32_t* __restrict__ max_bits, uint8_t* __restrict__ bmat,
                                                             float* __restrict__ qconst, float* __restrict__ ebound,
                                                             float* __restrict__ thr, uint32_t* __restrict__ cnt,
                                                             uint32_t* __restrict__ flags) {
    const int q = blockIdx.x;
    __shared__ double red[3][8];
    __shared__ double s_amax, s_C, s_G;
    uint8_t* b0 = bmat + (size_t)(3 * q) * Dp;
    if (q >= nq) {                                                    // padding query of the pass: zero limbs, never hits
        for (int j = threadIdx.x; j < 3 * Dp; j += blockDim.x) b0[j] = 0;
        if (threadIdx.x < 4) qconst[q * 4 + threadIdx.x] = 0.f;
        if (threadIdx.x == 0) { ebound[q] = 0.f; thr[q] = INFINITY; cnt[q] = 0; flags[q] = 0; }
        return;
    }
    const float* c0 = consts + (size_t)q * 3 * Dc;
    const float* c1 = c0 + Dc;
    const float* c2 = c1 + Dc;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    double amax = 0.0, C = 0.0, G = 0.0;
    for (int j = threadIdx.x; j < D; j += blockDim.x) {
        const double s = (double)c0[j], m = (double)c1[j], w = (double)c2[j];
        amax = fmax(amax, fabs(s * w));
        C = fma(m, w, C);
        G = fma(fmax(fabs(m), fabs(m + 255.0 * s)), fabs(w), G);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        amax = fmax(amax, __shfl_xor_sync(FPV_FULL_MASK, amax, o));
        C += __shfl_xor_sync(FPV_FULL_MASK, C, o);
        G += __shfl_xor_sync(FPV_FULL_MASK, G, o);
    }
    if (lane == 0) { red[0][warp] = amax; red[1][warp] = C; red[2][warp] = G; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double m = 0.0, c = 0.0, g = 0.0;
        for (int w2 = 0; w2 < W; ++w2) { m = fmax(m, red[0][w2]); c += red[1][w2]; g += red[2][w2]; }
        s_amax = m; s_C = c; s_G = g;
    }
    __syncthreads();
    // c and alpha as the floats the epilogue multiplies with; the limbs are computed against exactly these values
    const float c_f = s_amax > 0.0 ? (float)s_amax * 1.000001f : 0.f;                       // >= max |a_j|
    const float alpha_f = c_f > 0.f ? (float)(2.0 * (double)c_f / 16777215.0) * 1.0000002f : 1.0f;
    const double alpha = (double)alpha_f, cd = (double)c_f;
    for (int j = threadIdx.x; j < Dp; j += blockDim.x) {
        uint32_t Aj = 0;
        if (j < D) {
            const double a = (double)c0[j] * (double)c2[j] + cd;                            // in [0, 2c]
            long long v = __double2ll_rn(a / alpha);
            Aj = (uint32_t)(v < 0 ? 0 : (v > 16777215ll ? 16777215ll : v));
        }
        b0[j] = (uint8_t)(Aj & 0xFFu);
        b0[Dp + j] = (uint8_t)((Aj >> 8) & 0xFFu);
        b0[2 * Dp + j] = (uint8_t)(Aj >> 16);
    }
    if (threadIdx.x == 0) {
        const float rmax = __uint_as_float(max_bits[0]), imax = __uint_as_float(max_bits[1]);
        const float u = 5.9604645e-8f;
        const float absC = (float)fabs(s_C) * 1.000001f;
        const float inner = rmax * alpha_f * 0.5f + u * (10.1f * c_f * rmax + 3.0f * absC);
        float e;
        if (kind == FPV_SQ_DOT) e = 1.25f * (inner + 1.01f * ((float)D / 32.0f + 22.0f) * u * (float)s_G * 1.000001f);
        else e = 1.25f * (inner * imax + ((3.0f * (float)D / 64.0f + 36.0f) * 1.0001f + 5.0f) * u);
        const bool ok = e == e && e < INFINITY;
        qconst[q * 4 + 0] = alpha_f;
        qconst[q * 4 + 1] = -(float)s_C;
        qconst[q * 4 + 2] = 0.f;
        qconst[q * 4 + 3] = c_f;
        ebound[q] = ok ? e : 0.f;
        thr[q] = ok ? -INFINITY : INFINITY;                            // no usable bound: nothing passes, the SIMT scan answers
        cnt[q] = 0;
        flags[q] = ok ? 0u : 1u;
    }
}

// finish for DOT / COSINE: as sq_mma_finish_kernel, the re-score is the loop of sq_scan_kernel<KIND, true> (fpv_sq.cu)
template <int KIND>
__global__ void __launch_bounds__(256) sq_mma_finish_dc_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt,
                                                               const float* __restrict__ thr, const float* __restrict__ ebound,
                                                               uint32_t* __restrict__ flags, const float* __restrict__ consts,
                                                               const uint8_t* __restrict__ codes, int D, int Dc, int k,
                                                               int64_t id_base, float* __restrict__ out_dist,
                                                               int64_t* __restrict__ out_idx, int32_t* __restrict__ out_count) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);                         // [SQM_CAP]
    uint64_t* sel = keys + SQM_CAP;                                               // [SQM_RMAX]
    float* cs = reinterpret_cast<float*>(sel + SQM_RMAX);                         // [3][Dc]
    __shared__ uint32_t hist[256];
    __shared__ int s_bin, s_need, s_R, s_flag;
    const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const uint32_t c_raw = cnt[q];
    const int c = (int)min(c_raw, (uint32_t)SQM_CAP);
    const uint64_t* mine = cand + (size_t)q * SQM_CAP;
    for (int i = threadIdx.x; i < c; i += blockDim.x) keys[i] = mine[i];
    for (int j = threadIdx.x; j < 3 * Dc; j += blockDim.x) cs[j] = consts[(size_t)q * 3 * Dc + j];
    if (threadIdx.x == 0) { s_R = 0; s_flag = (c_raw > (uint32_t)SQM_CAP) || flags[q] != 0; }
    __syncthreads();
    const float t = thr[q], E = ebound[q];
    float a_k = INFINITY;
    if (c >= k) a_k = ordered_to_f32((uint32_t)(block_radix_select(keys, c, k, hist, &s_bin, &s_need) >> 32));
    const float limit = a_k + 2.0f * E;
    const bool certified = (t == -INFINITY) || (limit <= -t);
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const uint64_t key = keys[i];
        if (ordered_to_f32((uint32_t)(key >> 32)) <= limit) {
            const int pos = atomicAdd(&s_R, 1);
            if (pos < SQM_RMAX) sel[pos] = key;
        }
    }
    __syncthreads();
    const int R = s_R;
    if (threadIdx.x == 0) {
        if (!certified || R > SQM_RMAX) s_flag = 1;
        flags[q] = s_flag;
    }
    __syncthreads();
    if (s_flag) return;                                   // the SIMT scan answers this query (fpv_sq.cu, gated on flags)
    const float4* c0 = reinterpret_cast<const float4*>(cs);
    const float4* c1 = reinterpret_cast<const float4*>(cs + Dc);
    const float4* c2 = reinterpret_cast<const float4*>(cs + 2 * Dc);
    const int nchunk = Dc >> 4;
    for (int i = warp; i < R; i += W) {
        const uint32_t row = (uint32_t)sel[i];
        const uint4* rowp = reinterpret_cast<const uint4*>(codes + (size_t)row * D);
        float acc = 0.f, nrm = 0.f;
        for (int ch = lane; ch < nchunk; ch += 32) {
            const uint4 w = ldg_nc_u4(rowp + ch);
            const int f4 = ch * 4;
            sq_word<KIND>(w.x, c0[f4], c1[f4], c2[f4], acc, nrm);
            sq_word<KIND>(w.y, c0[f4 + 1], c1[f4 + 1], c2[f4 + 1], acc, nrm);
            sq_word<KIND>(w.z, c0[f4 + 2], c1[f4 + 2], c2[f4 + 2], acc, nrm);
            sq_word<KIND>(w.w, c0[f4 + 3], c1[f4 + 3], c2[f4 + 3], acc, nrm);
        }
        acc = warp_sum(acc);
        float d;
        if (KIND == FPV_SQ_DOT) d = -acc;
        else d = 1.0f - acc / (sqrtf(warp_sum(nrm)) + 1e-8f);
        if (lane == 0) keys[i] = make_key(d, row);
    }
    int P2 = 2; while (P2 < R) P2 <<= 1;
    __syncthreads();
    for (int i = R + threadIdx.x; i < P2; i += blockDim.x) keys[i] = FPV_KEY_MAX;
    __syncthreads();
    block_bitonic_sort(keys, P2);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = i < R;
        const uint64_t key = ok ? keys[i] : FPV_KEY_MAX;
        out_dist[(size_t)q * k + i] = ok ? ordered_to_f32((uint32_t)(key >> 32)) : INFINITY;
        out_idx[(size_t)q * k + i] = ok ? id_base + (int64_t)(uint32_t)key : -1;
    }
    if (out_count && threadIdx.x == 0) out_count[q] = min(R, k);
}

// ---- finish: window = a_k + 2E, exact re-score with the arithmetic of the SIMT scan (fpv_sq.cu), sort, emit ----------
__device__ __forceinline__ float sqm_u8f(uint32_t w, int b) { return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7540u + b)); }

__global__ void __launch_bounds__(256) sq_mma_finish_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt,
                                                            const float* __restrict__ thr, const float* __restrict__ ebound,
                                                            uint32_t* __restrict__ flags, const uint8_t* __restrict__ qcodes,
                                                            const float* __restrict__ scale, const uint8_t* __restrict__ codes,
                                                            int D, int k, int64_t id_base, float* __restrict__ out_dist,
                                                            int64_t* __restrict__ out_idx, int32_t* __restrict__ out_count) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);                         // [SQM_CAP]
    uint64_t* sel = keys + SQM_CAP;                                               // [SQM_RMAX]
    const int Dp = (D + 15) / 16 * 16;
    float* c1 = reinterpret_cast<float*>(sel + SQM_RMAX);                         // [Dp]  s_j
    float* c2 = c1 + Dp;                                                          // [Dp]  -2^23 s_j
    uint32_t* qw = reinterpret_cast<uint32_t*>(c2 + Dp);                          // [Dp / 4] query code words
    __shared__ uint32_t hist[256];
    __shared__ int s_bin, s_need, s_R, s_flag;
    const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const uint32_t c_raw = cnt[q];
    const int c = (int)min(c_raw, (uint32_t)SQM_CAP);
    const uint64_t* mine = cand + (size_t)q * SQM_CAP;
    for (int i = threadIdx.x; i < c; i += blockDim.x) keys[i] = mine[i];
    for (int j = threadIdx.x; j < Dp; j += blockDim.x) {
        const float s = j < D ? __fdiv_rn(scale[j], 255.0f) : 0.f;
        c1[j] = s;
        c2[j] = -8388608.0f * s;
    }
    for (int j = threadIdx.x; j < Dp / 4; j += blockDim.x) {
        uint32_t w = 0;
        for (int b = 0; b < 4; ++b) { const int jj = 4 * j + b; if (jj < D) w |= (uint32_t)qcodes[(size_t)q * D + jj] << (8 * b); }
        qw[j] = w;
    }
    if (threadIdx.x == 0) { s_R = 0; s_flag = (c_raw > (uint32_t)SQM_CAP) || flags[q] != 0; }
    __syncthreads();
    const float t = thr[q], E = ebound[q];
    float a_k = INFINITY;
    if (c >= k) a_k = ordered_to_f32((uint32_t)(block_radix_select(keys, c, k, hist, &s_bin, &s_need) >> 32));
    const float limit = a_k + 2.0f * E;
    const bool certified = (t == -INFINITY) || (limit <= -t);
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const uint64_t key = keys[i];
        if (ordered_to_f32((uint32_t)(key >> 32)) <= limit) {
            const int pos = atomicAdd(&s_R, 1);
            if (pos < SQM_RMAX) sel[pos] = key;
        }
    }
    __syncthreads();
    const int R = s_R;
    if (threadIdx.x == 0) {
        if (!certified || R > SQM_RMAX) s_flag = 1;
        flags[q] = s_flag;
    }
    __syncthreads();
    if (s_flag) return;                                   // the SIMT scan answers this query (fpv_sq.cu, gated on flags)
    const int nchunk = Dp >> 4;
    for (int i = warp; i < R; i += W) {
        const uint32_t row = (uint32_t)sel[i];
        const uint4* rowp = reinterpret_cast<const uint4*>(codes + (size_t)row * D);
        float a = 0.f, a2 = 0.f;                          // the two chains of sq_l2_tma_kernel, same element order
        for (int ch = lane; ch < nchunk; ch += 32) {
            const uint4 w4 = ldg_nc_u4(rowp + ch);
            const uint32_t ws[4] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const uint32_t ad = __vabsdiffu4(qw[ch * 4 + u], ws[u]);
#pragma unroll
                for (int b = 0; b < 4; ++b) {
                    const int j = ch * 16 + u * 4 + b;
                    const float tt = fmaf(sqm_u8f(ad, b), c1[j], c2[j]);
                    if (b & 1) a2 = fmaf(tt, tt, a2); else a = fmaf(tt, tt, a);
                }
            }
        }
        const float acc = warp_sum(a + a2);
        if (lane == 0) keys[i] = make_key(sqrtf(acc), row);
    }
    int P2 = 2; while (P2 < R) P2 <<= 1;
    __syncthreads();
    for (int i = R + threadIdx.x; i < P2; i += blockDim.x) keys[i] = FPV_KEY_MAX;
    __syncthreads();
    block_bitonic_sort(keys, P2);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = i < R;
        const uint64_t key = ok ? keys[i] : FPV_KEY_MAX;
        out_dist[(size_t)q * k + i] = ok ? ordered_to_f32((uint32_t)(key >> 32)) : INFINITY;
        out_idx[(size_t)q * k + i] = ok ? id_base + (int64_t)(uint32_t)key : -1;
    }
    if (out_count && threadIdx.x == 0) out_count[q] = min(R, k);
}

// the limb dots themselves, for the bit-exactness test: out [3][N] int32 for ONE query (limb l of row r at out[l*N + r])
__global__ void __launch_bounds__(256) sq_limb_dump_kernel(const uint8_t* __restrict__ bmat, int Dp, const uint8_t* __restrict__ codes,
                                                           int64_t N, int D, int32_t* __restrict__ out) {
    // reference evaluation on CUDA cores of what the MMA must produce; only used to cross-check the MMA (debug entry)
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    for (int64_t row = (int64_t)blockIdx.x * W + warp; row < N; row += (int64_t)gridDim.x * W) {
        int s0 = 0, s1 = 0, s2 = 0;
        for (int j = lane; j < D; j += 32) {
            const int b = codes[row * D + j];
            s0 += b * bmat[j]; s1 += b * bmat[Dp + j]; s2 += b * bmat[2 * Dp + j];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            s0 += __shfl_xor_sync(FPV_FULL_MASK, s0, o); s1 += __shfl_xor_sync(FPV_FULL_MASK, s1, o); s2 += __shfl_xor_sync(FPV_FULL_MASK, s2, o);
        }
        if (lane == 0) { out[row] = s0; out[N + row] = s1; out[2 * N + row] = s2; }
    }
}

struct SqmPlan { int Dp, Dc, passes; size_t off_bmat, off_qconst, off_eb, off_thr, off_cnt, off_flags, off_cand, off_consts, off_scan, scan_bytes, total; };

size_t sq_flagged_workspace(int64_t Q, int64_t N, int D, int k);
int sq_topk_flagged(int kind, const uint8_t* qcodes, int64_t q, const uint8_t* codes, int64_t n, int d, const float* min_vals,
                    const float* scale, int k, const uint32_t* mask_words, int64_t id_base, float* out_dist, int64_t* out_idx,
                    int32_t* out_count, const uint32_t* only_flagged, void* ws, size_t ws_bytes, cudaStream_t st);

static SqmPlan plan_sqm(int64_t Q, int64_t N, int D, int k) {
    SqmPlan pl{};
    pl.Dp = (D + SQM_KROW - 1) / SQM_KROW * SQM_KROW;
    pl.passes = (int)((Q + SQM_QB - 1) / SQM_QB);
    const int64_t Qp = (int64_t)pl.passes * SQM_QB;
    size_t o = 0;
    pl.off_bmat = o;   o += align_up((size_t)Qp * 3 * pl.Dp, 1024);
    pl.off_qconst = o; o += align_up((size_t)Qp * 16, 256);
    pl.off_eb = o;     o += align_up((size_t)Qp * 4, 256);
    pl.off_thr = o;    o += align_up((size_t)Qp * 4, 256);
    pl.off_cnt = o;    o += align_up((size_t)Qp * 4, 256);
    pl.off_flags = o;  o += align_up((size_t)Qp * 4, 256);
    pl.off_cand = o;   o += (size_t)Qp * SQM_CAP * 8;
    pl.Dc = (D + 15) / 16 * 16;
    pl.off_consts = o; o += align_up((size_t)SQM_QB * 3 * pl.Dc * 4, 256);       // DOT / COSINE: the scan's constants, one pass
    pl.off_scan = o;
    pl.scan_bytes = sq_flagged_workspace(Q, N, D, k);
    pl.total = o + pl.scan_bytes;
    return pl;
}

typedef CUresult (*SqmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int sqm_make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    static SqmEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<SqmEncodeFn>(p);
    }
    if (!fn) { set_error("sq_mma: cuTensorMapEncodeTiled entry point not available"); return FPV_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld};
    cuuint32_t box[2] = {(cuuint32_t)SQM_KROW, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("sq_mma: cuTensorMapEncodeTiled failed with %d", (int)r); return FPV_ERR_CUDA; }
    return FPV_OK;
}

}  // namespace fpv

using namespace fpv;

// C_row for every row of a code matrix (+ the maximum, a device float read by later searches): index build.
extern "C" int fpv_sq_row_term(const uint8_t* codes, int64_t n, int d, const float* scale, float* row_term, float* row_term_max,
                               void* stream) {
    FPV_REQUIRE(n >= 0 && d >= 1 && d <= 16384, "sq_row_term: bad shape n=%lld d=%d", (long long)n, d);
    FPV_REQUIRE(scale && row_term_max && (n == 0 || (codes && row_term)), "sq_row_term: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    FPV_CUDA(cudaMemsetAsync(row_term_max, 0, 4, st));
    if (n == 0) return FPV_OK;
    const int64_t blocks = std::min<int64_t>((n + 7) / 8, (int64_t)sm_count() * 8);
    const size_t smem = (size_t)d * 8;
    if (smem > 48 * 1024) FPV_CUDA(cudaFuncSetAttribute(sq_row_term_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sq_row_term_kernel<<<(unsigned)blocks, 256, smem, st>>>(codes, n, d, scale, row_term, reinterpret_cast<uint32_t*>(row_term_max));
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

// R_row = sum_j b_j and invn_row = 1 / (|decode(row)| + 1e-8) for every row, + their maxima (two device floats read by
// later searches): index build work for the DOT / COSINE tensor-core scans.
extern "C" int fpv_sq_row_terms_dc(const uint8_t* codes, int64_t n, int d, const float* min_vals, const float* scale,
                                   float* row_sum, float* row_invn, float* maxima, void* stream) {
    FPV_REQUIRE(n >= 0 && d >= 1 && d <= 16384, "sq_row_terms_dc: bad shape n=%lld d=%d", (long long)n, d);
    FPV_REQUIRE(min_vals && scale && maxima && (n == 0 || (codes && row_sum && row_invn)), "sq_row_terms_dc: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    FPV_CUDA(cudaMemsetAsync(maxima, 0, 8, st));
    if (n == 0) return FPV_OK;
    const int64_t blocks = std::min<int64_t>((n + 7) / 8, (int64_t)sm_count() * 8);
    const size_t smem = (size_t)d * 16;
    if (smem > 48 * 1024) FPV_CUDA(cudaFuncSetAttribute(sq_row_terms_dc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    sq_row_terms_dc_kernel<<<(unsigned)blocks, 256, smem, st>>>(codes, n, d, min_vals, scale, row_sum, row_invn,
                                                                reinterpret_cast<uint32_t*>(maxima));
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

extern "C" int fpv_sq_mma_supported(int64_t n, int d, int k) {
    return n >= 65536 && n < (1ll << 31) && d >= 16 && d % 16 == 0 && d <= SQM_MAX_KB * SQM_KROW && k >= 1 && k <= FPV_MAX_K &&
           4 * k <= SQM_CAP;
}

extern "C" size_t fpv_sq_mma_workspace(int64_t q, int64_t n, int d, int k) {
    if (q <= 0 || d <= 0 || k <= 0) return 256;
    return plan_sqm(q, n, d, k).total;
}

// Batched uint8-scalar L2 top-k on the int8 tensor cores.  qcodes [q][d] = the re-quantised queries
// (ScalarQuantizer.encode_query), codes [n][d], row_term / row_term_max from fpv_sq_row_term for THESE codes and scale.
// min_vals is only used by the SIMT fallback.  Results are those of fpv_sq_topk(FPV_SQ_L2) bit for bit.
static int sqm_run(int kind, const uint8_t* qcodes, int64_t q, const uint8_t* codes, int64_t n, int d, const float* min_vals,
                   const float* scale, const float* row_term, const float* row_term2, const float* row_term_max, int k,
                   const uint32_t* mask_words, int64_t id_base, float* out_dist, int64_t* out_idx,
                   int32_t* out_count, void* ws, size_t ws_bytes, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FPV_REQUIRE(kind == FPV_SQ_L2 || kind == FPV_SQ_DOT || kind == FPV_SQ_COSINE, "sq_mma: unknown kind %d", kind);
    FPV_REQUIRE(kind != FPV_SQ_COSINE || row_term2, "sq_mma: the cosine scan needs the per-row inverse norms");
    FPV_REQUIRE(q >= 1 && q <= 65535, "sq_mma: q=%lld outside [1,65535]", (long long)q);
    FPV_REQUIRE(fpv_sq_mma_supported(n, d, k), "sq_mma: unsupported shape n=%lld d=%d k=%d", (long long)n, d, k);
    FPV_REQUIRE(qcodes && codes && min_vals && scale && row_term && row_term_max && out_dist && out_idx, "sq_mma: null pointer");
    FPV_REQUIRE((reinterpret_cast<uintptr_t>(codes) & 15) == 0, "sq_mma: codes must be 16-byte aligned");
    SqmPlan pl = plan_sqm(q, n, d, k);
    if (!ws || ws_bytes < pl.total) { set_error("sq_mma: workspace %zu < %zu", ws_bytes, pl.total); return FPV_ERR_WORKSPACE; }
    FPV_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "sq_mma: workspace must be 256-byte aligned");
    char* w = static_cast<char*>(ws);
    static std::mutex attr_mutex;
    static bool attr_set[32];
    typedef void (*MainKernel)(const CUtensorMap, const CUtensorMap, SqmParams);
    const MainKernel main_kernel = kind == FPV_SQ_L2 ? sq_mma_kernel<FPV_SQ_L2> : kind == FPV_SQ_DOT ? sq_mma_kernel<FPV_SQ_DOT>
                                                                                                    : sq_mma_kernel<FPV_SQ_COSINE>;
    int dev_id = 0;
    FPV_CUDA(cudaGetDevice(&dev_id));
    {
        std::unique_lock<std::mutex> lk(attr_mutex);
        if (dev_id < 0 || dev_id >= 32 || !attr_set[dev_id]) {
            FPV_CUDA(cudaFuncSetAttribute(sq_mma_kernel<FPV_SQ_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SQM_SMEM));
            FPV_CUDA(cudaFuncSetAttribute(sq_mma_kernel<FPV_SQ_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SQM_SMEM));
            FPV_CUDA(cudaFuncSetAttribute(sq_mma_kernel<FPV_SQ_COSINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SQM_SMEM));
            FPV_CUDA(cudaFuncSetAttribute(tighten_kernel<SQM_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, SQM_CAP * 8));
            if (dev_id >= 0 && dev_id < 32) attr_set[dev_id] = true;
        }
    }
    CUtensorMap tmA;
    int rc = sqm_make_map(&tmA, codes, n, d, d, SQM_BM);
    if (rc != FPV_OK) return rc;
    const int nkb = pl.Dp / SQM_KROW;
    const int64_t tiles_total = (n + SQM_BM - 1) / SQM_BM;
    const size_t fin_smem = (size_t)(SQM_CAP + SQM_RMAX) * 8 + (size_t)((d + 15) / 16 * 16) * 9 + 64;
    const size_t fin_dc_smem = (size_t)(SQM_CAP + SQM_RMAX) * 8 + (size_t)3 * pl.Dc * 4;
    if (kind == FPV_SQ_L2) FPV_CUDA(cudaFuncSetAttribute(sq_mma_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
    else if (kind == FPV_SQ_DOT)
        FPV_CUDA(cudaFuncSetAttribute(sq_mma_finish_dc_kernel<FPV_SQ_DOT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_dc_smem));
    else
        FPV_CUDA(cudaFuncSetAttribute(sq_mma_finish_dc_kernel<FPV_SQ_COSINE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_dc_smem));
    float* consts = reinterpret_cast<float*>(w + pl.off_consts);
    for (int pass = 0; pass < pl.passes; ++pass) {
        const int64_t q0 = (int64_t)pass * SQM_QB;
        const int nq = (int)std::min<int64_t>(SQM_QB, q - q0);
        uint8_t* bmat = reinterpret_cast<uint8_t*>(w + pl.off_bmat) + (size_t)q0 * 3 * pl.Dp;
        float* qconst = reinterpret_cast<float*>(w + pl.off_qconst) + q0 * 4;
        float* eb = reinterpret_cast<float*>(w + pl.off_eb) + q0;
        float* thr = reinterpret_cast<float*>(w + pl.off_thr) + q0;
        uint32_t* cnt = reinterpret_cast<uint32_t*>(w + pl.off_cnt) + q0;
        uint32_t* flags = reinterpret_cast<uint32_t*>(w + pl.off_flags) + q0;
        uint64_t* cand = reinterpret_cast<uint64_t*>(w + pl.off_cand) + (size_t)q0 * SQM_CAP;
        if (kind == FPV_SQ_L2) {
            sq_mma_prep_kernel<<<SQM_QB, 256, 0, st>>>(qcodes + q0 * d, nq, d, pl.Dp, scale, reinterpret_cast<const uint32_t*>(row_term_max),
                                                       bmat, qconst, eb, thr, cnt, flags);
            FPV_LAUNCH_CHECK();
        } else {
            rc = sq_prep_launch(kind, qcodes + q0 * d, nq, d, pl.Dc, min_vals, scale, consts, st);
            if (rc != FPV_OK) return rc;
            sq_mma_prep_dc_kernel<<<SQM_QB, 256, 0, st>>>(kind, consts, nq, d, pl.Dc, pl.Dp, reinterpret_cast<const uint32_t*>(row_term_max),
                                                          bmat, qconst, eb, thr, cnt, flags);
            FPV_LAUNCH_CHECK();
        }
        CUtensorMap tmB;
        rc = sqm_make_map(&tmB, bmat, SQM_BN, pl.Dp, pl.Dp, SQM_BN);
        if (rc != FPV_OK) return rc;
        SqmParams p{};
        p.row_term = row_term; p.row_term2 = row_term2; p.mask = mask_words; p.qconst = qconst; p.thr = thr; p.cnt = cnt; p.cand = cand;
        p.N = n; p.nq = nq; p.nkb = nkb;
        // Scan order: a fixed permutation of the tiles (stride coprime to their number), so that the dense first slab --
        // the source of the first threshold -- is spread over the whole matrix instead of being its first 8192 rows (on
        // data stored in cluster or time order a prefix gives a useless threshold and the next slab overflows the lists).
        p.tiles_total = (int)tiles_total;
        {
            int64_t st = std::max<int64_t>(1, tiles_total / 64) | 1;
            auto gcd = [](int64_t a, int64_t b) { while (b) { const int64_t t = a % b; a = b; b = t; } return a; };
            while (gcd(st, tiles_total) != 1) st += 2;
            p.perm_stride = (int)(st % tiles_total == 0 ? 1 : st % tiles_total);
            if (tiles_total == 1) p.perm_stride = 1;
            static int perm_on = -1;                       // FPV_SQ_PERMUTE=0: scan in row order (A/B measurements)
            if (perm_on < 0) { const char* e = getenv("FPV_SQ_PERMUTE"); perm_on = (e && e[0] == '0') ? 0 : 1; }
            if (!perm_on) p.perm_stride = 1;
        }
        // slabs: the first one (<= 8192 rows) is dense -- every row is a candidate -- then each slab may be as large as
        // keeps the expected number of rows under the tightened threshold (~ slab * k / rows_seen, the window 2E is a
        // few 1e-6 of d^2) within half the candidate slots
        int64_t done = 0, slab = std::max<int64_t>(8192, 4 * (int64_t)k) / SQM_BM;
        const double growth = (double)(SQM_CAP / 2) / (1.5 * k);
        while (done < tiles_total) {
            int64_t take = std::min<int64_t>(slab, tiles_total - done);
            if (tiles_total - done - take < take / 2) take = tiles_total - done;
            p.tile0 = (int)done; p.ntiles = (int)take;
            const unsigned grid = (unsigned)std::min<int64_t>(take, sm_count());
            main_kernel<<<grid, SQM_THREADS, SQM_SMEM, st>>>(tmA, tmB, p);
            FPV_LAUNCH_CHECK();
            done += take;
            if (done < tiles_total) {
                tighten_kernel<SQM_CAP><<<(unsigned)nq, 256, SQM_CAP * 8, st>>>(cand, cnt, thr, eb, flags, k, nullptr);
                FPV_LAUNCH_CHECK();
            }
            slab = (int64_t)((double)done * growth);
            if (slab < 1) slab = 1;
        }
        if (kind == FPV_SQ_L2)
            sq_mma_finish_kernel<<<(unsigned)nq, 256, fin_smem, st>>>(cand, cnt, thr, eb, flags, qcodes + q0 * d, scale, codes, d, k, id_base,
                                                                       out_dist + q0 * k, out_idx + q0 * k, out_count ? out_count + q0 : nullptr);
        else if (kind == FPV_SQ_DOT)
            sq_mma_finish_dc_kernel<FPV_SQ_DOT><<<(unsigned)nq, 256, fin_dc_smem, st>>>(cand, cnt, thr, eb, flags, consts, codes, d, pl.Dc, k,
                id_base, out_dist + q0 * k, out_idx + q0 * k, out_count ? out_count + q0 : nullptr);
        else
            sq_mma_finish_dc_kernel<FPV_SQ_COSINE><<<(unsigned)nq, 256, fin_dc_smem, st>>>(cand, cnt, thr, eb, flags, consts, codes, d, pl.Dc, k,
                id_base, out_dist + q0 * k, out_idx + q0 * k, out_count ? out_count + q0 : nullptr);
        FPV_LAUNCH_CHECK();
    }
    // queries whose window did not fit (ties, adversarial data): the SIMT scan, gated on the device-side flags
    return sq_topk_flagged(kind, qcodes, q, codes, n, d, min_vals, scale, k, mask_words, id_base, out_dist, out_idx, out_count,
                           reinterpret_cast<const uint32_t*>(w + pl.off_flags), w + pl.off_scan, pl.scan_bytes, st);
}

extern "C" int fpv_sq_l2_mma_topk(const uint8_t* qcodes, int64_t q, const uint8_t* codes, int64_t n, int d, const float* min_vals,
                                  const float* scale, const float* row_term, const float* row_term_max, int k,
                                  const uint32_t* mask_words, int64_t id_base, float* out_dist, int64_t* out_idx,
                                  int32_t* out_count, void* ws, size_t ws_bytes, void* stream) {
    return sqm_run(FPV_SQ_L2, qcodes, q, codes, n, d, min_vals, scale, row_term, nullptr, row_term_max, k, mask_words, id_base, out_dist,
                   out_idx, out_count, ws, ws_bytes, stream);
}

// The same for distances_dot / distances_cosine (kind = FPV_SQ_DOT / FPV_SQ_COSINE): row_sum / row_invn / maxima from
// fpv_sq_row_terms_dc for THESE codes, min_vals and scale.  Results are those of fpv_sq_topk(kind) bit for bit.
extern "C" int fpv_sq_dc_mma_topk(int kind, const uint8_t* qcodes, int64_t q, const uint8_t* codes, int64_t n, int d,
                                  const float* min_vals, const float* scale, const float* row_sum, const float* row_invn,
                                  const float* maxima, int k, const uint32_t* mask_words, int64_t id_base, float* out_dist,
                                  int64_t* out_idx, int32_t* out_count, void* ws, size_t ws_bytes, void* stream) {
    FPV_REQUIRE(kind == FPV_SQ_DOT || kind == FPV_SQ_COSINE, "sq_dc_mma: kind must be FPV_SQ_DOT or FPV_SQ_COSINE");
    FPV_REQUIRE(row_sum && row_invn && maxima, "sq_dc_mma: null pointer");
    return sqm_run(kind, qcodes, q, codes, n, d, min_vals, scale, row_sum, row_invn, maxima, k, mask_words, id_base, out_dist, out_idx,
                   out_count, ws, ws_bytes, stream);
}

extern "C" size_t fpv_sq_mma_flags_offset(int64_t q, int64_t n, int d, int k) {
    if (q <= 0 || d <= 0 || k <= 0) return 0;
    return plan_sqm(q, n, d, k).off_flags;
}

// Test hook: the three limb dots of ONE query against every row, once from the tensor cores (out_mma [3][n] int32) and
// once from a plain CUDA-core loop (out_simt); limbs_out [3][Dp] receives the limb rows themselves so that the test can
// redo the dot products in int64 on the host (oracle.sq_limb_dots_int64).  Dp = d rounded up to 128.
extern "C" int fpv_sq_mma_limb_dots(const uint8_t* qcodes, const uint8_t* codes, int64_t n, int d, const float* scale,
                                    uint8_t* limbs_out, int32_t* out_mma, int32_t* out_simt, void* ws, size_t ws_bytes,
                                    void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    FPV_REQUIRE(n >= 1 && n < (1ll << 31) && d >= 16 && d % 16 == 0 && d <= SQM_MAX_KB * SQM_KROW, "sq_mma_limb_dots: bad shape");
    FPV_REQUIRE(qcodes && codes && scale && limbs_out && out_mma && out_simt, "sq_mma_limb_dots: null pointer");
    SqmPlan pl = plan_sqm(1, n, d, 1);
    if (!ws || ws_bytes < pl.total) { set_error("sq_mma_limb_dots: workspace %zu < %zu", ws_bytes, pl.total); return FPV_ERR_WORKSPACE; }
    char* w = static_cast<char*>(ws);
    uint8_t* bmat = reinterpret_cast<uint8_t*>(w + pl.off_bmat);
    float* qconst = reinterpret_cast<float*>(w + pl.off_qconst);
    float* eb = reinterpret_cast<float*>(w + pl.off_eb);
    float* thr = reinterpret_cast<float*>(w + pl.off_thr);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(w + pl.off_cnt);
    uint32_t* flags = reinterpret_cast<uint32_t*>(w + pl.off_flags);
    FPV_CUDA(cudaMemsetAsync(eb, 0, 4, st));                 // stands in for row_term_max (only the error bound uses it)
    sq_mma_prep_kernel<<<SQM_QB, 256, 0, st>>>(qcodes, 1, d, pl.Dp, scale, reinterpret_cast<const uint32_t*>(eb), bmat, qconst,
                                               eb, thr, cnt, flags);
    FPV_LAUNCH_CHECK();
    FPV_CUDA(cudaMemcpyAsync(limbs_out, bmat, (size_t)3 * pl.Dp, cudaMemcpyDeviceToDevice, st));
    FPV_CUDA(cudaFuncSetAttribute(sq_mma_kernel<FPV_SQ_L2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SQM_SMEM));
    CUtensorMap tmA, tmB;
    int rc = sqm_make_map(&tmA, codes, n, d, d, SQM_BM);
    if (rc != FPV_OK) return rc;
    rc = sqm_make_map(&tmB, bmat, SQM_BN, pl.Dp, pl.Dp, SQM_BN);
    if (rc != FPV_OK) return rc;
    SqmParams p{};
    p.row_term = reinterpret_cast<const float*>(out_simt);   // any readable [n] array: the dump path ignores its values
    p.qconst = qconst; p.thr = thr; p.cnt = cnt; p.cand = reinterpret_cast<uint64_t*>(w + pl.off_cand);
    p.N = n; p.nq = 1; p.nkb = pl.Dp / SQM_KROW; p.dump = out_mma;
    const int64_t tiles = (n + SQM_BM - 1) / SQM_BM;
    p.tile0 = 0; p.ntiles = (int)tiles; p.perm_stride = 1; p.tiles_total = (int)tiles;
    sq_mma_kernel<FPV_SQ_L2><<<(unsigned)std::min<int64_t>(tiles, sm_count()), SQM_THREADS, SQM_SMEM, st>>>(tmA, tmB, p);
    FPV_LAUNCH_CHECK();
    sq_limb_dump_kernel<<<(unsigned)std::min<int64_t>((n + 7) / 8, (int64_t)sm_count() * 8), 256, 0, st>>>(bmat, pl.Dp, codes, n, d, out_simt);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}
