// Scalar (uint8, min-max, truncating) quantizer: encode and the three quantized distance scans with fused top-k.
//
// Replaces ScalarQuantizer.encode (quantization.py:118-126), distances_l2 -> _sq_distances_l2_vectorized
// (:151-152, :217-236), distances_dot -> _sq_distances_dot_vectorized (:180-181, :239-251) and
// distances_cosine (:161-174).  The reference materialises N x D int16 and fp32 temporaries; here a warp
// streams one row of raw uint8 codes (128-bit loads, 512 codes per load instruction) and keeps everything in
// registers.  uint8 -> float uses the 0x4B000000 byte-permute trick (one PRMT, exact) instead of I2F.
//
// Per-dimension constants (computed once per query by sq_prep_kernel, kept in shared memory by the scan):
//   L2     c0 = qcode + 2^23,  c1 = scale/255                 d = sqrt(sum(((c0 - (2^23+code)) * c1)^2))
//   DOT    c0 = scale/255, c1 = min, c2 = decode(qcode)        d = -sum((code*c0 + c1) * c2)
//   COSINE c0, c1 as DOT, c2 = decode(qcode)/(|.|+1e-8)        d = 1 - sum(dec*c2)/(sqrt(sum(dec^2))+1e-8)
#include <algorithm>

#include "fpv_common.cuh"
#include "fpv_sq_common.cuh"

namespace fpv {

__global__ void sq_encode_kernel(const float* __restrict__ v, int64_t N, int D, int64_t ld,
                                 const float* __restrict__ mn, const float* __restrict__ sc, uint8_t* __restrict__ out) {
    const int64_t total = N * D;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = i / D;
        int j = (int)(i - row * D);
        float unit = __fdiv_rn(__fsub_rn(v[row * ld + j], mn[j]), sc[j]);   // (v - min) / scale
        float x = __fmul_rn(unit, 255.0f);
        x = fminf(fmaxf(x, 0.0f), 255.0f);                                   // np.clip
        out[i] = (uint8_t)(int)x;                                            // astype(uint8): truncation
    }
}

// D % 4 == 0, 16-byte aligned rows and parameters: four elements per thread (128-bit loads, one 32-bit store), 32-bit
// index arithmetic.  The element-per-thread form above divides a 64-bit index per element and runs at 0.29 of HBM.
// Same operations per element (IEEE division included): same codes.
__global__ void __launch_bounds__(256) sq_encode_vec_kernel(const float* __restrict__ v, uint32_t N, uint32_t Dq, int64_t ld,
                                                            const float* __restrict__ mn, const float* __restrict__ sc,
                                                            uint8_t* __restrict__ out) {
    const uint32_t total = N * Dq;                                           // < 2^32 (checked on the host)
    for (uint64_t i64 = blockIdx.x * blockDim.x + threadIdx.x; i64 < total; i64 += gridDim.x * blockDim.x) {   // no 32-bit wrap
        const uint32_t i = (uint32_t)i64;
        const uint32_t row = i / Dq, jq = i - row * Dq;
        const float4 x = ldg_nc_f4(reinterpret_cast<const float4*>(v + (int64_t)row * ld) + jq);
        const float4 m4 = __ldg(reinterpret_cast<const float4*>(mn) + jq);
        const float4 s4 = __ldg(reinterpret_cast<const float4*>(sc) + jq);
        const float xs[4] = {x.x, x.y, x.z, x.w}, ms[4] = {m4.x, m4.y, m4.z, m4.w}, ss[4] = {s4.x, s4.y, s4.z, s4.w};
        uint32_t packed = 0;
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            float y = __fmul_rn(__fdiv_rn(__fsub_rn(xs[e], ms[e]), ss[e]), 255.0f);      // (v - min) / scale * 255
            y = fminf(fmaxf(y, 0.0f), 255.0f);                                           // np.clip
            packed |= (uint32_t)(int)y << (8 * e);                                       // astype(uint8): truncation
        }
        reinterpret_cast<uint32_t*>(out)[(size_t)row * Dq + jq] = packed;
    }
}

// consts [Q][3][Dp] (Dp = D rounded up to 16, zero padded so padded columns contribute nothing)
__global__ void sq_prep_kernel(int kind, const uint8_t* __restrict__ qcodes, int D, int Dp,
                               const float* __restrict__ mn, const float* __restrict__ sc, float* __restrict__ consts) {
    const int64_t q = blockIdx.x;
    float* c0 = consts + (size_t)q * 3 * Dp;
    float* c1 = c0 + Dp;
    float* c2 = c1 + Dp;
    __shared__ float red[32];
    __shared__ float s_inv;
    float nrm = 0.f;
    for (int j = threadIdx.x; j < Dp; j += blockDim.x) {
        float a = 0.f, b = 0.f, c = 0.f;
        if (j < D) {
            float qc = (float)qcodes[q * D + j];
            float s255 = __fdiv_rn(sc[j], 255.0f);                           // scale / 255.0
            // L2: c0 = 2^23 + code (exact u8 <-> f32 trick), c1 = scale/255, c2 = -2^23 * c1 (exact: power-of-two scaling)
            if (kind == FPV_SQ_L2) { a = qc + 8388608.0f; b = s255; c = -8388608.0f * s255; }
            else {
                float qr = __fadd_rn(__fmul_rn(__fdiv_rn(qc, 255.0f), sc[j]), mn[j]);   // decode (quantization.py:136-137)
                a = s255; b = mn[j]; c = qr;
                nrm = fmaf(qr, qr, nrm);
            }
        } else if (kind == FPV_SQ_L2) {
            a = 8388608.0f;     // padded code bytes are 0 -> diff 0
        }
        c0[j] = a; c1[j] = b; c2[j] = c;
    }
    if (kind == FPV_SQ_COSINE) {
        nrm = warp_sum(nrm);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = nrm;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[w];
            s_inv = 1.0f / (sqrtf(t) + 1e-8f);
        }
        __syncthreads();
        for (int j = threadIdx.x; j < D; j += blockDim.x) c2[j] *= s_inv;
    }
}

// per-query constants of the DOT / COSINE (and L2) scans for the tensor-core path's finish kernel (fpv_sq_mma.cu)
int sq_prep_launch(int kind, const uint8_t* qcodes, int64_t q, int D, int Dp, const float* mn, const float* sc, float* consts,
                   cudaStream_t st) {
    sq_prep_kernel<<<(unsigned)q, 256, 0, st>>>(kind, qcodes, D, Dp, mn, sc, consts);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

struct SqParams {
    const float* consts;       // [Q][3][Dp]
    const uint8_t* codes;      // [N][D]
    const uint32_t* mask;
    uint64_t* partials;
    float* out_all;
    int64_t Q, N;
    int D, Dp, K, CAP, parts;
    const uint32_t* only_flagged;   // optional [Q]: a query with a zero entry already has its answer (tensor-core path)
};

// grid = (parts, Q), block = 256; one warp per row.  VEC: D % 16 == 0 and 16-byte aligned base.
template <int KIND, bool VEC>
__global__ void __launch_bounds__(256) sq_scan_kernel(SqParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* cs = reinterpret_cast<float*>(smem_raw);                         // [3][Dp]
    uint64_t* sel_base = reinterpret_cast<uint64_t*>(smem_raw + (size_t)3 * p.Dp * 4);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q = blockIdx.y;
    if (p.only_flagged && p.only_flagged[q] == 0) return;
    // Shared-memory copy of the constants, permuted for the access pattern below: lane l reads the four float4 of chunk
    // c = l + 32 i, i.e. addresses 64 bytes apart in the natural [chunk][4] order -- a 4-way bank conflict on every
    // LDS.128 (ncu: 16 M conflicts, LSU wavefronts 75 % busy at a quarter of the HBM rate).  Stored [4][chunk], the
    // lanes of a load are 16 bytes apart: conflict free.  Same values, same arithmetic.
    const int nchunk = p.Dp >> 4;                                           // 16-code chunks per row
    {
        const float4* src4 = reinterpret_cast<const float4*>(p.consts + (size_t)q * 3 * p.Dp);
        float4* cs4 = reinterpret_cast<float4*>(cs);
        const int per = p.Dp >> 2;                                          // float4 per constant array
        for (int i = threadIdx.x; i < 3 * per; i += blockDim.x) {
            const int arr = i / per, f = i - arr * per;                     // f = chunk * 4 + u
            cs4[arr * per + (f & 3) * nchunk + (f >> 2)] = src4[i];
        }
    }
    WarpSelect<1> sel;
    const bool select = p.K > 0;
    if (select) sel.init(sel_base + (size_t)warp * (p.K + p.CAP), p.K, p.CAP, lane);
    __syncthreads();
    const float4* c0 = reinterpret_cast<const float4*>(cs);
    const float4* c1 = reinterpret_cast<const float4*>(cs + p.Dp);
    const float4* c2 = reinterpret_cast<const float4*>(cs + 2 * p.Dp);
    // A warp takes R rows at a time: the 12 constant vectors of a chunk are read from shared memory ONCE for the R rows
    // (one row at a time the scan read 12 bytes of constants per code byte and ran at 0.11 of HBM: 28 ms per query over
    // 20M x 1024), and R 128-bit loads are in flight per lane.  Per row the element order is unchanged: same sums.
    constexpr int R = 4;
    for (int64_t row0 = ((int64_t)blockIdx.x * W + warp) * R; row0 < p.N; row0 += (int64_t)gridDim.x * W * R) {
        bool valid[R], work[R];
        bool any = false;
#pragma unroll
        for (int r = 0; r < R; ++r) {
            const int64_t row = row0 + r;
            valid[r] = row < p.N && (!p.mask || mask_bit(p.mask, row));
            work[r] = row < p.N && (valid[r] || p.out_all);
            any = any || work[r];
        }
        if (!any) continue;
        float acc[R], nrm[R];
#pragma unroll
        for (int r = 0; r < R; ++r) { acc[r] = 0.f; nrm[r] = 0.f; }
        auto fetch = [&](uint4 (&w)[R], int c) {
#pragma unroll
            for (int r = 0; r < R; ++r) {
                w[r] = make_uint4(0, 0, 0, 0);
                if (work[r]) {
                    const uint8_t* rp = p.codes + (row0 + r) * p.D;
                    if (VEC) {
                        w[r] = ldg_nc_u4(reinterpret_cast<const uint4*>(rp) + c);
                    } else {
                        uint32_t t[4] = {0, 0, 0, 0};
                        for (int b = 0; b < 16; ++b) {
                            int j = c * 16 + b;
                            if (j < p.D) t[b >> 2] |= (uint32_t)__ldg(rp + j) << (8 * (b & 3));
                        }
                        w[r] = make_uint4(t[0], t[1], t[2], t[3]);
                    }
                }
            }
        };
        auto consume = [&](const uint4 (&w)[R], int c) {
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float4 a = c0[u * nchunk + c], b = c1[u * nchunk + c], cc = c2[u * nchunk + c];
#pragma unroll
                for (int r = 0; r < R; ++r) {
                    const uint32_t word = u == 0 ? w[r].x : u == 1 ? w[r].y : u == 2 ? w[r].z : w[r].w;
                    sq_word<KIND>(word, a, b, cc, acc[r], nrm[r]);
                }
            }
        };
        uint4 wa[R];
        if (KIND == FPV_SQ_COSINE) {
            // (the second accumulator set of the cosine form leaves no registers for a second buffer: 101 registers and
            // 7.7 ms with it, 7.5 ms without, 20M x 1024)
            for (int c = lane; c < nchunk; c += 32) { fetch(wa, c); consume(wa, c); }
        } else {
            // the next chunk's codes are in flight while this one is multiplied (two register buffers, unrolled by two):
            // dot 7.0 -> 6.4 ms
            uint4 wb[R];
            int c = lane;
            if (c < nchunk) fetch(wa, c);
            while (c < nchunk) {
                if (c + 32 < nchunk) fetch(wb, c + 32);
                consume(wa, c);
                c += 32;
                if (c >= nchunk) break;
                if (c + 32 < nchunk) fetch(wa, c + 32);
                consume(wb, c);
                c += 32;
            }
        }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            if (!work[r]) continue;                                         // warp-uniform
            const int64_t row = row0 + r;
            const float a = warp_sum(acc[r]);
            float d;
            if (KIND == FPV_SQ_L2) d = sqrtf(a);
            else if (KIND == FPV_SQ_DOT) d = -a;
            else d = 1.0f - a / (sqrtf(warp_sum(nrm[r])) + 1e-8f);
            if (p.out_all && lane == 0) p.out_all[q * p.N + row] = d;
            if (select && valid[r]) sel.add_uniform(0, make_key(d, (uint32_t)row), lane);
        }
    }
    if (select) {
        sel.flush_all(lane);
        block_merge_store<1>(sel_base, p.K, p.CAP, 1, p.partials + ((size_t)q * p.parts + blockIdx.x) * p.K, 0);
    }
}

// L2 fast path (D <= 512*CPL, D % 16 == 0): the per-dimension constants live in REGISTERS (the shared-memory version
// needs 8 bytes of constants per code byte, more than the 128 B/clk an SM can read: ncu showed 94% smem wavefronts),
// and every warp keeps R rows = R*CPL 128-bit loads in flight.
template <int CPL, int R>
__global__ void __launch_bounds__(256) sq_l2_reg_kernel(SqParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sel_base = reinterpret_cast<uint64_t*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q = blockIdx.y;
    if (p.only_flagged && p.only_flagged[q] == 0) return;
    const int nchunk = p.Dp >> 4;
    // Per element: |q - b| for four codes at once (VABSDIFF4), PRMT to 2^23 + |q - b|, then
    // t = fma(2^23 + |q - b|, c1, -2^23 * c1) = round(|q - b| * c1) exactly -- the same single rounding as the
    // reference's (q - b) * (scale / 255) -- and acc = fma(t, t, acc): 3.25 instructions per code instead of 4.
    float c1r[CPL][16], c2r[CPL][16];
    uint32_t qw[CPL][4];
    {
        const float* c0 = p.consts + (size_t)q * 3 * p.Dp;
        const float* c1 = c0 + p.Dp;
        const float* c2 = c1 + p.Dp;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
#pragma unroll
            for (int u = 0; u < 4; ++u) qw[i][u] = 0u;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                c1r[i][e] = c < nchunk ? c1[c * 16 + e] : 0.0f;
                c2r[i][e] = c < nchunk ? c2[c * 16 + e] : 0.0f;
                const uint32_t code = c < nchunk ? (__float_as_uint(c0[c * 16 + e]) & 0xFFu) : 0u;   // mantissa of 2^23 + code
                qw[i][e >> 2] |= code << (8 * (e & 3));
            }
        }
    }
    WarpSelect<1> sel;
    const bool select = p.K > 0;
    if (select) sel.init(sel_base + (size_t)warp * (p.K + p.CAP), p.K, p.CAP, lane);
    for (int64_t row0 = ((int64_t)blockIdx.x * W + warp) * R; row0 < p.N; row0 += (int64_t)gridDim.x * W * R) {
        uint4 w[R][CPL];
#pragma unroll
        for (int r = 0; r < R; ++r)
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = lane + 32 * i;
                const int64_t row = row0 + r;
                w[r][i] = (row < p.N && c < nchunk) ? ldg_nc_u4(reinterpret_cast<const uint4*>(p.codes + row * p.D) + c)
                                                   : make_uint4(0, 0, 0, 0);
            }
#pragma unroll
        for (int r = 0; r < R; ++r) {
            float a4[4] = {0.f, 0.f, 0.f, 0.f};          // (four chains measured slower than one: 5.46 vs 5.05 ms; keep one)
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const uint32_t ws[4] = {w[r][i].x, w[r][i].y, w[r][i].z, w[r][i].w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t ad = __vabsdiffu4(qw[i][u], ws[u]);
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float t = fmaf(u8f(ad, b), c1r[i][u * 4 + b], c2r[i][u * 4 + b]);
                        a4[0] = fmaf(t, t, a4[0]);
                    }
                }
            }
            float acc = warp_sum((a4[0] + a4[1]) + (a4[2] + a4[3]));
            const int64_t row = row0 + r;
            if (row < p.N) {
                const float d = sqrtf(acc);
                if (p.out_all && lane == 0) p.out_all[q * p.N + row] = d;
                if (select && (!p.mask || mask_bit(p.mask, row))) sel.add_uniform(0, make_key(d, (uint32_t)row), lane);
            }
        }
    }
    if (select) {
        sel.flush_all(lane);
        block_merge_store<1>(sel_base, p.K, p.CAP, 1, p.partials + ((size_t)q * p.parts + blockIdx.x) * p.K, 0);
    }
}

// ---------------------------------------------------------------------------------------------- L2, TMA-staged rows
// Same arithmetic as sq_l2_reg_kernel, but the code rows no longer travel through registers: one producer thread
// streams tiles of whole rows (contiguous bytes) into a 3-stage shared-memory ring with cp.async.bulk + mbarriers, and
// the consumer warps read their rows from shared memory.  The register kernel could keep only R*CPL = 8 128-bit loads
// per lane in flight at 16 warps per SM (the per-dimension constants take 64 registers), which left it latency bound
// at ~64 % of HBM; here ~190 KB per SM are in flight regardless of occupancy.
constexpr int SQT_STAGES = 3;
constexpr int SQT_TILE_BYTES = 32768;
constexpr int SQT_CONSUMERS = 8;            // consumer warps; the next warp's lane 0 produces.  (7 + 1 warps = 256 threads would
                                            // avoid the ~40 bytes of spills of the 96-register cap, but measured 4.66 vs 4.50 ms:
                                            // the scan is bound by the consumers' instruction issue, not by memory)
constexpr int SQT_THREADS = 32 * (SQT_CONSUMERS + 1);

__device__ __forceinline__ uint32_t sq_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sq_mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t done = 0, spins = 0;
    while (true) {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(done) : "r"(bar), "r"(parity) : "memory");
        if (done) break;
        if (++spins > (1u << 24)) __trap();      // a stuck pipeline traps instead of hanging the GPU
    }
}

// Squared-distance bound that certainly contains every row whose (sqrt) distance can beat the selector's k-th key:
// a row is rejected on its squared sum alone (one compare); sqrt, key and the mask bit are only computed for the
// few rows that pass.  (Per-row bookkeeping was 80 of the 190 instructions per row and made the scan issue bound.)
__device__ __forceinline__ float sq_thr2(uint64_t tau) {
    if (tau == FPV_KEY_MAX) return INFINITY;
    const float d = ordered_to_f32((uint32_t)(tau >> 32));
    return d * d * 1.000001f + 1e-37f;
}

template <int CPL, bool ALL>
__global__ void __launch_bounds__(SQT_THREADS, 2) sq_l2_tma_kernel(SqParams p, int tile_rows) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* ring = smem_raw;                                                      // [STAGES][tile_rows * D]
    const size_t stage_bytes = (size_t)tile_rows * p.D;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_raw + SQT_STAGES * stage_bytes);   // full[S], empty[S]
    uint64_t* sel_base = bars + 2 * SQT_STAGES + 2;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int64_t q = blockIdx.y;
    if (p.only_flagged && p.only_flagged[q] == 0) return;     // uniform for the CTA, before any barrier
    const int nchunk = p.Dp >> 4;
    const uint32_t bar0 = sq_smem_u32(bars);
    if (threadIdx.x == 0) {
        for (int s = 0; s < SQT_STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * s), "r"(1) : "memory");
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar0 + 8 * (SQT_STAGES + s)), "r"(SQT_CONSUMERS) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const int64_t ntiles = (p.N + tile_rows - 1) / tile_rows;

    if (warp == SQT_CONSUMERS) {                        // ---------------- producer
        if (lane == 0) {
            int s = 0; uint32_t ph = 0;
            for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
                const int64_t row0 = t * tile_rows;
                const uint32_t bytes = (uint32_t)(min((int64_t)tile_rows, p.N - row0) * p.D);
                sq_mbar_wait(bar0 + 8 * (SQT_STAGES + s), ph ^ 1);
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8 * s), "r"(bytes) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(sq_smem_u32(ring + s * stage_bytes)), "l"(p.codes + row0 * p.D), "r"(bytes), "r"(bar0 + 8 * s)
                             : "memory");
                if (++s == SQT_STAGES) { s = 0; ph ^= 1; }
            }
        }
        return;
    }

    // ---------------- consumers: constants in registers exactly as in sq_l2_reg_kernel
    // (packed FFMA2 was tried for the two FMAs per code: 7.3 ms instead of 4.5 ms, it issues far below the scalar rate)
    float c1r[CPL][16], c2r[CPL][16];
    uint32_t qw[CPL][4];
    {
        const float* c0 = p.consts + (size_t)q * 3 * p.Dp;
        const float* c1 = c0 + p.Dp;
        const float* c2 = c1 + p.Dp;
#pragma unroll
        for (int i = 0; i < CPL; ++i) {
            const int c = lane + 32 * i;
#pragma unroll
            for (int u = 0; u < 4; ++u) qw[i][u] = 0u;
#pragma unroll
            for (int e = 0; e < 16; ++e) {
                c1r[i][e] = c < nchunk ? c1[c * 16 + e] : 0.0f;
                c2r[i][e] = c < nchunk ? c2[c * 16 + e] : 0.0f;
                const uint32_t code = c < nchunk ? (__float_as_uint(c0[c * 16 + e]) & 0xFFu) : 0u;
                qw[i][e >> 2] |= code << (8 * (e & 3));
            }
        }
    }
    WarpSelect<1> sel;
    const bool select = p.K > 0;
    if (select) sel.init(sel_base + (size_t)warp * (p.K + p.CAP), p.K, p.CAP, lane);
    float thr2 = INFINITY;
    int s = 0; uint32_t ph = 0;
    for (int64_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int64_t row0 = t * tile_rows;
        const int rows = (int)min((int64_t)tile_rows, p.N - row0);
        sq_mbar_wait(bar0 + 8 * s, ph);
        const unsigned char* tile = ring + s * stage_bytes;
        for (int r = warp; r < rows; r += SQT_CONSUMERS) {
            const uint4* rowp = reinterpret_cast<const uint4*>(tile + (size_t)r * p.D);
            uint4 w[CPL];
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const int c = lane + 32 * i;
                w[i] = c < nchunk ? rowp[c] : make_uint4(0, 0, 0, 0);
            }
            float a = 0.f, a2 = 0.f;                     // two dependent chains
#pragma unroll
            for (int i = 0; i < CPL; ++i) {
                const uint32_t ws[4] = {w[i].x, w[i].y, w[i].z, w[i].w};
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const uint32_t ad = __vabsdiffu4(qw[i][u], ws[u]);
#pragma unroll
                    for (int b = 0; b < 4; ++b) {
                        const float tt = fmaf(u8f(ad, b), c1r[i][u * 4 + b], c2r[i][u * 4 + b]);
                        if (b & 1) a2 = fmaf(tt, tt, a2); else a = fmaf(tt, tt, a);
                    }
                }
            }
            const float acc = warp_sum(a + a2);
            const int64_t row = row0 + r;
            if (ALL && lane == 0) p.out_all[q * p.N + row] = sqrtf(acc);
            if (select && acc <= thr2) {                 // uniform: every lane holds the same sum
                if (!p.mask || mask_bit(p.mask, row)) {
                    sel.add_uniform(0, make_key(sqrtf(acc), (uint32_t)row), lane);
                    thr2 = sq_thr2(sel.tau[0]);
                }
            }
        }
        __syncwarp();
        if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0 + 8 * (SQT_STAGES + s)) : "memory");
        if (++s == SQT_STAGES) { s = 0; ph ^= 1; }
    }
    if (select) sel.flush_all(lane);
    // block_merge_store() synchronises the whole CTA, which the producer warp has already left: merge here among the
    // consumer warps with a named barrier instead
    asm volatile("bar.sync 1, %0;" ::"n"(32 * SQT_CONSUMERS) : "memory");
    if (select && warp == 0) {
        uint64_t* dst = sel_base;
        for (int w2 = 1; w2 < SQT_CONSUMERS; ++w2) merge_sorted_into(dst, sel_base + (size_t)w2 * (p.K + p.CAP), p.K, lane);
        uint64_t* o = p.partials + ((size_t)q * p.parts + blockIdx.x) * p.K;
        for (int i = lane; i < p.K; i += 32) o[i] = dst[i];
    }
}

// FPV_SQ_TMA=0 keeps the register-staged kernel for every size (A/B measurements)
static bool sq_tma_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FPV_SQ_TMA"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

struct SqPlan { int K, CAP, parts, Dp; size_t off_const, off_part, total, smem; };
static SqPlan plan_sq(int64_t Q, int64_t N, int D, int k) {
    SqPlan pl{};
    pl.K = k > 0 ? sel_K(k) : 0;
    pl.CAP = k > 0 ? sel_CAP(pl.K) : 0;
    pl.Dp = (D + 15) / 16 * 16;
    pl.smem = (size_t)3 * pl.Dp * 4 + (size_t)8 * (pl.K + pl.CAP) * 8;
    int64_t want = (int64_t)sm_count() * 4;
    int64_t parts = Q > 0 ? (want + Q - 1) / Q : want;
    int64_t max_parts = (N + 7) / 8;
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    pl.parts = (int)parts;
    pl.off_const = 0;
    pl.off_part = align_up((size_t)(Q > 0 ? Q : 0) * 3 * pl.Dp * 4, 256);
    pl.total = pl.off_part + 256 + (size_t)(Q > 0 ? Q : 0) * pl.parts * pl.K * 8;
    return pl;
}

template <int KIND>
static int launch_sq(const SqParams& p, const SqPlan& pl, bool vec, cudaStream_t st) {
    dim3 grid(pl.parts, (unsigned)p.Q);
    if (vec) {
        if (pl.smem > 48 * 1024)
            FPV_CUDA(cudaFuncSetAttribute(sq_scan_kernel<KIND, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        sq_scan_kernel<KIND, true><<<grid, 256, pl.smem, st>>>(p);
    } else {
        if (pl.smem > 48 * 1024)
            FPV_CUDA(cudaFuncSetAttribute(sq_scan_kernel<KIND, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        sq_scan_kernel<KIND, false><<<grid, 256, pl.smem, st>>>(p);
    }
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

}  // namespace fpv

using namespace fpv;

extern "C" int fpv_sq_encode(const float* vectors, int64_t n, int d, int64_t ld, const float* min_vals,
                             const float* scale, uint8_t* out_codes, void* stream) {
    FPV_REQUIRE(n >= 0 && d >= 1 && ld >= d, "sq_encode: bad shape n=%lld d=%d ld=%lld", (long long)n, d, (long long)ld);
    if (n == 0) return FPV_OK;
    FPV_REQUIRE(vectors && min_vals && scale && out_codes, "sq_encode: null pointer");
    int64_t total = n * d;
    const uintptr_t al = reinterpret_cast<uintptr_t>(vectors) | reinterpret_cast<uintptr_t>(min_vals) |
                         reinterpret_cast<uintptr_t>(scale);
    if (d % 4 == 0 && ld % 4 == 0 && (al & 15) == 0 && (reinterpret_cast<uintptr_t>(out_codes) & 3) == 0 && total / 4 < (1ll << 32)) {
        int64_t vblocks = std::min<int64_t>((total / 4 + 255) / 256, (int64_t)sm_count() * 16);
        sq_encode_vec_kernel<<<(unsigned)vblocks, 256, 0, (cudaStream_t)stream>>>(vectors, (uint32_t)n, (uint32_t)(d / 4), ld, min_vals,
                                                                               scale, out_codes);
        FPV_LAUNCH_CHECK();
        return FPV_OK;
    }
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    sq_encode_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(vectors, n, d, ld, min_vals, scale, out_codes);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

extern "C" size_t fpv_sq_workspace(int64_t q, int64_t n, int d, int k) {
    if (d <= 0 || k < 0) return 256;
    return plan_sq(q, n, d, k).total;
}

static int sq_topk_impl(int kind, const uint8_t* qcodes, int64_t q, const uint8_t* codes, int64_t n, int d,
                        const float* min_vals, const float* scale, int k, const uint32_t* mask_words, int64_t id_base,
                        float* out_dist, int64_t* out_idx, int32_t* out_count, float* out_all, const uint32_t* only_flagged,
                        void* ws, size_t ws_bytes, cudaStream_t st) {
    FPV_REQUIRE(kind >= 0 && kind <= 2, "sq: unknown kind %d", kind);
    FPV_REQUIRE(q >= 0 && n >= 0 && d >= 1 && d <= 16384, "sq: bad shape q=%lld n=%lld d=%d", (long long)q, (long long)n, d);
    FPV_REQUIRE(k >= 0 && k <= FPV_MAX_K, "sq: k=%d outside [0,%d]", k, FPV_MAX_K);
    FPV_REQUIRE(k > 0 || out_all, "sq: nothing to do (k == 0 and out_all == NULL)");
    FPV_REQUIRE(n < (1ll << 32), "sq: N=%lld rows per call exceeds 2^32-1 (shard the database)", (long long)n);
    FPV_REQUIRE(q <= 65535, "sq: at most 65535 queries per call");
    if (q == 0) return FPV_OK;
    FPV_REQUIRE(qcodes && (codes || n == 0) && min_vals && scale, "sq: null pointer");
    FPV_REQUIRE(k == 0 || (out_dist && out_idx), "sq: null output");
    SqPlan pl = plan_sq(q, n, d, k);
    if (!ws || ws_bytes < pl.total) { set_error("sq: workspace %zu < %zu", ws_bytes, pl.total); return FPV_ERR_WORKSPACE; }
    FPV_REQUIRE(pl.smem <= (size_t)max_smem_optin(), "sq: d=%d k=%d needs %zu B shared memory", d, k, pl.smem);
    char* w = static_cast<char*>(ws);
    float* consts = reinterpret_cast<float*>(w + pl.off_const);
    uint64_t* partials = reinterpret_cast<uint64_t*>(w + pl.off_part);
    sq_prep_kernel<<<(unsigned)q, 256, 0, st>>>(kind, qcodes, d, pl.Dp, min_vals, scale, consts);
    FPV_LAUNCH_CHECK();
    SqParams p{};
    p.consts = consts; p.codes = codes; p.mask = mask_words; p.partials = partials; p.out_all = out_all;
    p.Q = q; p.N = n; p.D = d; p.Dp = pl.Dp; p.K = pl.K; p.CAP = pl.CAP; p.parts = pl.parts;
    p.only_flagged = only_flagged;
    const bool vec = (d % 16 == 0) && ((reinterpret_cast<uintptr_t>(codes) & 15) == 0);
    int rc;
    if (kind == FPV_SQ_L2 && vec && d <= 1024 && n >= 65536 && sq_tma_enabled()) {
        // large scans: rows staged through shared memory by cp.async.bulk (see sq_l2_tma_kernel)
        int tile_rows = SQT_TILE_BYTES / d / SQT_CONSUMERS * SQT_CONSUMERS;
        if (tile_rows < SQT_CONSUMERS) tile_rows = SQT_CONSUMERS;
        const size_t smem = (size_t)SQT_STAGES * tile_rows * d + (2 * SQT_STAGES + 2) * 8 + (size_t)8 * (pl.K + pl.CAP) * 8;
        int gx = (int)std::min<int64_t>(pl.parts, (int64_t)2 * sm_count());
        gx = (int)std::min<int64_t>(gx, (n + tile_rows - 1) / tile_rows);
        p.parts = gx;
        dim3 grid(gx, (unsigned)q);
        auto kern = d <= 512 ? (out_all ? sq_l2_tma_kernel<1, true> : sq_l2_tma_kernel<1, false>)
                             : (out_all ? sq_l2_tma_kernel<2, true> : sq_l2_tma_kernel<2, false>);
        FPV_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, SQT_THREADS, smem, st>>>(p, tile_rows);
        FPV_LAUNCH_CHECK();
        if (k > 0) return launch_finalize(partials, q, gx, pl.K, k, id_base, out_dist, out_idx, out_count, st, only_flagged);
        return FPV_OK;
    } else if (kind == FPV_SQ_L2 && vec && d <= 1024) {
        const size_t smem = (size_t)8 * (pl.K + pl.CAP) * 8;
        dim3 grid(pl.parts, (unsigned)q);
        if (d <= 512) {
            if (smem > 48 * 1024) FPV_CUDA(cudaFuncSetAttribute(sq_l2_reg_kernel<1, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sq_l2_reg_kernel<1, 4><<<grid, 256, smem, st>>>(p);
        } else {
            if (smem > 48 * 1024) FPV_CUDA(cudaFuncSetAttribute(sq_l2_reg_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            sq_l2_reg_kernel<2, 4><<<grid, 256, smem, st>>>(p);
        }
        FPV_LAUNCH_CHECK();
        rc = FPV_OK;
    } else if (kind == FPV_SQ_L2) rc = launch_sq<FPV_SQ_L2>(p, pl, vec, st);
    else if (kind == FPV_SQ_DOT) rc = launch_sq<FPV_SQ_DOT>(p, pl, vec, st);
    else rc = launch_sq<FPV_SQ_COSINE>(p, pl, vec, st);
    if (rc != FPV_OK) return rc;
    if (k > 0) return launch_finalize(partials, q, pl.parts, pl.K, k, id_base, out_dist, out_idx, out_count, st, only_flagged);
    return FPV_OK;
}

extern "C" int fpv_sq_topk(int kind, const uint8_t* qcodes, int64_t q, const uint8_t* codes, int64_t n, int d,
                           const float* min_vals, const float* scale, int k, const uint32_t* mask_words, int64_t id_base,
                           float* out_dist, int64_t* out_idx, int32_t* out_count, float* out_all,
                           void* ws, size_t ws_bytes, void* stream) {
    return sq_topk_impl(kind, qcodes, q, codes, n, d, min_vals, scale, k, mask_words, id_base, out_dist, out_idx, out_count, out_all,
                        nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

// the same scan restricted to the queries whose only_flagged entry is non-zero (fallback of fpv_sq_mma.cu)
namespace fpv {
size_t sq_flagged_workspace(int64_t Q, int64_t N, int D, int k) { return fpv_sq_workspace(Q, N, D, k); }
int sq_topk_flagged(int kind, const uint8_t* qcodes, int64_t q, const uint8_t* codes, int64_t n, int d, const float* min_vals,
                    const float* scale, int k, const uint32_t* mask_words, int64_t id_base, float* out_dist, int64_t* out_idx,
                    int32_t* out_count, const uint32_t* only_flagged, void* ws, size_t ws_bytes, cudaStream_t st) {
    return sq_topk_impl(kind, qcodes, q, codes, n, d, min_vals, scale, k, mask_words, id_base, out_dist, out_idx, out_count, nullptr,
                        only_flagged, ws, ws_bytes, st);
}
}  // namespace fpv
