// Binary quantizer: encode (bit pack) and the Hamming scan with fused top-k.
//
// Replaces BinaryQuantizer.encode (quantization.py:336-350), hamming_distances (:364-374, which materialises
// an N x D byte matrix through np.unpackbits) and search (:384-394).
//
// Layout: codes are [N][nbytes] uint8 exactly as np.packbits produces them (MSB-first inside a byte; bit
// order is irrelevant to XOR+popcount).  Fast path (nbytes = 16 * 2^j <= 512, 16-byte aligned base): L = nbytes/16
// lanes share one row, every load instruction is one fully coalesced 512-byte warp transaction covering
// 32/L consecutive rows; the per-lane partial popcounts are reduced with a transposed butterfly so that after
// L loads every lane owns the finished distance of one row (lane-per-row form for the selector).
// Generic path (any nbytes): one lane per row.
#include <algorithm>

#include "fpv_common.cuh"

namespace fpv {

// ---------------------------------------------------------------------------------------------------- encode
// one warp per row; lane handles one output byte at a time (8 dims), MSB-first like np.packbits.
__global__ void bq_encode_kernel(const float* __restrict__ v, int64_t N, int D, int64_t ld,
                                 const float* __restrict__ thr, uint8_t* __restrict__ out, int nbytes) {
    const int64_t total = N * nbytes;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        int64_t row = i / nbytes;
        int b = (int)(i - row * nbytes);
        const float* src = v + row * ld + b * 8;
        unsigned byte = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            int dim = b * 8 + j;
            if (dim < D && src[j] > thr[dim]) byte |= 0x80u >> j;
        }
        out[i] = (uint8_t)byte;
    }
}

// D % 8 == 0, 16-byte aligned rows and thresholds: one output byte per thread from two 128-bit loads, 32-bit index
// arithmetic (the form above: eight 4-byte loads 32 bytes apart across the lanes and a 64-bit division per byte, 0.36 of HBM)
__global__ void __launch_bounds__(256) bq_encode_vec_kernel(const float* __restrict__ v, uint32_t N, uint32_t nbytes, int64_t ld,
                                                            const float* __restrict__ thr, uint8_t* __restrict__ out) {
    const uint32_t total = N * nbytes;                                       // < 2^32 (checked on the host)
    for (uint64_t i64 = blockIdx.x * blockDim.x + threadIdx.x; i64 < total; i64 += gridDim.x * blockDim.x) {   // no 32-bit wrap
        const uint32_t i = (uint32_t)i64;
        const uint32_t row = i / nbytes, b = i - row * nbytes;
        const float4* src = reinterpret_cast<const float4*>(v + (int64_t)row * ld) + 2 * b;
        const float4 x0 = ldg_nc_f4(src), x1 = ldg_nc_f4(src + 1);
        const float4 t0 = __ldg(reinterpret_cast<const float4*>(thr) + 2 * b), t1 = __ldg(reinterpret_cast<const float4*>(thr) + 2 * b + 1);
        const unsigned byte = (x0.x > t0.x ? 0x80u : 0u) | (x0.y > t0.y ? 0x40u : 0u) | (x0.z > t0.z ? 0x20u : 0u) | (x0.w > t0.w ? 0x10u : 0u) |
                              (x1.x > t1.x ? 0x08u : 0u) | (x1.y > t1.y ? 0x04u : 0u) | (x1.z > t1.z ? 0x02u : 0u) | (x1.w > t1.w ? 0x01u : 0u);
        out[(size_t)row * nbytes + b] = (uint8_t)byte;
    }
}

// ---------------------------------------------------------------------------------------------------- scan
struct HamParams {
    const uint8_t* qbits;      // [Q][nbytes]
    const uint8_t* codes;      // [N][nbytes]
    const uint8_t* dimmask;    // [nbytes] valid-bit mask (first `dims` bits, MSB-first)
    const uint32_t* mask;
    uint64_t* partials;        // [Q][parts][K]
    float* out_all;            // [Q][N] or null
    int64_t Q, N;
    int nbytes, K, CAP, parts;
    const uint32_t* only_flagged;   // optional [Q]: queries with a zero entry already have their answer (tensor-core path)
};

// L lanes per row (L = nbytes / 16), QB queries share every code load.  grid = (parts, ceil(Q / QB)).
template <int L, int QB>
__global__ void __launch_bounds__(256) hamming_fast_kernel(HamParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sel_base = reinterpret_cast<uint64_t*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q0 = (int64_t)blockIdx.y * QB;
    const int nq = (int)min((int64_t)QB, p.Q - q0);
    if (p.only_flagged) {                                   // uniform for the CTA
        bool any = false;
        for (int q = 0; q < nq; ++q) any |= p.only_flagged[q0 + q] != 0;
        if (!any) return;
    }
    constexpr int RPL = 32 / L;                 // rows covered by one warp load
    const int sub = lane % L;                   // which 16-byte chunk of the row this lane owns
    uint4 qv[QB];
#pragma unroll
    for (int q = 0; q < QB; ++q)
        qv[q] = *reinterpret_cast<const uint4*>(p.qbits + (q0 + (q < nq ? q : 0)) * p.nbytes + sub * 16);
    const uint4 mv = *reinterpret_cast<const uint4*>(p.dimmask + sub * 16);

    WarpSelect<QB> sel;
    const bool select = p.K > 0;
    if (select) sel.init(sel_base + (size_t)warp * QB * (p.K + p.CAP), p.K, p.CAP, lane);

    const uint4* base = reinterpret_cast<const uint4*>(p.codes);
    const int64_t ngroups = (p.N + 31) / 32;    // 32 rows per warp iteration
    constexpr int B = L < 8 ? L : 8;            // loads per batch
    // a 32-row group is one batch: prefetch the NEXT group before computing this one.  Only for the single-query
    // (HBM-bound) form; the batched form is POPC-bound and would just lose occupancy to the extra registers.
    constexpr bool PF = L <= 8 && QB == 1;
    const int64_t gstep = (int64_t)gridDim.x * W;
    const int64_t last_row = p.N - 1;           // out-of-range rows read the last row (unpredicated loads), masked by `valid`
    uint4 nxt[B];
    int64_t g = (int64_t)blockIdx.x * W + warp;
    if (PF && g < ngroups) {
#pragma unroll
        for (int b = 0; b < B; ++b) nxt[b] = ldg_nc_u4(base + min(g * 32 + b * RPL + lane / L, last_row) * L + sub);
    }
    for (; g < ngroups; g += gstep) {
        const int64_t row0 = g * 32;
        int part[QB][L];
#pragma unroll
        for (int j0 = 0; j0 < L; j0 += B) {
            uint4 dv[B];
            if (PF) {
#pragma unroll
                for (int b = 0; b < B; ++b) dv[b] = nxt[b];
                if (g + gstep < ngroups) {
#pragma unroll
                    for (int b = 0; b < B; ++b)
                        nxt[b] = ldg_nc_u4(base + min((g + gstep) * 32 + b * RPL + lane / L, last_row) * L + sub);
                }
            } else {
#pragma unroll
                for (int b = 0; b < B; ++b)
                    dv[b] = ldg_nc_u4(base + min(row0 + (j0 + b) * RPL + lane / L, last_row) * L + sub);
            }
#pragma unroll
            for (int q = 0; q < QB; ++q)
#pragma unroll
                for (int b = 0; b < B; ++b)
                    part[q][j0 + b] = __popc((dv[b].x ^ qv[q].x) & mv.x) + __popc((dv[b].y ^ qv[q].y) & mv.y) +
                                      __popc((dv[b].z ^ qv[q].z) & mv.z) + __popc((dv[b].w ^ qv[q].w) & mv.w);
        }
        const int64_t row = row0 + sub * RPL + lane / L;
        const bool valid = row < p.N && (!p.mask || mask_bit(p.mask, row));
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            // transposed butterfly: L values on each of L lanes -> one total per lane; lane `sub` ends with load j = sub
#pragma unroll
            for (int s = L / 2; s >= 1; s >>= 1) {
                const bool upper = (sub & s) != 0;
#pragma unroll
                for (int i = 0; i < s; ++i) {
                    int send = upper ? part[q][i] : part[q][i + s];
                    int keep = upper ? part[q][i + s] : part[q][i];
                    part[q][i] = keep + __shfl_xor_sync(FPV_FULL_MASK, send, s);
                }
            }
            if (q < nq) {
                const float d = (float)part[q][0];
                if (p.out_all && row < p.N) p.out_all[(q0 + q) * p.N + row] = d;
                if (select) sel.add_lanes(q, make_key(d, (uint32_t)row), valid, lane);
            }
        }
    }
    if (select) {
        sel.flush_all(lane);
        block_merge_store<QB>(sel_base, p.K, p.CAP, nq, p.partials + ((size_t)q0 * p.parts + blockIdx.x) * p.K,
                              (size_t)p.parts * p.K);
    }
}

// any nbytes: lane per row, byte-granular loads (correct for every shape; the fast path covers the common widths)
__global__ void __launch_bounds__(256) hamming_generic_kernel(HamParams p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q = blockIdx.y;
    if (p.only_flagged && p.only_flagged[q] == 0) return;
    uint8_t* qm = smem_raw;                                   // [2][nbytes]: query bytes, valid mask
    const size_t qm_bytes = align_up((size_t)2 * p.nbytes, 16);
    uint64_t* sel_base = reinterpret_cast<uint64_t*>(smem_raw + qm_bytes);
    for (int i = threadIdx.x; i < p.nbytes; i += blockDim.x) {
        qm[i] = p.qbits[q * p.nbytes + i];
        qm[p.nbytes + i] = p.dimmask[i];
    }
    WarpSelect<1> sel;
    const bool select = p.K > 0;
    if (select) sel.init(sel_base + (size_t)warp * (p.K + p.CAP), p.K, p.CAP, lane);
    __syncthreads();
    const bool w32 = (p.nbytes % 4 == 0) && ((reinterpret_cast<uintptr_t>(p.codes) & 3) == 0);
    const int64_t ngroups = (p.N + 31) / 32;
    for (int64_t g = (int64_t)blockIdx.x * W + warp; g < ngroups; g += (int64_t)gridDim.x * W) {
        const int64_t row = g * 32 + lane;
        int c = 0;
        if (row < p.N) {
            const uint8_t* src = p.codes + row * p.nbytes;
            if (w32) {
                const uint32_t* s4 = reinterpret_cast<const uint32_t*>(src);
                const uint32_t* q4 = reinterpret_cast<const uint32_t*>(qm);
                const uint32_t* m4 = reinterpret_cast<const uint32_t*>(qm + p.nbytes);
                for (int j = 0; j < p.nbytes / 4; ++j) c += __popc((__ldg(s4 + j) ^ q4[j]) & m4[j]);
            } else {
                for (int j = 0; j < p.nbytes; ++j) c += __popc((unsigned)((__ldg(src + j) ^ qm[j]) & qm[p.nbytes + j]));
            }
        }
        const bool valid = row < p.N && (!p.mask || mask_bit(p.mask, row));
        const float d = (float)c;
        if (p.out_all && row < p.N) p.out_all[q * p.N + row] = d;
        if (select) sel.add_lanes(0, make_key(d, (uint32_t)row), valid, lane);
    }
    if (select) {
        sel.flush_all(lane);
        block_merge_store<1>(sel_base, p.K, p.CAP, 1, p.partials + ((size_t)q * p.parts + blockIdx.x) * p.K, 0);
    }
}

__global__ void dimmask_kernel(uint8_t* m, int nbytes, int dims) {
    for (int i = threadIdx.x; i < nbytes; i += blockDim.x) {
        int lo = i * 8;
        unsigned v;
        if (dims <= 0 || lo + 8 <= dims) v = 0xFFu;
        else if (lo >= dims) v = 0u;
        else v = (0xFFu << (8 - (dims - lo))) & 0xFFu;       // MSB-first: first (dims-lo) bits of the byte
        m[i] = (uint8_t)v;
    }
}

struct HamPlan { int K, CAP, parts, qb; size_t off_mask, off_part, total, smem; };
// qb = queries sharing one pass over the codes (fast path with <= 8 lanes per row only)
static int hamming_qb(int64_t Q, int nbytes, int K, int CAP) {
    if (nbytes > 128 || nbytes % 16 != 0 || (nbytes & (nbytes - 1)) != 0) return 1;
    int qb = Q >= 4 ? 4 : (Q >= 2 ? 2 : 1);
    while (qb > 1 && (size_t)8 * qb * (K + CAP) * 8 > 96 * 1024) qb >>= 1;
    return qb;
}
static HamPlan plan_hamming(int64_t Q, int64_t N, int nbytes, int k, int qb) {
    HamPlan pl{};
    pl.K = k > 0 ? sel_K(k) : 0;
    pl.CAP = k > 0 ? sel_CAP(pl.K) : 0;
    pl.qb = qb;
    const int64_t nqc = Q > 0 ? (Q + qb - 1) / qb : 1;
    int64_t want = (int64_t)sm_count() * 4;
    int64_t parts = (want + nqc - 1) / nqc;
    int64_t max_parts = (N + 255) / 256;
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    pl.parts = (int)parts;
    pl.off_mask = 0;
    pl.off_part = align_up((size_t)nbytes, 256);
    pl.total = pl.off_part + (size_t)(Q > 0 ? Q : 0) * pl.parts * pl.K * 8;
    pl.smem = (size_t)8 * qb * (pl.K + pl.CAP) * 8;
    return pl;
}

template <int L, int QB>
static int launch_fast_qb(const HamParams& p, const HamPlan& pl, cudaStream_t st) {
    if (pl.smem > 48 * 1024)
        FPV_CUDA(cudaFuncSetAttribute(hamming_fast_kernel<L, QB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
    hamming_fast_kernel<L, QB><<<dim3(pl.parts, (unsigned)((p.Q + QB - 1) / QB)), 256, pl.smem, st>>>(p);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}
template <int L>
static int launch_fast(const HamParams& p, const HamPlan& pl, cudaStream_t st) {
    if constexpr (L <= 8) {
        if (pl.qb == 4) return launch_fast_qb<L, 4>(p, pl, st);
        if (pl.qb == 2) return launch_fast_qb<L, 2>(p, pl, st);
    }
    return launch_fast_qb<L, 1>(p, pl, st);
}

}  // namespace fpv

using namespace fpv;

extern "C" int fpv_bq_encode(const float* vectors, int64_t n, int d, int64_t ld, const float* thresholds,
                             uint8_t* out_codes, void* stream) {
    FPV_REQUIRE(n >= 0 && d >= 1 && ld >= d, "bq_encode: bad shape n=%lld d=%d ld=%lld", (long long)n, d, (long long)ld);
    if (n == 0) return FPV_OK;
    FPV_REQUIRE(vectors && thresholds && out_codes, "bq_encode: null pointer");
    int nbytes = (d + 7) / 8;
    int64_t total = n * nbytes;
    if (d % 8 == 0 && ld % 4 == 0 && ((reinterpret_cast<uintptr_t>(vectors) | reinterpret_cast<uintptr_t>(thresholds)) & 15) == 0 &&
        total < (1ll << 32)) {
        int64_t vblocks = std::min<int64_t>((total + 255) / 256, (int64_t)sm_count() * 16);
        bq_encode_vec_kernel<<<(unsigned)vblocks, 256, 0, (cudaStream_t)stream>>>(vectors, (uint32_t)n, (uint32_t)nbytes, ld, thresholds,
                                                                               out_codes);
        FPV_LAUNCH_CHECK();
        return FPV_OK;
    }
    int64_t blocks = (total + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    bq_encode_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(vectors, n, d, ld, thresholds, out_codes, nbytes);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

extern "C" size_t fpv_hamming_workspace(int64_t q, int64_t n, int nbytes, int k) {
    if (nbytes <= 0 || k < 0) return 256;
    // the launch may or may not batch queries (alignment is only known at launch): size for the larger plan
    const int K = k > 0 ? sel_K(k) : 0;
    const size_t a = plan_hamming(q, n, nbytes, k, 1).total;
    const size_t b = plan_hamming(q, n, nbytes, k, hamming_qb(q, nbytes, K, K ? sel_CAP(K) : 0)).total;
    return a > b ? a : b;
}

static int hamming_topk_impl(const uint8_t* qbits, int64_t q, const uint8_t* codes, int64_t n, int nbytes, int dims,
                             int k, const uint32_t* mask_words, int64_t id_base,
                             float* out_dist, int64_t* out_idx, int32_t* out_count, float* out_all,
                             const uint32_t* only_flagged, void* ws, size_t ws_bytes, cudaStream_t st) {
    FPV_REQUIRE(q >= 0 && n >= 0 && nbytes >= 1, "hamming: bad shape q=%lld n=%lld nbytes=%d", (long long)q, (long long)n, nbytes);
    FPV_REQUIRE(k >= 0 && k <= FPV_MAX_K, "hamming: k=%d outside [0,%d]", k, FPV_MAX_K);
    FPV_REQUIRE(k > 0 || out_all, "hamming: nothing to do (k == 0 and out_all == NULL)");
    FPV_REQUIRE(n < (1ll << 32), "hamming: N=%lld rows per call exceeds 2^32-1 (shard the database)", (long long)n);
    FPV_REQUIRE(q <= 65535, "hamming: at most 65535 queries per call");
    FPV_REQUIRE(dims <= nbytes * 8, "hamming: dims=%d exceeds code width %d bits", dims, nbytes * 8);
    if (q == 0) return FPV_OK;
    FPV_REQUIRE(qbits && (codes || n == 0), "hamming: null pointer");
    FPV_REQUIRE(k == 0 || (out_dist && out_idx), "hamming: null output");
    const bool aligned = ((reinterpret_cast<uintptr_t>(codes) & 15) == 0) && ((reinterpret_cast<uintptr_t>(qbits) & 15) == 0);
    const int K0 = k > 0 ? sel_K(k) : 0;
    HamPlan pl = plan_hamming(q, n, nbytes, k, aligned ? hamming_qb(q, nbytes, K0, K0 ? sel_CAP(K0) : 0) : 1);
    if (!ws || ws_bytes < pl.total) { set_error("hamming: workspace %zu < %zu", ws_bytes, pl.total); return FPV_ERR_WORKSPACE; }
    char* w = static_cast<char*>(ws);
    uint8_t* dimmask = reinterpret_cast<uint8_t*>(w + pl.off_mask);
    uint64_t* partials = reinterpret_cast<uint64_t*>(w + pl.off_part);
    dimmask_kernel<<<1, 128, 0, st>>>(dimmask, nbytes, dims);
    FPV_LAUNCH_CHECK();
    HamParams p{};
    p.qbits = qbits; p.codes = codes; p.dimmask = dimmask; p.mask = mask_words; p.partials = partials;
    p.out_all = out_all; p.Q = q; p.N = n; p.nbytes = nbytes; p.K = pl.K; p.CAP = pl.CAP; p.parts = pl.parts;
    p.only_flagged = only_flagged;
    int rc = -1;
    if (aligned) {
        switch (nbytes) {
            case 16: rc = launch_fast<1>(p, pl, st); break;
            case 32: rc = launch_fast<2>(p, pl, st); break;
            case 64: rc = launch_fast<4>(p, pl, st); break;
            case 128: rc = launch_fast<8>(p, pl, st); break;
            case 256: rc = launch_fast<16>(p, pl, st); break;
            case 512: rc = launch_fast<32>(p, pl, st); break;
            default: break;
        }
    }
    if (rc == -1) {
        size_t smem = align_up((size_t)2 * nbytes, 16) + pl.smem;
        FPV_REQUIRE(smem <= (size_t)max_smem_optin(), "hamming: nbytes=%d k=%d needs %zu B shared memory", nbytes, k, smem);
        if (smem > 48 * 1024)
            FPV_CUDA(cudaFuncSetAttribute(hamming_generic_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        hamming_generic_kernel<<<dim3(pl.parts, (unsigned)q), 256, smem, st>>>(p);
        FPV_LAUNCH_CHECK();
        rc = FPV_OK;
    }
    if (rc != FPV_OK) return rc;
    if (k > 0) return launch_finalize(partials, q, pl.parts, pl.K, k, id_base, out_dist, out_idx, out_count, st, only_flagged);
    return FPV_OK;
}

extern "C" int fpv_hamming_topk(const uint8_t* qbits, int64_t q, const uint8_t* codes, int64_t n, int nbytes, int dims,
                                int k, const uint32_t* mask_words, int64_t id_base,
                                float* out_dist, int64_t* out_idx, int32_t* out_count, float* out_all,
                                void* ws, size_t ws_bytes, void* stream) {
    return hamming_topk_impl(qbits, q, codes, n, nbytes, dims, k, mask_words, id_base, out_dist, out_idx, out_count, out_all,
                             nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

// the same scan restricted to the queries whose only_flagged entry is non-zero (fallback of fpv_hamming_mma.cu)
namespace fpv {
size_t hamming_flagged_workspace(int64_t Q, int64_t N, int nbytes, int k) { return fpv_hamming_workspace(Q, N, nbytes, k); }
int hamming_topk_flagged(const uint8_t* qbits, int64_t q, const uint8_t* codes, int64_t n, int nbytes, int dims, int k,
                         const uint32_t* mask_words, int64_t id_base, float* out_dist, int64_t* out_idx, int32_t* out_count,
                         const uint32_t* only_flagged, void* ws, size_t ws_bytes, cudaStream_t st) {
    return hamming_topk_impl(qbits, q, codes, n, nbytes, dims, k, mask_words, id_base, out_dist, out_idx, out_count, nullptr,
                             only_flagged, ws, ws_bytes, st);
}
}  // namespace fpv
