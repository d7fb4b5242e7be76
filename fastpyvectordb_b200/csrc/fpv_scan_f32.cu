// Exact fp32 streaming scan with fused top-k (HBM-bound small-batch regime), row norms, full distance rows
// and the exact re-rank of gathered candidates.
//
// Replaces the NumPy bodies of _compute_distances_vectorized (parallel_search.py:105-134),
// search_parallel (:209-244), search_chunked_parallel (:326-368) and the re-rank block of
// ParallelCollection.search_hybrid (:919-934).
//
// Layout: db is [N][ld] fp32 row major in HBM, read exactly once per batch of QB queries with 128-bit
// coalesced loads (one warp per row, 512 B per load instruction).  The QB queries of a pass live in shared
// memory; nothing but the final top-k ever leaves the SM.
#include <algorithm>

#include "fpv_common.cuh"

namespace fpv {

// ----------------------------------------------------------------------------------------------------
// query preparation: qprep = q / (||q|| + 1e-10) for cosine (parallel_search.py:121), q otherwise;
// qsq = q.q (parallel_search.py:129).
// ----------------------------------------------------------------------------------------------------
__global__ void prep_queries_kernel(const float* __restrict__ q, int64_t Q, int D, int metric,
                                    float* __restrict__ qprep, float* __restrict__ qsq) {
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= Q) return;
    const float* src = q + row * D;
    float s = 0.f;
    for (int j = lane; j < D; j += 32) { float x = src[j]; s = fmaf(x, x, s); }
    s = warp_sum(s);
    float inv = 1.0f;
    if (metric == FPV_METRIC_COSINE) inv = 1.0f / (sqrtf(s) + 1e-10f);
    for (int j = lane; j < D; j += 32)
        qprep[row * D + j] = (metric == FPV_METRIC_COSINE) ? src[j] * inv : src[j];
    if (lane == 0) qsq[row] = s;
}

__global__ void row_sqnorm_kernel(const float* __restrict__ db, int64_t N, int D, int64_t ld,
                                  float* __restrict__ out) {
    const int lane = threadIdx.x & 31;
    const int64_t w0 = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int64_t stride = (int64_t)gridDim.x * (blockDim.x >> 5);
    const bool vec = (D % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(db) & 15) == 0);
    for (int64_t row = w0; row < N; row += stride) {
        const float* v = db + row * ld;
        float s = 0.f;
        if (vec) {
            const float4* v4 = reinterpret_cast<const float4*>(v);
            for (int c = lane; c < (D >> 2); c += 32) {
                float4 x = ldg_nc_f4(v4 + c);
                s = fmaf(x.x, x.x, s); s = fmaf(x.y, x.y, s); s = fmaf(x.z, x.z, s); s = fmaf(x.w, x.w, s);
            }
        } else {
            for (int j = lane; j < D; j += 32) { float x = v[j]; s = fmaf(x, x, s); }
        }
        s = warp_sum(s);
        if (lane == 0) out[row] = s;
    }
}

__device__ __forceinline__ float finish_distance(int metric, float dot, float vsq, float qsq) {
    if (metric == FPV_METRIC_COSINE) return 1.0f - dot / (sqrtf(vsq) + 1e-10f);
    if (metric == FPV_METRIC_L2) return sqrtf(fmaxf(qsq + vsq - 2.0f * dot, 0.0f));
    return -dot;
}

struct ScanF32Params {
    const float* qprep;     // [Q][D]
    const float* qsq;       // [Q]
    const float* db;
    const uint32_t* mask;
    const float* row_sq;    // may be null
    uint64_t* partials;     // [Q][parts][K]
    float* out_all;         // may be null: [Q][N]
    const uint32_t* flags;  // may be null: [Q], a CTA whose queries are all unflagged exits at once
    int64_t Q, N, ld;
    int D, metric, K, CAP, parts;
};

template <int QB>
__device__ __forceinline__ void fma_chunk(const float4& x, const float4* qs4, int c, int D4, float (&acc)[QB], float& vsq) {
    vsq = fmaf(x.x, x.x, vsq); vsq = fmaf(x.y, x.y, vsq); vsq = fmaf(x.z, x.z, vsq); vsq = fmaf(x.w, x.w, vsq);
#pragma unroll
    for (int q = 0; q < QB; ++q) {
        float4 y = qs4[q * D4 + c];
        acc[q] = fmaf(x.x, y.x, acc[q]); acc[q] = fmaf(x.y, y.y, acc[q]);
        acc[q] = fmaf(x.z, y.z, acc[q]); acc[q] = fmaf(x.w, y.w, acc[q]);
    }
}

// grid = (parts, ceil(Q/QB)); block = 256.  smem: QB*D4*16 bytes of queries + W*QB*(K+CAP) keys.
template <int QB, bool VEC>
__global__ void __launch_bounds__(256) scan_f32_kernel(ScanF32Params p) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int D4 = (p.D + 3) >> 2;
    float4* qs4 = reinterpret_cast<float4*>(smem_raw);
    uint64_t* sel_base = reinterpret_cast<uint64_t*>(smem_raw + (size_t)QB * D4 * sizeof(float4));
    const int64_t q0 = (int64_t)blockIdx.y * QB;
    const int nq = (int)min((int64_t)QB, p.Q - q0);
    if (p.flags) {          // fallback mode (fpv_gemm_topk.cu): only flagged queries are recomputed
        bool any = false;
        for (int q = 0; q < nq; ++q) any |= p.flags[q0 + q] != 0;
        if (!any) return;
    }

    {   // stage the queries of this pass (zero padded)
        float* qs = reinterpret_cast<float*>(qs4);
        for (int i = threadIdx.x; i < QB * D4 * 4; i += blockDim.x) {
            int q = i / (D4 * 4), j = i - q * (D4 * 4);
            qs[i] = (q < nq && j < p.D) ? p.qprep[(q0 + q) * p.D + j] : 0.f;
        }
    }
    float qsq[QB];
#pragma unroll
    for (int q = 0; q < QB; ++q) qsq[q] = (q < nq) ? p.qsq[q0 + q] : 0.f;

    WarpSelect<QB> sel;
    const bool select = p.K > 0;
    if (select) sel.init(sel_base + (size_t)warp * QB * (p.K + p.CAP), p.K, p.CAP, lane);
    __syncthreads();

    for (int64_t row = (int64_t)blockIdx.x * W + warp; row < p.N; row += (int64_t)gridDim.x * W) {
        if (p.mask && !mask_bit(p.mask, row)) continue;
        const float* v = p.db + row * p.ld;
        float acc[QB];
#pragma unroll
        for (int q = 0; q < QB; ++q) acc[q] = 0.f;
        float vsq = 0.f;
        int c = lane;
        for (; c + 96 < D4; c += 128) {       // four independent 128-bit loads in flight per lane
            float4 x0 = load4<VEC>(v, c, p.D), x1 = load4<VEC>(v, c + 32, p.D);
            float4 x2 = load4<VEC>(v, c + 64, p.D), x3 = load4<VEC>(v, c + 96, p.D);
            fma_chunk<QB>(x0, qs4, c, D4, acc, vsq);
            fma_chunk<QB>(x1, qs4, c + 32, D4, acc, vsq);
            fma_chunk<QB>(x2, qs4, c + 64, D4, acc, vsq);
            fma_chunk<QB>(x3, qs4, c + 96, D4, acc, vsq);
        }
        for (; c < D4; c += 32) {
            float4 x = load4<VEC>(v, c, p.D);
            fma_chunk<QB>(x, qs4, c, D4, acc, vsq);
        }
        vsq = p.row_sq ? __ldg(p.row_sq + row) : warp_sum(vsq);
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            float dot = warp_sum(acc[q]);
            if (q < nq) {
                float d = finish_distance(p.metric, dot, vsq, qsq[q]);
                if (p.out_all && lane == 0) p.out_all[(q0 + q) * p.N + row] = d;
                if (select) sel.add_uniform(q, make_key(d, (uint32_t)row), lane);
            }
        }
    }
    if (select) {
        sel.flush_all(lane);
        block_merge_store<QB>(sel_base, p.K, p.CAP, nq,
                              p.partials + ((size_t)q0 * p.parts + blockIdx.x) * p.K, (size_t)p.parts * p.K);
    }
}

struct ScanPlan { int QB, parts, K, CAP, nqc; size_t smem, off_qprep, off_qsq, off_part, total; };

static ScanPlan plan_scan_f32(int64_t Q, int64_t N, int D, int k) {
    ScanPlan pl{};
    pl.K = k > 0 ? sel_K(k) : 0;
    pl.CAP = k > 0 ? sel_CAP(pl.K) : 0;
    const int D4 = (D + 3) / 4;
    const size_t budget = 160 * 1024;
    int QB = 1;
    while (QB < 8 && QB < Q) QB <<= 1;                    // smallest pow2 >= Q, at most 8
    auto smem_of = [&](int qb) { return (size_t)qb * D4 * 16 + (size_t)8 * qb * (pl.K + pl.CAP) * 8; };
    while (QB > 1 && smem_of(QB) > budget) QB >>= 1;
    pl.QB = QB;
    pl.smem = smem_of(QB);
    pl.nqc = (int)((Q + QB - 1) / QB);
    int sms = sm_count();
    int64_t want = (int64_t)sms * 4;                      // CTAs in flight
    int64_t parts = pl.nqc > 0 ? (want + pl.nqc - 1) / pl.nqc : 1;
    int64_t max_parts = (N + 7) / 8;                      // at least one row per warp
    if (parts > max_parts) parts = max_parts;
    if (parts < 1) parts = 1;
    pl.parts = (int)parts;
    pl.off_qprep = 0;
    pl.off_qsq = align_up((size_t)Q * D * 4, 256);
    pl.off_part = pl.off_qsq + align_up((size_t)Q * 4, 256);
    pl.total = pl.off_part + (size_t)Q * pl.parts * pl.K * 8;
    return pl;
}

template <int QB>
static int launch_scan(const ScanF32Params& p, const ScanPlan& pl, bool vec, cudaStream_t st) {
    dim3 grid(pl.parts, pl.nqc);
    if (vec) {
        if (pl.smem > 48 * 1024)
            FPV_CUDA(cudaFuncSetAttribute(scan_f32_kernel<QB, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        scan_f32_kernel<QB, true><<<grid, 256, pl.smem, st>>>(p);
    } else {
        if (pl.smem > 48 * 1024)
            FPV_CUDA(cudaFuncSetAttribute(scan_f32_kernel<QB, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)pl.smem));
        scan_f32_kernel<QB, false><<<grid, 256, pl.smem, st>>>(p);
    }
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

static int run_scan_f32(const float* queries, int64_t Q, const float* db, int64_t N, int D, int64_t ld, int metric,
                        int k, const uint32_t* mask, const float* row_sq, int64_t id_base, float* out_dist,
                        int64_t* out_idx, int32_t* out_count, float* out_all, void* ws, size_t ws_bytes,
                        cudaStream_t st, const uint32_t* flags = nullptr) {
    FPV_REQUIRE(Q >= 0 && N >= 0 && D >= 1 && ld >= D, "scan_f32: bad shape Q=%lld N=%lld D=%d ld=%lld",
                (long long)Q, (long long)N, D, (long long)ld);
    FPV_REQUIRE(metric >= 0 && metric <= 2, "scan_f32: unknown metric %d", metric);
    FPV_REQUIRE(k >= 0 && k <= FPV_MAX_K, "scan_f32: k=%d outside [0,%d]", k, FPV_MAX_K);
    FPV_REQUIRE(N < (1ll << 32), "scan_f32: N=%lld rows per call exceeds 2^32-1 (shard the database)", (long long)N);
    FPV_REQUIRE(D <= 16384, "scan_f32: D=%d > 16384", D);
    if (Q == 0) return FPV_OK;
    FPV_REQUIRE(queries && (db || N == 0), "scan_f32: null pointer");
    FPV_REQUIRE(k == 0 || (out_dist && out_idx), "scan_f32: null output");
    ScanPlan pl = plan_scan_f32(Q, N, D, k);
    if (ws_bytes < pl.total || !ws) { set_error("scan_f32: workspace %zu < %zu", ws_bytes, pl.total); return FPV_ERR_WORKSPACE; }
    FPV_REQUIRE(pl.smem <= (size_t)max_smem_optin(), "scan_f32: D=%d k=%d needs %zu B of shared memory", D, k, pl.smem);
    char* w = static_cast<char*>(ws);
    float* qprep = reinterpret_cast<float*>(w + pl.off_qprep);
    float* qsq = reinterpret_cast<float*>(w + pl.off_qsq);
    uint64_t* partials = reinterpret_cast<uint64_t*>(w + pl.off_part);
    prep_queries_kernel<<<(unsigned)((Q + 7) / 8), 256, 0, st>>>(queries, Q, D, metric, qprep, qsq);
    FPV_LAUNCH_CHECK();
    ScanF32Params p{};
    p.qprep = qprep; p.qsq = qsq; p.db = db; p.mask = mask; p.row_sq = row_sq; p.partials = partials;
    p.out_all = out_all; p.flags = flags; p.Q = Q; p.N = N; p.ld = ld; p.D = D; p.metric = metric; p.K = pl.K; p.CAP = pl.CAP;
    p.parts = pl.parts;
    const bool vec = (D % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(db) & 15) == 0);
    int rc;
    switch (pl.QB) {
        case 1: rc = launch_scan<1>(p, pl, vec, st); break;
        case 2: rc = launch_scan<2>(p, pl, vec, st); break;
        case 4: rc = launch_scan<4>(p, pl, vec, st); break;
        default: rc = launch_scan<8>(p, pl, vec, st); break;
    }
    if (rc != FPV_OK) return rc;
    if (k > 0) return launch_finalize(partials, Q, pl.parts, pl.K, k, id_base, out_dist, out_idx, out_count, st, flags);
    return FPV_OK;
}

// ----------------------------------------------------------------------------------------------------
// exact re-rank of candidates: one CTA per query, one warp per candidate row, block bitonic sort.
// ----------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) rerank_kernel(const float* __restrict__ queries, const float* __restrict__ db,
                                                     int64_t N, int D, int64_t ld, int metric,
                                                     const int64_t* __restrict__ cand, int C, int P, int k,
                                                     const float* __restrict__ row_sq, int64_t id_base,
                                                     float* __restrict__ out_dist, int64_t* __restrict__ out_idx,
                                                     int32_t* __restrict__ out_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem_raw);             // [P]
    float* qs = reinterpret_cast<float*>(smem_raw + (size_t)P * 8);     // [D]
    __shared__ float s_qsq, s_inv;
    __shared__ int s_cnt;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q = blockIdx.x;
    const float* qsrc = queries + q * D;
    if (warp == 0) {
        float s = 0.f;
        for (int j = lane; j < D; j += 32) { float x = qsrc[j]; s = fmaf(x, x, s); }
        s = warp_sum(s);
        if (lane == 0) { s_qsq = s; s_inv = (metric == FPV_METRIC_COSINE) ? 1.0f / (sqrtf(s) + 1e-10f) : 1.0f; s_cnt = 0; }
    }
    __syncthreads();
    const int D4 = (D + 3) >> 2;
    for (int j = threadIdx.x; j < D4 * 4; j += blockDim.x)
        qs[j] = j < D ? ((metric == FPV_METRIC_COSINE) ? qsrc[j] * s_inv : qsrc[j]) : 0.f;
    for (int i = threadIdx.x; i < P; i += blockDim.x) keys[i] = FPV_KEY_MAX;
    __syncthreads();
    const bool vec = rows_vectorizable(db, D, ld);
    const float4* qs4 = reinterpret_cast<const float4*>(qs);
    for (int c = warp; c < C; c += W) {
        int64_t row = cand[q * C + c];
        if (row < 0 || row >= N) continue;
        const float* v = db + row * ld;
        if (metric == FPV_METRIC_L2_DIFF) {                          // uniform: explicit differences, canonical chunk order
            float s = 0.f;
            for (int cc = lane; cc < D4; cc += 32) {
                const float4 x = vec ? load4<true>(v, cc, D) : load4<false>(v, cc, D);
                const float4 y = qs4[cc];
                const float a = x.x - y.x, b = x.y - y.y, c2 = x.z - y.z, d2 = x.w - y.w;
                s = fmaf(a, a, s); s = fmaf(b, b, s); s = fmaf(c2, c2, s); s = fmaf(d2, d2, s);
            }
            s = warp_sum(s);
            if (lane == 0) keys[c] = make_key(sqrtf(s), (uint32_t)row);
            continue;
        }
        const float dot = canonical_dot(v, qs4, D, vec, lane);       // same order as the scan kernel: bit-identical
        float vsq;
        if (row_sq) vsq = __ldg(row_sq + row);
        else {
            float s = 0.f;
            for (int cc = lane; cc < D4; cc += 32) {
                const float4 x = vec ? load4<true>(v, cc, D) : load4<false>(v, cc, D);
                s = fmaf(x.x, x.x, s); s = fmaf(x.y, x.y, s); s = fmaf(x.z, x.z, s); s = fmaf(x.w, x.w, s);
            }
            vsq = warp_sum(s);
        }
        if (lane == 0) keys[c] = make_key(finish_distance(metric, dot, vsq, s_qsq), (uint32_t)row);
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1)
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
                bool up = (lo & size) == 0;
                uint64_t x = keys[lo], y = keys[hi];
                if ((x > y) == up) { keys[lo] = y; keys[hi] = x; }
            }
            __syncthreads();
        }
    // candidate ids are expected to be unique per query (they come from a top-k list)
    int cnt = 0;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        uint64_t key = i < P ? keys[i] : FPV_KEY_MAX;
        bool ok = key != FPV_KEY_MAX;
        out_dist[q * k + i] = ok ? ordered_to_f32((uint32_t)(key >> 32)) : INFINITY;
        out_idx[q * k + i] = ok ? id_base + (int64_t)(uint32_t)key : -1;
        cnt += ok;
    }
    if (out_count) {
        if (cnt) atomicAdd(&s_cnt, cnt);
        __syncthreads();
        if (threadIdx.x == 0) out_count[q] = s_cnt;
    }
}

// full rows of explicit-difference L2 distances: grid = (row blocks, Q), one warp per row, the query in shared memory
__global__ void __launch_bounds__(256) l2_diff_rows_kernel(const float* __restrict__ queries, const float* __restrict__ db,
                                                           int64_t N, int D, int64_t ld, float* __restrict__ out_all) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    float* qs = reinterpret_cast<float*>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q = blockIdx.y;
    const int D4 = (D + 3) >> 2;
    for (int j = threadIdx.x; j < D4 * 4; j += blockDim.x) qs[j] = j < D ? queries[q * D + j] : 0.f;
    __syncthreads();
    const bool vec = rows_vectorizable(db, D, ld);
    const float4* qs4 = reinterpret_cast<const float4*>(qs);
    for (int64_t row = (int64_t)blockIdx.x * W + warp; row < N; row += (int64_t)gridDim.x * W) {
        const float* v = db + row * ld;
        float s = 0.f;
        for (int cc = lane; cc < D4; cc += 32) {
            const float4 x = vec ? load4<true>(v, cc, D) : load4<false>(v, cc, D);
            const float4 y = qs4[cc];
            const float a = x.x - y.x, b = x.y - y.y, c2 = x.z - y.z, d2 = x.w - y.w;
            s = fmaf(a, a, s); s = fmaf(b, b, s); s = fmaf(c2, c2, s); s = fmaf(d2, d2, s);
        }
        s = warp_sum(s);
        if (lane == 0) out_all[q * N + row] = sqrtf(s);
    }
}

size_t scan_f32_flagged_workspace(int64_t Q, int64_t N, int D, int k) { return plan_scan_f32(Q, N, D, k).total; }
int scan_f32_flagged(const float* queries, int64_t Q, const float* db, int64_t N, int D, int64_t ld, int metric, int k,
                     const float* row_sq, int64_t id_base, const uint32_t* flags, const uint32_t* mask_words, float* out_dist,
                     int64_t* out_idx, int32_t* out_count, void* ws, size_t ws_bytes, cudaStream_t st) {
    return run_scan_f32(queries, Q, db, N, D, ld, metric, k, mask_words, row_sq, id_base, out_dist, out_idx, out_count, nullptr,
                        ws, ws_bytes, st, flags);
}

}  // namespace fpv

using namespace fpv;

extern "C" int fpv_row_sqnorm_f32(const float* db, int64_t n, int d, int64_t ld, float* row_sq, void* stream) {
    FPV_REQUIRE(n >= 0 && d >= 1 && ld >= d, "row_sqnorm: bad shape n=%lld d=%d ld=%lld", (long long)n, d, (long long)ld);
    if (n == 0) return FPV_OK;
    FPV_REQUIRE(db && row_sq, "row_sqnorm: null pointer");
    int64_t blocks = (n + 7) / 8;
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    row_sqnorm_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(db, n, d, ld, row_sq);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

extern "C" size_t fpv_scan_f32_workspace(int64_t q, int64_t n, int d, int k) {
    if (q <= 0 || d <= 0 || k < 0) return 256;
    return plan_scan_f32(q, n, d, k).total;
}

extern "C" int fpv_scan_f32_topk(const float* queries, int64_t q, const float* db, int64_t n, int d, int64_t ld,
                                 int metric, int k, const uint32_t* mask_words, const float* row_sq, int64_t id_base,
                                 float* out_dist, int64_t* out_idx, int32_t* out_count,
                                 void* ws, size_t ws_bytes, void* stream) {
    FPV_REQUIRE(k >= 1, "scan_f32_topk: k=%d must be >= 1", k);
    return run_scan_f32(queries, q, db, n, d, ld, metric, k, mask_words, row_sq, id_base, out_dist, out_idx,
                        out_count, nullptr, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int fpv_distances_f32(const float* queries, int64_t q, const float* db, int64_t n, int d, int64_t ld,
                                 int metric, const float* row_sq, float* out_all, void* ws, size_t ws_bytes, void* stream) {
    FPV_REQUIRE(out_all || q == 0 || n == 0, "distances_f32: null output");
    if (metric == FPV_METRIC_L2_DIFF) {
        FPV_REQUIRE(q >= 0 && n >= 0 && d >= 1 && d <= 16384 && ld >= d && q <= 65535, "distances_f32: bad shape");
        if (q == 0 || n == 0) return FPV_OK;
        FPV_REQUIRE(queries && db, "distances_f32: null pointer");
        const int64_t blocks = std::min<int64_t>((n + 7) / 8, (int64_t)sm_count() * 8);
        l2_diff_rows_kernel<<<dim3((unsigned)blocks, (unsigned)q), 256, (size_t)((d + 3) / 4 * 4) * 4, (cudaStream_t)stream>>>(
            queries, db, n, d, ld, out_all);
        FPV_LAUNCH_CHECK();
        return FPV_OK;
    }
    return run_scan_f32(queries, q, db, n, d, ld, metric, 0, nullptr, row_sq, 0, nullptr, nullptr, nullptr,
                        out_all, ws, ws_bytes, (cudaStream_t)stream);
}

extern "C" int fpv_rerank_f32(const float* queries, int64_t q, const float* db, int64_t n, int d, int64_t ld, int metric,
                              const int64_t* cand_idx, int c, int k, const float* row_sq, int64_t id_base,
                              float* out_dist, int64_t* out_idx, int32_t* out_count, void* stream) {
    FPV_REQUIRE(q >= 0 && n >= 0 && d >= 1 && ld >= d && c >= 1 && k >= 1 && k <= c,
                "rerank: bad shape q=%lld n=%lld d=%d c=%d k=%d", (long long)q, (long long)n, d, c, k);
    FPV_REQUIRE(metric >= 0 && metric <= FPV_METRIC_L2_DIFF, "rerank: unknown metric %d", metric);
    FPV_REQUIRE(n < (1ll << 32), "rerank: N too large");
    if (q == 0) return FPV_OK;
    FPV_REQUIRE(queries && db && cand_idx && out_dist && out_idx, "rerank: null pointer");
    int P = next_pow2(c);
    if (P < 2) P = 2;
    size_t smem = (size_t)P * 8 + (size_t)((d + 3) / 4 * 4) * 4;
    FPV_REQUIRE(smem <= (size_t)max_smem_optin(), "rerank: c=%d d=%d needs %zu B shared memory", c, d, smem);
    if (smem > 48 * 1024)
        FPV_CUDA(cudaFuncSetAttribute(rerank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    rerank_kernel<<<(unsigned)q, 256, smem, (cudaStream_t)stream>>>(queries, db, n, d, ld, metric, cand_idx, c, P, k,
                                                                     row_sq, id_base, out_dist, out_idx, out_count);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}
