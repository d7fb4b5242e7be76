// Shared device/host helpers for the FastPyVectorDB B200 search kernels.
//
// Ordering rule used EVERYWHERE (scan kernels, GEMM epilogue, per-CTA merge, cross-shard merge):
//   key = (ordered(distance) << 32) | row      smaller key == better
// i.e. ascending distance, ties broken by LOWEST row index.  The reference's own
// order among equal distances is arbitrary (np.argpartition, parallel_search.py:230-231).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stddef.h>
#include <math.h>

#include "../../include/fpv_b200.h"

#define FPV_KEY_MAX 0xFFFFFFFFFFFFFFFFull
#define FPV_FULL_MASK 0xFFFFFFFFu

namespace fpv {

// ---------------------------------------------------------------- host side error plumbing
void set_error(const char* fmt, ...);
int cuda_fail(cudaError_t e, const char* what);
int sm_count();
int max_smem_optin();
#define FPV_CUDA(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) return ::fpv::cuda_fail(e__, #x); } while (0)
void note_launch();
// after every <<<>>>: count the launch (fpv_launch_count) and surface launch-configuration errors
#define FPV_LAUNCH_CHECK() do { ::fpv::note_launch(); FPV_CUDA(cudaGetLastError()); } while (0)
#define FPV_REQUIRE(c, ...) do { if (!(c)) { ::fpv::set_error(__VA_ARGS__); return FPV_ERR_INVALID; } } while (0)

#ifdef __CUDACC__
#define FPV_HD __host__ __device__
#else
#define FPV_HD
#endif
FPV_HD static inline int next_pow2(int v) { int p = 1; while (p < v) p <<= 1; return p; }
// selector geometry for a requested k
FPV_HD static inline int sel_K(int k) { int p = next_pow2(k); return p < 32 ? 32 : p; }
FPV_HD static inline int sel_CAP(int K) { return K < 64 ? 64 : K; }
FPV_HD static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// the common tail: merge per-CTA partial lists into the final (dist, idx) rows
int launch_finalize(const uint64_t* partials, int64_t Q, int n_parts, int K, int k, int64_t id_base,
                    float* out_dist, int64_t* out_idx, int32_t* out_count, cudaStream_t st,
                    const uint32_t* only_flagged = nullptr);

#ifdef __CUDACC__
// ---------------------------------------------------------------- keys
__device__ __forceinline__ uint32_t f32_to_ordered(float d) {
    if (d != d) return 0xFFFFFFFFu;          // NaN sorts last (NumPy puts NaN last too)
    if (d == 0.0f) d = 0.0f;                 // -0 -> +0 so equal distances tie on the index
    uint32_t u = __float_as_uint(d);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float ordered_to_f32(uint32_t o) {
    uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
    return __uint_as_float(u);
}
__device__ __forceinline__ uint64_t make_key(float d, uint32_t row) {
    return ((uint64_t)f32_to_ordered(d) << 32) | (uint64_t)row;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FPV_FULL_MASK, v, o);
    return v;
}

__device__ __forceinline__ uint4 ldg_nc_u4(const uint4* p) {
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float4 ldg_nc_f4(const float4* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
    return r;
}

// ---------------------------------------------------------------- canonical exact fp32 dot product
// Every kernel that produces an exact fp32 distance (scan, re-rank, GEMM finish) accumulates in THIS order, so a
// row's distance is bit-identical no matter which path computed it: lane l takes the float4 chunks l, l+32, ...
// (x, y, z, w FMAs in that order), then the xor-butterfly warp_sum.
template <bool VEC>
__device__ __forceinline__ float4 load4(const float* v, int c, int D) {
    if (VEC) return ldg_nc_f4(reinterpret_cast<const float4*>(v) + c);
    float4 r;
    int j = c * 4;
    r.x = j < D ? __ldg(v + j) : 0.f;
    r.y = j + 1 < D ? __ldg(v + j + 1) : 0.f;
    r.z = j + 2 < D ? __ldg(v + j + 2) : 0.f;
    r.w = j + 3 < D ? __ldg(v + j + 3) : 0.f;
    return r;
}
// q4: the query, zero padded to whole float4 chunks (shared memory)
__device__ __forceinline__ float canonical_dot(const float* v, const float4* q4, int D, bool vec, int lane) {
    const int D4 = (D + 3) >> 2;
    float acc = 0.f;
    int c = lane;
    for (; c + 96 < D4; c += 128) {          // four 128-bit loads in flight per lane; the FMA order is unchanged
        const float4 x0 = vec ? load4<true>(v, c, D) : load4<false>(v, c, D);
        const float4 x1 = vec ? load4<true>(v, c + 32, D) : load4<false>(v, c + 32, D);
        const float4 x2 = vec ? load4<true>(v, c + 64, D) : load4<false>(v, c + 64, D);
        const float4 x3 = vec ? load4<true>(v, c + 96, D) : load4<false>(v, c + 96, D);
        float4 y = q4[c];
        acc = fmaf(x0.x, y.x, acc); acc = fmaf(x0.y, y.y, acc); acc = fmaf(x0.z, y.z, acc); acc = fmaf(x0.w, y.w, acc);
        y = q4[c + 32];
        acc = fmaf(x1.x, y.x, acc); acc = fmaf(x1.y, y.y, acc); acc = fmaf(x1.z, y.z, acc); acc = fmaf(x1.w, y.w, acc);
        y = q4[c + 64];
        acc = fmaf(x2.x, y.x, acc); acc = fmaf(x2.y, y.y, acc); acc = fmaf(x2.z, y.z, acc); acc = fmaf(x2.w, y.w, acc);
        y = q4[c + 96];
        acc = fmaf(x3.x, y.x, acc); acc = fmaf(x3.y, y.y, acc); acc = fmaf(x3.z, y.z, acc); acc = fmaf(x3.w, y.w, acc);
    }
    for (; c < D4; c += 32) {
        const float4 x = vec ? load4<true>(v, c, D) : load4<false>(v, c, D);
        const float4 y = q4[c];
        acc = fmaf(x.x, y.x, acc); acc = fmaf(x.y, y.y, acc); acc = fmaf(x.z, y.z, acc); acc = fmaf(x.w, y.w, acc);
    }
    return warp_sum(acc);
}
// Two rows at once (the gather of the re-rank kernels is latency bound: one row per warp keeps 4, then 1, then 1
// loads in flight).  Same chunk order and FMA order per row as canonical_dot, so the results are bit-identical;
// only the loads are issued earlier: 8 in flight, then the (up to 3 per row) tail chunks together.
__device__ __forceinline__ void canonical_dot2(const float* v0, const float* v1, const float4* q4, int D, bool vec, int lane,
                                               float& r0, float& r1) {
    if (!vec) { r0 = canonical_dot(v0, q4, D, false, lane); r1 = canonical_dot(v1, q4, D, false, lane); return; }
    const int D4 = (D + 3) >> 2;
    const float4* p0 = reinterpret_cast<const float4*>(v0);
    const float4* p1 = reinterpret_cast<const float4*>(v1);
    float a0 = 0.f, a1 = 0.f;
    int c = lane;
#define FPV_FMA4(acc, vv, qq) acc = fmaf(vv.x, qq.x, acc); acc = fmaf(vv.y, qq.y, acc); acc = fmaf(vv.z, qq.z, acc); acc = fmaf(vv.w, qq.w, acc)
    for (; c + 96 < D4; c += 128) {
        const float4 x0 = ldg_nc_f4(p0 + c), x1 = ldg_nc_f4(p0 + c + 32), x2 = ldg_nc_f4(p0 + c + 64), x3 = ldg_nc_f4(p0 + c + 96);
        const float4 z0 = ldg_nc_f4(p1 + c), z1 = ldg_nc_f4(p1 + c + 32), z2 = ldg_nc_f4(p1 + c + 64), z3 = ldg_nc_f4(p1 + c + 96);
        float4 y = q4[c];
        FPV_FMA4(a0, x0, y); FPV_FMA4(a1, z0, y);
        y = q4[c + 32];
        FPV_FMA4(a0, x1, y); FPV_FMA4(a1, z1, y);
        y = q4[c + 64];
        FPV_FMA4(a0, x2, y); FPV_FMA4(a1, z2, y);
        y = q4[c + 96];
        FPV_FMA4(a0, x3, y); FPV_FMA4(a1, z3, y);
    }
    {
        const bool h0 = c < D4, h1 = c + 32 < D4, h2 = c + 64 < D4;     // at most three chunks are left per lane
        const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
        const float4 x0 = h0 ? ldg_nc_f4(p0 + c) : zero, x1 = h1 ? ldg_nc_f4(p0 + c + 32) : zero, x2 = h2 ? ldg_nc_f4(p0 + c + 64) : zero;
        const float4 z0 = h0 ? ldg_nc_f4(p1 + c) : zero, z1 = h1 ? ldg_nc_f4(p1 + c + 32) : zero, z2 = h2 ? ldg_nc_f4(p1 + c + 64) : zero;
        if (h0) { const float4 y = q4[c]; FPV_FMA4(a0, x0, y); FPV_FMA4(a1, z0, y); }
        if (h1) { const float4 y = q4[c + 32]; FPV_FMA4(a0, x1, y); FPV_FMA4(a1, z1, y); }
        if (h2) { const float4 y = q4[c + 64]; FPV_FMA4(a0, x2, y); FPV_FMA4(a1, z2, y); }
    }
#undef FPV_FMA4
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        a0 += __shfl_xor_sync(FPV_FULL_MASK, a0, o);
        a1 += __shfl_xor_sync(FPV_FULL_MASK, a1, o);
    }
    r0 = a0; r1 = a1;
}
__device__ __forceinline__ bool rows_vectorizable(const float* db, int D, int64_t ld) {
    return (D % 4 == 0) && (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(db) & 15) == 0);
}

// ---------------------------------------------------------------- warp-level bitonic networks over shared memory
__device__ __forceinline__ void bitonic_sort_warp(uint64_t* a, int n, int lane) {
    for (int size = 2; size <= n; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = lane; t < (n >> 1); t += 32) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = (lo & size) == 0;
                uint64_t x = a[lo], y = a[hi];
                if ((x > y) == up) { a[lo] = y; a[hi] = x; }
            }
            __syncwarp();
        }
    }
}
// a[0..n) bitonic -> ascending
__device__ __forceinline__ void bitonic_merge_warp(uint64_t* a, int n, int lane) {
    for (int stride = n >> 1; stride > 0; stride >>= 1) {
        for (int t = lane; t < (n >> 1); t += 32) {
            int lo = 2 * t - (t & (stride - 1));
            int hi = lo + stride;
            uint64_t x = a[lo], y = a[hi];
            if (x > y) { a[lo] = y; a[hi] = x; }
        }
        __syncwarp();
    }
}
// dst, src sorted ascending (K each) -> dst = K smallest of the union, ascending
__device__ __forceinline__ void merge_sorted_into(uint64_t* dst, const uint64_t* src, int K, int lane) {
    for (int i = lane; i < K; i += 32) {
        uint64_t x = dst[i], y = src[K - 1 - i];
        dst[i] = x < y ? x : y;
    }
    __syncwarp();
    bitonic_merge_warp(dst, K, lane);
}

// ---------------------------------------------------------------- per-warp streaming top-K selector
// One instance per warp, QB independent queries.  Shared memory per (warp, query): K "best" keys kept
// sorted + CAP candidate slots.  A key is a candidate only if it beats tau = current K-th best, so
// after warm-up almost every element costs one compare (+ one ballot for the lane-per-row form).
template <int QB>
struct WarpSelect {
    uint64_t* mem;
    int K, CAP;
    uint64_t tau[QB];
    int ncand[QB];

    __device__ __forceinline__ uint64_t* best(int q) const { return mem + (size_t)q * (K + CAP); }
    __device__ __forceinline__ uint64_t* cand(int q) const { return best(q) + K; }

    __device__ __forceinline__ void init(uint64_t* warp_mem, int K_, int CAP_, int lane) {
        mem = warp_mem; K = K_; CAP = CAP_;
#pragma unroll
        for (int q = 0; q < QB; ++q) {
            tau[q] = FPV_KEY_MAX; ncand[q] = 0;
            uint64_t* b = best(q);
            for (int i = lane; i < K; i += 32) b[i] = FPV_KEY_MAX;
        }
        __syncwarp();
    }
    __device__ __forceinline__ void flush(int q, int lane) {
        uint64_t* b = best(q);
        uint64_t* c = cand(q);
        for (int i = ncand[q] + lane; i < CAP; i += 32) c[i] = FPV_KEY_MAX;
        __syncwarp();
        bitonic_sort_warp(c, CAP, lane);
        merge_sorted_into(b, c, K, lane);     // CAP >= K: only the K smallest candidates matter
        tau[q] = b[K - 1];
        ncand[q] = 0;
    }
    // every lane passes the SAME key (warp-per-row kernels)
    __device__ __forceinline__ void add_uniform(int q, uint64_t key, int lane) {
        if (key < tau[q]) {
            if (lane == 0) cand(q)[ncand[q]] = key;
            if (++ncand[q] == CAP) flush(q, lane);
        }
    }
    // every lane passes its OWN key (lane-per-row kernels)
    __device__ __forceinline__ void add_lanes(int q, uint64_t key, bool valid, int lane) {
        bool pass = valid && key < tau[q];
        unsigned m = __ballot_sync(FPV_FULL_MASK, pass);
        if (m) {
            if (pass) cand(q)[ncand[q] + __popc(m & ((1u << lane) - 1u))] = key;
            ncand[q] += __popc(m);
            if (ncand[q] > CAP - 32) flush(q, lane);
        }
    }
    __device__ __forceinline__ void flush_all(int lane) {
#pragma unroll
        for (int q = 0; q < QB; ++q) flush(q, lane);
    }
};

// After every warp called flush_all(): merge the W per-warp lists of each query into warp 0's list and
// store K sorted keys per query to out[q * out_stride .. +K).  Contains the needed __syncthreads().
template <int QB>
__device__ __forceinline__ void block_merge_store(uint64_t* sel_base, int K, int CAP, int nq,
                                                  uint64_t* out, size_t out_stride) {
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const size_t per_warp = (size_t)QB * (K + CAP);
    if (nq == 1 && W > 2 && (W & (W - 1)) == 0) {
        // one query, W lists: pairwise tree (log2 W levels, all warps busy) instead of W - 1 merges on warp 0 -- with 32
        // warps the serial form was a ~20 us single-warp tail at the end of every CTA
        for (int s = W >> 1; s >= 1; s >>= 1) {
            if (warp < s) merge_sorted_into(sel_base + (size_t)warp * per_warp, sel_base + (size_t)(warp + s) * per_warp, K, lane);
            __syncthreads();
        }
        for (int i = threadIdx.x; i < K; i += blockDim.x) out[i] = sel_base[i];
        return;
    }
    for (int q = warp; q < nq; q += W) {
        uint64_t* dst = sel_base + (size_t)q * (K + CAP);
        for (int w = 1; w < W; ++w)
            merge_sorted_into(dst, sel_base + w * per_warp + (size_t)q * (K + CAP), K, lane);
        uint64_t* o = out + (size_t)q * out_stride;
        for (int i = lane; i < K; i += 32) o[i] = dst[i];
    }
}

// Consumer side of fpv_peer_put (fpv_peer.cu): block until the first n flag words of this rank's region have reached
// `epoch` (wrap-safe compare), i.e. until every peer's contribution has landed in local memory.  Contains a
// __syncthreads: call from uniform control flow.  A peer that never arrives traps instead of hanging the GPU.
__device__ __forceinline__ void peer_wait(const uint32_t* flags, int n, uint32_t epoch) {
    if (!flags) return;
    if ((int)threadIdx.x < n) {
        const volatile uint32_t* f = flags + threadIdx.x;
        uint32_t spins = 0;
        while ((int32_t)(*f - epoch) < 0) {
            __nanosleep(40);
            if (++spins > (1u << 25)) __trap();
        }
    }
    __syncthreads();
}

__device__ __forceinline__ bool mask_bit(const uint32_t* __restrict__ mask, int64_t row) {
    return (__ldg(mask + (row >> 5)) >> (row & 31)) & 1u;
}
#endif  // __CUDACC__

}  // namespace fpv
