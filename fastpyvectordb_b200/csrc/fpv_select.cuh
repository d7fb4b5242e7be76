// Block-level selection helpers shared by the filter-then-re-rank paths (fpv_gemm_topk.cu, fpv_sq_mma.cu).
#pragma once
#include "fpv_common.cuh"

namespace fpv {

__device__ __forceinline__ void block_bitonic_sort(uint64_t* keys, int P) {
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
}


// ---- between slabs / after the last slab: radix-select based tighten and finish -----------------------------------
// Value (high 32 bits) of the kth smallest (1-based) of the c keys in shared memory; the returned key has that value in
// its high word and zeros below.  blockDim.x >= 256; contains __syncthreads (call it from uniform control flow).
//
// Radix select on the values RELATIVE TO THEIR MINIMUM, starting at the highest bit in which the keys actually differ:
// the approximate scores of one query share their sign / exponent / leading mantissa bits, so a select over the raw
// bytes put almost every key of the first passes into one or two bins -- 32-way same-address shared atomics, which is
// what the first version of this routine spent its time on (47 us per 4096 queries x 2048 keys; the tighten kernel
// ran three times per search).  Here the first pass spreads the keys over all 256 bins and every later pass only
// touches the few keys of one bin.
__device__ __forceinline__ uint64_t block_radix_select(const uint64_t* keys, int c, int kth, uint32_t* hist, int* s_bin, int* s_need) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (int i = threadIdx.x; i < c; i += blockDim.x) { const uint32_t hi = (uint32_t)(keys[i] >> 32); mn = min(mn, hi); mx = max(mx, hi); }
    mn = __reduce_min_sync(FPV_FULL_MASK, mn);
    mx = __reduce_max_sync(FPV_FULL_MASK, mx);
    if (threadIdx.x == 0) { hist[0] = 0xFFFFFFFFu; hist[1] = 0u; }
    __syncthreads();
    if (lane == 0) { atomicMin(&hist[0], mn); atomicMax(&hist[1], mx); }
    __syncthreads();
    mn = hist[0]; mx = hist[1];
    __syncthreads();
    const uint32_t range = mx - mn;
    if (range == 0u) return (uint64_t)mn << 32;
    int shift = 31 - __clz(range) - 7;                    // first digit = the top 8 bits of the range
    if (shift < 0) shift = 0;
    uint32_t base = 0u;                                   // lower end (relative to mn) of the bin selected so far
    int need = kth;
    while (true) {
        if (threadIdx.x < 256) hist[threadIdx.x] = 0;
        __syncthreads();
        for (int i = threadIdx.x; i < c; i += blockDim.x) {
            const uint32_t v = (uint32_t)(keys[i] >> 32) - mn;
            if (v >= base) {
                const uint32_t bin = (v - base) >> shift;
                if (bin < 256u) atomicAdd(&hist[bin], 1u);
            }
        }
        __syncthreads();
        if (warp == 0) {
            uint32_t h[8], sum = 0;
#pragma unroll
            for (int b = 0; b < 8; ++b) { h[b] = hist[lane * 8 + b]; sum += h[b]; }
            uint32_t incl = sum;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { uint32_t t = __shfl_up_sync(FPV_FULL_MASK, incl, o); if (lane >= o) incl += t; }
            const uint32_t excl = incl - sum;
            if (excl < (uint32_t)need && (uint32_t)need <= incl) {
                uint32_t run = excl;
#pragma unroll
                for (int b = 0; b < 8; ++b) {
                    if ((uint32_t)need <= run + h[b]) { *s_bin = lane * 8 + b; *s_need = need - (int)run; break; }
                    run += h[b];
                }
            }
        }
        __syncthreads();
        base += (uint32_t)(*s_bin) << shift;
        need = *s_need;
        __syncthreads();
        if (shift == 0) break;
        shift = shift > 8 ? shift - 8 : 0;
    }
    return (uint64_t)(mn + base) << 32;
}

// Between slabs.  With a_k = the k-th best approximate value seen so far, every row that can still end up in the exact
// top-k has approx <= a_k + 2E (a_k only decreases as more rows are seen), so that is the tightest threshold the
// certificate allows: keep exactly those candidates and raise the threshold to it.  (Keeping a fixed number of
// candidates instead — the first version — needed 2-4x more slots than this to leave room for the 2E margin and
// produced 2-3.5x more epilogue hits per slab.)
// `approx_out` (row-sharded search, after the LAST slab): additionally writes the k smallest approximate values of this
// shard (ordered-uint32 form, any order, padded with ordered(+inf)) to approx_out[q][0..k) -- what the other ranks
// need to find the GLOBAL k-th approximate value.
// blockDim.x = any multiple of 32 >= 256; dynamic smem = CAP * 8 bytes.  sample_groups: as in tighten_warp_kernel.
// thr_shift: added to the bound that becomes the next threshold (not to the bound the candidates are kept under).
// Integer metrics scanned in row order pass -1: a later row that only TIES the k-th value loses to the rows already
// held (lower index), so later slabs need strictly smaller values and tie groups do not pile up in the lists.
template <int CAP>
__global__ void __launch_bounds__(1024) tighten_kernel(uint64_t* __restrict__ cand, uint32_t* __restrict__ cnt,
                                                            float* __restrict__ thr, const float* __restrict__ ebound,
                                                            uint32_t* __restrict__ flags, int k,
                                                            uint32_t* __restrict__ approx_out, float thr_shift = 0.0f,
                                                            int sample_groups = 0) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);
    __shared__ uint32_t hist[256];
    __shared__ int s_bin, s_need, s_pos, s_low;
    const int q = blockIdx.x;
    const uint32_t c_raw = sample_groups > 0 ? (uint32_t)sample_groups : cnt[q];
    const int c = (int)min(c_raw, (uint32_t)CAP);
    if (c_raw > (uint32_t)CAP && threadIdx.x == 0) flags[q] = 1;   // overflow: the exact scan answers this query
    uint64_t* mine = cand + (size_t)q * CAP;
    uint32_t* aout = approx_out ? approx_out + (size_t)q * k : nullptr;
    const uint32_t ORD_INF = f32_to_ordered(INFINITY);
    if (c <= k) {                                        // fewer than k candidates so far: nothing to drop (uniform per CTA)
        if (aout)
            for (int i = threadIdx.x; i < k; i += blockDim.x) aout[i] = i < c ? (uint32_t)(mine[i] >> 32) : ORD_INF;
        if (sample_groups > 0 && threadIdx.x == 0) cnt[q] = 0;
        return;
    }
    for (int i = threadIdx.x; i < c; i += blockDim.x) keys[i] = mine[i];
    if (threadIdx.x == 0) { s_pos = 0; s_low = 0; }
    __syncthreads();
    const uint64_t kth = block_radix_select(keys, c, k, hist, &s_bin, &s_need);
    const uint32_t kth_v = (uint32_t)(kth >> 32);
    const float bound = ordered_to_f32(kth_v) + 2.0f * ebound[q];      // in approx = -score units
    if (sample_groups > 0) {                             // group-best keys of a sampling slab: only the threshold is kept
        if (aout) {                                      // ... and its k smallest values, for the other shards
            for (int i = threadIdx.x; i < c; i += blockDim.x) {
                const uint32_t v = (uint32_t)(keys[i] >> 32);
                if (v < kth_v) aout[atomicAdd(&s_low, 1)] = v;
            }
            __syncthreads();
            for (int i = s_low + threadIdx.x; i < k; i += blockDim.x) aout[i] = kth_v;
        }
        if (threadIdx.x == 0) { cnt[q] = 0; thr[q] = fmaxf(thr[q], -bound); }
        return;
    }
    const int lane = threadIdx.x & 31;
    for (int i0 = 0; i0 < c; i0 += blockDim.x) {                              // uniform trip count: the ballots need every lane
        const int i = i0 + threadIdx.x;
        const uint64_t key = i < c ? keys[i] : FPV_KEY_MAX;
        const uint32_t v = (uint32_t)(key >> 32);
        const bool keep = i < c && ordered_to_f32(v) <= bound;
        const uint32_t m = __ballot_sync(FPV_FULL_MASK, keep);          // one shared atomic per warp, not per key
        if (m) {
            int pos = 0;
            if (lane == 0) pos = atomicAdd(&s_pos, __popc(m));
            pos = __shfl_sync(FPV_FULL_MASK, pos, 0);
            if (keep) mine[pos + __popc(m & ((1u << lane) - 1u))] = key;
        }
        if (aout && i < c && v < kth_v) aout[atomicAdd(&s_low, 1)] = v;  // strictly below the k-th value: fewer than k of them
    }
    __syncthreads();
    if (aout)
        for (int i = s_low + threadIdx.x; i < k; i += blockDim.x) aout[i] = kth_v;   // the remaining slots tie on the k-th value
    if (threadIdx.x == 0) {
        cnt[q] = (uint32_t)s_pos;
        // epilogue keeps rows with score >= thr  <=>  approx <= bound.  Thresholds only tighten: a bound taken from the
        // whole job (row-sharded search, gemm_global_thr_kernel) can be below what this shard's own candidates give
        thr[q] = fmaxf(thr[q], -(bound + thr_shift));
    }
}


// ---- warp-per-query tighten ------------------------------------------------------------------------------------------
// The block-per-query form above spends its time in __syncthreads (about 40 barriers around 4-8 keys per thread:
// 40-50 us per 4096 queries, three times per search).  Here one WARP owns a query: the 32-bit values live in the warp's
// slice of shared memory, every pass is warp-synchronous, and twelve queries are in flight per SM.
__device__ __forceinline__ uint32_t warp_radix_select(const uint32_t* vals, int c, int kth, uint32_t mn, uint32_t mx,
                                                      uint32_t* hist, int lane) {
    const uint32_t range = mx - mn;
    if (range == 0u) return mn;
    int shift = 31 - __clz(range) - 7;
    if (shift < 0) shift = 0;
    uint32_t base = 0u;
    int need = kth;
    while (true) {
#pragma unroll
        for (int b = 0; b < 8; ++b) hist[lane + 32 * b] = 0u;
        __syncwarp();
        for (int i = lane; i < c; i += 32) {
            const uint32_t v = vals[i] - mn;
            if (v >= base) {
                const uint32_t bin = (v - base) >> shift;
                if (bin < 256u) atomicAdd(&hist[bin], 1u);
            }
        }
        __syncwarp();
        uint32_t h[8], sum = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) { h[b] = hist[lane * 8 + b]; sum += h[b]; }
        uint32_t incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t t = __shfl_up_sync(FPV_FULL_MASK, incl, o); if (lane >= o) incl += t; }
        const uint32_t excl = incl - sum;
        const bool mine = excl < (uint32_t)need && (uint32_t)need <= incl;
        int bin = 0, rest = 0;
        if (mine) {
            uint32_t run = excl;
#pragma unroll
            for (int b = 0; b < 8; ++b) {
                if ((uint32_t)need <= run + h[b]) { bin = lane * 8 + b; rest = need - (int)run; break; }
                run += h[b];
            }
        }
        const uint32_t who = __ballot_sync(FPV_FULL_MASK, mine);
        const int src = who ? __ffs(who) - 1 : 0;
        bin = __shfl_sync(FPV_FULL_MASK, bin, src);
        rest = __shfl_sync(FPV_FULL_MASK, rest, src);
        __syncwarp();
        if (!who) return mx;                              // kth beyond the keys in range (cannot happen for kth <= c)
        base += (uint32_t)bin << shift;
        need = rest;
        if (shift == 0) break;
        shift = shift > 8 ? shift - 8 : 0;
    }
    return mn + base;
}

// grid = ceil(Q / 4), block = 128 (4 warps, one query each); dynamic smem = 4 * (CAP + 256) * 4 bytes.
// sample_groups > 0: the list holds `sample_groups` group-best keys of a sampling slab (not candidates): only the
// threshold is derived from them and the list is emptied.
template <int CAP>
__global__ void __launch_bounds__(128) tighten_warp_kernel(uint64_t* __restrict__ cand, uint32_t* __restrict__ cnt,
                                                           float* __restrict__ thr, const float* __restrict__ ebound,
                                                           uint32_t* __restrict__ flags, int k, uint32_t* __restrict__ approx_out,
                                                           int Q, int sample_groups) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t* vals = reinterpret_cast<uint32_t*>(sm_raw) + (size_t)warp * (CAP + 256);
    uint32_t* hist = vals + CAP;
    const int q = blockIdx.x * 4 + warp;
    if (q >= Q) return;
    const uint32_t c_raw = sample_groups > 0 ? (uint32_t)sample_groups : cnt[q];
    const int c = (int)min(c_raw, (uint32_t)CAP);
    if (c_raw > (uint32_t)CAP && lane == 0) flags[q] = 1;              // overflow: the exact scan answers this query
    uint64_t* mine = cand + (size_t)q * CAP;
    uint32_t* aout = approx_out ? approx_out + (size_t)q * k : nullptr;
    const uint32_t ORD_INF = f32_to_ordered(INFINITY);
    if (c <= k) {                                                        // fewer than k candidates: nothing to drop
        if (aout)
            for (int i = lane; i < k; i += 32) aout[i] = i < c ? (uint32_t)(mine[i] >> 32) : ORD_INF;
        if (sample_groups > 0 && lane == 0) cnt[q] = 0;
        return;
    }
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    // eight independent loads in flight per lane: with one the loop was bound by the L2 latency (78 us per launch)
    for (int i0 = lane; i0 < c; i0 += 256) {
        uint32_t v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const int i = i0 + 32 * u; v[u] = i < c ? (uint32_t)(mine[i] >> 32) : 0xFFFFFFFFu; }
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int i = i0 + 32 * u;
            if (i < c) { vals[i] = v[u]; mn = min(mn, v[u]); mx = max(mx, v[u]); }
        }
    }
    mn = __reduce_min_sync(FPV_FULL_MASK, mn);
    mx = __reduce_max_sync(FPV_FULL_MASK, mx);
    __syncwarp();
    const uint32_t kth_v = warp_radix_select(vals, c, k, mn, mx, hist, lane);
    const float bound = ordered_to_f32(kth_v) + 2.0f * ebound[q];      // in approx = -score units
    if (sample_groups > 0) {
        if (aout) {                                                      // the k smallest group values, for the other shards
            int low = 0;
            for (int i0 = 0; i0 < c; i0 += 32) {
                const int i = i0 + lane;
                const uint32_t v = i < c ? vals[i] : 0xFFFFFFFFu;
                const bool lowv = i < c && v < kth_v;
                const uint32_t ml = __ballot_sync(FPV_FULL_MASK, lowv);
                if (lowv) aout[low + __popc(ml & ((1u << lane) - 1u))] = v;
                low += __popc(ml);
            }
            for (int i = low + lane; i < k; i += 32) aout[i] = kth_v;
        }
        if (lane == 0) { cnt[q] = 0; thr[q] = fmaxf(thr[q], -bound); }
        return;
    }
    // in-place, order-preserving compaction: position written <= position read, and an iteration reads its 32 keys
    // before it writes any
    int kept = 0, low = 0;
    for (int i0 = 0; i0 < c; i0 += 128) {                               // four loads in flight per lane
        uint64_t key4[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { const int i = i0 + 32 * u + lane; key4[u] = i < c ? mine[i] : FPV_KEY_MAX; }
        __syncwarp();                                                    // every read of this round precedes its writes
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + 32 * u + lane;
            const uint64_t key = key4[u];
            const uint32_t v = (uint32_t)(key >> 32);
            const bool keep = i < c && ordered_to_f32(v) <= bound;
            const bool lowv = i < c && v < kth_v;
            const uint32_t m = __ballot_sync(FPV_FULL_MASK, keep);
            const uint32_t ml = __ballot_sync(FPV_FULL_MASK, lowv);
            if (keep) mine[kept + __popc(m & ((1u << lane) - 1u))] = key;
            if (aout && lowv) aout[low + __popc(ml & ((1u << lane) - 1u))] = v;   // strictly below the k-th value: fewer than k
            kept += __popc(m);
            low += __popc(ml);
        }
    }
    if (aout)
        for (int i = low + lane; i < k; i += 32) aout[i] = kth_v;        // the remaining slots tie on the k-th value
    if (lane == 0) {
        cnt[q] = (uint32_t)kept;
        thr[q] = fmaxf(thr[q], -bound);                  // epilogue keeps rows with score >= thr  <=>  approx <= bound;
    }                                                    // thresholds only tighten (see tighten_kernel)
}

}  // namespace fpv
