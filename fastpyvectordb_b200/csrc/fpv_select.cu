// Error plumbing, the partial-list finalize kernel shared by every scan, and the cross-shard k-way merge
// (replaces _merge_top_k, parallel_search.py:137-156).
#include <atomic>
#include <cstdarg>
#include <cstdio>
#include <string>

#include "fpv_common.cuh"

namespace fpv {

static thread_local std::string g_err;
static std::atomic<long long> g_launches{0};
void note_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

void set_error(const char* fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof(buf), fmt, ap);
    va_end(ap);
    g_err = buf;
}
int cuda_fail(cudaError_t e, const char* what) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), what);
    return FPV_ERR_CUDA;
}
int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 148;   // B200; keeps *_workspace() usable on a box without a GPU
    }
    return n;
}
int max_smem_optin() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess ||
        cudaDeviceGetAttribute(&n, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return 227 * 1024;
    }
    return n;
}

// ------------------------------------------------------------------------------------------------
// finalize: one CTA per query reduces n_parts sorted partial lists (K keys each) to the final k rows.
// ------------------------------------------------------------------------------------------------
// grid = (groups, Q).  CTA (g, q) folds the lists  first = g*span + i*stride  (i < n, inside [0, n_parts)) of query q.
// emit == 0: the merged list is written back over list slot g*span (level 1 of a two-level reduction, in place);
// emit == 1: the first k keys become the (distance, id) output rows.
__global__ void __launch_bounds__(1024) finalize_kernel(uint64_t* __restrict__ partials, int n_parts, int span, int stride,
                                                       int K, int k, int64_t id_base, int emit,
                                                       float* __restrict__ out_dist, int64_t* __restrict__ out_idx,
                                                       int32_t* __restrict__ out_count,
                                                       const uint32_t* __restrict__ only_flagged) {
    // Every partial list is already sorted, so no selector is needed: warp w folds its share of the lists into its own
    // K-list with the bitonic merge step (min against the reversed list, log2(K) compare-exchange stages), then the
    // W lists are merged pairwise in a tree.  (The first version streamed all n_parts*K keys through WarpSelect:
    // 160 us for 592 lists.)
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* lists = reinterpret_cast<uint64_t*>(smem_raw);           // [W][K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const int64_t q = blockIdx.x;                          // grid = (Q, groups): Q can exceed the 65535 limit of grid.y
    if (only_flagged && only_flagged[q] == 0) return;      // this query already has its (certified) answer
    uint64_t* qbase = partials + (size_t)q * n_parts * K;
    const int first = blockIdx.y * span;
    const int n = min(span, n_parts - first) <= 0 ? 0 : (min(span, n_parts - first) + stride - 1) / stride;
    const uint64_t* src = qbase + (size_t)first * K;
    uint64_t* mine = lists + (size_t)warp * K;
    for (int i = lane; i < K; i += 32) mine[i] = warp < n ? src[(size_t)warp * stride * K + i] : FPV_KEY_MAX;
    __syncwarp();
    for (int part = warp + W; part < n; part += W) merge_sorted_into(mine, src + (size_t)part * stride * K, K, lane);
    __syncthreads();
    for (int s = W >> 1; s >= 1; s >>= 1) {
        if (warp < s) merge_sorted_into(mine, lists + (size_t)(warp + s) * K, K, lane);
        __syncthreads();
    }
    if (!emit) {
        if (n > 0) for (int i = threadIdx.x; i < K; i += blockDim.x) qbase[(size_t)first * K + i] = lists[i];
        return;
    }
    if (warp == 0) {
        uint64_t* dst = lists;
        int cnt = 0;
        for (int i = lane; i < k; i += 32) {
            uint64_t key = dst[i];
            bool ok = key != FPV_KEY_MAX;
            out_dist[q * k + i] = ok ? ordered_to_f32((uint32_t)(key >> 32)) : INFINITY;
            out_idx[q * k + i] = ok ? id_base + (int64_t)(uint32_t)key : -1;
            cnt += ok;
        }
        if (out_count) {
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) cnt += __shfl_xor_sync(FPV_FULL_MASK, cnt, o);
            if (lane == 0) out_count[q] = cnt;
        }
    }
}

int launch_finalize(const uint64_t* partials, int64_t Q, int n_parts, int K, int k, int64_t id_base,
                    float* out_dist, int64_t* out_idx, int32_t* out_count, cudaStream_t st,
                    const uint32_t* only_flagged) {
    if (Q <= 0) return FPV_OK;
    uint64_t* lists = const_cast<uint64_t*>(partials);     // level 1 folds in place (the partial lists are scratch)
    auto warps_for = [&](int64_t ctas, int n_lists) {
        int W = 1;     // few CTAs x many lists: wide CTAs; many CTAs: the grid supplies the parallelism
        while (W < 32 && W < n_lists && ctas * W < 8192 && (size_t)(2 * W) * K * sizeof(uint64_t) <= 96 * 1024) W *= 2;
        return W;
    };
    auto launch = [&](int groups, int span, int stride, int n_lists, int emit) -> int {
        const int W = warps_for((int64_t)groups * Q, n_lists);
        const size_t smem = (size_t)W * K * sizeof(uint64_t);
        if (smem > 48 * 1024)
            FPV_CUDA(cudaFuncSetAttribute(finalize_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        finalize_kernel<<<dim3((unsigned)Q, groups), W * 32, smem, st>>>(lists, n_parts, span, stride, K, k, id_base, emit,
                                                                       out_dist, out_idx, out_count, only_flagged);
        FPV_LAUNCH_CHECK();
        return FPV_OK;
    };
    // Single-query scans leave ~600 lists: one CTA folding them all is a 70 us serial tail, so fold in two levels
    // (16 CTAs x ~37 lists, then one CTA x 16 lists).
    if (n_parts > 64 && Q <= 32) {
        const int groups = 16, span = (n_parts + groups - 1) / groups;
        int rc = launch(groups, span, 1, span, 0);
        if (rc != FPV_OK) return rc;
        return launch(1, n_parts, span, groups, 1);
    }
    return launch(1, n_parts, 1, n_parts, 1);
}

// ------------------------------------------------------------------------------------------------
// cross-shard merge.  Entries are (float dist, int64 global id); ordering (dist, id) with a 96-bit compare.
// One CTA per query; bitonic sort of the padded candidate set in shared memory.
// ------------------------------------------------------------------------------------------------
struct MergeEnt { uint32_t d; uint32_t pad; int64_t id; };
__device__ __forceinline__ bool ent_less(const MergeEnt& a, const MergeEnt& b) {
    return a.d < b.d || (a.d == b.d && a.id < b.id);
}

// 8-byte wire format of one candidate for the cross-GPU exchange: ordered(distance) << 32 | row LOCAL to the shard
// (the receiver adds the shard's first row), FPV_KEY_MAX = empty slot.  Half the bytes of (int64 id, fp32 distance).
__global__ void pack_topk_kernel(const float* __restrict__ dist, const int64_t* __restrict__ idx, int64_t Q, int k_in,
                                 int k_pad, int64_t id_base, uint64_t* __restrict__ out) {
    const int64_t total = Q * k_pad;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
        const int64_t q = i / k_pad;
        const int j = (int)(i - q * k_pad);
        uint64_t key = FPV_KEY_MAX;
        if (j < k_in) {
            const int64_t id = idx[q * k_in + j];
            if (id >= 0) key = ((uint64_t)f32_to_ordered(dist[q * k_in + j]) << 32) | (uint64_t)(uint32_t)(id - id_base);
        }
        out[i] = key;
    }
}

__global__ void __launch_bounds__(256) merge_kernel(const float* __restrict__ dist, const int64_t* __restrict__ idx,
                                                    const uint64_t* __restrict__ packed, const int64_t* __restrict__ bases,
                                                    int shards, int64_t Q, int k_in, int k_out, int P,
                                                    float* __restrict__ out_dist, int64_t* __restrict__ out_idx,
                                                    int32_t* __restrict__ out_count) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    MergeEnt* e = reinterpret_cast<MergeEnt*>(smem_raw);
    __shared__ int total_cnt;
    const int64_t q = blockIdx.x;
    const int total = shards * k_in;
    for (int i = threadIdx.x; i < P; i += blockDim.x) {
        MergeEnt v;
        v.d = 0xFFFFFFFFu; v.pad = 0; v.id = INT64_MAX;
        if (i < total) {
            int s = i / k_in, j = i - s * k_in;
            size_t off = ((size_t)s * Q + q) * k_in + j;
            if (packed) {
                const uint64_t key = packed[off];
                if (key != FPV_KEY_MAX) { v.d = (uint32_t)(key >> 32); v.id = bases[s] + (int64_t)(uint32_t)key; }
            } else {
                int64_t id = idx[off];
                if (id >= 0) { v.d = f32_to_ordered(dist[off]); v.id = id; }
            }
        }
        e[i] = v;
    }
    __syncthreads();
    for (int size = 2; size <= P; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int t = threadIdx.x; t < (P >> 1); t += blockDim.x) {
                int lo = 2 * t - (t & (stride - 1));
                int hi = lo + stride;
                bool up = (lo & size) == 0;
                MergeEnt x = e[lo], y = e[hi];
                if (ent_less(y, x) == up) { e[lo] = y; e[hi] = x; }
            }
            __syncthreads();
        }
    }
    int cnt = 0;
    for (int i = threadIdx.x; i < k_out; i += blockDim.x) {
        bool ok = i < P && e[i].id != INT64_MAX;
        out_dist[q * k_out + i] = ok ? ordered_to_f32(e[i].d) : INFINITY;
        out_idx[q * k_out + i] = ok ? e[i].id : -1;
        cnt += ok;
    }
    if (out_count) {
        if (threadIdx.x == 0) total_cnt = 0;
        __syncthreads();
        if (cnt) atomicAdd(&total_cnt, cnt);
        __syncthreads();
        if (threadIdx.x == 0) out_count[q] = total_cnt;
    }
}

// Merge of SORTED packed lists (what every rank holds after the all-gather): no sorting network.  Every entry finds
// its final position directly: rank(x) = number of entries, over all lists, that precede x in (distance, global id)
// order = its own position + one binary search per other list.  One CTA per query, one thread per entry; the old
// form (bitonic sort of next_pow2(shards * k) 16-byte records: 55 stages at 8 shards x k = 100) was the largest
// fixed cost of the row-sharded search after the exact re-rank.
__global__ void __launch_bounds__(256) merge_rank_kernel(const uint64_t* __restrict__ packed, const int64_t* __restrict__ bases,
                                                         int shards, int64_t Q, int k_in, int k_out,
                                                         float* __restrict__ out_dist, int64_t* __restrict__ out_idx,
                                                         int32_t* __restrict__ out_count,
                                                         const uint32_t* __restrict__ wait_flags, uint32_t epoch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    peer_wait(wait_flags, shards, epoch);                               // peer-memory exchange: the lists arrive by NVLink stores
    uint32_t* dv = reinterpret_cast<uint32_t*>(smem_raw);             // [shards][k_in] ordered distances (0xFFFFFFFF = empty)
    uint32_t* rv = dv + (size_t)shards * k_in;                        // [shards][k_in] shard-local rows
    __shared__ int s_valid;
    const int64_t q = blockIdx.x;
    const int total = shards * k_in;
    if (threadIdx.x == 0) s_valid = 0;
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int s = i / k_in, j = i - s * k_in;
        const uint64_t key = __ldcg(reinterpret_cast<const unsigned long long*>(packed) + ((size_t)s * Q + q) * k_in + j);   // not through L1
        // an empty slot sorts after every real entry (a real distance never has the all-ones pattern: NaN keys included,
        // they carry a real row and are distinguished by rv below)
        dv[i] = key == FPV_KEY_MAX ? 0xFFFFFFFFu : (uint32_t)(key >> 32);
        rv[i] = (uint32_t)key;
    }
    __syncthreads();
    int valid = 0;
    // entry j of a list already has j entries of its own list before it, so only j < k_out can make the answer
    for (int i = threadIdx.x; i < total; i += blockDim.x) {
        const int s = i / k_in, j = i - s * k_in;
        const uint32_t d = dv[i];
        const bool empty = d == 0xFFFFFFFFu && rv[i] == 0xFFFFFFFFu;
        if (empty) continue;
        ++valid;
        if (j >= k_out) continue;
        const int64_t id = bases[s] + (int64_t)rv[i];
        int rank = j;
        for (int t = 0; t < shards && rank < k_out; ++t) {
            if (t == s) continue;
            const uint32_t* lst = dv + (size_t)t * k_in;
            // entries of list t with a strictly smaller distance: binary search on 32-bit values; the search never needs
            // to look past position k_out - rank (beyond it this entry is out of the answer anyway)
            int lo = 0, hi = min(k_in, k_out - rank);
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (lst[mid] < d) lo = mid + 1; else hi = mid;
            }
            // ties on the distance: ordered by global id (a short run, usually empty)
            const uint32_t* rl = rv + (size_t)t * k_in;
            const int64_t bt = bases[t];
            while (lo < k_in && lst[lo] == d && !(d == 0xFFFFFFFFu && rl[lo] == 0xFFFFFFFFu) && bt + (int64_t)rl[lo] < id) ++lo;
            rank += lo;
        }
        if (rank < k_out) {
            out_dist[q * k_out + rank] = ordered_to_f32(d);
            out_idx[q * k_out + rank] = id;
        }
    }
    if (valid) atomicAdd(&s_valid, valid);
    __syncthreads();
    const int filled = min(s_valid, k_out);
    for (int i = filled + threadIdx.x; i < k_out; i += blockDim.x) { out_dist[q * k_out + i] = INFINITY; out_idx[q * k_out + i] = -1; }
    if (out_count && threadIdx.x == 0) out_count[q] = filled;
}

}  // namespace fpv

using namespace fpv;

extern "C" int fpv_abi_version(void) { return FPV_ABI_VERSION; }
extern "C" const char* fpv_last_error(void) { return fpv::g_err.c_str(); }
extern "C" long long fpv_launch_count(void) { return fpv::g_launches.load(std::memory_order_relaxed); }

extern "C" int fpv_merge_topk(const float* dist, const int64_t* idx, int shards, int64_t q, int k_in, int k_out,
                              float* out_dist, int64_t* out_idx, int32_t* out_count, void* stream) {
    FPV_REQUIRE(shards >= 1 && q >= 0 && k_in >= 1 && k_out >= 1, "merge: bad shape shards=%d q=%lld k_in=%d k_out=%d",
                shards, (long long)q, k_in, k_out);
    FPV_REQUIRE(dist && idx && out_dist && out_idx, "merge: null pointer");
    if (q == 0) return FPV_OK;
    int P = next_pow2(shards * k_in);
    if (P < 2) P = 2;
    size_t smem = (size_t)P * sizeof(MergeEnt);
    FPV_REQUIRE(smem <= (size_t)max_smem_optin(), "merge: shards*k_in=%d too large for one CTA", shards * k_in);
    if (smem > 48 * 1024)
        FPV_CUDA(cudaFuncSetAttribute(merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_kernel<<<(unsigned)q, 256, smem, (cudaStream_t)stream>>>(dist, idx, nullptr, nullptr, shards, q, k_in, k_out, P,
                                                                    out_dist, out_idx, out_count);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

extern "C" int fpv_pack_topk(const float* dist, const int64_t* idx, int64_t q, int k_in, int k_pad, int64_t id_base,
                             uint64_t* out_packed, void* stream) {
    FPV_REQUIRE(q >= 0 && k_in >= 0 && k_pad >= 1 && k_pad >= k_in, "pack_topk: bad shape q=%lld k_in=%d k_pad=%d",
                (long long)q, k_in, k_pad);
    if (q == 0) return FPV_OK;
    FPV_REQUIRE(out_packed && (k_in == 0 || (dist && idx)), "pack_topk: null pointer");
    int64_t blocks = (q * k_pad + 255) / 256;
    int64_t cap = (int64_t)sm_count() * 8;
    if (blocks > cap) blocks = cap;
    pack_topk_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(dist, idx, q, k_in, k_pad, id_base, out_packed);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

static int merge_packed_impl(const uint64_t* packed, const int64_t* shard_bases, int shards, int64_t q, int k_in, int k_out,
                             float* out_dist, int64_t* out_idx, int32_t* out_count, const uint32_t* wait_flags, uint32_t epoch,
                             void* stream) {
    FPV_REQUIRE(shards >= 1 && q >= 0 && k_in >= 1 && k_out >= 1, "merge_packed: bad shape shards=%d q=%lld k_in=%d k_out=%d",
                shards, (long long)q, k_in, k_out);
    FPV_REQUIRE(packed && shard_bases && out_dist && out_idx, "merge_packed: null pointer");
    if (q == 0) return FPV_OK;
    // every list must be sorted ascending by (distance, row) with the empty slots (FPV_KEY_MAX) last: what
    // fpv_pack_topk makes of any top-k output of this library
    const size_t smem = (size_t)shards * k_in * sizeof(uint64_t);
    FPV_REQUIRE(smem <= (size_t)max_smem_optin(), "merge_packed: shards*k_in=%d too large for one CTA", shards * k_in);
    if (smem > 48 * 1024)
        FPV_CUDA(cudaFuncSetAttribute(merge_rank_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    merge_rank_kernel<<<(unsigned)q, 256, smem, (cudaStream_t)stream>>>(packed, shard_bases, shards, q, k_in, k_out, out_dist,
                                                                         out_idx, out_count, wait_flags, epoch);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

extern "C" int fpv_merge_packed(const uint64_t* packed, const int64_t* shard_bases, int shards, int64_t q, int k_in, int k_out,
                                float* out_dist, int64_t* out_idx, int32_t* out_count, void* stream) {
    return merge_packed_impl(packed, shard_bases, shards, q, k_in, k_out, out_dist, out_idx, out_count, nullptr, 0u, stream);
}

// The same merge as the consumer of a peer-memory exchange (fpv_peer_put): `packed` is this rank's gather area, and the
// kernel first waits until the `shards` flag words at wait_flags have reached `epoch`.
extern "C" int fpv_merge_packed_peer(const uint64_t* packed, const int64_t* shard_bases, int shards, int64_t q, int k_in, int k_out,
                                     const uint32_t* wait_flags, uint32_t epoch, float* out_dist, int64_t* out_idx,
                                     int32_t* out_count, void* stream) {
    FPV_REQUIRE(wait_flags && shards <= 256, "merge_packed_peer: null flags");
    return merge_packed_impl(packed, shard_bases, shards, q, k_in, k_out, out_dist, out_idx, out_count, wait_flags, epoch, stream);
}
