// Large-batch exact search on the 5th-gen tensor cores: TMA -> shared memory -> tcgen05.mma (TF32 straight from the
// fp32 database, or BF16 from a shadow copy) -> TMEM accumulators -> fused threshold-filter epilogue, followed by a
// certified exact fp32 re-rank.  The Q x N distance matrix never exists anywhere (parallel_search.py:279-290 writes
// it to memory and then loops over its rows in Python, :296-309).
//
// Replaces ParallelSearchEngine.search_batch_parallel (parallel_search.py:246-311); small batches use it too when the
// bf16 shadow exists (half the bytes of the fp32 scan), and the filter_mask of search_parallel is applied in the epilogue.
//
// Exactness.  Tensor-core products are approximate, so the first pass only *filters*: for every query it collects ALL
// rows whose approximate value is below a per-query threshold that is tightened between row slabs.  With E a rigorous
// bound on |approx - exact| (TF32: eps * |q| * max|v|, 2^-11 on the pre-rounded query + 2^-10 on the truncated database
// element; BF16: the Cauchy-Schwarz bound on the MEASURED rounding residuals of the query and of the shadow copy),
// every row of the true top-k has approx <= a_k + 2E (a_k = k-th best approximate value, which only decreases as rows
// are seen), so gemm_tighten2_kernel sets the threshold to exactly a_k + 2E and keeps the candidates below it; every
// row that is NOT in the candidate buffer is strictly above the threshold of its time, hence above the final
// a_k + 2E.  gemm_finish2_kernel re-ranks the rows below the final limit in exact fp32 (same formula, row norms and
// summation order as the scan kernel) and sorts by (distance, index).  A query whose buffer overflowed (or whose
// certificate fails) is flagged and recomputed by the exact fp32 scan kernel on the device (fpv_scan_f32.cu) — no
// host round trip, never an approximate answer.
//
// Kernel shape: persistent, one CTA per SM, 384 threads = warp0 TMA producer, warp1 MMA issuer, warp2 TMEM allocator,
// warps 4-11 epilogue (TMEM lane quarter = warp % 4, column half = (warp-4)/4, thread <-> one query row x 128
// accumulator columns).  Tile 128 queries x 256 database rows per CTA, K in 128-byte (SWIZZLE_128B) blocks, 192 KB
// smem ring, two 256-column TMEM accumulators so the epilogue of tile i overlaps the MMAs of tile i+1.  With more than
// one query block two CTAs form a cluster and share a 256 x 256 tile (tcgen05 cta_group::2, see GemmCfg).
// Measured on B200 (profiles/r01_gemm_final_ncu_summary.txt): tensor pipe 99.7 % active in the large slabs.
//
// Lessons recorded from the ncu captures and phase timers of this round (profiles/, tools/trace_gemm.py):
//  * the CTAs that share a database tile run in lockstep; without rotating the K order per CTA they all missed on
//    the same L2 lines at once and each went to HBM (56 GB read for a 2.6 GB slab);
//  * one global atomic per hit serialised the epilogue (13 us/tile): hits are now counted in a mask pass and
//    reserved with ONE atomic per thread per tile, after the accumulator has been handed back to the MMA warp;
//  * compare+select+or per element made the epilogue issue-latency bound (two warps per scheduler): the hit mask is
//    now built from the sign bits of (score - threshold) with funnel shifts, 2-3 instructions per element;
//  * hit capture through local memory or single-column TMEM loads cost 250+ cycles per hit (the L1 beside a 192 KB
//    ring cannot hold the stacks): hits are walked from a shared-memory staging column;
//  * with `if (lane == 0)` around the role loops every TMA / MMA operand went through an ELECT + R2UR loop and the MMA
//    warp, which shares its scheduler with two epilogue warps, could not issue a K block per 512 cycles: the role
//    loops are warp-uniform now (83 % -> 99.7 % tensor pipe).
#include <cuda.h>
#include <cuda_bf16.h>

#include <mutex>

#include "fpv_common.cuh"
#include "fpv_select.cuh"
#include "fpv_tc.cuh"

namespace fpv {

constexpr int BM = 128;                 // queries per tile (UMMA M)
constexpr int BN = 256;                 // database rows per tile (UMMA N)
constexpr int KROW = 128;               // bytes of K per smem row (one swizzle span)
constexpr int A_BYTES = BM * KROW;      // 16 KB
constexpr int TMEM_COLS = 512;
constexpr int GEMM_CAP = 4096;          // candidate slots per query
constexpr int GEMM_MAX_K = 256;
constexpr int GEMM_THREADS = 384;        // warps 0-3: TMA / MMA / TMEM alloc / spare, warps 4-11: epilogue
constexpr int EPI_THREADS = 256;
constexpr int HIT_BUF = 8;                 // hits a thread can capture per tile on the fast path (steady state: ~0.1)
constexpr int STG_WORDS = 16;              // accumulators a thread stages at a time (half of a 32-column chunk)
// NCTA = 1: one CTA per 128 x 256 tile.  NCTA = 2: a CTA pair (cluster of 2, tcgen05 cta_group::2) per 256 x 256 tile;
// each CTA stages its own 128 query rows and HALF of the database tile, so the L2 -> SM operand traffic per MMA
// drops from 48 KB to 32 KB per K block (the phase timers showed the single-CTA MMA thread waiting ~2.5K of every
// ~7.3K cycles for operands) and the ring is 6 stages deep instead of 4.
template <int NCTA>
struct GemmCfg {
    static constexpr int STAGES = NCTA == 2 ? 6 : 4;
    static constexpr int B_ROWS = BN / NCTA;                  // database rows this CTA stages per tile
    static constexpr int B_BYTES = B_ROWS * KROW;             // 32 KB / 16 KB
    static constexpr int RING = STAGES * (A_BYTES + B_BYTES); // 192 KB either way
    // operand ring | aux[2][BN] | barriers + TMEM slot (256 B) | staging [STG_WORDS][256] u32 | captured keys [HIT_BUF][256] u64
    static constexpr size_t SMEM = (size_t)RING + 2 * BN * 4 + 256 + (size_t)STG_WORDS * EPI_THREADS * 4 +
                                   (size_t)HIT_BUF * EPI_THREADS * 8;
    static_assert(16 * STAGES + 32 + 4 <= 256, "barrier block");
    static_assert(SMEM <= 232448, "shared memory budget");    // the dynamic window is 1024-byte aligned (checked in the kernel)
};

// c_format F32 (1<<4), a/b format (F16=0, BF16=1, TF32=2) at bits 7 / 10, K-major both, N>>3 at 17, M>>4 at 24
// (M = 128 per CTA; 256 for the pair instruction)
__host__ __device__ constexpr uint32_t make_idesc(int fmt, int m) {
    return (1u << 4) | ((uint32_t)fmt << 7) | ((uint32_t)fmt << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

struct GemmParams {
    const float* aux;       // per database row: 1/(|v|+eps) (cosine), |v|^2 (l2), unused (ip)
    const uint32_t* mask;   // optional row filter: bit (row & 31) of word (row >> 5) set = row may be returned
    const float* thr;       // [Qp] score threshold per query (-inf at the start)
    uint32_t* cnt;          // [Qp] candidates appended so far (may exceed CAP: overflow)
    uint64_t* cand;         // [Qp][CAP] keys: ordered(-score) << 32 | row
    int64_t N;              // rows in the database
    int Q;                  // valid queries
    int m_blocks;           // ceil(Q / 128)
    int tile0, ntiles;      // database tile range of this slab (tiles of 256 rows)
    int tile_stride;        // 1; the sampling slab takes every tile_stride-th tile so that it spans the whole database
    int nkb;                // K blocks of 128 bytes
    int metric;
    int sample;             // 1: sampling slab -- the epilogue writes group-best keys to fixed slots, no candidates
    int slab;               // slab ordinal (only used by the FPV_GEMM_TRACE experiment build)
    int debug;              // FPV_GEMM_TRACE experiments: bit 0 = skip the epilogue work, bit 1 = always load the same tiles
};

template <int METRIC>
__device__ __forceinline__ float score_of(float acc, float aux) {
    if (METRIC == FPV_METRIC_COSINE) return acc * aux;              // cos = q^.v / (|v| + eps)
    if (METRIC == FPV_METRIC_L2) return fmaf(2.0f, acc, -aux);      // -( |v|^2 - 2 q.v ) = q.q - d^2
    return acc;                                                      // q.v
}

// Two passes over the accumulator tile (TMEM reads are cheap, global atomics are not):
//   pass 1 builds one 32-bit pass mask per 32-column chunk, no memory traffic;
//   then ONE atomicAdd per thread reserves its slots in the query's candidate buffer;
//   pass 2 re-reads only the chunks in which some lane of the warp has a hit and stores the keys.
constexpr int EPI_COLS = BN / 2;            // each epilogue warp owns 32 query rows x 128 accumulator columns
constexpr int EPI_CHUNKS = EPI_COLS / 32;

// Hit mask of one 32-column chunk: bit j set iff score(j) >= thr.  The test is done on the SIGN of t = score - thr
// and the sign bits are collected with funnel shifts (one SHF per element) on four independent chains: 2 (cosine,
// ip) or 3 (l2) instructions per element instead of ~9 for compare + select + or (ncu: the epilogue warps were
// issue-latency bound, two warps per scheduler on one dependent chain).  t = +0 when score == thr (hit), -inf for
// thr = +inf (padding rows), +inf for thr = -inf (first slab: everything hits).
template <int METRIC, bool FULL>
__device__ __forceinline__ uint32_t chunk_mask(const uint32_t (&r)[32], const float* aux32, float thr, int col0, int ncols) {
    const float4* a4 = reinterpret_cast<const float4*>(aux32);
    const float nthr = -thr;
    uint32_t ch[4] = {0u, 0u, 0u, 0u};
#pragma unroll
    for (int j4 = 0; j4 < 8; ++j4) {
        const float4 a = METRIC == FPV_METRIC_IP ? make_float4(0.f, 0.f, 0.f, 0.f) : a4[j4];
        const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int j = j4 * 4 + u;
            const float acc = __uint_as_float(r[j]);
            float t;
            if (METRIC == FPV_METRIC_COSINE) t = fmaf(acc, av[u], nthr);
            else if (METRIC == FPV_METRIC_L2) t = fmaf(2.0f, acc, -av[u]) + nthr;
            else t = acc + nthr;
            ch[j >> 3] = __funnelshift_l(__float_as_uint(t), ch[j >> 3], 1);     // (chain << 1) | sign(t)
        }
    }
    // chain c holds elements 8c..8c+7, element 8c+i at bit 7-i  ->  reversed mask R: element j at bit 31-j
    const uint32_t rev = (ch[0] << 24) | (ch[1] << 16) | (ch[2] << 8) | ch[3];
    uint32_t m = __brev(~rev);
    if (!FULL) {
        const int left = ncols - col0;                                            // valid columns in this chunk
        m &= left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
    }
    return m;
}

// Walk the hits of one half (16 columns) of a chunk whose accumulators are in registers: a lane with a hit stages
// its 16 values in shared memory (column layout [j][thread]: conflict free for any j) and visits the set bits of
// its mask with a dynamically indexed load -- ~10 instructions per hit, no warp-collective work.  Shared, not
// local, memory: with 200 KB of operand ring the L1 left over is too small for 256 threads' stacks, and the
// phase timers showed 250+ cycles per hit for local-memory and single-column-TMEM forms alike.
template <int METRIC, int HALF, class Emit>
__device__ __forceinline__ void walk_half(const uint32_t (&r)[32], uint32_t* stg, const float* aux32, uint32_t m,
                                          uint32_t colbase, Emit&& emit) {
    uint32_t mh = (m >> (16 * HALF)) & 0xFFFFu;
    if (mh == 0) return;
#pragma unroll
    for (int j = 0; j < STG_WORDS; ++j) stg[j * EPI_THREADS] = r[HALF * 16 + j];
    while (mh) {
        const int j = __ffs(mh) - 1;
        mh &= mh - 1;
        const int jj = HALF * 16 + j;
        const float s = score_of<METRIC>(__uint_as_float(stg[j * EPI_THREADS]), METRIC == FPV_METRIC_IP ? 0.f : aux32[jj]);
        emit(((uint64_t)f32_to_ordered(-s) << 32) | (uint64_t)(colbase + jj));
    }
}

// Dense form of the above for pass 2 (first slabs: most of a half's 32 x 16 elements are hits).  The per-lane walk
// stores 8 bytes per lane to 32 different candidate rows, so every 32-byte sector is written four times and the LSU
// (about 2 cycles per sector) bounds the tile: 58K cycles for an all-hit tile.  Here all lanes stage their 16 values,
// then the warp writes the rows two at a time, 16 lanes per row: one row's hits are one contiguous run of keys.
// `pos` is the lane's next slot in its own row, `q` its query row; both are read from the owning lane by shuffle.
template <int METRIC, int HALF>
__device__ __forceinline__ void dense_half(const GemmParams& p, const uint32_t (&r)[32], uint32_t* stg, const float* aux32,
                                           uint32_t m, uint32_t colbase, uint32_t& pos, int q, int lane) {
    const uint32_t mh = (m >> (16 * HALF)) & 0xFFFFu;
#pragma unroll
    for (int j = 0; j < STG_WORDS; ++j) stg[j * EPI_THREADS] = r[HALF * 16 + j];
    __syncwarp();
    const int sub = lane >> 4, j = lane & 15, jj = HALF * 16 + j;
    const float aux = METRIC == FPV_METRIC_IP ? 0.f : aux32[jj];
    const uint32_t* col = stg - lane + j * EPI_THREADS;               // staged column j of this warp's 32 rows
    const uint32_t below = (1u << j) - 1u;
#pragma unroll 4
    for (int rr = 0; rr < 32; rr += 2) {
        const int row = rr + sub;
        const uint32_t mr = __shfl_sync(FPV_FULL_MASK, mh, row);
        const uint32_t pr = __shfl_sync(FPV_FULL_MASK, pos, row);
        if ((mr >> j) & 1u) {
            const float s = score_of<METRIC>(__uint_as_float(col[row]), aux);
            const uint32_t slot = pr + (uint32_t)__popc(mr & below);
            if (slot < (uint32_t)GEMM_CAP)
                p.cand[(size_t)(q - lane + row) * GEMM_CAP + slot] = ((uint64_t)f32_to_ordered(-s) << 32) | (uint64_t)(colbase + jj);
        }
    }
    pos += (uint32_t)__popc(mh);
    __syncwarp();                                                     // the staging columns are reused by the next half
}

// pass 1: keep the keys of this chunk's hits (shared memory, [slot][thread]); a lane that would overflow HIT_BUF
// only counts -- the tile then takes the two-pass path.
template <int METRIC>
__device__ __forceinline__ void capture_chunk(const uint32_t (&r)[32], uint32_t* stg, uint64_t* hks, const float* aux32,
                                              uint32_t m, uint32_t colbase, uint32_t& nh) {
    if (m == 0) return;
    const uint32_t cnt = (uint32_t)__popc(m);
    if (nh + cnt <= (uint32_t)HIT_BUF) {
        uint32_t w = nh;
        auto put = [&](uint64_t key) { hks[w * EPI_THREADS] = key; ++w; };
        walk_half<METRIC, 0>(r, stg, aux32, m, colbase, put);
        walk_half<METRIC, 1>(r, stg, aux32, m, colbase, put);
    }
    nh += cnt;
}

// `bar` is a shared::cluster address: the accumulator-empty barrier lives in the leader CTA of a pair
__device__ __forceinline__ void release_accumulator(uint32_t bar, int lane) {
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive_cluster(bar);
}

// The 32 rows of a chunk start at a multiple of 32, so their filter bits are exactly one word (same for all lanes).
__device__ __forceinline__ uint32_t row_filter_word(const GemmParams& p, uint32_t first_row) {
    return (int64_t)first_row < p.N ? __ldg(p.mask + (first_row >> 5)) : 0u;
}

// `taddr` / `auxs` / `col0` already point at this warp's 128-column half of the tile.  Releases the accumulator
// (arrive on `bar_release`) as early as possible: the MMA of the tile after next is waiting for it.
template <int METRIC, bool FULL>
__device__ __forceinline__ void epilogue_tile(const GemmParams& p, uint32_t taddr, const float* auxs, int col0, int ncols,
                                              float thr, int q, int64_t n0, uint32_t bar_release, int lane,
                                              uint32_t* stg, uint64_t* hks) {
    static_assert(EPI_CHUNKS == 4, "mask registers below assume four 32-column chunks per warp");
    uint32_t mk0 = 0, mk1 = 0, mk2 = 0, mk3 = 0;
    uint32_t nh = 0;
    const uint32_t gcol = (uint32_t)(n0 + col0);
    {   // pass 1, software pipelined: the TMEM load of the next chunk is in flight while this one is scored.
        // The chunk loops are deliberately NOT unrolled: the fully unrolled kernel was 400 KB of SASS and the
        // epilogue warps stalled on instruction fetch (ncu: stall_no_inst).
        uint32_t ra[32], rb[32];
        TMEM_LD32(ra, taddr);
#pragma unroll 1
        for (int cp = 0; cp < EPI_CHUNKS / 2; ++cp) {
            const int c = 2 * cp;
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            TMEM_LD32(rb, taddr + (c + 1) * 32);
            uint32_t m0 = chunk_mask<METRIC, FULL>(ra, auxs + c * 32, thr, col0 + c * 32, ncols);
            if (p.mask) m0 &= row_filter_word(p, gcol + c * 32);
            capture_chunk<METRIC>(ra, stg, hks, auxs + c * 32, m0, gcol + c * 32, nh);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (cp + 1 < EPI_CHUNKS / 2) TMEM_LD32(ra, taddr + (c + 2) * 32);
            uint32_t m1 = chunk_mask<METRIC, FULL>(rb, auxs + (c + 1) * 32, thr, col0 + (c + 1) * 32, ncols);
            if (p.mask) m1 &= row_filter_word(p, gcol + (c + 1) * 32);
            capture_chunk<METRIC>(rb, stg, hks, auxs + (c + 1) * 32, m1, gcol + (c + 1) * 32, nh);
            if (cp == 0) { mk0 = m0; mk1 = m1; } else { mk2 = m0; mk3 = m1; }
        }
    }
    const uint32_t total = nh;
    uint64_t* dst = p.cand + (size_t)q * GEMM_CAP;
    if (__all_sync(FPV_FULL_MASK, total <= (uint32_t)HIT_BUF)) {
        // fast path (every slab but the first two): nothing more is needed from TMEM -> hand it back to the MMA
        // warp BEFORE the global atomic and the stores (ncu: the MMA thread was spinning on this barrier)
        release_accumulator(bar_release, lane);
        if (total) {
            uint32_t pos = atomicAdd(p.cnt + q, total);
            for (uint32_t i = 0; i < total; ++i)
                if (pos + i < (uint32_t)GEMM_CAP) dst[pos + i] = hks[i * EPI_THREADS];
        }
        return;
    }
    uint32_t pos = 0;
    if (total) pos = atomicAdd(p.cnt + q, total);
#pragma unroll 1
    for (int c = 0; c < EPI_CHUNKS; ++c) {
        const uint32_t m = c == 0 ? mk0 : (c == 1 ? mk1 : (c == 2 ? mk2 : mk3));
        if (__any_sync(FPV_FULL_MASK, m != 0)) {
            uint32_t r[32];
            TMEM_LD32(r, taddr + c * 32);
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            auto put = [&](uint64_t key) { if (pos < (uint32_t)GEMM_CAP) dst[pos] = key; ++pos; };
            // more than a quarter of a half's 512 elements hit: cooperative row-contiguous stores
            if (__reduce_add_sync(FPV_FULL_MASK, (unsigned)__popc(m & 0xFFFFu)) >= 128u)
                dense_half<METRIC, 0>(p, r, stg, auxs + c * 32, m, gcol + c * 32, pos, q, lane);
            else walk_half<METRIC, 0>(r, stg, auxs + c * 32, m, gcol + c * 32, put);
            if (__reduce_add_sync(FPV_FULL_MASK, (unsigned)__popc(m >> 16)) >= 128u)
                dense_half<METRIC, 1>(p, r, stg, auxs + c * 32, m, gcol + c * 32, pos, q, lane);
            else walk_half<METRIC, 1>(r, stg, auxs + c * 32, m, gcol + c * 32, put);
        }
    }
    release_accumulator(bar_release, lane);
}

// Sampling slab (the first launch of a search).  No threshold exists yet, and writing every score of a dense first
// slab to the candidate lists (then radix-selecting 2048 keys per query) cost 48 + 47 us per 4096 queries.  The only
// thing the first slab has to deliver is an upper bound on the final k-th value, and the k-th best of ANY k distinct
// rows is one: each thread reduces its 128 accumulator columns to 16 group maxima (groups of 8 columns, two
// instructions per element, no memory traffic until the end) and stores them as keys in fixed slots -- S/8 keys per
// query for S sample rows; the k-th best of those group maxima is a valid bound (each comes from a different row) and
// with S/8 >= 2k groups it sits within ~1.15x of the exact k-th quantile of the sample.  The sample rows are scanned
// again by the first filtering slab (S/N of the work).
constexpr int SAMPLE_GW = 8;                                  // columns per group
constexpr int SAMPLE_KEYS = EPI_COLS / SAMPLE_GW;             // 16 keys per thread and tile
template <int METRIC, bool FULL>
__device__ __forceinline__ void sample_tile(const GemmParams& p, uint32_t taddr, const float* auxs, int col0, int ncols, int q,
                                            int64_t n0, uint32_t bar_release, int lane, int slot0) {
    const uint32_t gcol = (uint32_t)(n0 + col0);
    float gm[SAMPLE_KEYS];
#pragma unroll
    for (int g = 0; g < SAMPLE_KEYS; ++g) gm[g] = -INFINITY;
#pragma unroll 1
    for (int c = 0; c < EPI_CHUNKS; ++c) {
        uint32_t r[32];
        TMEM_LD32(r, taddr + c * 32);
        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
        uint32_t ok = 0xFFFFFFFFu;
        if (p.mask) ok = row_filter_word(p, gcol + c * 32);
        if (!FULL) {
            const int left = ncols - (col0 + c * 32);
            ok &= left >= 32 ? 0xFFFFFFFFu : (left <= 0 ? 0u : ((1u << left) - 1u));
        }
        const float4* a4 = reinterpret_cast<const float4*>(auxs + c * 32);
#pragma unroll
        for (int j4 = 0; j4 < 8; ++j4) {
            const float4 a = METRIC == FPV_METRIC_IP ? make_float4(0.f, 0.f, 0.f, 0.f) : a4[j4];
            const float av[4] = {a.x, a.y, a.z, a.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = j4 * 4 + u;
                float sc = score_of<METRIC>(__uint_as_float(r[j]), av[u]);
                if (p.mask || !FULL) sc = ((ok >> j) & 1u) ? sc : -INFINITY;
                r[j] = __float_as_uint(sc);
            }
        }
#pragma unroll
        for (int g = 0; g < 32 / SAMPLE_GW; ++g) {
            float m = __uint_as_float(r[g * SAMPLE_GW]);
#pragma unroll
            for (int e = 1; e < SAMPLE_GW; ++e) m = fmaxf(m, __uint_as_float(r[g * SAMPLE_GW + e]));
            // c is a runtime loop variable (the loop is deliberately not unrolled: instruction-fetch stalls), so select
            // the destination register with predicated moves instead of a dynamic index
#pragma unroll
            for (int cc = 0; cc < EPI_CHUNKS; ++cc)
                if (cc == c) gm[cc * (32 / SAMPLE_GW) + g] = m;
        }
    }
    release_accumulator(bar_release, lane);
    if (q < p.Q) {
        uint4* dst = reinterpret_cast<uint4*>(p.cand + (size_t)q * GEMM_CAP + slot0);
#pragma unroll
        for (int g = 0; g < SAMPLE_KEYS; g += 2) {
            const uint64_t k0 = ((uint64_t)f32_to_ordered(-gm[g]) << 32) | (uint32_t)(slot0 + g);
            const uint64_t k1 = ((uint64_t)f32_to_ordered(-gm[g + 1]) << 32) | (uint32_t)(slot0 + g + 1);
            dst[g >> 1] = make_uint4((uint32_t)k0, (uint32_t)(k0 >> 32), (uint32_t)k1, (uint32_t)(k1 >> 32));
        }
    }
}

#ifdef FPV_GEMM_TRACE
// experiment-only per-CTA phase timers (cycles): [0] epilogue: aux staging + named barrier, [1] epilogue: wait for the
// accumulator, [2] epilogue: tile processing, [3] MMA thread: wait tempty, [4] MMA thread: wait full, [5] tiles
__device__ unsigned long long g_trace[8 * 148 * 8];   // [slab][cta][counter]
#endif

template <int KIND, int METRIC, int NCTA>   // KIND 0: TF32 operands (fp32 in memory), 1: BF16 operands
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_filter_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, GemmParams p) {
    using Cfg = GemmCfg<NCTA>;
    constexpr int STAGES = Cfg::STAGES, B_BYTES = Cfg::B_BYTES;
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) __trap();                 // SWIZZLE_128B operand tiles need the 1024-byte aligned window
    const uint32_t sA = base, sB = base + STAGES * A_BYTES;
    constexpr uint32_t off_aux = Cfg::RING;
    uint8_t* gen = smem_raw;
    float* auxs = reinterpret_cast<float*>(gen + off_aux);                  // [2][BN]
    const uint32_t bars = base + off_aux + 2 * BN * 4;
    const uint32_t bar_full = bars, bar_empty = bars + 8 * STAGES, bar_tfull = bars + 16 * STAGES, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(gen + off_aux + 2 * BN * 4 + 16 * STAGES + 32);

    // warp / rank / tmem_base go through a lane-0 shuffle so that the compiler knows they are warp-uniform: the role
    // loops below then live on the uniform datapath.  (With `if (lane == 0)` around them every TMA / MMA operand was
    // converted by an ELECT + R2UR loop, ~95 instructions per K block, and the phase timers showed the MMA warp --
    // which shares its scheduler with two epilogue warps -- unable to issue one K block per 512 cycles.)
    const int warp = __shfl_sync(FPV_FULL_MASK, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const uint32_t rank = NCTA == 2 ? __shfl_sync(FPV_FULL_MASK, cluster_ctarank(), 0) : 0u;   // 0 = leader: owns full / tempty, issues the MMAs
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(bar_full + 8 * s, 1); mbar_init(bar_empty + 8 * s, 1); }
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, 8 * NCTA); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        if (NCTA == 2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(TMEM_COLS) : "memory");
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
        }
    }
    tc_fence_before();
    __syncthreads();
    if (NCTA == 2) cluster_sync_all();          // the peer's barriers are initialised before anything signals them
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(FPV_FULL_MASK, *tmem_slot, 0);
    // work items: (query block [pair], database tile); a pair walks the same sequence in both CTAs
    const int mgroups = p.m_blocks / NCTA;
    const int total = mgroups * p.ntiles;
    const int first = blockIdx.x / NCTA, step = gridDim.x / NCTA;
    constexpr int KELEMS = KIND == 0 ? 32 : 64;     // elements per 128-byte K block

    if (warp == 0) {                                // ---------------- TMA producer (both CTAs of a pair)
        // the whole warp walks the loop (uniform control flow); one elected lane issues the asynchronous operations
        int stage = 0; uint32_t phase = 0;
        const uint32_t full_leader = NCTA == 2 ? mapa(bar_full, 0) : bar_full;
        for (int t = first; t < total; t += step) {
            const int mg = t % mgroups, nt = p.tile0 + (t / mgroups) * p.tile_stride;
            int mb = mg * NCTA + (int)rank;
#ifdef FPV_GEMM_TRACE
            int nt_load = (p.debug & 2) ? p.tile0 : nt;
            if (p.debug & 2) mb = (int)rank;
#else
            const int nt_load = nt;
#endif
            // The CTAs that share a database tile (one per query block) run in lockstep; without a stagger they
            // all miss on the same L2 lines at the same instant and every one of them goes to HBM (measured: 22x
            // refetch).  Rotating the K order by query block makes them touch different lines at any instant, so
            // one CTA's fill serves the others.  The accumulation order does not matter to the filter pass.
            // (t % nkb also spreads the CTAs that share a QUERY block over the K blocks, which matters when there
            // is a single query block and all 148 CTAs would otherwise hammer the same L2 lines of A.)
            // Both CTAs of a pair must use the same order: the rotation is a function of the pair's work item.
            const int rot = mgroups >= 8 ? (int)(((int64_t)mg * p.nkb) / mgroups) : t % p.nkb;
            for (int kb0 = 0; kb0 < p.nkb; ++kb0) {
                int kb = kb0 + rot; if (kb >= p.nkb) kb -= p.nkb;
                mbar_wait(bar_empty + 8 * stage, phase ^ 1);
                if (elect_one()) {
                    if (NCTA == 2) {
                        if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * (A_BYTES + B_BYTES));   // both CTAs' bytes
                        tma_load_2d_pair(sA + stage * A_BYTES, &tmA, full_leader + 8 * stage, kb * KELEMS, mb * BM);
                        tma_load_2d_pair(sB + stage * B_BYTES, &tmB, full_leader + 8 * stage, kb * KELEMS,
                                         nt_load * BN + (int)rank * Cfg::B_ROWS);
                    } else {
                        mbar_expect_tx(bar_full + 8 * stage, A_BYTES + B_BYTES);
                        tma_load_2d(sA + stage * A_BYTES, &tmA, bar_full + 8 * stage, kb * KELEMS, mb * BM);
                        tma_load_2d(sB + stage * B_BYTES, &tmB, bar_full + 8 * stage, kb * KELEMS, nt_load * BN);
                    }
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (rank == 0) {                            // ---------------- MMA issuer (the leader CTA of a pair)
            constexpr uint32_t idesc = make_idesc(KIND == 0 ? 2 : 1, BM * NCTA);
            int stage = 0; uint32_t phase = 0; int as = 0; uint32_t aphase = 0;
#ifdef FPV_GEMM_TRACE
            unsigned long long w_te = 0, w_fu = 0;
#endif
            for (int t = first; t < total; t += step) {
#ifdef FPV_GEMM_TRACE
                long long c0 = clock64();
#endif
                mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
#ifdef FPV_GEMM_TRACE
                w_te += clock64() - c0;
#endif
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + as * BN;
                for (int kb = 0; kb < p.nkb; ++kb) {
#ifdef FPV_GEMM_TRACE
                    long long c1 = clock64();
#endif
                    mbar_wait(bar_full + 8 * stage, phase);
#ifdef FPV_GEMM_TRACE
                    w_fu += clock64() - c1;
#endif
                    tc_fence_after();
                    if (elect_one()) {
                        const uint64_t ad = make_smem_desc(sA + stage * A_BYTES), bd = make_smem_desc(sB + stage * B_BYTES);
#pragma unroll
                        for (int k = 0; k < KROW / 32; ++k) {     // 32 bytes of K per instruction
                            if (NCTA == 2) tc_mma_pair(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0, KIND == 0);
                            else tc_mma(d_tmem, ad + 2 * k, bd + 2 * k, idesc, (kb | k) != 0, KIND == 0);
                        }
                        // frees the smem slot (in both CTAs of a pair) when these MMAs retire
                        if (NCTA == 2) tc_commit_pair(bar_empty + 8 * stage); else tc_commit(bar_empty + 8 * stage);
                    }
                    __syncwarp();
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
                // accumulator complete (each CTA of a pair holds its own 128 query rows of it)
                if (elect_one()) { if (NCTA == 2) tc_commit_pair(bar_tfull + 8 * as); else tc_commit(bar_tfull + 8 * as); }
                __syncwarp();
                as ^= 1; if (as == 0) aphase ^= 1;
            }
#ifdef FPV_GEMM_TRACE
            if (lane == 0) { g_trace[(p.slab & 7) * 148 * 8 + blockIdx.x * 8 + 3] = w_te; g_trace[(p.slab & 7) * 148 * 8 + blockIdx.x * 8 + 4] = w_fu; }
#endif
        }
    } else if (warp >= 4) {                         // ---------------- epilogue: TMEM -> registers -> filter
        // 8 warps: TMEM lane quarter = warp % 4 (hardware rule), column half = (warp - 4) / 4
        const int quarter = warp & 3, half = (warp - 4) >> 2, et = threadIdx.x - 128;       // et in [0, 256)
        const int col0 = half * EPI_COLS;
        uint32_t* stg = reinterpret_cast<uint32_t*>(gen + off_aux + 2 * BN * 4 + 256) + et;                  // [STG_WORDS][256]
        uint64_t* hks = reinterpret_cast<uint64_t*>(gen + off_aux + 2 * BN * 4 + 256 + STG_WORDS * EPI_THREADS * 4) + et;  // [HIT_BUF][256]
        int as = 0; uint32_t aphase = 0;
        const uint32_t tempty_leader = NCTA == 2 ? mapa(bar_tempty, 0) : bar_tempty;
#ifdef FPV_GEMM_TRACE
        unsigned long long e_aux = 0, e_wait = 0, e_work = 0, e_tiles = 0;
#endif
        for (int t = first; t < total; t += step) {
#ifdef FPV_GEMM_TRACE
            long long c0 = clock64();
#endif
            const int mb = (t % mgroups) * NCTA + (int)rank, nt = p.tile0 + (t / mgroups) * p.tile_stride;
            const int64_t n0 = (int64_t)nt * BN;
            const int q = mb * BM + quarter * 32 + lane;
            const float thr = q < p.Q ? __ldg(p.thr + q) : INFINITY;
            float* a_s = auxs + as * BN;
            if (p.aux) a_s[et] = __ldg(p.aux + min(n0 + et, p.N - 1));
            asm volatile("bar.sync 1, 256;" ::: "memory");
#ifdef FPV_GEMM_TRACE
            long long c1 = clock64();
#endif
            mbar_wait(bar_tfull + 8 * as, aphase);
#ifdef FPV_GEMM_TRACE
            long long c2 = clock64();
#endif
            tc_fence_after();
            const int ncols = (int)min((int64_t)BN, p.N - n0);
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * BN + col0;
            const float* a_h = a_s + col0;
#ifdef FPV_GEMM_TRACE
            if (p.debug & 1) release_accumulator(tempty_leader + 8 * as, lane);
            else if ((p.debug & 4) || ((p.debug & 8) && quarter == 1) || ((p.debug & 16) && quarter == 0) ||
                     ((p.debug & 32) && quarter >= 2)) {          // TMEM loads only
                uint32_t r[32]; uint32_t acc = 0;
                for (int c = 0; c < EPI_CHUNKS; ++c) {
                    TMEM_LD32(r, taddr + c * 32);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    acc ^= r[0] ^ r[31];
                }
                if (acc == 0x12345678u) p.cnt[q] = 0;
                release_accumulator(tempty_leader + 8 * as, lane);
            } else
#endif
            if (p.sample) {
                const int slot0 = ((t / mgroups) * 2 + half) * SAMPLE_KEYS;
                if (ncols == BN) sample_tile<METRIC, true>(p, taddr, a_h, col0, ncols, q, n0, tempty_leader + 8 * as, lane, slot0);
                else sample_tile<METRIC, false>(p, taddr, a_h, col0, ncols, q, n0, tempty_leader + 8 * as, lane, slot0);
            } else if (ncols == BN) epilogue_tile<METRIC, true>(p, taddr, a_h, col0, ncols, thr, q, n0, tempty_leader + 8 * as, lane, stg, hks);
            else epilogue_tile<METRIC, false>(p, taddr, a_h, col0, ncols, thr, q, n0, tempty_leader + 8 * as, lane, stg, hks);
            as ^= 1; if (as == 0) aphase ^= 1;
#ifdef FPV_GEMM_TRACE
            e_aux += c1 - c0; e_wait += c2 - c1; e_work += clock64() - c2; ++e_tiles;
#endif
        }
#ifdef FPV_GEMM_TRACE
        if (threadIdx.x == 128) {
            g_trace[(p.slab & 7) * 148 * 8 + blockIdx.x * 8 + 0] = e_aux; g_trace[(p.slab & 7) * 148 * 8 + blockIdx.x * 8 + 1] = e_wait; g_trace[(p.slab & 7) * 148 * 8 + blockIdx.x * 8 + 2] = e_work;
            g_trace[(p.slab & 7) * 148 * 8 + blockIdx.x * 8 + 5] = e_tiles;
        }
#endif
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (NCTA == 2) cluster_sync_all();          // nobody leaves (or frees TMEM) while the peer can still touch this CTA
    if (warp == 2) {
        tc_fence_after();
        if (NCTA == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(TMEM_COLS) : "memory");
    }
}

// ----------------------------------------------------------------------------------------------- helper kernels
// qprep: exact fp32 queries (cosine: normalised like parallel_search.py:121,271); qa: the tensor-core operand copy,
// rounded to TF32 (or BF16) to nearest and zero padded to Qp rows; per-query |q|^2, error bound E, thr = -inf, cnt = 0.
__global__ void gemm_prep_kernel(const float* __restrict__ q, int Q, int Qp, int D, int Dp, int metric, int kind,
                                 float eps, float vmax, float db_err_abs, float db_err_rel, float* __restrict__ qprep,
                                 void* __restrict__ qa, float* __restrict__ qsq, float* __restrict__ ebound,
                                 float* __restrict__ thr, uint32_t* __restrict__ cnt, uint32_t* __restrict__ flags) {
    const int lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (row >= Qp) return;
    float s = 0.f;
    if (row < Q)
        for (int j = lane; j < D; j += 32) { float x = q[(size_t)row * D + j]; s = fmaf(x, x, s); }
    s = warp_sum(s);
    const float inv = (metric == FPV_METRIC_COSINE) ? 1.0f / (sqrtf(s) + 1e-10f) : 1.0f;
    float lo2 = 0.f;                                 // |x - rounded(x)|^2: the query's own rounding error, measured
    for (int j = lane; j < Dp; j += 32) {
        float x = (row < Q && j < D) ? q[(size_t)row * D + j] * inv : 0.f;
        if (j < D) qprep[(size_t)row * D + j] = x;
        float xr;
        if (kind == 0) {
            uint32_t u = __float_as_uint(x);
            u = (u + 0x00000FFFu + ((u >> 13) & 1u)) & 0xFFFFE000u;          // round to nearest even at 10 mantissa bits
            xr = __uint_as_float(u);
            reinterpret_cast<float*>(qa)[(size_t)row * Dp + j] = xr;
        } else {
            const __nv_bfloat16 b = __float2bfloat16_rn(x);
            xr = __bfloat162float(b);
            reinterpret_cast<__nv_bfloat16*>(qa)[(size_t)row * Dp + j] = b;
        }
        lo2 = fmaf(x - xr, x - xr, lo2);
    }
    lo2 = warp_sum(lo2);
    if (lane == 0) {
        const float qn = metric == FPV_METRIC_COSINE ? 1.0f : sqrtf(s);     // norm of the vector that is multiplied
        qsq[row] = s;
        float e;
        if (db_err_abs > 0.f) {
            // Measured bound (Cauchy-Schwarz, still rigorous): |q~.v~ - q.v| <= |q - q~| |v~| + |q| |v - v~| + accumulation,
            // with |q - q~| computed above, |v - v~| <= db_err_abs (max over rows, from the index build), |v~| <= 1.002 |v|.
            // About 1.7x tighter than the worst-case per-element bound, so fewer rows reach the exact re-rank.
            const float lo = sqrtf(lo2) * 1.0001f, slack = 3e-4f;
            if (metric == FPV_METRIC_COSINE) e = lo * 1.002f * (1.0f + db_err_rel) + db_err_rel * 1.0001f + slack;  // score / |v|
            else e = lo * vmax * 1.002f + qn * db_err_abs * 1.0001f + slack * qn * vmax;
            if (metric == FPV_METRIC_L2) e *= 2.0f;                          // score = 2 q.v - |v|^2
        } else {
            if (metric == FPV_METRIC_COSINE) e = eps * 1.0001f;              // |q^| = 1, score scaled by 1/|v|
            else if (metric == FPV_METRIC_L2) e = 2.0f * eps * qn * vmax;
            else e = eps * qn * vmax;
        }
        ebound[row] = e;
        thr[row] = -INFINITY;
        cnt[row] = 0;
        flags[row] = 0;
    }
}

__device__ __forceinline__ float finish_distance_g(int metric, float dot, float vsq, float qsq) {
    if (metric == FPV_METRIC_COSINE) return 1.0f - dot / (sqrtf(vsq) + 1e-10f);
    if (metric == FPV_METRIC_L2) return sqrtf(fmaxf(qsq + vsq - 2.0f * dot, 0.0f));
    return -dot;
}

constexpr int FIN_RMAX = 2048;      // rows re-ranked exactly per query at most; beyond that the query falls back

__global__ void __launch_bounds__(1024) gemm_finish2_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt,
                                                           const float* __restrict__ thr, const float* __restrict__ ebound,
                                                           uint32_t* __restrict__ flags, const float* __restrict__ qprep,
                                                           const float* __restrict__ qsq, const float* __restrict__ db,
                                                           const float* __restrict__ row_sq, int D, int64_t ld, int metric,
                                                           int k, int64_t id_base, float* __restrict__ out_dist,
                                                           int64_t* __restrict__ out_idx, int32_t* __restrict__ out_count,
                                                           const uint32_t* __restrict__ approx_all, int shards, int Q,
                                                           const uint32_t* __restrict__ wait_flags, uint32_t epoch) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);                                   // [GEMM_CAP]
    uint64_t* sel = keys + GEMM_CAP;                                                        // [FIN_RMAX]
    float* qs = reinterpret_cast<float*>(sm_raw + (size_t)(GEMM_CAP + FIN_RMAX) * 8);       // [D4 * 4]
    __shared__ uint32_t hist[256];
    __shared__ int s_bin, s_need, s_R, s_flag;
    const int q = blockIdx.x, lane = threadIdx.x & 31, warp = threadIdx.x >> 5, W = blockDim.x >> 5;
    const uint32_t c_raw = cnt[q];
    const int c = (int)min(c_raw, (uint32_t)GEMM_CAP);
    const uint64_t* mine = cand + (size_t)q * GEMM_CAP;
    // Row-sharded search: a_k is the k-th best approximate value over ALL shards, selected from the k best values of
    // every shard (approx_all [shards][Q][k], gathered by the caller).  Each shard then re-ranks only its own rows
    // below the global limit, so the gather cost of the whole job is that of one GPU, divided by the shard count.
    float a_k = INFINITY;
    peer_wait(wait_flags, shards, epoch);                // peer-memory exchange: the shards' values arrive by NVLink stores
    if (approx_all) {
        const int tot = shards * k;                      // <= GEMM_CAP (checked on the host)
        for (int i = threadIdx.x; i < tot; i += blockDim.x) {
            const int sh = i / k, j = i - sh * k;
            keys[i] = ((uint64_t)__ldcg(approx_all + ((size_t)sh * Q + q) * k + j) << 32) | (uint32_t)i;     // not through L1
        }
        __syncthreads();
        a_k = ordered_to_f32((uint32_t)(block_radix_select(keys, tot, k, hist, &s_bin, &s_need) >> 32));
        __syncthreads();
    }
    for (int i = threadIdx.x; i < c; i += blockDim.x) keys[i] = mine[i];
    const int D4 = (D + 3) >> 2;
    for (int j = threadIdx.x; j < D4 * 4; j += 256) qs[j] = j < D ? qprep[(size_t)q * D + j] : 0.f;
    if (threadIdx.x == 0) { s_R = 0; s_flag = (c_raw > (uint32_t)GEMM_CAP) || flags[q] != 0; }
    __syncthreads();
    const float t = thr[q];
    const float E = ebound[q];
    if (!approx_all && c >= k) a_k = ordered_to_f32((uint32_t)(block_radix_select(keys, c, k, hist, &s_bin, &s_need) >> 32));
    const float limit = a_k + 2.0f * E;       // every true top-k row has approx value <= limit ...
    // ... and every row outside the buffer failed `score >= thr`, i.e. is STRICTLY above -thr, so equality is fine
    const bool certified = (t == -INFINITY) || (limit <= -t);
    // The gather of the candidate rows is what this kernel costs (R x D x 4 bytes per query), so re-rank in two stages:
    // first only the rows within a_k + 1.25E; if their exact k-th distance d_k is <= a_k + 0.25E then every other row
    // (approx > a_k + 1.25E, hence exact > a_k + 0.25E >= d_k) is provably out and the second stage is skipped.
    // (the second stage needs the exact k-th distance of the WHOLE job, so a shard of a row-sharded search takes all
    // of its rows below the limit in one stage: they are 1/shards of the window)
    const float limit1 = approx_all ? limit : a_k + 1.25f * E;
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const uint64_t key = keys[i];
        if (ordered_to_f32((uint32_t)(key >> 32)) <= limit1) {
            const int pos = atomicAdd(&s_R, 1);
            if (pos < FIN_RMAX) sel[pos] = key;
        }
    }
    __syncthreads();
    const int R1 = s_R;
    __syncthreads();
    for (int i = threadIdx.x; i < c; i += blockDim.x) {
        const uint64_t key = keys[i];
        const float a = ordered_to_f32((uint32_t)(key >> 32));
        if (a > limit1 && a <= limit) {
            const int pos = atomicAdd(&s_R, 1);
            if (pos < FIN_RMAX) sel[pos] = key;
        }
    }
    __syncthreads();
    const int R = s_R;
    if (threadIdx.x == 0) {
        if (!certified || R > FIN_RMAX) s_flag = 1;
        flags[q] = s_flag;
    }
    __syncthreads();
    if (s_flag) return;                                  // the exact scan fallback writes this query
    const float my_qsq = qsq[q];
    const bool vec = rows_vectorizable(db, D, ld);
    auto rerank_range = [&](int lo, int hi) {            // exact fp32 distance, one warp per PAIR of candidate rows
        for (int i = lo + 2 * warp; i < hi; i += 2 * W) {
            const bool two = i + 1 < hi;
            const uint32_t row0 = (uint32_t)sel[i], row1 = two ? (uint32_t)sel[i + 1] : row0;
            float dot0, dot1;
            canonical_dot2(db + (size_t)row0 * ld, db + (size_t)row1 * ld, reinterpret_cast<const float4*>(qs), D, vec, lane,
                           dot0, dot1);
            if (lane == 0) {
                keys[i] = make_key(finish_distance_g(metric, dot0, __ldg(row_sq + row0), my_qsq), row0);
                if (two) keys[i + 1] = make_key(finish_distance_g(metric, dot1, __ldg(row_sq + row1), my_qsq), row1);
            }
        }
    };
    rerank_range(0, R1);
    int P2 = 2; while (P2 < R1) P2 <<= 1;
    for (int i = R1 + threadIdx.x; i < P2; i += blockDim.x) keys[i] = FPV_KEY_MAX;
    __syncthreads();
    block_bitonic_sort(keys, P2);
    // exact distances live in the distance domain, a_k in the approx-score domain: compare through the same map
    bool need2 = R > R1;
    if (need2 && R1 >= k) {
        // d_k as an approx-domain value: cosine a = d - 1, l2 a = d^2 - |q|^2, ip a = d
        const float dk = ordered_to_f32((uint32_t)(keys[k - 1] >> 32));
        const float dk_a = metric == FPV_METRIC_COSINE ? dk - 1.0f : (metric == FPV_METRIC_L2 ? dk * dk - my_qsq : dk);
        need2 = !(dk_a <= a_k + 0.25f * E - 1e-6f * fmaxf(1.0f, fabsf(a_k)));
    }
    __syncthreads();                                     // everyone has read keys[k-1] before stage 2 overwrites the padding
    if (need2) {                                         // uniform per CTA: keys[k-1], a_k, E are block-wide values
        rerank_range(R1, R);                             // stage-1 results stay in keys[0, R1); sel[] is untouched
        P2 = 2; while (P2 < R) P2 <<= 1;
        for (int i = R + threadIdx.x; i < P2; i += blockDim.x) keys[i] = FPV_KEY_MAX;
        __syncthreads();
        block_bitonic_sort(keys, P2);
    }
    const int R_out = need2 ? R : R1;
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = i < R_out;
        const uint64_t key = ok ? keys[i] : FPV_KEY_MAX;
        out_dist[(size_t)q * k + i] = ok ? ordered_to_f32((uint32_t)(key >> 32)) : INFINITY;
        out_idx[(size_t)q * k + i] = ok ? id_base + (int64_t)(uint32_t)key : -1;
    }
    if (out_count && threadIdx.x == 0) out_count[q] = min(R_out, k);
}

// fp32 -> bf16 shadow copy of the database (index build)
__global__ void to_bf16_kernel(const float* __restrict__ src, __nv_bfloat16* __restrict__ dst, int64_t n) {
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
        dst[i] = __float2bfloat16_rn(src[i]);
}

// Row-sharded search, between the sampling slab and the filtering slabs: sample_all [shards][Q][k] holds every shard's
// k best group values (each the best approximate value of a different group of rows).  Their k-th smallest g_k is
// reached by at least k rows of the job, so the job's k-th approximate value a_k <= g_k and every row of the true
// top-k has approx <= g_k + 2E: the first threshold of EVERY shard (E is the same everywhere: the bounds are maxima
// over the shards).  One CTA per query; waits for the peer stores when the values arrive over NVLink.
__global__ void __launch_bounds__(256) gemm_global_thr_kernel(const uint32_t* __restrict__ sample_all, int shards, int Q, int k,
                                                              const float* __restrict__ ebound, float* __restrict__ thr,
                                                              const uint32_t* __restrict__ wait_flags, uint32_t epoch) {
    // a WARP per query (eight per CTA): shards * k values in the warp's slice of shared memory, warp-synchronous radix
    // select (a CTA per query spent its 30 us per 4096 queries in __syncthreads)
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int tot = shards * k;
    uint32_t* vals = reinterpret_cast<uint32_t*>(sm_raw) + (size_t)warp * (tot + 256);
    uint32_t* hist = vals + tot;
    peer_wait(wait_flags, shards, epoch);
    const int q = blockIdx.x * 8 + warp;
    if (q >= Q) return;
    uint32_t mn = 0xFFFFFFFFu, mx = 0u;
    for (int i = lane; i < tot; i += 32) {
        const int sh = i / k, j = i - sh * k;
        const uint32_t v = __ldcg(sample_all + ((size_t)sh * Q + q) * k + j);                 // not through L1
        vals[i] = v;
        mn = min(mn, v); mx = max(mx, v);
    }
    mn = __reduce_min_sync(FPV_FULL_MASK, mn);
    mx = __reduce_max_sync(FPV_FULL_MASK, mx);
    __syncwarp();
    const float g_k = ordered_to_f32(warp_radix_select(vals, tot, k, mn, mx, hist, lane));
    if (lane == 0) thr[q] = fmaxf(thr[q], -(g_k + 2.0f * ebound[q]));                       // +inf (fewer than k groups): no change
}

// ----------------------------------------------------------------------------------------------- host orchestration
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D row-major [rows][cols] tensor, box = 128 bytes of K x box_rows rows, SWIZZLE_128B, zero fill out of bounds
static int make_map(CUtensorMap* m, const void* ptr, int kind, int64_t rows, int64_t cols, int64_t ld_elems, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) { set_error("gemm: cuTensorMapEncodeTiled entry point not available"); return FPV_ERR_CUDA; }
    const int esz = kind == 0 ? 4 : 2;
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld_elems * esz};
    cuuint32_t box[2] = {(cuuint32_t)(KROW / esz), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, kind == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr),
                    dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("gemm: cuTensorMapEncodeTiled failed with %d", (int)r); return FPV_ERR_CUDA; }
    return FPV_OK;
}

// FPV_GEMM_PAIR=0 keeps the one-CTA kernel for every shape (A/B measurements)
static bool pair_mode_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FPV_GEMM_PAIR"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

// rows of the first slab (every row of it is a candidate): FPV_GEMM_SLAB0 overrides for experiments
// FPV_GEMM_SAMPLE=0 keeps the dense first slab (A/B measurements)
static bool sample_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FPV_GEMM_SAMPLE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

// FPV_GEMM_SAMPLE_STRIDE=0 samples the first tiles instead of every (tiles / ts)-th one (A/B measurements)
static bool sample_stride_enabled() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("FPV_GEMM_SAMPLE_STRIDE"); v = (e && e[0] == '0') ? 0 : 1; }
    return v != 0;
}

static int64_t first_slab_rows(int k) {
    static int64_t env = -1;
    if (env < 0) { const char* e = getenv("FPV_GEMM_SLAB0"); env = e ? atoll(e) : 0; }
    int64_t rows = env > 0 ? env : 2048;
    const int64_t need = ((int64_t)2 * k + BN - 1) / BN * BN;          // at least 2k rows, whole tiles
    if (rows < need) rows = need;
    return (rows + BN - 1) / BN * BN;
}

struct GemmPlan {
    int Qp, Dp, keep, K_sel, esz;
    size_t off_qprep, off_qa, off_qsq, off_eb, off_thr, off_cnt, off_flags, off_cand, off_scan, scan_bytes, total;
};

// candidates kept per query between slabs: the gap between the k-th and the keep-th approximate value must exceed
// twice the error bound for the certificate to hold, so the coarser BF16 pass keeps twice as many as TF32.
static int keep_for(int k, int kind) { int v = next_pow2((kind == 0 ? 2 : 4) * k + 1); return v < 32 ? 32 : v; }

}  // namespace fpv

// exact scan restricted to flagged queries (fpv_scan_f32.cu)
namespace fpv {
size_t scan_f32_flagged_workspace(int64_t Q, int64_t N, int D, int k);
int scan_f32_flagged(const float* queries, int64_t Q, const float* db, int64_t N, int D, int64_t ld, int metric, int k,
                     const float* row_sq, int64_t id_base, const uint32_t* flags, const uint32_t* mask_words, float* out_dist,
                     int64_t* out_idx, int32_t* out_count, void* ws, size_t ws_bytes, cudaStream_t st);

static GemmPlan plan_gemm(int64_t Q, int64_t N, int D, int k, int kind) {
    GemmPlan pl{};
    pl.esz = kind == 0 ? 4 : 2;
    pl.Qp = Q <= BM ? BM : (int)((Q + 2 * BM - 1) / (2 * BM) * (2 * BM));   // more than one query block: whole CTA pairs
    const int kel = KROW / pl.esz;
    pl.Dp = (D + kel - 1) / kel * kel;                    // operand copy of the queries is padded to whole K blocks
    pl.keep = keep_for(k, kind);
    size_t o = 0;
    pl.off_qprep = o; o += align_up((size_t)pl.Qp * D * 4, 1024);
    pl.off_qa = o;    o += align_up((size_t)pl.Qp * pl.Dp * pl.esz, 1024);
    pl.off_qsq = o;   o += align_up((size_t)pl.Qp * 4, 256);
    pl.off_eb = o;    o += align_up((size_t)pl.Qp * 4, 256);
    pl.off_thr = o;   o += align_up((size_t)pl.Qp * 4, 256);
    pl.off_cnt = o;   o += align_up((size_t)pl.Qp * 4, 256);
    pl.off_flags = o; o += align_up((size_t)pl.Qp * 4, 256);
    pl.off_cand = o;  o += (size_t)pl.Qp * GEMM_CAP * 8;
    pl.off_scan = o;
    pl.scan_bytes = scan_f32_flagged_workspace(Q, N, D, k);
    pl.total = o + pl.scan_bytes;
    return pl;
}
}  // namespace fpv

using namespace fpv;

#ifdef FPV_GEMM_TRACE
extern "C" __attribute__((visibility("default"))) int fpv_debug_trace(unsigned long long* host_out) {
    return (int)cudaMemcpyFromSymbol(host_out, fpv::g_trace, sizeof(unsigned long long) * 8 * 148 * 8);
}
#endif

// ---- measurement hooks: CUDA events around the filter launches of the most recent call
namespace {
constexpr int PROF_MAX = 32;
bool g_prof_on = false;
int g_prof_n = 0;
cudaEvent_t g_prof_ev[2 * PROF_MAX];
bool g_prof_made = false;
}  // namespace

extern "C" int fpv_gemm_profile(int enable) {
    if (enable && !g_prof_made) {
        for (int i = 0; i < 2 * PROF_MAX; ++i) FPV_CUDA(cudaEventCreate(&g_prof_ev[i]));
        g_prof_made = true;
    }
    g_prof_on = enable != 0;
    g_prof_n = 0;
    return FPV_OK;
}

extern "C" int fpv_gemm_profile_read(float* filter_ms, int* filter_launches) {
    FPV_REQUIRE(filter_ms && filter_launches, "gemm_profile_read: null pointer");
    float total = 0.f;
    for (int i = 0; i < g_prof_n; ++i) {
        FPV_CUDA(cudaEventSynchronize(g_prof_ev[2 * i + 1]));
        float ms = 0.f;
        FPV_CUDA(cudaEventElapsedTime(&ms, g_prof_ev[2 * i], g_prof_ev[2 * i + 1]));
        total += ms;
    }
    *filter_ms = total;
    *filter_launches = g_prof_n;
    return FPV_OK;
}

extern "C" int fpv_to_bf16(const float* src, void* dst, int64_t n, void* stream) {
    FPV_REQUIRE(n >= 0, "to_bf16: negative size");
    if (n == 0) return FPV_OK;
    FPV_REQUIRE(src && dst, "to_bf16: null pointer");
    to_bf16_kernel<<<(unsigned)std::min<int64_t>((n + 255) / 256, (int64_t)sm_count() * 16), 256, 0, (cudaStream_t)stream>>>(
        src, reinterpret_cast<__nv_bfloat16*>(dst), n);
    FPV_LAUNCH_CHECK();
    return FPV_OK;
}

extern "C" size_t fpv_gemm_topk_workspace(int64_t q, int64_t n, int d, int k, int kind) {
    if (q <= 0 || d <= 0 || k <= 0) return 256;
    return plan_gemm(q, n, d, k, kind).total;
}

extern "C" size_t fpv_gemm_topk_flags_offset(int64_t q, int64_t n, int d, int k, int kind) {
    if (q <= 0 || d <= 0 || k <= 0) return 0;
    return plan_gemm(q, n, d, k, kind).off_flags;
}

// ---- phases.  The single-GPU search runs PHASE_FILTER | PHASE_FINISH in one call.  The row-sharded search splits
// them around a collective: PHASE_FILTER (+ the local k best approximate values -> approx_out), all-gather, then
// PHASE_FINISH with the gathered values (approx_all): every shard re-ranks only its rows below the GLOBAL limit.
// With more than one shard of >= ~70K rows the filter itself is split once more around a small exchange (PHASE_SAMPLE:
// prep + sampling slab + the k best group values -> approx_out; all-gather; PHASE_SLABS: the k-th best group value of
// the WHOLE job becomes every shard's first threshold, then the filtering slabs): with its own sample only, every shard
// restarts from a threshold G times looser than the job's and appends G times more candidates per slab.
enum { PHASE_FILTER = 1, PHASE_FINISH = 2, PHASE_SAMPLE = 4, PHASE_SLABS = 8 };

struct GemmCall {
    const float* queries; int64_t q; const float* db; const void* db_lowp; int64_t n; int d; int metric; int k; int kind;
    const float* row_sq; const float* aux; float vmax, db_err_abs, db_err_rel; const uint32_t* mask_words; int64_t id_base;
    float* out_dist; int64_t* out_idx; int32_t* out_count; void* ws; size_t ws_bytes; cudaStream_t st;
    uint32_t* approx_out;            // PHASE_FILTER / PHASE_SAMPLE / PHASE_SLABS, optional: [q][k]
    const uint32_t* approx_all;      // PHASE_FINISH / PHASE_SLABS, optional: [shards][q][k]
    int shards;
    const uint32_t* wait_flags;      // optional: approx_all arrives by peer stores, wait for these `shards` flag words
    uint32_t epoch;
};

static int gemm_run(const GemmCall& c, int phases) {
    cudaStream_t st = c.st;
    const int64_t q = c.q, n = c.n;
    const int d = c.d, metric = c.metric, k = c.k, kind = c.kind;
    FPV_REQUIRE(kind == 0 || kind == 1, "gemm: kind must be 0 (tf32) or 1 (bf16)");
    FPV_REQUIRE(metric >= 0 && metric <= 2, "gemm: unknown metric %d", metric);
    FPV_REQUIRE(q >= 1 && q <= (1 << 20) && n >= 1 && n < (1ll << 31), "gemm: bad shape q=%lld n=%lld", (long long)q, (long long)n);
    FPV_REQUIRE(k >= 1 && k <= GEMM_MAX_K, "gemm: k=%d outside [1,%d]", k, GEMM_MAX_K);
    FPV_REQUIRE(d >= 4 && d % (kind == 0 ? 4 : 8) == 0 && d <= 16384, "gemm: d=%d must be a multiple of %d", d, kind == 0 ? 4 : 8);
    FPV_REQUIRE(c.queries && c.db && c.row_sq, "gemm: null pointer");
    FPV_REQUIRE(!(phases & PHASE_FINISH) || (c.out_dist && c.out_idx), "gemm: null output pointer");
    FPV_REQUIRE(kind == 0 || c.db_lowp, "gemm: bf16 pass needs the shadow copy");
    FPV_REQUIRE(metric == FPV_METRIC_IP || c.aux, "gemm: aux array required for cosine / l2");
    FPV_REQUIRE((reinterpret_cast<uintptr_t>(c.db) & 15) == 0 && (reinterpret_cast<uintptr_t>(c.db_lowp) & 15) == 0,
                "gemm: database must be 16-byte aligned");
    FPV_REQUIRE(!c.approx_all || (c.shards >= 1 && (int64_t)c.shards * k <= GEMM_CAP),
                "gemm: shards * k = %lld exceeds %d", (long long)c.shards * k, GEMM_CAP);
    const int filter_phases = phases & (PHASE_FILTER | PHASE_SAMPLE | PHASE_SLABS);
    if (g_prof_on && (phases & (PHASE_FILTER | PHASE_SAMPLE))) g_prof_n = 0;
    GemmPlan pl = plan_gemm(q, n, d, k, kind);
    if (!c.ws || c.ws_bytes < pl.total) { set_error("gemm: workspace %zu < %zu", c.ws_bytes, pl.total); return FPV_ERR_WORKSPACE; }
    FPV_REQUIRE((reinterpret_cast<uintptr_t>(c.ws) & 255) == 0, "gemm: workspace must be 256-byte aligned");
    char* w = static_cast<char*>(c.ws);
    float* qprep = reinterpret_cast<float*>(w + pl.off_qprep);
    void* qa = w + pl.off_qa;
    float* qsq = reinterpret_cast<float*>(w + pl.off_qsq);
    float* eb = reinterpret_cast<float*>(w + pl.off_eb);
    float* thr = reinterpret_cast<float*>(w + pl.off_thr);
    uint32_t* cnt = reinterpret_cast<uint32_t*>(w + pl.off_cnt);
    uint32_t* flags = reinterpret_cast<uint32_t*>(w + pl.off_flags);
    uint64_t* cand = reinterpret_cast<uint64_t*>(w + pl.off_cand);

    constexpr int MAX_DEV = 32;
    static std::mutex attr_mutex;                  // callers may search from several host threads
    std::unique_lock<std::mutex> attr_lock(attr_mutex);
    static bool attr_set[MAX_DEV][2][2][3];
    static int groups_cache[MAX_DEV][2][3];
    static bool tighten_set[MAX_DEV];
    static int fin_smem_set[MAX_DEV];
    int dev_id = 0;
    FPV_CUDA(cudaGetDevice(&dev_id));
    const bool cacheable = dev_id >= 0 && dev_id < MAX_DEV;

    if (filter_phases) {
        // error bound of one product term: query pre-rounded to nearest (2^-11 TF32 / 2^-9 BF16), database element truncated
        // by the TF32 datapath (2^-10) or rounded to BF16 (2^-9); + fp32 accumulation slack.
        const float eps = kind == 0 ? 1.65e-3f : 4.2e-3f;
        FPV_REQUIRE(!(c.db_err_abs > 0.f) || (kind == 1 && c.db_err_rel > 0.f), "gemm: measured error bounds are for the bf16 pass");
        if (!(phases & PHASE_SLABS)) {                 // PHASE_SLABS continues a PHASE_SAMPLE call: the operands are in ws
            gemm_prep_kernel<<<(pl.Qp + 7) / 8, 256, 0, st>>>(c.queries, (int)q, pl.Qp, d, pl.Dp, metric, kind, eps, c.vmax, c.db_err_abs,
                                                              c.db_err_rel, qprep, qa, qsq, eb, thr, cnt, flags);
            FPV_LAUNCH_CHECK();
        }

        CUtensorMap tmA, tmB;
        int rc = make_map(&tmA, qa, kind, pl.Qp, pl.Dp, pl.Dp, BM);
        if (rc != FPV_OK) return rc;
        // CTA pairs whenever there is more than one query block (a single block keeps the one-CTA kernel)
        const int ncta = (pl.Qp / BM) % 2 == 0 && pair_mode_enabled() ? 2 : 1;
        rc = make_map(&tmB, kind == 0 ? (const void*)c.db : c.db_lowp, kind, n, d, d, BN / ncta);
        if (rc != FPV_OK) return rc;

        typedef void (*FilterKernel)(const CUtensorMap, const CUtensorMap, GemmParams);
        static const FilterKernel kernels[2][2][3] = {
            {{gemm_filter_kernel<0, FPV_METRIC_COSINE, 1>, gemm_filter_kernel<0, FPV_METRIC_L2, 1>, gemm_filter_kernel<0, FPV_METRIC_IP, 1>},
             {gemm_filter_kernel<1, FPV_METRIC_COSINE, 1>, gemm_filter_kernel<1, FPV_METRIC_L2, 1>, gemm_filter_kernel<1, FPV_METRIC_IP, 1>}},
            {{gemm_filter_kernel<0, FPV_METRIC_COSINE, 2>, gemm_filter_kernel<0, FPV_METRIC_L2, 2>, gemm_filter_kernel<0, FPV_METRIC_IP, 2>},
             {gemm_filter_kernel<1, FPV_METRIC_COSINE, 2>, gemm_filter_kernel<1, FPV_METRIC_L2, 2>, gemm_filter_kernel<1, FPV_METRIC_IP, 2>}}};
        const FilterKernel filter = kernels[ncta - 1][kind][metric];
        const size_t filter_smem = ncta == 2 ? GemmCfg<2>::SMEM : GemmCfg<1>::SMEM;
        // function attributes and the cluster occupancy are per device and per kernel: set / query them once (they cost
        // several microseconds of host time per call, which is visible in small-batch latency)
        if (!cacheable || !attr_set[dev_id][ncta - 1][kind][metric]) {
            FPV_CUDA(cudaFuncSetAttribute(filter, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)filter_smem));
            if (cacheable) attr_set[dev_id][ncta - 1][kind][metric] = true;
        }
        constexpr size_t TW_SMEM = (size_t)4 * (GEMM_CAP + 256) * 4;        // tighten_warp_kernel: 4 warps x (values + histogram)
        if (!cacheable || !tighten_set[dev_id]) {
            FPV_CUDA(cudaFuncSetAttribute(tighten_warp_kernel<GEMM_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)TW_SMEM));
            FPV_CUDA(cudaFuncSetAttribute(tighten_kernel<GEMM_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, GEMM_CAP * 8));
            if (cacheable) tighten_set[dev_id] = true;
        }
        cudaLaunchConfig_t cfg{};
        cudaLaunchAttribute cluster_attr{};
        cluster_attr.id = cudaLaunchAttributeClusterDimension;
        cluster_attr.val.clusterDim.x = (unsigned)ncta; cluster_attr.val.clusterDim.y = 1; cluster_attr.val.clusterDim.z = 1;
        cfg.blockDim = dim3(GEMM_THREADS); cfg.dynamicSmemBytes = filter_smem; cfg.stream = st;
        cfg.attrs = &cluster_attr; cfg.numAttrs = 1;
        int max_groups = sm_count();                   // co-resident CTAs (ncta == 1) or CTA pairs (ncta == 2)
        if (ncta == 2) {
            if (cacheable && groups_cache[dev_id][kind][metric] > 0) {
                max_groups = groups_cache[dev_id][kind][metric];
            } else {
                cfg.gridDim = dim3(2 * (unsigned)sm_count());
                FPV_CUDA(cudaOccupancyMaxActiveClusters(&max_groups, filter, &cfg));
                FPV_REQUIRE(max_groups >= 1, "gemm: no CTA pair fits on this device");
                if (cacheable) groups_cache[dev_id][kind][metric] = max_groups;
            }
        }
        const int kel = KROW / pl.esz;
        GemmParams p{};
        p.aux = metric == FPV_METRIC_IP ? nullptr : c.aux;
        p.mask = c.mask_words;
        p.thr = thr; p.cnt = cnt; p.cand = cand; p.N = n; p.Q = (int)q; p.m_blocks = pl.Qp / BM;
        p.nkb = (d + kel - 1) / kel; p.metric = metric;
#ifdef FPV_GEMM_TRACE
        { const char* e = getenv("FPV_GEMM_DEBUG"); p.debug = e ? atoi(e) : 0; }
#endif
        const int64_t tiles_total = (n + BN - 1) / BN;
        const int mgroups = p.m_blocks / ncta;
        auto launch_filter = [&](int64_t tile0, int64_t take, int sample) -> int {
            p.tile0 = (int)tile0; p.ntiles = (int)take; p.sample = sample;
            // the sample is spread over the whole database (every (tiles / ts)-th tile): a prefix would give a useless
            // first threshold on data stored in cluster or time order
            p.tile_stride = sample && sample_stride_enabled() ? (int)std::max<int64_t>(1, tiles_total / std::max<int64_t>(1, take)) : 1;
            const int64_t work = (int64_t)mgroups * take;
            cfg.gridDim = dim3((unsigned)(ncta * std::min<int64_t>(work, max_groups)));
            const bool prof = g_prof_on && g_prof_n < PROF_MAX;
            if (prof) FPV_CUDA(cudaEventRecord(g_prof_ev[2 * g_prof_n], st));
            FPV_CUDA(cudaLaunchKernelEx(&cfg, filter, tmA, tmB, p));
            FPV_LAUNCH_CHECK();
            if (prof) { FPV_CUDA(cudaEventRecord(g_prof_ev[2 * g_prof_n + 1], st)); ++g_prof_n; }
            return FPV_OK;
        };
        // few queries: a warp per query leaves the GPU empty and each lane walks 128 keys (Q = 64: 40 us for the 4096
        // group keys of the sampling slab), so a whole CTA takes the query (one round of loads, block-wide select)
        const bool tighten_by_cta = q <= 2 * (int64_t)sm_count();
        auto launch_tighten = [&](uint32_t* approx_out, int sample_groups) -> int {
            if (tighten_by_cta)
                tighten_kernel<GEMM_CAP><<<(unsigned)q, 512, (size_t)GEMM_CAP * 8, st>>>(cand, cnt, thr, eb, flags, k, approx_out, 0.0f,
                                                                                      sample_groups);
            else
                tighten_warp_kernel<GEMM_CAP><<<(unsigned)((q + 3) / 4), 128, TW_SMEM, st>>>(cand, cnt, thr, eb, flags, k, approx_out, (int)q,
                                                                                          sample_groups);
            FPV_LAUNCH_CHECK();
            return FPV_OK;
        };
        // ---- sampling slab: one wave of work (at least 32 tiles = 8192 rows and 2k groups of 8 rows, at most 128 tiles:
        // GEMM_CAP / 32 keys), reduced to group maxima in the epilogue -> the first threshold (see sample_tile).
        // 8192 rows, not 4096: the first filtering slab then sees ~1.6 % hits instead of 3.3 %, which keeps its epilogue on
        // the fast path (<= HIT_BUF hits per thread and tile); at 3.3 % two thirds of the warps took the two-pass path and
        // the slab ran at half speed (150 us for 18K rows against 28 us more for the larger sample).
        int64_t ts = (max_groups + mgroups - 1) / mgroups;
        ts = std::max<int64_t>(ts, 32);
        ts = std::max<int64_t>(ts, ((int64_t)2 * k * SAMPLE_GW + BN - 1) / BN);
        ts = std::min<int64_t>(ts, GEMM_CAP / (2 * SAMPLE_KEYS));
        ts = std::min<int64_t>(ts, tiles_total);
        const bool sampling = sample_enabled() && tiles_total > 2 * ts;
        // pl.keep is the budgeted number of rows inside the 2E window (the measured counts are about a third of it; growth
        // factors of 9-14 measured no faster than 7, and 20 overflows the buffers: 92 % of the queries fall back); a slab
        // that is (growth-1) times the rows seen so far then adds <= ~3072 hits per query to a 4096-slot buffer
        const double growth = 1.0 + 3072.0 / pl.keep;
        int64_t done = 0, slab;
        FPV_REQUIRE(sampling || !(phases & (PHASE_SAMPLE | PHASE_SLABS)), "gemm: shard of %lld rows is too small for the sample exchange",
                    (long long)n);
        if (phases & PHASE_SLABS) {
            // the sample of the WHOLE job (shards x ts tiles) gives the first threshold
            FPV_REQUIRE(c.approx_all && c.shards >= 1 && (int64_t)c.shards * k <= GEMM_CAP, "gemm: bad gathered sample");
            const size_t gt_smem = (size_t)8 * (c.shards * k + 256) * 4;
            if (gt_smem > 48 * 1024)
                FPV_CUDA(cudaFuncSetAttribute(gemm_global_thr_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gt_smem));
            gemm_global_thr_kernel<<<(unsigned)((q + 7) / 8), 256, gt_smem, st>>>(c.approx_all, c.shards, (int)q, k, eb, thr,
                                                                                  c.wait_flags, c.epoch);
            FPV_LAUNCH_CHECK();
            slab = (int64_t)((double)ts * c.shards / 1.35 * (growth - 1.0));
            if (slab < ts) slab = ts;
        } else if (sampling) {
            int rc2 = launch_filter(0, ts, 1);
            if (rc2 != FPV_OK) return rc2;
            rc2 = launch_tighten((phases & PHASE_SAMPLE) ? c.approx_out : nullptr, (int)(ts * 2 * SAMPLE_KEYS));
            if (rc2 != FPV_OK) return rc2;
            if (phases & PHASE_SAMPLE) return FPV_OK;
            // the k-th best of the group maxima is as tight as the exact k-th best of ~1/1.35 of the sample rows
            slab = (int64_t)((double)ts / 1.35 * (growth - 1.0));
            if (slab < ts) slab = ts;
        } else {
            slab = first_slab_rows(k) / BN;          // dense first slab: every row is a candidate
        }
        p.slab = 0;
        while (done < tiles_total) {
            int64_t take = std::min<int64_t>(slab, tiles_total - done);
            // a remainder of less than half a slab joins this one: one launch + one tighten less, for at most 1.5x the
            // budgeted hits (the budget itself is ~2x the measured counts)
            if (tiles_total - done - take < take / 2) take = tiles_total - done;
            p.slab += (done > 0);
            int rc2 = launch_filter(done, take, 0);
            if (rc2 != FPV_OK) return rc2;
            done += take;
            // between slabs: raise the threshold to a_k + 2E.  After the last slab only the sharded search tightens
            // (it needs the local k best approximate values for the exchange and a compact candidate list).
            if (done < tiles_total || c.approx_out) {
                rc2 = launch_tighten(done < tiles_total ? nullptr : c.approx_out, 0);
                if (rc2 != FPV_OK) return rc2;
            }
            slab = (int64_t)((double)done * (growth - 1.0));
            if (slab < 1) slab = 1;
        }
    }
    if (!(phases & PHASE_FINISH)) return FPV_OK;
    const size_t fin_smem = (size_t)(GEMM_CAP + FIN_RMAX) * 8 + (size_t)((d + 3) / 4 * 4) * 4;
    FPV_REQUIRE(fin_smem <= (size_t)max_smem_optin(), "gemm: d=%d too large for the finish kernel", d);
    if (!cacheable || fin_smem_set[dev_id] < (int)fin_smem) {
        FPV_CUDA(cudaFuncSetAttribute(gemm_finish2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)fin_smem));
        if (cacheable) fin_smem_set[dev_id] = (int)fin_smem;
    }
    // few queries: one CTA per query cannot fill the GPU, so give each CTA 32 warps for the row gather (measured at
    // Q = 64: 69 us with 8 warps, the largest item after the filter itself)
    const int fin_threads = q <= 2 * (int64_t)sm_count() ? 1024 : 256;
    gemm_finish2_kernel<<<(unsigned)q, fin_threads, fin_smem, st>>>(cand, cnt, thr, eb, flags, qprep, qsq, c.db, c.row_sq, d, d, metric, k,
                                                            c.id_base, c.out_dist, c.out_idx, c.out_count, c.approx_all,
                                                            c.shards, (int)q, c.wait_flags, c.epoch);
    FPV_LAUNCH_CHECK();
    // exact fp32 scan for the queries whose certificate failed (normally none): decided on the device
    return scan_f32_flagged(c.queries, q, c.db, n, d, d, metric, k, c.row_sq, c.id_base, flags, c.mask_words, c.out_dist, c.out_idx,
                            c.out_count, w + pl.off_scan, pl.scan_bytes, st);
}

// kind 0: TF32 tensor-core pass straight from the fp32 rows (db_lowp ignored); kind 1: BF16 pass over db_lowp, a
// [n][d] bf16 shadow copy made by fpv_to_bf16.  aux: per-row 1/(|v|+1e-10) for cosine, row_sq for l2, NULL for ip.
// vmax = max row norm (error bound).  Requires 16 <= q, k <= 256, d % 4 == 0 (TF32) or d % 8 == 0 (BF16).
extern "C" int fpv_gemm_topk_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                                 int metric, int k, int kind, const float* row_sq, const float* aux, float vmax,
                                 float db_err_abs, float db_err_rel, const uint32_t* mask_words, int64_t id_base,
                                 float* out_dist, int64_t* out_idx, int32_t* out_count, void* ws, size_t ws_bytes,
                                 void* stream) {
    GemmCall c{queries, q, db, db_lowp, n, d, metric, k, kind, row_sq, aux, vmax, db_err_abs, db_err_rel, mask_words, id_base,
               out_dist, out_idx, out_count, ws, ws_bytes, (cudaStream_t)stream, nullptr, nullptr, 0, nullptr, 0u};
    return gemm_run(c, PHASE_FILTER | PHASE_FINISH);
}

// Row-sharded search, phase 1 (this GPU's rows only): tensor-core filter; leaves the candidate lists in `ws` and writes
// the k best approximate values of this shard per query to approx_out [q][k] (uint32, order-preserving encoding,
// padded with +inf).  The caller all-gathers approx_out over the shards and calls fpv_gemm_finish_sharded_f32 with
// the SAME ws.  vmax / db_err_* must be the maxima over ALL shards (the error bound must hold on every shard).
extern "C" int fpv_gemm_filter_sharded_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                                           int metric, int k, int kind, const float* row_sq, const float* aux, float vmax,
                                           float db_err_abs, float db_err_rel, const uint32_t* mask_words, uint32_t* approx_out,
                                           void* ws, size_t ws_bytes, void* stream) {
    FPV_REQUIRE(approx_out, "gemm_filter_sharded: null approx_out");
    GemmCall c{queries, q, db, db_lowp, n, d, metric, k, kind, row_sq, aux, vmax, db_err_abs, db_err_rel, mask_words, 0,
               nullptr, nullptr, nullptr, ws, ws_bytes, (cudaStream_t)stream, approx_out, nullptr, 0, nullptr, 0u};
    return gemm_run(c, PHASE_FILTER);
}

// Phase 1 split around the sample exchange (shards of >= 70K rows).  First half: prep + sampling slab; sample_out
// [q][k] = this shard's k best group values (same encoding as approx_out).  Second half (SAME ws, same arguments):
// sample_all [shards][q][k] = the gathered first halves (a plain device buffer, or this rank's peer-memory gather area
// with wait_flags / epoch as in fpv_gemm_finish_sharded_peer_f32; wait_flags NULL otherwise) -> global first
// threshold, filtering slabs, approx_out as fpv_gemm_filter_sharded_f32.
extern "C" int fpv_gemm_sample_sharded_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                                           int metric, int k, int kind, const float* row_sq, const float* aux, float vmax,
                                           float db_err_abs, float db_err_rel, const uint32_t* mask_words, uint32_t* sample_out,
                                           void* ws, size_t ws_bytes, void* stream) {
    FPV_REQUIRE(sample_out, "gemm_sample_sharded: null sample_out");
    GemmCall c{queries, q, db, db_lowp, n, d, metric, k, kind, row_sq, aux, vmax, db_err_abs, db_err_rel, mask_words, 0,
               nullptr, nullptr, nullptr, ws, ws_bytes, (cudaStream_t)stream, sample_out, nullptr, 0, nullptr, 0u};
    return gemm_run(c, PHASE_SAMPLE);
}

extern "C" int fpv_gemm_slabs_sharded_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                                          int metric, int k, int kind, const float* row_sq, const float* aux, float vmax,
                                          float db_err_abs, float db_err_rel, const uint32_t* mask_words, const uint32_t* sample_all,
                                          int shards, const uint32_t* wait_flags, uint32_t epoch, uint32_t* approx_out, void* ws,
                                          size_t ws_bytes, void* stream) {
    FPV_REQUIRE(sample_all && approx_out && shards >= 1 && shards <= 256, "gemm_slabs_sharded: null pointer");
    GemmCall c{queries, q, db, db_lowp, n, d, metric, k, kind, row_sq, aux, vmax, db_err_abs, db_err_rel, mask_words, 0,
               nullptr, nullptr, nullptr, ws, ws_bytes, (cudaStream_t)stream, approx_out, sample_all, shards, wait_flags, epoch};
    return gemm_run(c, PHASE_SLABS);
}

// Row-sharded search, phase 2: approx_all [shards][q][k] = the gathered phase-1 outputs.  Selects the k-th best
// approximate value of the WHOLE job per query, re-ranks this shard's candidates below (that + 2E) in exact fp32 and
// writes this shard's (distance, global id) list [q][k] ordered by (distance, id), padded with (+inf, -1).  Merging
// the shards' lists (fpv_merge_packed) gives the exact global top-k.
extern "C" int fpv_gemm_finish_sharded_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                                           int metric, int k, int kind, const float* row_sq, const uint32_t* mask_words,
                                           int64_t id_base, const uint32_t* approx_all, int shards, float* out_dist,
                                           int64_t* out_idx, int32_t* out_count, void* ws, size_t ws_bytes, void* stream) {
    FPV_REQUIRE(approx_all && shards >= 1, "gemm_finish_sharded: null approx_all");
    GemmCall c{queries, q, db, db_lowp, n, d, metric, k, kind, row_sq, row_sq /* aux unused */, 0.f, 0.f, 0.f, mask_words, id_base,
               out_dist, out_idx, out_count, ws, ws_bytes, (cudaStream_t)stream, nullptr, approx_all, shards, nullptr, 0u};
    return gemm_run(c, PHASE_FINISH);
}

// Phase 2 as the consumer of a peer-memory exchange (fpv_peer_put): approx_all is this rank's gather area; the kernel
// waits until the `shards` flag words at wait_flags have reached `epoch` before it reads it.
extern "C" int fpv_gemm_finish_sharded_peer_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                                                int metric, int k, int kind, const float* row_sq, const uint32_t* mask_words,
                                                int64_t id_base, const uint32_t* approx_all, int shards, const uint32_t* wait_flags,
                                                uint32_t epoch, float* out_dist, int64_t* out_idx, int32_t* out_count, void* ws,
                                                size_t ws_bytes, void* stream) {
    FPV_REQUIRE(approx_all && wait_flags && shards >= 1 && shards <= 256, "gemm_finish_sharded_peer: null pointer");
    GemmCall c{queries, q, db, db_lowp, n, d, metric, k, kind, row_sq, row_sq /* aux unused */, 0.f, 0.f, 0.f, mask_words, id_base,
               out_dist, out_idx, out_count, ws, ws_bytes, (cudaStream_t)stream, nullptr, approx_all, shards, wait_flags, epoch};
    return gemm_run(c, PHASE_FINISH);
}
