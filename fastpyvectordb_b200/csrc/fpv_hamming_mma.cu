// Batched Hamming scan on the int8 tensor cores.
//
// Replaces BinaryQuantizer.hamming_distances + the caller's top-k (quantization.py:356-394) for BATCHES of queries.
// The CUDA-core scan (fpv_hamming.cu) is HBM bound for one query and POPC-issue bound beyond ~3 queries per pass
// (measured r1: 16 queries over 20M x 1024 bits = 4 passes, 3.55 ms, 11 % of HBM).  Here one pass serves 31 queries:
//
//     popc(x ^ q) = popc(x) + popc(q) - 2 popc(x & q),      popc(x & q) = sum_bits x_b q_b
//
// is an integer dot product once the bits are bytes.  The packed codes stay packed in HBM and in the TMA-staged shared
// memory tile; the expansion happens inside the SM, straight into TENSOR MEMORY: an expander thread owns one database
// row, turns 16 packed bytes into 128 operand bytes with 7 shifts + 8 byte-permutes per 32-bit word (PRMT with
// sign replication: byte b of the result is 0xFF iff bit t of byte b is set -- read as int8 that is -1) and writes them
// with one tcgen05.st: TMEM lane = row, 32 columns = 128 operand bytes, exactly the layout tcgen05.mma reads its A
// operand from (A-from-TMEM form), so the expanded data never touches shared memory.  The queries (31 + one all-ones
// row whose dot product is -popc(x)) are expanded once to u8 {0,1} rows and stay resident as the B operand.
// D = -popc(x & q) exactly (s8 x u8 -> s32); the epilogue forms the integer distance and filters it against the
// query's threshold; hits go to the candidate list, which a radix select tightens between row slabs (the same
// filter-then-select structure as the float and uint8-scalar paths).  Distances are integers: no error bound, and a
// later row that only ties the k-th value loses to the lower row ids already held (tighten_kernel thr_shift = -1).
//
// Kernel shape: persistent, one CTA per SM, 896 threads: warp 0 TMA producer (raw tiles, SWIZZLE_128B so that the
// expanders' 16-byte reads are conflict free), warp 1 MMA issuer, warp 2 TMEM allocator, warps 4-19 expanders,
// warps 20-27 epilogue.
#include <cuda.h>

#include <algorithm>
#include <cstdlib>
#include <mutex>

#include "fpv_common.cuh"
#include "fpv_select.cuh"
#include "fpv_tc.cuh"

namespace fpv {

constexpr int HM_BM = 128;                 // database rows per tile (UMMA M)
constexpr int HM_N = 32;                   // B rows: 31 queries + the all-ones row
constexpr int HM_QB = 31;                  // queries per pass
constexpr int HM_ONES = 31;                // index of the all-ones row
constexpr int HM_RAW_STAGES = 8;           // ring slots reserved; 128 KB of raw tiles in flight per SM (8 x 16 KB or 4 x 32 KB):
                                           // with 3 x 16 KB the TMA round trip (~2.7 us under load) bounded the scan at ~1.3 ms
constexpr int HM_SB_KB = 4;                // K blocks per operand stage ("super block"): one handshake + one tcgen05.commit per
                                           // 64 packed bytes of every row; with one per K block the scan was bound by those (1.3 of 1.6 ms)
constexpr int HM_A_STAGES = 3;             // operand stages in TMEM, 128 columns each
constexpr int HM_ACC_COLS = 32;
constexpr int HM_A_COL0 = 2 * HM_ACC_COLS;
constexpr int HM_TMEM_COLS = 512;             // 2 x 32 accumulator columns + 8 x 32 operand columns (power of two)
constexpr int HM_CAP = 16384;
constexpr int HM_EXP_WARPS = 16;             // expander warps: 4 TMEM lane quarters x the 4 K blocks of a super block
constexpr int HM_EPI_WARPS = 8;              // two groups of four (one per TMEM lane quarter): queries 0-15 and 16-30
constexpr int HM_THREADS = 32 * (4 + HM_EXP_WARPS + HM_EPI_WARPS);
constexpr int HM_SORT_MAX = 4096;
constexpr int HM_MAX_NBYTES = 256;

struct HmParams {
    const uint32_t* mask;       // optional row filter
    const int* pq;              // [QB] popc(q & dimmask)
    const float* thr;           // [QB] -(bound)
    uint32_t* cnt;              // [QB]
    uint64_t* cand;             // [QB][HM_CAP]   ordered(float(distance)) << 32 | row
    int64_t N;
    int nq, nbytes, nkb;        // nkb = nbytes / 16 K blocks per tile
    int raw_stages;             // raw tiles in flight (<= HM_RAW_STAGES)
    int tile0, ntiles;
    int* dump;                  // test hook: [32][N] raw accumulators
    int debug;                  // experiments (FPV_HAM_DEBUG): 1 = no expansion, 2 = no MMAs, 4 = no per-query epilogue work
};

// A from tensor memory, B from shared memory
__device__ __forceinline__ void tc_mma_i8_ta(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
// c_format S32 (2 << 4), a_format signed 8 bit (1 << 7), b_format unsigned (0), K-major, N >> 3 at 17, M >> 4 at 24
constexpr uint32_t HM_IDESC = (2u << 4) | (1u << 7) | ((uint32_t)(HM_N >> 3) << 17) | ((uint32_t)(HM_BM >> 4) << 24);

// prmt.b32 with the sign-replicate bit (8) set in every selector nibble: result byte b = 0xFF if the msb of byte b of x
// is set, else 0x00.  (__byte_perm() masks the selector to three bits per nibble and cannot express this.)
__device__ __forceinline__ uint32_t prmt_sign(uint32_t x) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(x), "r"(0u), "r"(0xBA98u));
    return d;
}

#define TMEM_ST32(taddr, r)                                                                                              \
    asm volatile("tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "                                                         \
                 "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28," \
                 "%29,%30,%31,%32};"                                                                                     \
                 ::"r"(taddr), "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]),     \
                   "r"(r[8]), "r"(r[9]), "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]),           \
                   "r"(r[16]), "r"(r[17]), "r"(r[18]), "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]),         \
                   "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]), "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])          \
                 : "memory")

__global__ void __launch_bounds__(HM_THREADS, 1)
ham_mma_kernel(const __grid_constant__ CUtensorMap tmRaw, const __grid_constant__ CUtensorMap tmB, HmParams p) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    const uint32_t base = smem_u32(smem_raw);
    if (base & 1023u) __trap();
    const int raw_bytes = HM_BM * p.nbytes;                               // one raw tile (16 KB per 128-byte column block)
    const uint32_t sRaw = base;
    const uint32_t sB = base + p.raw_stages * raw_bytes;                  // [nkb][32 rows x 128 B]
    const uint32_t off_bar = p.raw_stages * raw_bytes + p.nkb * (HM_N * 128);
    const uint32_t bars = base + off_bar;
    const uint32_t bar_rfull = bars, bar_rempty = bars + 8 * HM_RAW_STAGES;
    const uint32_t bar_afull = bars + 16 * HM_RAW_STAGES, bar_aempty = bar_afull + 8 * HM_A_STAGES;
    const uint32_t bar_b = bar_aempty + 8 * HM_A_STAGES, bar_tfull = bar_b + 8, bar_tempty = bar_tfull + 16;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_raw + off_bar + 304);
    int* qci = reinterpret_cast<int*>(smem_raw + off_bar + 320);           // [32] popc(q), [32] integer bound
    const int warp = __shfl_sync(FPV_FULL_MASK, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < HM_RAW_STAGES; ++s) { mbar_init(bar_rfull + 8 * s, 1); mbar_init(bar_rempty + 8 * s, HM_EXP_WARPS); }
        for (int s = 0; s < HM_A_STAGES; ++s) { mbar_init(bar_afull + 8 * s, HM_EXP_WARPS); mbar_init(bar_aempty + 8 * s, 1); }
        mbar_init(bar_b, 1);
        for (int a = 0; a < 2; ++a) { mbar_init(bar_tfull + 8 * a, 1); mbar_init(bar_tempty + 8 * a, HM_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (threadIdx.x < 32) {
        const int q = threadIdx.x;
        const int pq = q < p.nq ? p.pq[q] : 0;
        const float b = q < p.nq ? -p.thr[q] : -1.0f;                      // padding queries never hit
        const int bound = b >= 2.0e9f ? 0x3FFFFFFF : (b < -1.0f ? -1 : (int)floorf(b));
        qci[q] = pq;
        qci[32 + q] = bound - pq;                                           // hit  <=>  popc(x) - 2 popc(x & q) <= bound - popc(q)
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(HM_TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = __shfl_sync(FPV_FULL_MASK, *tmem_slot, 0);
    const int ncb = p.nbytes >> 7;                                          // 128-byte column blocks per row

    if (warp == 0) {                                    // ---------------- TMA producer: expanded queries once, raw tiles
        if (elect_one()) {
            mbar_expect_tx(bar_b, (uint32_t)(p.nkb * HM_N * 128));
            for (int kb = 0; kb < p.nkb; ++kb) tma_load_2d(sB + kb * (HM_N * 128), &tmB, bar_b, kb * 128, 0);
        }
        __syncwarp();
        int rs = 0; uint32_t rph = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const int row0 = (p.tile0 + t) * HM_BM;
            mbar_wait(bar_rempty + 8 * rs, rph ^ 1);
            if (elect_one()) {
                mbar_expect_tx(bar_rfull + 8 * rs, (uint32_t)raw_bytes);
                for (int cb = 0; cb < ncb; ++cb)
                    tma_load_2d(sRaw + rs * raw_bytes + cb * (HM_BM * 128), &tmRaw, bar_rfull + 8 * rs, cb * 128, row0);
            }
            __syncwarp();
            if (++rs == p.raw_stages) { rs = 0; rph ^= 1; }
        }
    } else if (warp == 1) {                             // ---------------- MMA issuer
        mbar_wait(bar_b, 0);
        tc_fence_after();
        uint32_t g = 0; int as = 0; uint32_t aphase = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            mbar_wait(bar_tempty + 8 * as, aphase ^ 1);
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + as * HM_ACC_COLS;
            for (int sb = 0; sb < p.nkb / HM_SB_KB; ++sb, ++g) {
                const int s = g % HM_A_STAGES;
                if (!(p.debug & 32)) mbar_wait(bar_afull + 8 * s, (g / HM_A_STAGES) & 1);
                tc_fence_after();
                if (elect_one()) {
                    if (!(p.debug & 2))
#pragma unroll
                    for (int kl = 0; kl < HM_SB_KB; ++kl) {
                        const int kb = sb * HM_SB_KB + kl;
                        const uint64_t bd = make_smem_desc(sB + kb * (HM_N * 128));
                        const uint32_t a_tmem = tmem_base + HM_A_COL0 + s * (32 * HM_SB_KB) + kl * 32;
#pragma unroll
                        for (int k = 0; k < 4; ++k)                   // 32 operand bytes (= 8 TMEM columns) of K per instruction
                            tc_mma_i8_ta(d_tmem, a_tmem + k * 8, bd + 2 * k, HM_IDESC, (kb | k) != 0);
                    }
                    tc_commit(bar_aempty + 8 * s);
                }
                __syncwarp();
            }
            if (elect_one()) tc_commit(bar_tfull + 8 * as);
            __syncwarp();
            as ^= 1; if (as == 0) aphase ^= 1;
        }
    } else if (warp >= 4 && warp < 4 + HM_EXP_WARPS) {  // ---------------- expanders: packed bits -> int8 operand in TMEM
        // 16 warps: TMEM lane quarter = warp % 4 (hardware rule for tcgen05.st); warp class (warp - 4) / 4 expands K block
        // `cls` of every super block, so all 16 warps fill one operand stage together and arrive on its barrier once
        const int e = warp - 4, quarter = e & 3, cls = e >> 2;             // cls = K block of every super block this warp expands
        const int r = quarter * 32 + lane;                                  // row of the tile = TMEM lane
        const uint32_t row_off = (uint32_t)(r >> 3) * 1024u + (uint32_t)(r & 7) * 128u;
        int rs = 0; uint32_t rph = 0; uint32_t g0 = 0;
        const int nsb = p.nkb / HM_SB_KB;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x, g0 += nsb) {
            mbar_wait(bar_rfull + 8 * rs, rph);
            const uint32_t tile = sRaw + rs * raw_bytes;
            for (int sb = 0; sb < nsb; ++sb) {
                const uint32_t g = g0 + sb;
                const int s = g % HM_A_STAGES;
                const int kb = sb * HM_SB_KB + cls;
                // 16 packed bytes of this row: column block kb / 8, 16-byte chunk kb % 8 (XOR-swizzled by the row)
                const uint32_t addr = tile + (uint32_t)(kb >> 3) * (HM_BM * 128) + row_off + (uint32_t)(((kb & 7) ^ (r & 7)) << 4);
                uint32_t w[4] = {0u, 0u, 0u, 0u};
                if (!(p.debug & 64))
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(addr));
                uint32_t o[32];
#pragma unroll
                for (int wi = 0; wi < 4; ++wi)
#pragma unroll
                    for (int tb = 0; tb < 8; ++tb)                    // byte b of the result = 0xFF iff bit tb of byte b of w
                        o[wi * 8 + tb] = prmt_sign(w[wi] << (7 - tb));
                if (p.debug & 32) continue;
                mbar_wait(bar_aempty + 8 * s, ((g / HM_A_STAGES) & 1) ^ 1);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + HM_A_COL0 + s * (32 * HM_SB_KB) + cls * 32;
                if (!(p.debug & 1)) {
                    TMEM_ST32(taddr, o);
                    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                }
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(bar_afull + 8 * s);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_rempty + 8 * rs);               // this warp is done reading the raw tile
            if (++rs == p.raw_stages) { rs = 0; rph ^= 1; }
        }
    } else if (warp >= 4 + HM_EXP_WARPS) {              // ---------------- epilogue: integer distances, threshold filter
        const int quarter = warp & 3, group = (warp - 4 - HM_EXP_WARPS) >> 2;     // group 0: queries 0-15, group 1: 16-30
        int as = 0; uint32_t aphase = 0;
        for (int t = blockIdx.x; t < p.ntiles; t += gridDim.x) {
            const int64_t row = (int64_t)(p.tile0 + t) * HM_BM + quarter * 32 + lane;
            const bool valid = row < p.N && (!p.mask || mask_bit(p.mask, row));
            mbar_wait(bar_tfull + 8 * as, aphase);
            tc_fence_after();
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + as * HM_ACC_COLS;
            uint32_t acc[32];
            if (!(p.debug & 16)) {
                TMEM_LD32(acc, taddr);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(bar_tempty + 8 * as);
            as ^= 1; if (as == 0) aphase ^= 1;
            if (p.dump) {                                                   // uniform; test hook only
                if (row < p.N && group == 0)
#pragma unroll
                    for (int c = 0; c < 32; ++c) p.dump[(size_t)c * p.N + row] = (int)acc[c];
                continue;
            }
            if (p.debug & 4) continue;
            const int px = -(int)acc[HM_ONES];                              // popc(x & dimmask)
            // Pass 1: the 16 hit tests of this warp's queries, independent of each other (one broadcast LDS, one IMAD, one
            // compare per query; the votes are OR-ed).  A vote + branch per query made every query a ~60-cycle dependent
            // chain (43 us per query per 20M-row scan); hits are rare after the first slab, so one branch per tile remains.
            uint32_t hitbits = 0;
#pragma unroll
            for (int j = 0; j < 16; ++j) {
                const int qi = group * 16 + j;
                const int acc_q = group == 0 ? (int)acc[j] : (int)acc[16 + j];
                const bool hit = valid && qi < p.nq && px + 2 * acc_q <= qci[32 + qi];   // popc(x) - 2 popc(x & q) <= bound - popc(q)
                hitbits |= hit ? (1u << j) : 0u;
            }
            if (__any_sync(FPV_FULL_MASK, hitbits != 0)) {
#pragma unroll 1
                for (int j = 0; j < 16; ++j) {
                    const uint32_t m = __ballot_sync(FPV_FULL_MASK, (hitbits >> j) & 1u);
                    if (m) {
                        const int qi = group * 16 + j;
                        const bool hit = (hitbits >> j) & 1u;
                        // accumulator j of this group: select without a dynamic register index
                        int acc_q = 0;
#pragma unroll
                        for (int jj = 0; jj < 16; ++jj) if (jj == j) acc_q = group == 0 ? (int)acc[jj] : (int)acc[16 + jj];
                        const int ham = px + qci[qi] + 2 * acc_q;           // popc(x) + popc(q) - 2 popc(x & q)
                        const int leader = __ffs(m) - 1;
                        uint32_t pos = 0;
                        if (lane == leader) pos = atomicAdd(p.cnt + qi, (uint32_t)__popc(m));
                        pos = __shfl_sync(FPV_FULL_MASK, pos, leader) + (uint32_t)__popc(m & ((1u << lane) - 1u));
                        if (hit && pos < (uint32_t)HM_CAP)
                            p.cand[(size_t)qi * HM_CAP + pos] = ((uint64_t)f32_to_ordered((float)ham) << 32) | (uint64_t)(uint32_t)row;
                    }
                }
            }
        }
    }
    __syncwarp();
    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(HM_TMEM_COLS) : "memory");
    }
}

// Expanded query rows (u8 0/1, operand order of the expanders), popc(q & dimmask); resets the per-query state.
// Operand index of bit t (LSB = 0) of packed byte j:  (j / 16) * 128 + (((j % 16) / 4) * 8 + t) * 4 + j % 4.
__global__ void __launch_bounds__(256) ham_prep_kernel(const uint8_t* __restrict__ qbits, int nq, int nbytes, int dims,
                                                       uint8_t* __restrict__ bmat, int* __restrict__ pq, float* __restrict__ thr,
                                                       float* __restrict__ ebound, uint32_t* __restrict__ cnt,
                                                       uint32_t* __restrict__ flags) {
    const int qi = blockIdx.x;                                              // 0..31
    const int kexp = nbytes * 8;
    uint8_t* row = bmat + (size_t)qi * kexp;
    __shared__ int s_pop;
    if (threadIdx.x == 0) s_pop = 0;
    __syncthreads();
    const bool real = qi < nq, ones = qi == HM_ONES;
    int pop = 0;
    for (int j = threadIdx.x; j < nbytes; j += blockDim.x) {
        unsigned dm = 0xFFu;                                                // valid bits of this byte: dim = 8 j + (7 - t) < dims
        if (dims > 0) {
            const int left = dims - 8 * j;
            dm = left >= 8 ? 0xFFu : (left <= 0 ? 0u : (0xFFu << (8 - left)) & 0xFFu);
        }
        const unsigned b = ones ? dm : (real ? ((unsigned)qbits[(size_t)qi * nbytes + j] & dm) : 0u);
        pop += __popc(b);
        const int basee = (j >> 4) * 128 + ((j & 15) >> 2) * 32 + (j & 3);
#pragma unroll
        for (int t = 0; t < 8; ++t) row[basee + t * 4] = (uint8_t)((b >> t) & 1u);
    }
    if (pop) atomicAdd(&s_pop, pop);
    __syncthreads();
    if (threadIdx.x == 0 && qi < HM_QB + 1) {
        pq[qi] = s_pop;
        thr[qi] = real ? -INFINITY : INFINITY;
        ebound[qi] = 0.f;
        cnt[qi] = 0;
        flags[qi] = 0;
    }
}

// one CTA per query: the list (tightened once more by the caller) is sorted by (distance, row) and the first k are emitted
__global__ void __launch_bounds__(1024) ham_finish_kernel(const uint64_t* __restrict__ cand, const uint32_t* __restrict__ cnt,
                                                         uint32_t* __restrict__ flags, int k, int64_t id_base,
                                                         float* __restrict__ out_dist, int64_t* __restrict__ out_idx,
                                                         int32_t* __restrict__ out_count) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    uint64_t* keys = reinterpret_cast<uint64_t*>(sm_raw);                   // [HM_SORT_MAX]
    const int q = blockIdx.x;
    const uint32_t c_raw = cnt[q];
    if (c_raw > (uint32_t)HM_SORT_MAX || flags[q] != 0) {                  // a tie group of thousands, or an overflow
        if (threadIdx.x == 0) flags[q] = 1;                                 // -> the CUDA-core scan answers this query
        return;
    }
    const int c = (int)c_raw;
    int P2 = 2; while (P2 < c) P2 <<= 1;
    const uint64_t* mine = cand + (size_t)q * HM_CAP;
    for (int i = threadIdx.x; i < P2; i += blockDim.x) keys[i] = i < c ? mine[i] : FPV_KEY_MAX;
    __syncthreads();
    block_bitonic_sort(keys, P2);
    for (int i = threadIdx.x; i < k; i += blockDim.x) {
        const bool ok = i < c;
        const uint64_t key = ok ? keys[i] : FPV_KEY_MAX;
        out_dist[(size_t)q * k + i] = ok ? ordered_to_f32((uint32_t)(key >> 32)) : INFINITY;
        out_idx[(size_t)q * k + i] = ok ? id_base + (int64_t)(uint32_t)key : -1;
    }
    if (out_count && threadIdx.x == 0) out_count[q] = min(c, k);
}

struct HmPlan { int passes, kexp; size_t off_bmat, off_pq, off_thr, off_eb, off_cnt, off_flags, off_cand, off_scan, scan_bytes, total; };

size_t hamming_flagged_workspace(int64_t Q, int64_t N, int nbytes, int k);
int hamming_topk_flagged(const uint8_t* qbits, int64_t q, const uint8_t* codes, int64_t n, int nbytes, int dims, int k,
                         const uint32_t* mask_words, int64_t id_base, float* out_dist, int64_t* out_idx, int32_t* out_count,
                         const uint32_t* only_flagged, void* ws, size_t ws_bytes, cudaStream_t st);

static HmPlan plan_hm(int64_t Q, int64_t N, int nbytes, int k) {
    HmPlan pl{};
    pl.kexp = nbytes * 8;
    pl.passes = (int)((Q + HM_QB - 1) / HM_QB);
    const size_t P = (size_t)pl.passes;
    size_t o = 0;
    pl.off_bmat = o;  o += align_up(P * HM_N * pl.kexp, 1024);
    pl.off_pq = o;    o += align_up(P * 32 * 4, 256);
    pl.off_thr = o;   o += align_up(P * 32 * 4, 256);
    pl.off_eb = o;    o += align_up(P * 32 * 4, 256);
    pl.off_cnt = o;   o += align_up(P * 32 * 4, 256);
    pl.off_flags = o; o += align_up(P * 32 * 4, 256);
    pl.off_cand = o;  o += P * 32 * HM_CAP * 8;
    pl.off_scan = o;
    pl.scan_bytes = hamming_flagged_workspace(Q, N, nbytes, k);
    pl.total = o + pl.scan_bytes + (size_t)Q * 4 + 256;
    return pl;
}

typedef CUresult (*HmEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static int hm_make_map(CUtensorMap* m, const void* ptr, int64_t rows, int64_t cols, int box_rows) {
    static HmEncodeFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<HmEncodeFn>(p);
    }
    if (!fn) { set_error("hamming_mma: cuTensorMapEncodeTiled entry point not available"); return FPV_ERR_CUDA; }
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols};
    cuuint32_t box[2] = {128u, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { set_error("hamming_mma: cuTensorMapEncodeTiled failed with %d", (int)r); return FPV_ERR_CUDA; }
    return FPV_OK;
}

static int hm_raw_stages(int nbytes) { return std::min(HM_RAW_STAGES, (128 * 1024) / (HM_BM * nbytes)); }
static size_t hm_smem(int nbytes) {
    static_assert(16 * HM_RAW_STAGES + 16 * HM_A_STAGES + 8 + 32 <= 304, "barrier block");
    return (size_t)hm_raw_stages(nbytes) * HM_BM * nbytes + (size_t)(nbytes / 16) * HM_N * 128 + 320 + 64 * 4 + 64;
}

}  // namespace fpv

using namespace fpv;

extern "C" int fpv_hamming_mma_supported(int64_t q, int64_t n, int nbytes, int k) {
    return q >= 4 && n >= 65536 && n < (1ll << 31) && (nbytes == 128 || nbytes == 256) && k >= 1 && k <= FPV_MAX_K && 4 * k <= HM_CAP;
}

extern "C" size_t fpv_hamming_mma_workspace(int64_t q, int64_t n, int nbytes, int k) {
    if (q <= 0 || nbytes <= 0 || k <= 0) return 256;
    return plan_hm(q, n, nbytes, k).total;
}

static int hm_run(const uint8_t* qbits, int64_t q, const uint8_t* codes, int64_t n, int nbytes, int dims, int k,
                  const uint32_t* mask_words, int64_t id_base, float* out_dist, int64_t* out_idx, int32_t* out_count,
                  int* dump, void* ws, size_t ws_bytes, cudaStream_t st) {
    HmPlan pl = plan_hm(q, n, nbytes, k);
    if (!ws || ws_bytes < pl.total) { set_error("hamming_mma: workspace %zu < %zu", ws_bytes, pl.total); return FPV_ERR_WORKSPACE; }
    FPV_REQUIRE((reinterpret_cast<uintptr_t>(ws) & 255) == 0, "hamming_mma: workspace must be 256-byte aligned");
    char* w = static_cast<char*>(ws);
    const size_t smem = hm_smem(nbytes);
    static std::mutex attr_mutex;
    {
        std::unique_lock<std::mutex> lk(attr_mutex);
        FPV_CUDA(cudaFuncSetAttribute(ham_mma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FPV_CUDA(cudaFuncSetAttribute(tighten_kernel<HM_CAP>, cudaFuncAttributeMaxDynamicSharedMemorySize, HM_CAP * 8));
        FPV_CUDA(cudaFuncSetAttribute(ham_finish_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, HM_SORT_MAX * 8));
    }
    CUtensorMap tmRaw;
    int rc = hm_make_map(&tmRaw, codes, n, nbytes, HM_BM);
    if (rc != FPV_OK) return rc;
    const int64_t tiles_total = (n + HM_BM - 1) / HM_BM;
    for (int pass = 0; pass < pl.passes; ++pass) {
        const int64_t q0 = (int64_t)pass * HM_QB;
        const int nq = (int)std::min<int64_t>(HM_QB, q - q0);
        uint8_t* bmat = reinterpret_cast<uint8_t*>(w + pl.off_bmat) + (size_t)pass * HM_N * pl.kexp;
        int* pq = reinterpret_cast<int*>(w + pl.off_pq) + pass * 32;
        float* thr = reinterpret_cast<float*>(w + pl.off_thr) + pass * 32;
        float* eb = reinterpret_cast<float*>(w + pl.off_eb) + pass * 32;
        uint32_t* cnt = reinterpret_cast<uint32_t*>(w + pl.off_cnt) + pass * 32;
        uint32_t* flags = reinterpret_cast<uint32_t*>(w + pl.off_flags) + pass * 32;
        uint64_t* cand = reinterpret_cast<uint64_t*>(w + pl.off_cand) + (size_t)pass * 32 * HM_CAP;
        ham_prep_kernel<<<HM_N, 256, 0, st>>>(qbits + q0 * nbytes, nq, nbytes, dims, bmat, pq, thr, eb, cnt, flags);
        FPV_LAUNCH_CHECK();
        CUtensorMap tmB;
        rc = hm_make_map(&tmB, bmat, HM_N, pl.kexp, HM_N);
        if (rc != FPV_OK) return rc;
        HmParams p{};
        p.mask = mask_words; p.pq = pq; p.thr = thr; p.cnt = cnt; p.cand = cand; p.N = n; p.nq = nq; p.nbytes = nbytes;
        p.nkb = nbytes / 16; p.dump = dump; p.raw_stages = hm_raw_stages(nbytes);
        { const char* e = getenv("FPV_HAM_DEBUG"); p.debug = e ? atoi(e) : 0; }
        if (dump) {
            p.tile0 = 0; p.ntiles = (int)tiles_total;
            ham_mma_kernel<<<(unsigned)std::min<int64_t>(tiles_total, sm_count()), HM_THREADS, smem, st>>>(tmRaw, tmB, p);
            FPV_LAUNCH_CHECK();
            return FPV_OK;
        }
        // slabs in ROW ORDER (the tie rule of the thresholds depends on it): a dense first slab, then as many rows as keep
        // the expected number of strictly better rows (~ slab * k / rows_seen) within half the candidate slots
        int64_t done = 0, slab = std::max<int64_t>(8192, 4 * (int64_t)k) / HM_BM;
        const double growth = (double)(HM_CAP / 2) / (1.5 * k);
        while (done < tiles_total) {
            int64_t take = std::min<int64_t>(slab, tiles_total - done);
            if (tiles_total - done - take < take / 2) take = tiles_total - done;
            p.tile0 = (int)done; p.ntiles = (int)take;
            ham_mma_kernel<<<(unsigned)std::min<int64_t>(take, sm_count()), HM_THREADS, smem, st>>>(tmRaw, tmB, p);
            FPV_LAUNCH_CHECK();
            done += take;
            // keep the values <= k-th value; later rows must be strictly better (ties lose to the lower rows held)
            tighten_kernel<HM_CAP><<<(unsigned)nq, 256, HM_CAP * 8, st>>>(cand, cnt, thr, eb, flags, k, nullptr, -1.0f);
            FPV_LAUNCH_CHECK();
            slab = (int64_t)((double)done * growth);
            if (slab < 1) slab = 1;
        }
        ham_finish_kernel<<<(unsigned)nq, 1024, HM_SORT_MAX * 8, st>>>(cand, cnt, flags, k, id_base, out_dist + q0 * k, out_idx + q0 * k,
                                                                    out_count ? out_count + q0 : nullptr);
        FPV_LAUNCH_CHECK();
    }
    // gather the per-pass flags (32 per pass) into one [q] array for the gated CUDA-core fallback
    uint32_t* qflags = reinterpret_cast<uint32_t*>(w + pl.off_scan + pl.scan_bytes);
    for (int pass = 0; pass < pl.passes; ++pass) {
        const int64_t q0 = (int64_t)pass * HM_QB;
        const int nq = (int)std::min<int64_t>(HM_QB, q - q0);
        FPV_CUDA(cudaMemcpyAsync(qflags + q0, reinterpret_cast<uint32_t*>(w + pl.off_flags) + pass * 32, (size_t)nq * 4,
                                 cudaMemcpyDeviceToDevice, st));
    }
    return hamming_topk_flagged(qbits, q, codes, n, nbytes, dims, k, mask_words, id_base, out_dist, out_idx, out_count, qflags,
                                w + pl.off_scan, pl.scan_bytes, st);
}

// Batched Hamming top-k on the int8 tensor cores: qbits [q][nbytes] packed query bits, codes [n][nbytes].  Same results
// as fpv_hamming_topk (integer distances as float32, ties by lowest row).  Requires fpv_hamming_mma_supported.
extern "C" int fpv_hamming_mma_topk(const uint8_t* qbits, int64_t q, const uint8_t* codes, int64_t n, int nbytes, int dims, int k,
                                    const uint32_t* mask_words, int64_t id_base, float* out_dist, int64_t* out_idx,
                                    int32_t* out_count, void* ws, size_t ws_bytes, void* stream) {
    FPV_REQUIRE(q >= 1 && q <= 65535, "hamming_mma: q=%lld outside [1,65535]", (long long)q);
    FPV_REQUIRE(fpv_hamming_mma_supported(std::max<int64_t>(q, 4), n, nbytes, k), "hamming_mma: unsupported shape n=%lld nbytes=%d k=%d",
                (long long)n, nbytes, k);
    FPV_REQUIRE(dims >= 0 && dims <= nbytes * 8, "hamming_mma: dims=%d exceeds code width", dims);
    FPV_REQUIRE(qbits && codes && out_dist && out_idx, "hamming_mma: null pointer");
    FPV_REQUIRE((reinterpret_cast<uintptr_t>(codes) & 15) == 0, "hamming_mma: codes must be 16-byte aligned");
    return hm_run(qbits, q, codes, n, nbytes, dims, k, mask_words, id_base, out_dist, out_idx, out_count, nullptr, ws, ws_bytes,
                  (cudaStream_t)stream);
}

// Test hook: the raw s32 accumulators of the first (up to 31) queries + the all-ones row, out [32][n]:
// out[qi][row] = -popc(x_row & q_qi & dimmask), out[31][row] = -popc(x_row & dimmask).
extern "C" int fpv_hamming_mma_dots(const uint8_t* qbits, int64_t q, const uint8_t* codes, int64_t n, int nbytes, int dims,
                                    int32_t* out, void* ws, size_t ws_bytes, void* stream) {
    FPV_REQUIRE(q >= 1 && q <= HM_QB && n >= 1 && n < (1ll << 31) && (nbytes == 128 || nbytes == 256), "hamming_mma_dots: bad shape");
    FPV_REQUIRE(qbits && codes && out, "hamming_mma_dots: null pointer");
    return hm_run(qbits, q, codes, n, nbytes, dims, 1, nullptr, 0, nullptr, nullptr, nullptr, out, ws, ws_bytes, (cudaStream_t)stream);
}
