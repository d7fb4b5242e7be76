/* fpv_b200.h — C-ABI of the B200-native search hot path of FastPyVectorDB.
 *
 * The reference (jcolano/fastpyvectordb) is pure Python/NumPy and has no FFI of its own; the boundary it
 * offers for this path is the Python call surface of parallel_search.py / quantization.py.
 * Each entry point below replaces the NumPy body of the reference function it cites and is what a ctypes
 * binding inside those functions would call (INTEGRATION.md shows the stub).
 *
 * Conventions
 *  - every pointer is a DEVICE pointer on the current CUDA device (the Python host layer owns all memory
 *    as torch tensors); nothing is allocated or freed across this ABI, scratch comes in as (ws, ws_bytes)
 *    sized by the matching *_workspace() call;
 *  - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises;
 *  - return value: FPV_OK or an FPV_ERR_* code, text via fpv_last_error() (thread local); never throws;
 *  - results are ordered by (distance ascending, row index ascending) — the reference's order among equal
 *    distances is arbitrary (np.argpartition), ours is deterministic and shard-count invariant;
 *  - out_dist is [Q][k] float32, out_idx is [Q][k] int64 (= id_base + local row), rows that do not exist
 *    (k > number of permitted rows) are (inf, -1); out_count (may be NULL) is [Q] int32 valid entries;
 *  - mask_words (may be NULL): bit i of a little-endian uint32 word array, 1 = row i is permitted
 *    (the filter_mask of parallel_search.py:212-217; pack with np.packbits(bitorder="little")).
 */
#ifndef FPV_B200_H
#define FPV_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FPV_ABI_VERSION 1

#if defined(__GNUC__)
#define FPV_API __attribute__((visibility("default")))
#else
#define FPV_API
#endif

#define FPV_OK 0
#define FPV_ERR_INVALID 1      /* bad argument (shape, alignment, k, metric ...) */
#define FPV_ERR_CUDA 2         /* a CUDA runtime call failed */
#define FPV_ERR_WORKSPACE 3    /* ws_bytes smaller than *_workspace() */
#define FPV_ERR_UNSUPPORTED 4  /* valid request this build has no kernel for */

#define FPV_METRIC_COSINE 0    /* 1 - cos, eps 1e-10 on both norms (parallel_search.py:119-126) */
#define FPV_METRIC_L2 1        /* sqrt(max(q.q + v.v - 2 q.v, 0))   (parallel_search.py:127-132) */
#define FPV_METRIC_IP 2        /* -q.v                               (parallel_search.py:133-134) */
/* sqrt(sum_j (v_j - q_j)^2), the explicit-difference form of _compute_distances_chunk (parallel_search.py:92-95) and of
 * Collection.brute_force_search (vectordb_optimized.py:679-680): exactly 0 for a duplicate of the query, where the
 * expanded form above carries ~3e-4 of cancellation noise.  Accepted by fpv_distances_f32 and fpv_rerank_f32 only. */
#define FPV_METRIC_L2_DIFF 3

#define FPV_SQ_L2 0            /* quantization.py:217-236 */
#define FPV_SQ_DOT 1           /* quantization.py:239-251 */
#define FPV_SQ_COSINE 2        /* quantization.py:154-174 */

#define FPV_MAX_K 1024

FPV_API int fpv_abi_version(void);
FPV_API const char* fpv_last_error(void);
/* number of kernels this library has launched in this process (bench.py reports it as gpu_launches) */
FPV_API long long fpv_launch_count(void);

/* ---- float32 brute force (parallel_search.py) ------------------------------------------------------- */

/* row_sq[i] = sum_j db[i][j]^2 — the np.einsum('ij,ij->i') of parallel_search.py:123,130,275,285, computed
 * once per resident database instead of once per call. */
FPV_API int fpv_row_sqnorm_f32(const float* db, int64_t n, int d, int64_t ld, float* row_sq, void* stream);

/* Exact fp32 streaming scan + fused top-k: the body of ParallelSearchEngine.search_parallel /
 * search_chunked_parallel (parallel_search.py:209-244, 326-368) and, for small batches, of
 * search_batch_parallel (:259-311).  queries [Q][d] row major, db [n][ld].  row_sq may be NULL (computed on
 * the fly).  Distances follow _compute_distances_vectorized (:105-134). */
FPV_API size_t fpv_scan_f32_workspace(int64_t q, int64_t n, int d, int k);
FPV_API int fpv_scan_f32_topk(const float* queries, int64_t q, const float* db, int64_t n, int d, int64_t ld,
                      int metric, int k, const uint32_t* mask_words, const float* row_sq, int64_t id_base,
                      float* out_dist, int64_t* out_idx, int32_t* out_count,
                      void* ws, size_t ws_bytes, void* stream);

/* Large-batch exact search on the tensor cores (tcgen05 + TMEM + TMA), the GEMM regime of
 * ParallelSearchEngine.search_batch_parallel (parallel_search.py:246-311): approximate TF32 (kind 0, straight from
 * the fp32 rows) or BF16 (kind 1, over the db_lowp shadow copy made by fpv_to_bf16) filter pass with a fused
 * threshold epilogue, then a certified exact fp32 re-rank; queries whose certificate fails are recomputed by the
 * exact scan on the device.  Results are identical in kind to fpv_scan_f32_topk (exact fp32, (distance, index) order).
 * aux: per-row 1/(sqrt(row_sq)+1e-10) for cosine, row_sq for l2, NULL for ip; vmax = max row norm.
 * db_err_abs / db_err_rel (kind 1 only, 0 = unknown): max over rows of |v - bf16(v)| and of |v - bf16(v)| / |v|,
 * measured when the shadow copy is made; with them the filter's error bound is the Cauchy-Schwarz bound on the
 * measured rounding errors instead of the worst case per element (fewer rows reach the exact re-rank).
 * Requires 1 <= k <= 256, d % 4 == 0 (TF32) or d % 8 == 0 (BF16), 16-byte aligned database. */
FPV_API size_t fpv_gemm_topk_workspace(int64_t q, int64_t n, int d, int k, int kind);
FPV_API int fpv_gemm_topk_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                      int metric, int k, int kind, const float* row_sq, const float* aux, float vmax,
                      float db_err_abs, float db_err_rel, const uint32_t* mask_words, int64_t id_base,
                      float* out_dist, int64_t* out_idx, int32_t* out_count, void* ws, size_t ws_bytes, void* stream);
/* The same search over a ROW-SHARDED database (one shard per GPU): the chunk -> local top-k -> merge structure of
 * search_chunked_parallel (parallel_search.py:335-363) with GPUs as chunks, split around the one exchange it needs so
 * that per-query work is not repeated on every shard.
 *   phase 1  fpv_gemm_filter_sharded_f32: tensor-core filter of this shard's rows; leaves the candidate lists in ws
 *            and writes the shard's k best APPROXIMATE values per query to approx_out [q][k] (uint32, order
 *            preserving, padded with +inf).  vmax / db_err_* must be the maxima over ALL shards.
 *   --       the caller all-gathers approx_out over the shards (q*k*4 bytes per shard)
 *   phase 2  fpv_gemm_finish_sharded_f32 (same ws): approx_all [shards][q][k]; selects the k-th best approximate
 *            value of the whole job, re-ranks only this shard's rows below (that + 2E) in exact fp32 and writes the
 *            shard's (distance, global id) lists [q][k] padded with (+inf, -1); shards * k <= 4096.
 *   --       fpv_pack_topk, all-gather, fpv_merge_packed give every rank the exact global top-k. */
FPV_API int fpv_gemm_filter_sharded_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                      int metric, int k, int kind, const float* row_sq, const float* aux, float vmax,
                      float db_err_abs, float db_err_rel, const uint32_t* mask_words, uint32_t* approx_out,
                      void* ws, size_t ws_bytes, void* stream);
/* Phase 1 split once more around a small exchange (every shard >= 70K rows): the first half runs the sampling slab and
 * writes the shard's k best GROUP values to sample_out [q][k] (encoding of approx_out); the caller all-gathers them;
 * the second half (same ws, same arguments) takes sample_all [shards][q][k], makes the k-th best group value of the
 * WHOLE job every shard's first threshold and runs the filtering slabs -> approx_out.  With its own sample only, a shard
 * starts from a threshold `shards` times looser than the job's and appends `shards` times more candidates per slab.
 * wait_flags / epoch: sample_all is a peer-memory gather area (fpv_peer_put); NULL / 0 for a plain buffer. */
FPV_API int fpv_gemm_sample_sharded_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                      int metric, int k, int kind, const float* row_sq, const float* aux, float vmax,
                      float db_err_abs, float db_err_rel, const uint32_t* mask_words, uint32_t* sample_out,
                      void* ws, size_t ws_bytes, void* stream);
FPV_API int fpv_gemm_slabs_sharded_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                      int metric, int k, int kind, const float* row_sq, const float* aux, float vmax,
                      float db_err_abs, float db_err_rel, const uint32_t* mask_words, const uint32_t* sample_all, int shards,
                      const uint32_t* wait_flags, uint32_t epoch, uint32_t* approx_out, void* ws, size_t ws_bytes, void* stream);
FPV_API int fpv_gemm_finish_sharded_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                      int metric, int k, int kind, const float* row_sq, const uint32_t* mask_words, int64_t id_base,
                      const uint32_t* approx_all, int shards, float* out_dist, int64_t* out_idx, int32_t* out_count,
                      void* ws, size_t ws_bytes, void* stream);
/* Byte offset inside ws of the uint32 [q] array that is 1 for every query of the last fpv_gemm_topk_f32 call that
 * failed its certificate and was recomputed by the exact scan (diagnostics / tests). */
FPV_API size_t fpv_gemm_topk_flags_offset(int64_t q, int64_t n, int d, int k, int kind);
/* Measurement hooks (bench.py's roofline): with enable != 0 every fpv_gemm_topk_f32 call records CUDA events around
 * its tensor-core filter launches on the call's stream.  fpv_gemm_profile_read waits for the last recorded call and
 * returns the summed duration (ms) and the number of those launches. */
FPV_API int fpv_gemm_profile(int enable);
FPV_API int fpv_gemm_profile_read(float* filter_ms, int* filter_launches);
/* fp32 -> bf16 (round to nearest even) shadow copy used by kind 1 above. */
FPV_API int fpv_to_bf16(const float* src, void* dst, int64_t n, void* stream);

/* Full 1 x N distance rows (no selection): _compute_distances_vectorized (parallel_search.py:105-134) for
 * each of the Q queries; out_all is [Q][n] float32. */
FPV_API int fpv_distances_f32(const float* queries, int64_t q, const float* db, int64_t n, int d, int64_t ld,
                      int metric, const float* row_sq, float* out_all, void* ws, size_t ws_bytes, void* stream);

/* Exact fp32 re-rank of gathered candidates: the pattern of ParallelCollection.search_hybrid
 * (parallel_search.py:919-934) and the "exact re-rank of top-100 candidates" step of the quantized scans.
 * cand_idx [Q][c] int64 local rows (-1 = empty slot).  Output k <= c rows per query. */
FPV_API int fpv_rerank_f32(const float* queries, int64_t q, const float* db, int64_t n, int d, int64_t ld, int metric,
                   const int64_t* cand_idx, int c, int k, const float* row_sq, int64_t id_base,
                   float* out_dist, int64_t* out_idx, int32_t* out_count, void* stream);

/* k-way merge of per-shard sorted lists: _merge_top_k (parallel_search.py:137-156).  dist/idx are
 * [shards][Q][k_in]; idx < 0 marks an empty slot. */
FPV_API int fpv_merge_topk(const float* dist, const int64_t* idx, int shards, int64_t q, int k_in, int k_out,
                   float* out_dist, int64_t* out_idx, int32_t* out_count, void* stream);

/* The same merge for the multi-GPU exchange (the thread-pool + _merge_top_k of search_chunked_parallel,
 * parallel_search.py:338-363, with GPUs as chunks): fpv_pack_topk turns a rank's local (distance, global id) lists into
 * the 8-byte wire format  ordered(distance) << 32 | (id - id_base)  padded to k_pad columns with 0xFF..FF; after the
 * all-gather fpv_merge_packed merges the [shards][Q][k_in] keys, adding shard_bases[s] (device int64 [shards]) back.
 * Every packed list must be sorted by (distance, row) with its empty slots last (what fpv_pack_topk makes of any
 * top-k output of this library): the merge ranks every entry by binary search instead of sorting. */
FPV_API int fpv_pack_topk(const float* dist, const int64_t* idx, int64_t q, int k_in, int k_pad, int64_t id_base,
                  uint64_t* out_packed, void* stream);
FPV_API int fpv_merge_packed(const uint64_t* packed, const int64_t* shard_bases, int shards, int64_t q, int k_in, int k_out,
                     float* out_dist, int64_t* out_idx, int32_t* out_count, void* stream);

/* ---- exchange over NVLink peer memory (csrc/fpv_peer.cu): the all-gathers of the row-sharded search without a
 * collective library.  Every rank allocates one IPC-exportable region (fpv_peer_alloc), the 64-byte handles are
 * exchanged once by the host and opened (fpv_peer_open).  fpv_peer_put stores a buffer into slot `rank` of EVERY
 * region in `peers` (device array of region base pointers) with P2P stores and then sets flag word `rank` of every
 * region to `epoch`; the *_peer consumers below wait on their LOCAL flag words inside the kernel and read the
 * gathered data from local memory.  Epochs must increase by one per exchange and alternate between two data areas. */
FPV_API int fpv_peer_alloc(size_t bytes, void** out_ptr, unsigned char* handle64);
FPV_API int fpv_peer_open(const unsigned char* handle64, void** out_ptr);
FPV_API int fpv_peer_close(void* ptr);
FPV_API int fpv_peer_free(void* ptr);
FPV_API int fpv_peer_put(const void* src, size_t nbytes, void* const* peers, int shards, int rank, size_t data_off,
                 size_t slot_bytes, size_t flag_off, uint32_t epoch, uint32_t* done_counter, void* stream);
FPV_API int fpv_merge_packed_peer(const uint64_t* packed, const int64_t* shard_bases, int shards, int64_t q, int k_in, int k_out,
                          const uint32_t* wait_flags, uint32_t epoch, float* out_dist, int64_t* out_idx,
                          int32_t* out_count, void* stream);
FPV_API int fpv_gemm_finish_sharded_peer_f32(const float* queries, int64_t q, const float* db, const void* db_lowp, int64_t n, int d,
                          int metric, int k, int kind, const float* row_sq, const uint32_t* mask_words, int64_t id_base,
                          const uint32_t* approx_all, int shards, const uint32_t* wait_flags, uint32_t epoch,
                          float* out_dist, int64_t* out_idx, int32_t* out_count, void* ws, size_t ws_bytes, void* stream);

/* ---- binary quantizer (quantization.py:282-407) -------------------------------------------------------- */

/* BinaryQuantizer.encode (:336-350): bit = v > thr, packed MSB-first into ceil(d/8) bytes per row. */
FPV_API int fpv_bq_encode(const float* vectors, int64_t n, int d, int64_t ld, const float* thresholds,
                  uint8_t* out_codes, void* stream);

/* BinaryQuantizer.hamming_distances + search (:356-394): popcount(q XOR row) over the first `dims` bits
 * (dims <= 0: all nbytes*8 bits).  qbits [Q][nbytes], codes [n][nbytes].  k == 0 skips the selection;
 * out_all (may be NULL) receives the full [Q][n] float32 distance rows like hamming_distances() returns. */
FPV_API size_t fpv_hamming_workspace(int64_t q, int64_t n, int nbytes, int k);
FPV_API int fpv_hamming_topk(const uint8_t* qbits, int64_t q, const uint8_t* codes, int64_t n, int nbytes, int dims,
                     int k, const uint32_t* mask_words, int64_t id_base,
                     float* out_dist, int64_t* out_idx, int32_t* out_count, float* out_all,
                     void* ws, size_t ws_bytes, void* stream);

/* The Hamming scan for query BATCHES on the int8 tensor cores: one pass over the packed codes serves 31
 * queries.  popc(x ^ q) = popc(x) + popc(q) - 2 popc(x & q); the packed bits are expanded to int8 operand bytes inside
 * the SM, straight into tensor memory (tcgen05.st), and popc(x & q) is an exact s8 x u8 -> s32 tcgen05.mma with the A
 * operand read from TMEM; the epilogue filters the integer distances against per-query thresholds that a radix select
 * tightens between row slabs.  Same results as fpv_hamming_topk (ties by lowest row); queries whose tie group does not
 * fit are recomputed by that scan on the device.  Requires nbytes in {128, 256}, n >= 65536, k <= 1024. */
FPV_API int fpv_hamming_mma_supported(int64_t q, int64_t n, int nbytes, int k);
FPV_API size_t fpv_hamming_mma_workspace(int64_t q, int64_t n, int nbytes, int k);
FPV_API int fpv_hamming_mma_topk(const uint8_t* qbits, int64_t q, const uint8_t* codes, int64_t n, int nbytes, int dims, int k,
                         const uint32_t* mask_words, int64_t id_base, float* out_dist, int64_t* out_idx,
                         int32_t* out_count, void* ws, size_t ws_bytes, void* stream);
/* Test hook: raw s32 accumulators, out [32][n]: out[i][row] = -popc(x_row & q_i & dimmask) for i < q (q <= 31),
 * out[31][row] = -popc(x_row & dimmask). */
FPV_API int fpv_hamming_mma_dots(const uint8_t* qbits, int64_t q, const uint8_t* codes, int64_t n, int nbytes, int dims,
                         int32_t* out, void* ws, size_t ws_bytes, void* stream);

/* ---- product quantizer (quantization.py:414-615) -------------------------------------------------------- */

/* ProductQuantizer.encode (:520-539): first-min argmin over centroids per subspace. */
FPV_API int fpv_pq_encode(const float* vectors, int64_t n, int d, int64_t ld, const float* codebooks, int m, int kc,
                  uint8_t* out_codes, void* stream);
/* ProductQuantizer.build_lookup_table (:551-562) for Q queries: lut [Q][m][kc]. */
FPV_API int fpv_pq_build_lut(const float* codebooks, int m, int kc, int dsub, const float* queries, int64_t q,
                     float* lut, void* stream);
/* ProductQuantizer.distances_with_table + search (:571-597): sqrt(sum_m lut[m][code]) accumulated
 * sequentially in m (bit-identical to the reference), in-kernel bitmask, fused top-k. */
FPV_API size_t fpv_pq_adc_workspace(int64_t q, int64_t n, int m, int kc, int k);
FPV_API int fpv_pq_adc_topk(const float* lut, int64_t q, const uint8_t* codes, int64_t n, int m, int kc,
                    int k, const uint32_t* mask_words, int64_t id_base,
                    float* out_dist, int64_t* out_idx, int32_t* out_count, float* out_all,
                    void* ws, size_t ws_bytes, void* stream);

/* Conflict-free variant of the ADC scan for ProductQuantizer.search (:580-597).  fpv_pq_pack rewrites the [n][m] code
 * matrix once (m % 16 == 0) into the lane-rotated order the kernel reads; the scan then does bank-conflict-free
 * table lookups.  The per-row sum order differs from the reference's m = 0..M-1, so distances agree to fp32 rounding
 * rather than bit for bit (the exact-order kernel above stays behind distances_with_table()).
 * fpv_pq_adc_packed_workspace returns 0 when the shape is not supported (use fpv_pq_adc_topk then). */
FPV_API int fpv_pq_pack(const uint8_t* codes, int64_t n, int m, uint8_t* out_packed, void* stream);
FPV_API size_t fpv_pq_adc_packed_workspace(int64_t q, int64_t n, int m, int kc, int k);
FPV_API int fpv_pq_adc_packed_topk(const float* lut, int64_t q, const uint8_t* packed, int64_t n, int m, int kc,
                           int k, const uint32_t* mask_words, int64_t id_base,
                           float* out_dist, int64_t* out_idx, int32_t* out_count,
                           void* ws, size_t ws_bytes, void* stream);

/* ---- scalar (uint8) quantizer (quantization.py:64-276) ----------------------------------------------------- */

/* ScalarQuantizer.encode (:118-126): clip((v - min) / scale * 255, 0, 255) truncated to uint8. */
FPV_API int fpv_sq_encode(const float* vectors, int64_t n, int d, int64_t ld, const float* min_vals,
                  const float* scale, uint8_t* out_codes, void* stream);
/* ScalarQuantizer.distances_l2 / distances_dot / distances_cosine (:145-181) on already-encoded queries
 * qcodes [Q][d] (the reference re-quantises the query first, :151), fused top-k and/or full rows. */
FPV_API size_t fpv_sq_workspace(int64_t q, int64_t n, int d, int k);
FPV_API int fpv_sq_topk(int kind, const uint8_t* qcodes, int64_t q, const uint8_t* codes, int64_t n, int d,
                const float* min_vals, const float* scale, int k, const uint32_t* mask_words, int64_t id_base,
                float* out_dist, int64_t* out_idx, int32_t* out_count, float* out_all,
                void* ws, size_t ws_bytes, void* stream);

/* The L2 scan for query BATCHES on the int8 tensor cores (tcgen05.mma kind::i8): ScalarQuantizer.distances_l2
 * (quantization.py:145-152, 217-236) + the caller's top-k for up to 16 queries per pass over the codes.  The weighted
 * distance is expanded as  d^2 = A_q + C_row - 2 sum_j a_j b_j;  the cross term is three EXACT u8 x u8 -> s32 limb dot
 * products per (row, query) on the tensor cores (a_j in 24-bit fixed point), C_row comes from fpv_sq_row_term (once per
 * code matrix), and a certified window of rows is re-scored with the arithmetic of fpv_sq_topk, so the results equal
 * fpv_sq_topk(FPV_SQ_L2) bit for bit (queries whose window does not fit are recomputed by that scan on the device).
 * Requires fpv_sq_mma_supported(n, d, k): n >= 65536, d % 16 == 0, d <= 1024, k <= 1024, 16-byte aligned codes. */
FPV_API int fpv_sq_row_term(const uint8_t* codes, int64_t n, int d, const float* scale, float* row_term, float* row_term_max,
                    void* stream);
FPV_API int fpv_sq_mma_supported(int64_t n, int d, int k);
FPV_API size_t fpv_sq_mma_workspace(int64_t q, int64_t n, int d, int k);
FPV_API int fpv_sq_l2_mma_topk(const uint8_t* qcodes, int64_t q, const uint8_t* codes, int64_t n, int d, const float* min_vals,
                       const float* scale, const float* row_term, const float* row_term_max, int k,
                       const uint32_t* mask_words, int64_t id_base, float* out_dist, int64_t* out_idx, int32_t* out_count,
                       void* ws, size_t ws_bytes, void* stream);
/* The same for ScalarQuantizer.distances_dot (quantization.py:176-181, 239-251; kind FPV_SQ_DOT) and distances_cosine
 * (:154-174; FPV_SQ_COSINE).  The signed weights a_j = (scale_j/255) w_j (w = decoded, for cosine normalised, query) are
 * shifted by c = max |a_j| so that the limbs stay unsigned:  sum_j a_j b_j = alpha T - c R_row  with T from the three
 * exact limb dots and R_row = sum_j b_j;  cosine multiplies by invn_row = 1 / (|decode(row)| + 1e-8).  R_row, invn_row
 * and their maxima (two device floats) come from fpv_sq_row_terms_dc, once per code matrix.  Results equal
 * fpv_sq_topk(kind) bit for bit; same shape limits and workspace as the L2 entry. */
FPV_API int fpv_sq_row_terms_dc(const uint8_t* codes, int64_t n, int d, const float* min_vals, const float* scale,
                        float* row_sum, float* row_invn, float* maxima, void* stream);
FPV_API int fpv_sq_dc_mma_topk(int kind, const uint8_t* qcodes, int64_t q, const uint8_t* codes, int64_t n, int d,
                       const float* min_vals, const float* scale, const float* row_sum, const float* row_invn,
                       const float* maxima, int k, const uint32_t* mask_words, int64_t id_base, float* out_dist,
                       int64_t* out_idx, int32_t* out_count, void* ws, size_t ws_bytes, void* stream);
/* Byte offset inside ws of the uint32 [q] flags of the last fpv_sq_*_mma_topk call (1 = answered by the SIMT scan). */
FPV_API size_t fpv_sq_mma_flags_offset(int64_t q, int64_t n, int d, int k);
/* Test hook: the three limb dot products of ONE query against every row, from the tensor cores (out_mma [3][n] int32)
 * and from a CUDA-core loop (out_simt), plus the limb rows themselves (limbs_out [3][d rounded up to 128]); ws as
 * fpv_sq_mma_workspace(1, n, d, 1). */
FPV_API int fpv_sq_mma_limb_dots(const uint8_t* qcodes, const uint8_t* codes, int64_t n, int d, const float* scale,
                         uint8_t* limbs_out, int32_t* out_mma, int32_t* out_simt, void* ws, size_t ws_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* FPV_B200_H */
