/* sky_search.h -- C ABI of the B200-native similarity-search engine (libskysearch.so).
 *
 * Drop-in boundary for ONE path of teaghan/sky_embeddings: utils/similarity.py as driven by
 * similarity_search.py / sky_sim_search.py.  The reference has no FFI layer (it is pure Python,
 * SURVEY.md section 8(b)); each entry point below cites the reference code it replaces, and
 * INTEGRATION.md shows the ctypes stub a maintainer of the reference would add.
 *
 * Conventions
 *  - every function returns int: 0 = ok, negative = error (text via sky_last_error());
 *  - no exceptions, no torch / C++ types across the ABI;
 *  - every pointer is a DEVICE pointer unless its name starts with h_;
 *  - every call takes a cudaStream_t (passed as void*) and is asynchronous on it, except the
 *    *_host variants, which copy host<->device and synchronise the stream before returning;
 *  - a bank handle is bound to one device; calls on one handle must be externally serialised.
 *  - there is NO CPU fallback: without a CUDA device every compute call fails with SKY_ERR_CUDA.
 */
#ifndef SKY_SEARCH_H_
#define SKY_SEARCH_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKY_ABI_VERSION 1

#if defined(__GNUC__)
#define SKY_API __attribute__((visibility("default")))
#else
#define SKY_API
#endif

/* status codes */
#define SKY_OK            0
#define SKY_ERR_ARG      -1   /* invalid argument (incl. unknown metric: the reference raises
                                 UnboundLocalError there, utils/similarity.py:250-259) */
#define SKY_ERR_CUDA     -2   /* CUDA runtime / driver error */
#define SKY_ERR_STATE    -3   /* call order violated (e.g. search before finalize) */
#define SKY_ERR_NOMEM    -4
#define SKY_ERR_UNSUPPORTED -5

/* element types of bank storage / upload sources */
#define SKY_F32   0
#define SKY_BF16  1

/* metrics: utils/similarity.py:149-212; selection order :20-29 (cosine descending, others ascending) */
#define SKY_COSINE 0
#define SKY_MSE    1
#define SKY_MAE    2

/* combine over the L tokens of an item: utils/similarity.py:262-267 */
#define SKY_MEAN 0
#define SKY_MIN  1
#define SKY_MAX  2

/* token selection applied to encoder output [B, tokens, D]: utils/similarity.py:55-63, :87-95 */
#define SKY_TOK_ALL      0   /* rows are used as given (tokens == L) */
#define SKY_TOK_CLS      1   /* keep token 0                      (cls_token=True)  -> L = 1 */
#define SKY_TOK_PATCHES  2   /* drop the first num_extra_tokens   (max_pool=False)  -> L = tokens - extra */
#define SKY_TOK_MAXPOOL  3   /* drop extras, max over patches     (max_pool=True)   -> L = 1 */

/* kernel path */
#define SKY_PATH_AUTO    0
#define SKY_PATH_SIMT    1   /* HBM-streaming CUDA-core kernel, bulk-copy staged (any dtype / metric / weights) */
#define SKY_PATH_TENSOR  2   /* tcgen05 contraction (bf16 bank, cosine / MSE) */
#define SKY_PATH_BATCH   4   /* GEMM-shaped tcgen05 kernel for large query batches (phased bounds) */
#define SKY_PATH_GENERIC 3   /* generic CUDA-core kernel (any L, any D); also the fallback of SIMT */

typedef struct sky_bank sky_bank_t;

/* Text of the last error on the calling thread. */
SKY_API const char* sky_last_error(void);
SKY_API int sky_abi_version(void);

/* ---- bank: the device-resident, normalised embedding bank -------------------------------------
 * Replaces the per-batch loader + token select + first-batch normalisation of mae_simsearch
 * (utils/similarity.py:71-102): rows are normalised ONCE at upload instead of once per search.   */

/* n_items bank items of L tokens x D features, stored as `dtype` tiles in HBM. */
SKY_API int sky_bank_create(sky_bank_t** bank, int device, int64_t n_items, int L, int D, int dtype);
SKY_API int sky_bank_destroy(sky_bank_t* bank);

/* mean / unbiased std over (items, tokens) of the FIRST batch, after token selection
 * (utils/similarity.py:98-100).  src: [n_items, src_tokens, D] of src_dtype. */
SKY_API int sky_bank_fit_norm(sky_bank_t* bank, const void* src, int src_dtype, int64_t n_items,
                      int src_tokens, int token_mode, int num_extra_tokens, void* stream);
/* explicit statistics (mu[D], sigma[D]; the +1e-8 of :101-102 is added inside). */
SKY_API int sky_bank_set_norm(sky_bank_t* bank, const float* mu, const float* sigma, void* stream);
SKY_API int sky_bank_get_norm(const sky_bank_t* bank, float* mu, float* sigma, void* stream);

/* token-select, normalise ((x-mu)/(sigma+1e-8), utils/similarity.py:102), cast and store items
 * [item0, item0+n_items); also accumulates the per-row squared norms the tensor path needs.
 * Without fit_norm/set_norm rows are stored un-normalised (compute_similarity semantics). */
SKY_API int sky_bank_upload(sky_bank_t* bank, const void* src, int src_dtype, int64_t item0, int64_t n_items,
                    int src_tokens, int token_mode, int num_extra_tokens, void* stream);
SKY_API int sky_bank_finalize(sky_bank_t* bank, void* stream);
/* shrink / regrow the ACTIVE item count within the capacity given at creation (a streaming caller
 * reuses one bank for batches of varying size; get_train_samples, utils/similarity.py:4-14). */
SKY_API int sky_bank_resize(sky_bank_t* bank, int64_t n_items);
/* stored (normalised, rounded) rows back as f32 [n_items, L, D] -- tests / snapshots. */
SKY_API int sky_bank_download(const sky_bank_t* bank, int64_t item0, int64_t n_items, float* dst, void* stream);
SKY_API int sky_bank_info(const sky_bank_t* bank, int64_t* n_items, int* L, int* D, int* dtype);

/* ---- query preparation ----------------------------------------------------------------------
 * determine_target_features (utils/similarity.py:134-147) fused with the target half of the
 * first-batch normalisation (:101): targets [T_rows, D] f32 (token-selected, flattened) ->
 * t[D] = mean rows, w[D] = 1/std^2 normalised to sum 1 (ones if use_weights == 0, :246-247).
 * bank may be NULL (no normalisation). */
SKY_API int sky_query_from_targets(const sky_bank_t* bank, const float* targets, int64_t T_rows, int D,
                           int use_weights, float* t_out, float* w_out, void* stream);

/* ---- search -----------------------------------------------------------------------------------
 * compute_similarity + update_best_scores over the whole bank (utils/similarity.py:105-110,
 * :214-268, :18-35) for Q queries at once, with no score matrix written to HBM.
 *  t[Q,D], w[Q,D] f32 (w NULL = ones, i.e. use_weights=False); n_top_sims 0 = None (:257-259).
 *  out_scores[Q,k] f32 best-first, out_idx[Q,k] i64 bank item index (+ idx_offset);
 *  when the bank has fewer than k items the tail is padded like the reference's initial fill
 *  (:66: -inf for cosine, +inf otherwise) with index -1.  NaN scores rank as the largest value.
 *  Precision: SKY_PATH_SIMT / GENERIC keep the queries in fp32 (fp32 bank: 1e-5 relative to the reference).  The tensor
 *  paths (TENSOR / BATCH; AUTO picks them only for Q > 4 on a bf16 bank) round the query operands to bf16 (t; with
 *  weights w*t and w) and derive the per-query constants from the rounded operands, so scores are true cosines /
 *  squared distances of the rounded vectors: within ~1e-3 of a typical score of the bank (measured bounds: DESIGN.md
 *  section 5).  One search at a time per bank handle: the candidate state lives in the handle's scratch. */
SKY_API int sky_search(sky_bank_t* bank, const float* t, const float* w, int Q, int metric, int combine,
               int n_top_sims, int k, int64_t idx_offset, float* out_scores, int64_t* out_idx,
               int path, void* stream);
/* same, host buffers in and out (pinned or pageable); H2D + D2H happen inside the call. */
SKY_API int sky_search_host(sky_bank_t* bank, const float* h_t, const float* h_w, int Q, int metric,
                    int combine, int n_top_sims, int k, int64_t idx_offset, float* h_out_scores,
                    int64_t* h_out_idx, int path, void* stream);

/* compute_similarity only (utils/similarity.py:214-268): scores of items [item0, item0+n_items)
 * for Q queries -> out_scores[Q, n_items] f32.  Used by the Python mirror of compute_similarity. */
SKY_API int sky_score(sky_bank_t* bank, const float* t, const float* w, int Q, int metric, int combine,
              int n_top_sims, int64_t item0, int64_t n_items, float* out_scores, void* stream);

/* ---- pixel-space masked-MSE search (BASELINE config 5) -----------------------------------------
 * The reference has no pixel-space search; SURVEY.md section 8(d) defines it from the reference's
 * weighted_MSE (utils/similarity.py:174-192) with the weights replaced by a validity mask and the
 * NaN handling / normaliser of the MAE loss (utils/mim_vit.py:482-486, :509-519):
 *   valid = ~isnan(q) & ~isnan(x);  m = valid * qmask;  score = sum m (q-x)^2 / (sum m + 1e-5), ascending.
 * The bank holds raw cutouts [n_items, C, H, W] fp32 with their NaNs (cutouts / ra / dec layout of
 * data_processing/utils.py:346-350); it is a sky_bank_t handle (destroy / info / resize / download /
 * profile apply; the embedding calls refuse it).  src of upload may be a device or host pointer. */
SKY_API int sky_pixel_bank_create(sky_bank_t** bank, int device, int64_t n_items, int C, int H, int W);
SKY_API int sky_pixel_bank_upload(sky_bank_t* bank, const float* src, int64_t item0, int64_t n_items, void* stream);
/* q[Q, C*H*W] f32 (NaN = missing), qmask[Q, C*H*W] u8 (1 = compare) or NULL = all ones. */
SKY_API int sky_search_pixels(sky_bank_t* bank, const float* q, const unsigned char* qmask, int Q, int k,
                      int64_t idx_offset, float* out_scores, int64_t* out_idx, void* stream);
SKY_API int sky_score_pixels(sky_bank_t* bank, const float* q, const unsigned char* qmask, int Q, int64_t item0,
                     int64_t n_items, float* out_scores, void* stream);

/* ---- pixel-side preparation in front of the bank -------------------------
 * S/N pre-filter of the h5 search driver: calculate_snr (utils/misc.py:119-163) per image and channel,
 *   snr = mean(central n x n pixels) / (std(all other pixels, ddof 0) + 1e-8), NaN pixels propagating as in numpy;
 * out_snr[n_items, C] (may be NULL) and out_min[n_items] = nanmin over the first n_min_channels channels (may be NULL;
 * similarity_search.py:127 takes the first five).  The caller keeps rows with lo < out_min < hi (:130).
 * cutouts: device pointer [n_items, C, H, W] f32, H == W.  device < 0: taken from the pointer. */
SKY_API int sky_pixel_snr(const float* cutouts, int64_t n_items, int C, int H, int W, int n_central_pix,
                  int n_min_channels, float* out_snr, float* out_min, int device, void* stream);
/* same over items [item0, item0 + n_items) of a pixel bank (its cutouts are [C, H, W] with C*H*W pixels). */
SKY_API int sky_pixel_bank_snr(const sky_bank_t* bank, int C, int H, int W, int64_t item0, int64_t n_items,
                       int n_central_pix, int n_min_channels, float* out_snr, float* out_min, void* stream);
/* FITS-tile streaming: cutout i = clip(tile[:, h0_i : h0_i + size, w0_i : w0_i + size]) with (h0, w0) = coords[i]
 * (overlapping_cutouts, utils/dataloaders.py:511-536; clipping :657-661; the coordinate list of
 * generate_overlap_coords :481-509 comes from the host).  tile [C, H, W] f32, coords [n, 2] i32, out [n, C, size, size];
 * pixel_min / pixel_max = NaN disables that bound; NaN pixels stay NaN. */
SKY_API int sky_tile_cutouts(const float* tile, int C, int H, int W, const int32_t* coords, int64_t n, int size,
                     float pixel_min, float pixel_max, float* out, int device, void* stream);
/* h5 item path (utils/dataloaders.py:291-300, :685-700): clip, and central size x size crop of [n, C, Hs, Ws]. */
SKY_API int sky_center_clip(const float* src, int64_t n, int C, int Hs, int Ws, int size, float pixel_min,
                    float pixel_max, float* out, int device, void* stream);

/* merge R candidate lists per query into one top-k_out, best first: the shard merge after the
 * NCCL all-gather, and the running merge of update_best_scores (utils/similarity.py:18-35).
 * scores[R,Q,k_in] f32, idx[R,Q,k_in] i64 (idx < 0 = empty slot). */
SKY_API int sky_merge_candidates(const float* scores, const int64_t* idx, int R, int Q, int k_in, int k_out,
                         int metric, float* out_scores, int64_t* out_idx, int device, void* stream);

/* same, the R lists living in R per-rank blocks: list r starts stride_* ELEMENTS after list r-1 (the
 * gathered buffer of ONE all-gather that carries a rank's scores and indices together). */
SKY_API int sky_merge_candidates_strided(const float* scores, const int64_t* idx, int R, int Q, int k_in,
                                 int64_t stride_scores, int64_t stride_idx, int k_out, int metric,
                                 float* out_scores, int64_t* out_idx, int device, void* stream);

/* ---- candidate exchange of a row-sharded search over peer memory (NVLink) -------------------------------------
 * The reference's running top-k is a chunk-wise merge (utils/similarity.py:18-35), so a bank sharded by rows over the
 * GPUs of one box needs one exchange per search: every rank's [Q, k] candidates to every rank, then the merge above.
 * Instead of an NCCL all-gather in front of sky_merge_candidates_strided, every rank owns a buffer all peers map:
 * sky_exchange_merge pushes the local block into every peer with plain stores + a per-query flag, then one CTA per
 * query waits for its R flags and merges -- two kernels of this library on the search stream, no collective launch.
 * Setup (once): create on every rank, exchange the 64-byte handles with any transport (torch.distributed
 * all_gather), open.  All ranks must call sky_exchange_merge the same number of times (it is a collective).
 * One process per GPU: sky_exchange_open (cudaIpc).  Several ranks in one process: sky_exchange_open_local. */
typedef struct sky_exchange sky_exchange_t;
SKY_API int sky_exchange_create(sky_exchange_t** x, int device, int rank, int world, int max_Q, int max_k);
SKY_API int sky_exchange_handle_bytes(void);
SKY_API int sky_exchange_handle(sky_exchange_t* x, void* h_handle);
SKY_API int sky_exchange_open(sky_exchange_t* x, const void* h_handles /* [world][sky_exchange_handle_bytes()] */);
SKY_API int sky_exchange_open_local(sky_exchange_t* x, void* const* peer_ptrs /* [world] device pointers */);
SKY_API void* sky_exchange_local_ptr(sky_exchange_t* x);
SKY_API int sky_exchange_destroy(sky_exchange_t* x);
/* scores[Q,k] f32 / idx[Q,k] i64 (global indices, idx < 0 = empty slot): this rank's best-first candidates ->
 * out_scores / out_idx [Q, k_out]: the global top-k_out, identical on every rank. */
SKY_API int sky_exchange_merge(sky_exchange_t* x, const float* scores, const int64_t* idx, int Q, int k, int k_out,
                       int metric, float* out_scores, int64_t* out_idx, void* stream);
/* sky_search over this rank's shard + the exchange in one call: the shard merge kernel writes its [Q, k] result
 * straight into every peer (no separate push), then the flag-waiting merge yields the GLOBAL top-k on every rank.
 * idx_offset = first global row of the shard.  Collective: every rank calls it with the same Q and k. */
SKY_API int sky_search_sharded(sky_bank_t* bank, sky_exchange_t* x, const float* t, const float* w, int Q, int metric,
                       int combine, int n_top_sims, int k, int64_t idx_offset, float* out_scores, int64_t* out_idx,
                       int path, void* stream);

/* Timing of the dominant (scoring) kernel of each search with CUDA events recorded on the launch
 * stream right around it: enable, run searches, then read (#launches, total ms); read synchronises. */
SKY_API int sky_profile_enable(sky_bank_t* bank, int enable);
SKY_API int sky_profile_read(sky_bank_t* bank, int64_t* launches, double* total_ms, int reset);

/* kernels launched by this library on the calling thread since the last reset (bench accounting). */
SKY_API int64_t sky_launch_count(int reset);

#ifdef __cplusplus
}
#endif
#endif /* SKY_SEARCH_H_ */
