// bank.cuh -- the bank handle and internal launcher declarations.
#pragma once
#include <cuda.h>

#include <vector>

#include "common.cuh"

struct sky_exchange;      // exchange.cu

struct sky_bank {
    int device = 0;
    int num_sms = 0;
    int64_t n_items = 0;
    int64_t capacity = 0;   // items allocated
    int L = 1;          // tokens kept per item
    int D = 0;          // feature dimension
    int Dp = 0;         // row stride in elements (D padded to a multiple of 64)
    int dtype = SKY_F32;
    int64_t rows = 0;       // n_items * L
    int64_t rows_pad = 0;   // rows rounded up to the 128-row tensor tile
    void* data = nullptr;       // [rows_pad, Dp] row-major, normalised + rounded
    float* rownorm = nullptr;   // [rows_pad] sum of squares of the stored row
    float* mu = nullptr;        // [D]
    float* sigma = nullptr;     // [D]
    float* sp = nullptr;        // [D] sigma + 1e-8
    bool has_norm = false;
    bool finalized = false;
    bool pixel = false;         // pixel-space bank: data is row-major [n_items][D] fp32 with NaNs kept
    // grow-only scratch
    void* ws = nullptr;
    size_t ws_bytes = 0;
    void* ws2 = nullptr;        // small persistent scratch (query packing, statistics)
    size_t ws2_bytes = 0;
    // TMA descriptor of the bank (bf16 only), built at finalize
    CUtensorMap tmap_bank;
    bool tmap_ready = false;
    // optional timing of the dominant (scoring) kernel: CUDA event pairs on the launch stream
    bool profile = false;
    std::vector<cudaEvent_t>* prof_events = nullptr;   // [2 * launches]
};

namespace sky {

int ensure_ws(sky_bank* b, size_t bytes);
int ensure_ws2(sky_bank* b, size_t bytes);
void prof_mark(const sky_bank* b, cudaStream_t st);   // call right before and right after the scoring kernel

// bank.cu
int launch_ingest(const void* src, int src_dtype, int64_t n_items, int src_tokens, int token_mode,
                  int num_extra, int L, int D, int Dp, const float* mu, const float* sp, void* dst,
                  int dst_dtype, float* rownorm, int64_t dst_row0, cudaStream_t st);
int launch_col_stats(const float* x, int64_t n_rows, int D, int ld, int tiled_kblocks, const float* mu_in,
                     const float* sp_in, float* mean_out, float* std_out, cudaStream_t st);
int launch_add_eps(const float* sigma, float* sp, int D, cudaStream_t st);
int launch_finish_weights(const float* std_in, int D, int use_weights, float* w_out, cudaStream_t st);
int launch_download(const void* data, int dtype, int64_t row0, int64_t nrows, int D, int Dp, float* dst,
                    cudaStream_t st);

// Search state shared by the scorers and the merge (all device memory inside bank->ws).
struct SearchState {
    uint64_t* lists = nullptr;   // [P][Qtot][cap]
    int* counts = nullptr;       // [P][Qtot]
    uint32_t* gtop = nullptr;    // [p_stride][Qtot]  best key per (CTA, query)
    uint32_t* gtau = nullptr;    // [Qtot]            min over CTAs, maintained by reducer CTAs
    int P = 0;                   // number of CTAs (lists) of the scorer
    int p_stride = 0;            // P rounded up to 32
    int Qtot = 0;
    int cap = 0;
    int k = 0;
    int use_gtau = 0;
};

// simt_search.cu
struct SimtArgs {
    const void* bank; int dtype;
    int64_t row0;
    int64_t n_items; int L; int D; int Dp;
    const float* t; const float* w; int Q;
    int metric, combine, n_top;
    // emit mode (sky_score): scores of items [item0, item0+n) -> emit[q*n + (item-item0)]
    float* emit; int64_t item0; int64_t n;
};
int simt_grid(const sky_bank* b, int metric, int L, int64_t n_items, int qc, int n_top, int* grid, size_t* smem);
int simt_pick_qc(int Dp);
int launch_simt_search(const sky_bank* b, const SimtArgs& a, const SearchState& s, int grid, int qc, size_t smem,
                       cudaStream_t st);

// stream_search.cu
int stream_pick_qc(int Q);
bool stream_supported(const sky_bank* b, int qc);
int stream_grid(const sky_bank* b, int64_t row_lo, int64_t row_hi);
int debug_stream_stats(unsigned long long* h_out, int reset);
int launch_stream_search(const sky_bank* b, const SimtArgs& a, const SearchState& s, int grid, int qc, cudaStream_t st);

// pixel_search.cu
int pixel_pick_qc(int Q);
int pixel_grid(const sky_bank* b, int64_t n_rows, int qc);
int launch_pixel_fold(const float* q, const unsigned char* mask, int64_t n, int D, float* qp, int* excl, cudaStream_t st);
int launch_pixel_search(const sky_bank* b, const float* qp, const int* excl, int Q, int64_t row_lo, int64_t row_hi,
                        const SearchState& s, int grid, int qc, float* emit, cudaStream_t st);

// pixel_prep.cu
int launch_snr(const float* img, int64_t n_items, int C, int H, int W, int n_central, int n_min_channels, float* out_snr,
               float* out_min, cudaStream_t st);
int launch_tile_cutouts(const float* tile, int C, int H, int W, const int* coords, int64_t n, int size, float lo, float hi,
                        float* out, cudaStream_t st);
int launch_center_clip(const float* src, int64_t n, int C, int Hs, int Ws, int size, float lo, float hi, float* out,
                       cudaStream_t st);

// tc_search.cu
bool tc_supported(const sky_bank* b, int metric, bool weighted, int n_top);
int tc_grid(const sky_bank* b);
int tc_make_bank_tmap(sky_bank* b);
int launch_tc_search(sky_bank* b, const float* t, int Q, int metric, const SearchState& s, cudaStream_t st);
size_t tc_scratch_bytes(const sky_bank* b, int Q);
int debug_read_trace(unsigned long long* h_out, int n);
int debug_read_epi(unsigned long long* h_out);

// tc_weighted.cu
bool tc_weighted_supported(const sky_bank* b, int metric, bool weighted, int n_top);
size_t tc_weighted_scratch_bytes(const sky_bank* b);
int tc_weighted_grid(const sky_bank* b);
int debug_read_tw_trace(unsigned long long* h_out, int n);
int launch_tc_weighted(sky_bank* b, const float* t, const float* w, int Q, int metric, const SearchState& s, cudaStream_t st);

// tc_batch.cu
int debug_read_tb_trace(long long* h_out, int n);
bool tc_batch_supported(const sky_bank* b, int metric, bool weighted, int n_top, int k);
int launch_tc_batch(sky_bank* b, const float* t, int Q, int metric, int k, int64_t idx_offset, float* out_scores,
                    int64_t* out_idx, cudaStream_t st);

// exchange.cu: where a sharded search delivers its [Q, k] result (every peer's slot of this rank + a flag per query)
constexpr int kMaxPeers = 16;
struct XchgTarget {
    unsigned char* base[kMaxPeers];   // mapped base of every rank's exchange buffer (base[rank] = local)
    int world = 0, rank = 0;
    size_t slot_units = 0;            // int64 units of one rank's block
    int max_Q = 0, parity = 0;
    unsigned seq = 0;
    bool fused = false;               // set by the launcher that delivered the result itself
};

int xchg_begin(sky_exchange* x, int Q, int k, XchgTarget* xt);
void xchg_local_slot(const XchgTarget& xt, int Q, int k, float** scores, int64_t** idx);
int launch_xchg_push(const XchgTarget& xt, const float* scores, const int64_t* idx, int Q, int k, int skip_self, int device, cudaStream_t st);
int launch_xchg_merge(const XchgTarget& xt, int Q, int k, int k_out, int metric, float* out_scores, int64_t* out_idx, int device, cudaStream_t st);
int xchg_device(const sky_exchange* x);

// merge.cu
int launch_init_state(const SearchState& s, int p_active, cudaStream_t st);
int launch_merge_lists(const SearchState& s, int metric, int64_t idx_offset, float* out_scores, int64_t* out_idx,
                       cudaStream_t st, XchgTarget* xt = nullptr);
int launch_merge_candidates(const float* scores, const int64_t* idx, int R, int Q, int k_in, int64_t stride_s,
                            int64_t stride_i, int k_out, int metric, float* out_scores, int64_t* out_idx, cudaStream_t st);

}  // namespace sky
