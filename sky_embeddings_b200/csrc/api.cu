// api.cu -- the extern "C" boundary of libskysearch.so (see include/sky_search.h).
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>

#include "bank.cuh"

namespace sky {

static thread_local char g_err[512] = "";
static thread_local int64_t g_launches = 0;

int set_error(int code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return code;
}
void count_launch(int n) { g_launches += n; }
#ifdef SKY_EXPERIMENTS
int env_knob(const char* name, int dflt) {
    const char* e = getenv(name);
    return e ? atoi(e) : dflt;
}
#endif

static int grow(void** p, size_t* have, size_t want) {
    if (*have >= want) return SKY_OK;
    if (*p) {
        SKY_CUDA(cudaDeviceSynchronize());   // earlier work may still use the old buffer
        SKY_CUDA(cudaFree(*p));
        *p = nullptr;
        *have = 0;
    }
    want = static_cast<size_t>(round_up(static_cast<int64_t>(want), 1 << 20));
    cudaError_t e = cudaMalloc(p, want);
    if (e != cudaSuccess) {
        *p = nullptr;
        return set_error(SKY_ERR_NOMEM, "cudaMalloc(%zu) failed: %s", want, cudaGetErrorString(e));
    }
    *have = want;
    return SKY_OK;
}
int ensure_ws(sky_bank* b, size_t bytes) { return grow(&b->ws, &b->ws_bytes, bytes); }
int ensure_ws2(sky_bank* b, size_t bytes) { return grow(&b->ws2, &b->ws2_bytes, bytes); }

void prof_mark(const sky_bank* b, cudaStream_t st) {
    if (!b->profile || !b->prof_events || b->prof_events->size() >= 16384) return;
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    cudaEventRecord(e, st);
    b->prof_events->push_back(e);
}

constexpr int kAutoStreamMaxQ = 4;   // widest single pass of K1

static bool valid_metric(int m) { return m == SKY_COSINE || m == SKY_MSE || m == SKY_MAE; }
static bool valid_combine(int c) { return c == SKY_MEAN || c == SKY_MIN || c == SKY_MAX; }

static int token_out(int token_mode, int src_tokens, int num_extra) {
    switch (token_mode) {
        case SKY_TOK_ALL: return src_tokens;
        case SKY_TOK_CLS: return 1;
        case SKY_TOK_PATCHES: return src_tokens - num_extra;
        case SKY_TOK_MAXPOOL: return 1;
        default: return -1;
    }
}

struct DeviceGuard {
    int prev = -1;
    bool ok = false;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) == cudaSuccess && cudaSetDevice(dev) == cudaSuccess) ok = true;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

// Carve the candidate-list state for a search out of bank->ws.
static int plan_state(sky_bank* b, int P, int Q, int k, int p_active, SearchState* s) {
    s->P = P;
    s->p_stride = static_cast<int>(round_up(P, 32));
    s->Qtot = Q;
    s->k = k;
    // room for the k best plus one prune interval; more slack means fewer prunes
    int cap = k + kPruneSlack;
    const int want = (k < 256) ? 1024 : 2 * k + kPruneSlack;
    const size_t budget = static_cast<size_t>(3) << 30;
    if (static_cast<size_t>(P) * Q * want * sizeof(uint64_t) <= budget) cap = want;
    s->cap = cap;
    s->use_gtau = (k <= p_active) ? 1 : 0;
    const size_t lists_b = static_cast<size_t>(P) * Q * cap * sizeof(uint64_t);
    const size_t counts_b = static_cast<size_t>(round_up(static_cast<int64_t>(P) * Q * sizeof(int), 256));
    const size_t gtop_b = static_cast<size_t>(round_up(static_cast<int64_t>(Q) * s->p_stride * sizeof(uint32_t), 256));
    const size_t gtau_b = static_cast<size_t>(round_up(static_cast<int64_t>(Q) * sizeof(uint32_t), 256));
    int rc = ensure_ws(b, lists_b + counts_b + gtop_b + gtau_b + 256);
    if (rc) return rc;
    unsigned char* p = reinterpret_cast<unsigned char*>(b->ws);
    s->lists = reinterpret_cast<uint64_t*>(p);
    s->counts = reinterpret_cast<int*>(p + lists_b);
    s->gtop = reinterpret_cast<uint32_t*>(p + lists_b + counts_b);
    s->gtau = reinterpret_cast<uint32_t*>(p + lists_b + counts_b + gtop_b);
    return SKY_OK;
}

}  // namespace sky

using namespace sky;

extern "C" {

const char* sky_last_error(void) { return g_err; }
int sky_abi_version(void) { return SKY_ABI_VERSION; }
int64_t sky_launch_count(int reset) {
    int64_t v = g_launches;
    if (reset) g_launches = 0;
    return v;
}

int sky_bank_create(sky_bank_t** out, int device, int64_t n_items, int L, int D, int dtype) {
    if (!out) return set_error(SKY_ERR_ARG, "bank out-pointer is NULL");
    *out = nullptr;
    if (n_items < 0 || L < 1 || D < 1) return set_error(SKY_ERR_ARG, "bad bank shape n_items=%lld L=%d D=%d", (long long)n_items, L, D);
    if (dtype != SKY_F32 && dtype != SKY_BF16) return set_error(SKY_ERR_ARG, "bad bank dtype %d", dtype);
    if (n_items * L >= 0xFFFFFFFFll) return set_error(SKY_ERR_UNSUPPORTED, "a bank shard is limited to 2^32-2 rows");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_error(SKY_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    if (device < 0 || device >= ndev) return set_error(SKY_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard g(device);
    if (!g.ok) return set_error(SKY_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    sky_bank* b = new (std::nothrow) sky_bank();
    if (!b) return set_error(SKY_ERR_NOMEM, "host allocation failed");
    b->device = device;
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete b; return set_error(SKY_ERR_CUDA, "cudaGetDeviceProperties failed"); }
    if (prop.major != 10) { delete b; return set_error(SKY_ERR_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor); }
    b->num_sms = prop.multiProcessorCount;
    b->n_items = n_items; b->capacity = n_items; b->L = L; b->D = D; b->dtype = dtype;
    b->Dp = static_cast<int>(round_up(D, kKBlock));
    b->rows = n_items * L;
    b->rows_pad = round_up(b->rows > 0 ? b->rows : 1, kTileRows);
    const size_t esz = dtype == SKY_BF16 ? 2 : 4;
    const size_t data_b = static_cast<size_t>(b->rows_pad) * b->Dp * esz;
    cudaError_t e = cudaMalloc(&b->data, data_b);
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->rownorm), static_cast<size_t>(b->rows_pad) * sizeof(float));
    if (e == cudaSuccess) e = cudaMalloc(reinterpret_cast<void**>(&b->mu), 3 * static_cast<size_t>(D) * sizeof(float));
    if (e != cudaSuccess) {
        int rc = set_error(SKY_ERR_NOMEM, "bank allocation of %zu B failed: %s", data_b, cudaGetErrorString(e));
        sky_bank_destroy(b);
        return rc;
    }
    b->sigma = b->mu + D;
    b->sp = b->sigma + D;
    // the pad rows of the last tile must be finite zeros for the tensor path: clear that tile
    const size_t pad_rows = static_cast<size_t>(b->rows_pad - b->rows);
    if (pad_rows) {
        const size_t tile_elems = static_cast<size_t>(kTileRows) * b->Dp;
        cudaMemset(reinterpret_cast<unsigned char*>(b->data) + (data_b - tile_elems * esz), 0, tile_elems * esz);
        cudaMemset(b->rownorm + b->rows, 0, pad_rows * sizeof(float));
    }
    *out = b;
    return SKY_OK;
}

int sky_bank_destroy(sky_bank_t* b) {
    if (!b) return SKY_OK;
    DeviceGuard g(b->device);
    if (b->data) cudaFree(b->data);
    if (b->rownorm) cudaFree(b->rownorm);
    if (b->mu) cudaFree(b->mu);
    if (b->ws) cudaFree(b->ws);
    if (b->ws2) cudaFree(b->ws2);
    if (b->prof_events) {
        for (cudaEvent_t e : *b->prof_events) cudaEventDestroy(e);
        delete b->prof_events;
    }
    delete b;
    return SKY_OK;
}

int sky_bank_info(const sky_bank_t* b, int64_t* n_items, int* L, int* D, int* dtype) {
    if (!b) return set_error(SKY_ERR_ARG, "bank is NULL");
    if (n_items) *n_items = b->n_items;
    if (L) *L = b->L;
    if (D) *D = b->D;
    if (dtype) *dtype = b->dtype;
    return SKY_OK;
}

int sky_bank_fit_norm(sky_bank_t* b, const void* src, int src_dtype, int64_t n_items, int src_tokens,
                      int token_mode, int num_extra_tokens, void* stream) {
    if (!b || !src) return set_error(SKY_ERR_ARG, "NULL argument");
    if (src_dtype != SKY_F32 && src_dtype != SKY_BF16) return set_error(SKY_ERR_ARG, "bad source dtype %d", src_dtype);
    if (token_out(token_mode, src_tokens, num_extra_tokens) != b->L)
        return set_error(SKY_ERR_ARG, "token selection yields %d tokens, bank expects L=%d", token_out(token_mode, src_tokens, num_extra_tokens), b->L);
    if (n_items < 1) return set_error(SKY_ERR_ARG, "fit_norm needs at least one item");
    DeviceGuard g(b->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const int64_t nrows = n_items * b->L;
    // token-selected first batch as f32 tiles (+ throw-away row norms)
    const size_t tmp_b = static_cast<size_t>(round_up(nrows, kTileRows)) * b->Dp * sizeof(float);
    int rc = ensure_ws2(b, tmp_b + static_cast<size_t>(nrows) * sizeof(float) + 256);
    if (rc) return rc;
    float* tmp = reinterpret_cast<float*>(b->ws2);
    float* tmp_norm = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(b->ws2) + round_up(static_cast<int64_t>(tmp_b), 256));
    rc = launch_ingest(src, src_dtype, n_items, src_tokens, token_mode, num_extra_tokens, b->L, b->D, b->Dp,
                       nullptr, nullptr, tmp, SKY_F32, tmp_norm, 0, st);
    if (rc) return rc;
    rc = launch_col_stats(tmp, nrows, b->D, b->Dp, b->Dp / kKBlock, nullptr, nullptr, b->mu, b->sigma, st);
    if (rc) return rc;
    rc = launch_add_eps(b->sigma, b->sp, b->D, st);
    if (rc) return rc;
    b->has_norm = true;
    return SKY_OK;
}

int sky_bank_set_norm(sky_bank_t* b, const float* mu, const float* sigma, void* stream) {
    if (!b || !mu || !sigma) return set_error(SKY_ERR_ARG, "NULL argument");
    DeviceGuard g(b->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SKY_CUDA(cudaMemcpyAsync(b->mu, mu, b->D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SKY_CUDA(cudaMemcpyAsync(b->sigma, sigma, b->D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    int rc = launch_add_eps(b->sigma, b->sp, b->D, st);
    if (rc) return rc;
    b->has_norm = true;
    return SKY_OK;
}

int sky_bank_get_norm(const sky_bank_t* b, float* mu, float* sigma, void* stream) {
    if (!b || !mu || !sigma) return set_error(SKY_ERR_ARG, "NULL argument");
    if (!b->has_norm) return set_error(SKY_ERR_STATE, "bank has no normalisation statistics");
    DeviceGuard g(b->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SKY_CUDA(cudaMemcpyAsync(mu, b->mu, b->D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    SKY_CUDA(cudaMemcpyAsync(sigma, b->sigma, b->D * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return SKY_OK;
}

int sky_bank_upload(sky_bank_t* b, const void* src, int src_dtype, int64_t item0, int64_t n_items, int src_tokens,
                    int token_mode, int num_extra_tokens, void* stream) {
    if (!b || (!src && n_items > 0)) return set_error(SKY_ERR_ARG, "NULL argument");
    if (src_dtype != SKY_F32 && src_dtype != SKY_BF16) return set_error(SKY_ERR_ARG, "bad source dtype %d", src_dtype);
    if (item0 < 0 || n_items < 0 || item0 + n_items > b->n_items)
        return set_error(SKY_ERR_ARG, "items [%lld, %lld) outside the bank (%lld items)", (long long)item0, (long long)(item0 + n_items), (long long)b->n_items);
    if (b->pixel) return set_error(SKY_ERR_STATE, "this is a pixel bank: use sky_pixel_bank_upload");
    if (token_out(token_mode, src_tokens, num_extra_tokens) != b->L)
        return set_error(SKY_ERR_ARG, "token selection yields %d tokens, bank expects L=%d", token_out(token_mode, src_tokens, num_extra_tokens), b->L);
    DeviceGuard g(b->device);
    b->finalized = false;
    return launch_ingest(src, src_dtype, n_items, src_tokens, token_mode, num_extra_tokens, b->L, b->D, b->Dp,
                         b->has_norm ? b->mu : nullptr, b->has_norm ? b->sp : nullptr, b->data, b->dtype, b->rownorm,
                         item0 * b->L, static_cast<cudaStream_t>(stream));
}

int sky_bank_finalize(sky_bank_t* b, void* stream) {
    (void)stream;
    if (!b) return set_error(SKY_ERR_ARG, "bank is NULL");
    if (b->pixel) return SKY_OK;
    DeviceGuard g(b->device);
    if (b->dtype == SKY_BF16 && !b->tmap_ready) {
        int rc = tc_make_bank_tmap(b);
        if (rc) return rc;
    }
    b->finalized = true;
    return SKY_OK;
}

#ifdef SKY_EXPERIMENTS
/* experiment builds only (not part of the public header): kernel timelines, see tc_search.cu */
__attribute__((visibility("default"))) int sky_debug_trace(unsigned long long* h_out, int n) { return debug_read_trace(h_out, n); }

__attribute__((visibility("default"))) int sky_debug_stream_stats(unsigned long long* h_out, int reset) { return debug_stream_stats(h_out, reset); }

__attribute__((visibility("default"))) int sky_debug_tb_trace(long long* h_out, int n) { return debug_read_tb_trace(h_out, n); }

__attribute__((visibility("default"))) int sky_debug_epi(unsigned long long* h_out) { return debug_read_epi(h_out); }

__attribute__((visibility("default"))) int sky_debug_tw_trace(unsigned long long* h_out, int n) { return debug_read_tw_trace(h_out, n); }
#endif

int sky_profile_enable(sky_bank_t* b, int enable) {
    if (!b) return set_error(SKY_ERR_ARG, "bank is NULL");
    if (!b->prof_events) b->prof_events = new std::vector<cudaEvent_t>();
    b->profile = enable != 0;
    return SKY_OK;
}

int sky_profile_read(sky_bank_t* b, int64_t* launches, double* total_ms, int reset) {
    if (!b || !launches || !total_ms) return set_error(SKY_ERR_ARG, "NULL argument");
    *launches = 0;
    *total_ms = 0.0;
    if (!b->prof_events) return SKY_OK;
    DeviceGuard g(b->device);
    std::vector<cudaEvent_t>& ev = *b->prof_events;
    for (size_t i = 0; i + 1 < ev.size(); i += 2) {
        SKY_CUDA(cudaEventSynchronize(ev[i + 1]));
        float ms = 0.f;
        SKY_CUDA(cudaEventElapsedTime(&ms, ev[i], ev[i + 1]));
        *total_ms += ms;
        *launches += 1;
    }
    if (reset) {
        for (cudaEvent_t e : ev) cudaEventDestroy(e);
        ev.clear();
    }
    return SKY_OK;
}

int sky_bank_resize(sky_bank_t* b, int64_t n_items) {
    if (!b) return set_error(SKY_ERR_ARG, "bank is NULL");
    if (n_items < 0 || n_items > b->capacity)
        return set_error(SKY_ERR_ARG, "resize to %lld items exceeds the capacity %lld", (long long)n_items, (long long)b->capacity);
    b->n_items = n_items;
    b->rows = n_items * b->L;
    // rows_pad (allocation and TMA extent) keeps covering the capacity; rows >= b->rows are masked
    return SKY_OK;
}

int sky_bank_download(const sky_bank_t* b, int64_t item0, int64_t n_items, float* dst, void* stream) {
    if (!b || !dst) return set_error(SKY_ERR_ARG, "NULL argument");
    if (item0 < 0 || n_items < 0 || item0 + n_items > b->n_items) return set_error(SKY_ERR_ARG, "item range outside the bank");
    DeviceGuard g(b->device);
    if (b->pixel) {
        SKY_CUDA(cudaMemcpyAsync(dst, reinterpret_cast<const float*>(b->data) + static_cast<size_t>(item0) * b->D,
                                 static_cast<size_t>(n_items) * b->D * sizeof(float), cudaMemcpyDeviceToDevice,
                                 static_cast<cudaStream_t>(stream)));
        return SKY_OK;
    }
    return launch_download(b->data, b->dtype, item0 * b->L, n_items * b->L, b->D, b->Dp, dst, static_cast<cudaStream_t>(stream));
}

int sky_query_from_targets(const sky_bank_t* b, const float* targets, int64_t T_rows, int D, int use_weights,
                           float* t_out, float* w_out, void* stream) {
    if (!targets || !t_out || !w_out) return set_error(SKY_ERR_ARG, "NULL argument");
    if (T_rows < 1 || D < 1) return set_error(SKY_ERR_ARG, "bad target shape");
    if (b && b->D != D) return set_error(SKY_ERR_ARG, "target D=%d differs from bank D=%d", D, b->D);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    // without a bank the kernels run on the device that owns `targets`
    int dev = b ? b->device : -1;
    if (dev < 0) {
        cudaPointerAttributes pa;
        if (cudaPointerGetAttributes(&pa, targets) != cudaSuccess || pa.type != cudaMemoryTypeDevice)
            return set_error(SKY_ERR_ARG, "targets is not a device pointer (there is no CPU fallback)");
        dev = pa.device;
    }
    DeviceGuard g(dev);
    if (!g.ok) return set_error(SKY_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    const bool nrm = b && b->has_norm;
    // std goes to w_out first, then is turned into weights in place
    int rc = launch_col_stats(targets, T_rows, D, D, 0, nrm ? b->mu : nullptr, nrm ? b->sp : nullptr, t_out, w_out, st);
    if (rc) return rc;
    return launch_finish_weights(w_out, D, use_weights, w_out, st);
}

static int search_impl(sky_bank_t* b, const float* t, const float* w, int Q, int metric, int combine, int n_top_sims,
                       int k, int64_t idx_offset, float* out_scores, int64_t* out_idx, int path, cudaStream_t st,
                       XchgTarget* xt = nullptr) {
    if (!b || !t || !out_scores || !out_idx) return set_error(SKY_ERR_ARG, "NULL argument");
    if (!valid_metric(metric)) return set_error(SKY_ERR_ARG, "unknown metric %d (the reference accepts cosine, MSE, MAE)", metric);
    if (!valid_combine(combine)) return set_error(SKY_ERR_ARG, "unknown combine %d", combine);
    if (Q < 1 || k < 1) return set_error(SKY_ERR_ARG, "Q and k must be positive");
    if (n_top_sims < 0 || n_top_sims > b->L) return set_error(SKY_ERR_ARG, "n_top_sims=%d out of range for L=%d", n_top_sims, b->L);
    if (b->pixel) return set_error(SKY_ERR_STATE, "this is a pixel bank: use sky_search_pixels");
    if (!b->finalized) return set_error(SKY_ERR_STATE, "bank is not finalized");
    const bool tc_ok = tc_supported(b, metric, w != nullptr, n_top_sims);
    // queries with per-feature weights: second contraction against the squared tile, made on chip (K2w)
    const bool tw_ok = tc_weighted_supported(b, metric, w != nullptr, n_top_sims);
    if (path == SKY_PATH_TENSOR && !tc_ok && !tw_ok)
        return set_error(SKY_ERR_UNSUPPORTED, "tensor path needs a bf16 bank, L=1, cosine/MSE, no n_top_sims");
    // AUTO keeps Q <= kAutoStreamMaxQ on the streaming kernel K1 (one bank pass, fp32 queries): the tensor paths
    // round the query operands to bf16, and the reference's own regime (one query, or a handful) must not score
    // differently depending on the batch size
    if (tw_ok && (path == SKY_PATH_TENSOR || (path == SKY_PATH_AUTO && Q > kAutoStreamMaxQ))) {
        SearchState sw{};
        const int grid = tc_weighted_grid(b);
        int rcw = plan_state(b, grid, Q, k, grid, &sw);
        if (rcw) return rcw;
        rcw = ensure_ws2(b, tc_weighted_scratch_bytes(b));
        if (rcw) return rcw;
        rcw = launch_tc_weighted(b, t, w, Q, metric, sw, st);      // its packing kernel zeroes the exchange state
        if (rcw) return rcw;
        return launch_merge_lists(sw, metric, idx_offset, out_scores, out_idx, st, xt);
    }
    // small query batches stay on the streaming SIMT kernel (HBM bound there); larger ones need MMA
    const bool use_tc = tc_ok && (path == SKY_PATH_TENSOR || (path == SKY_PATH_AUTO && Q > kAutoStreamMaxQ));
    // large query batches: the GEMM-shaped kernel (bank tile reused by all query groups through L2)
    const bool batch_ok = tc_batch_supported(b, metric, w != nullptr, n_top_sims, k);
    if (path == SKY_PATH_BATCH && !batch_ok)
        return set_error(SKY_ERR_UNSUPPORTED, "batched tensor path needs a finalized bf16 bank, L=1, cosine/MSE, no weights, k <= 4096");
    bool use_batch = batch_ok && path == SKY_PATH_BATCH;
    if (batch_ok && path == SKY_PATH_AUTO && Q > 128) {
        // K2 streams the bank once per 64 queries (HBM bound); K2b is tensor bound but pays a launch + merge per
        // phase.  Pick by a two-line cost model (constants from the measurements in DESIGN.md section 5).
        const double bank_bytes = static_cast<double>(b->rows) * b->Dp * 2.0;
        const double t_stream = ((Q + 63) / 64) * (bank_bytes / 6.0e12 + 40e-6) + 50e-6;
        const int64_t tiles = (b->rows + kTileRows - 1) / kTileRows;
        int phases = 0;
        for (int64_t done = 0, per = 1; done * b->num_sms < tiles; ++phases) { done += per; if (phases >= 1) per *= 4; }
        const double t_batch = 2.0 * Q * static_cast<double>(b->rows) * b->Dp / 1.1e15 + phases * 0.1e-3 + 0.1e-3;
        // measured: 100k x 768, Q = 512: K2 0.51 ms, K2b 0.41 ms; 1M, Q = 4096: K2 19.0 ms, K2b 7.6 ms
        use_batch = t_batch < t_stream;
    }
    if (use_batch) return launch_tc_batch(b, t, Q, metric, k, idx_offset, out_scores, out_idx, st);

    SearchState s{};
    int rc;
    if (use_tc) {
        const int grid = tc_grid(b);
        rc = plan_state(b, grid, Q, k, grid, &s);
        if (rc) return rc;
        rc = ensure_ws2(b, tc_scratch_bytes(b, Q));
        if (rc) return rc;
        rc = launch_tc_search(b, t, Q, metric, s, st);              // its packing kernel zeroes the exchange state
        if (rc) return rc;
    } else {
        SimtArgs a;
        a.bank = b->data; a.dtype = b->dtype; a.row0 = 0; a.n_items = b->n_items; a.L = b->L; a.D = b->D; a.Dp = b->Dp;
        a.t = t; a.w = w; a.Q = Q; a.metric = metric; a.combine = combine; a.n_top = n_top_sims;
        a.emit = nullptr; a.item0 = 0; a.n = 0;
        const int sqc = stream_pick_qc(Q);
        if (path != SKY_PATH_GENERIC && b->rows > 0 && stream_supported(b, sqc)) {
            // K1: bulk-copy staged streaming scorer
            const int grid = stream_grid(b, 0, b->rows);
            rc = plan_state(b, grid, Q, k, grid, &s);
            if (rc) return rc;
            rc = launch_init_state(s, grid, st);
            if (rc) return rc;
            rc = launch_stream_search(b, a, s, grid, sqc, st);
            if (rc) return rc;
        } else {
            // generic CUDA-core scorer: any L, any D
            const int qc = (Q == 1) ? 1 : simt_pick_qc(b->Dp);
            int grid = 1;
            size_t smem = 0;
            rc = simt_grid(b, metric, b->L, b->n_items, qc, n_top_sims, &grid, &smem);
            if (rc) return rc;
            rc = plan_state(b, grid, Q, k, grid, &s);
            if (rc) return rc;
            rc = launch_init_state(s, grid, st);
            if (rc) return rc;
            rc = launch_simt_search(b, a, s, grid, qc, smem, st);
            if (rc) return rc;
        }
    }
    return launch_merge_lists(s, metric, idx_offset, out_scores, out_idx, st, xt);
}

int sky_search(sky_bank_t* b, const float* t, const float* w, int Q, int metric, int combine, int n_top_sims, int k,
               int64_t idx_offset, float* out_scores, int64_t* out_idx, int path, void* stream) {
    if (!b) return set_error(SKY_ERR_ARG, "bank is NULL");
    DeviceGuard g(b->device);
    return search_impl(b, t, w, Q, metric, combine, n_top_sims, k, idx_offset, out_scores, out_idx, path,
                       static_cast<cudaStream_t>(stream));
}

int sky_search_sharded(sky_bank_t* b, sky_exchange_t* x, const float* t, const float* w, int Q, int metric, int combine,
                       int n_top_sims, int k, int64_t idx_offset, float* out_scores, int64_t* out_idx, int path, void* stream) {
    if (!b || !x) return set_error(SKY_ERR_ARG, "bank / exchange is NULL");
    if (xchg_device(x) != b->device) return set_error(SKY_ERR_ARG, "the exchange lives on device %d, the bank on %d", xchg_device(x), b->device);
    DeviceGuard g(b->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    XchgTarget xt;
    int rc = xchg_begin(x, Q, k, &xt);
    if (rc) return rc;
    // the shard's result is delivered into this rank's slot of every peer: by the shard merge kernel itself where the
    // path ends in one (fused), else it lands in the local slot first and one push kernel forwards it
    float* ls = nullptr;
    int64_t* li = nullptr;
    xchg_local_slot(xt, Q, k, &ls, &li);
    rc = search_impl(b, t, w, Q, metric, combine, n_top_sims, k, idx_offset, ls, li, path, st, &xt);
    if (rc) return rc;
    if (!xt.fused) {
        rc = launch_xchg_push(xt, ls, li, Q, k, /*skip_self=*/1, b->device, st);
        if (rc) return rc;
    }
    return launch_xchg_merge(xt, Q, k, k, metric, out_scores, out_idx, b->device, st);
}

int sky_search_host(sky_bank_t* b, const float* h_t, const float* h_w, int Q, int metric, int combine, int n_top_sims,
                    int k, int64_t idx_offset, float* h_out_scores, int64_t* h_out_idx, int path, void* stream) {
    if (!b || !h_t || !h_out_scores || !h_out_idx) return set_error(SKY_ERR_ARG, "NULL argument");
    if (Q < 1 || k < 1) return set_error(SKY_ERR_ARG, "Q and k must be positive");
    DeviceGuard g(b->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    const size_t qb = static_cast<size_t>(Q) * b->D * sizeof(float);
    const size_t sb = static_cast<size_t>(Q) * k * sizeof(float);
    const size_t ib = static_cast<size_t>(Q) * k * sizeof(int64_t);
    // Pinned (page-locked, mapped) host buffers are read and written by the kernels directly: the packing kernel pulls
    // the queries over PCIe and the merge kernel posts the [Q, k] result rows to the host, so the call is the search's
    // own launches plus one stream synchronisation -- no staging copies (four copy operations and their launch gaps,
    // ~40 us of a 0.3 ms BASELINE config 2 step).  Pageable buffers take the staged route below.
    {
        auto mapped = [](const void* h) -> void* {
            cudaPointerAttributes a;
            if (cudaPointerGetAttributes(&a, h) != cudaSuccess) { cudaGetLastError(); return nullptr; }
            return (a.type == cudaMemoryTypeHost) ? a.devicePointer : nullptr;
        };
        void* m_t = mapped(h_t);
        void* m_w = h_w ? mapped(h_w) : nullptr;
        void* m_s = mapped(h_out_scores);
        void* m_i = mapped(h_out_idx);
        if (m_t && (m_w || !h_w) && m_s && m_i) {
            // the tensor paths read the queries once (packing kernel); the streaming kernels read them in every CTA,
            // which is slow over PCIe (measured +0.11 ms on a Q = 1 search): those get one staged copy of the queries
            const bool packed_once = b->dtype == SKY_BF16 && b->L == 1 && metric != SKY_MAE && n_top_sims == 0 &&
                                     (path == SKY_PATH_TENSOR || path == SKY_PATH_BATCH || (path == SKY_PATH_AUTO && Q > kAutoStreamMaxQ));
            const float* q_t = static_cast<const float*>(m_t);
            const float* q_w = static_cast<const float*>(m_w);
            if (!packed_once) {
                size_t scratch0 = tc_scratch_bytes(b, Q);
                if (tc_weighted_scratch_bytes(b) > scratch0) scratch0 = tc_weighted_scratch_bytes(b);
                const size_t o0 = static_cast<size_t>(round_up(static_cast<int64_t>(scratch0), 256));
                int rcs = ensure_ws2(b, o0 + 2 * round_up(qb, 256));
                if (rcs) return rcs;
                float* d_t0 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(b->ws2) + o0);
                float* d_w0 = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(b->ws2) + o0 + round_up(qb, 256));
                SKY_CUDA(cudaMemcpyAsync(d_t0, h_t, qb, cudaMemcpyHostToDevice, st));
                if (h_w) SKY_CUDA(cudaMemcpyAsync(d_w0, h_w, qb, cudaMemcpyHostToDevice, st));
                q_t = d_t0;
                q_w = h_w ? d_w0 : nullptr;
            }
            int rc0 = search_impl(b, q_t, q_w, Q, metric, combine, n_top_sims, k,
                                  idx_offset, static_cast<float*>(m_s), static_cast<int64_t*>(m_i), path, st);
            if (rc0) return rc0;
            SKY_CUDA(cudaStreamSynchronize(st));
            return SKY_OK;
        }
    }
    // staging lives behind the tensor-path scratch in ws2
    size_t scratch = tc_scratch_bytes(b, Q);
    if (tc_weighted_scratch_bytes(b) > scratch) scratch = tc_weighted_scratch_bytes(b);
    const size_t off0 = static_cast<size_t>(round_up(static_cast<int64_t>(scratch), 256));
    int rc = ensure_ws2(b, off0 + 2 * round_up(qb, 256) + round_up(sb, 256) + round_up(ib, 256));
    if (rc) return rc;
    unsigned char* p = reinterpret_cast<unsigned char*>(b->ws2) + off0;
    float* d_t = reinterpret_cast<float*>(p); p += round_up(qb, 256);
    float* d_w = reinterpret_cast<float*>(p); p += round_up(qb, 256);
    int64_t* d_i = reinterpret_cast<int64_t*>(p); p += round_up(ib, 256);
    float* d_s = reinterpret_cast<float*>(p);
    SKY_CUDA(cudaMemcpyAsync(d_t, h_t, qb, cudaMemcpyHostToDevice, st));
    if (h_w) SKY_CUDA(cudaMemcpyAsync(d_w, h_w, qb, cudaMemcpyHostToDevice, st));
    rc = search_impl(b, d_t, h_w ? d_w : nullptr, Q, metric, combine, n_top_sims, k, idx_offset, d_s, d_i, path, st);
    if (rc) return rc;
    SKY_CUDA(cudaMemcpyAsync(h_out_scores, d_s, sb, cudaMemcpyDeviceToHost, st));
    SKY_CUDA(cudaMemcpyAsync(h_out_idx, d_i, ib, cudaMemcpyDeviceToHost, st));
    SKY_CUDA(cudaStreamSynchronize(st));
    return SKY_OK;
}

int sky_score(sky_bank_t* b, const float* t, const float* w, int Q, int metric, int combine, int n_top_sims,
              int64_t item0, int64_t n_items, float* out_scores, void* stream) {
    if (!b || !t || !out_scores) return set_error(SKY_ERR_ARG, "NULL argument");
    if (!valid_metric(metric)) return set_error(SKY_ERR_ARG, "unknown metric %d (the reference accepts cosine, MSE, MAE)", metric);
    if (!valid_combine(combine)) return set_error(SKY_ERR_ARG, "unknown combine %d", combine);
    if (Q < 1) return set_error(SKY_ERR_ARG, "Q must be positive");
    if (n_top_sims < 0 || n_top_sims > b->L) return set_error(SKY_ERR_ARG, "n_top_sims=%d out of range for L=%d", n_top_sims, b->L);
    if (item0 < 0 || n_items < 0 || item0 + n_items > b->n_items) return set_error(SKY_ERR_ARG, "item range outside the bank");
    if (b->pixel) return set_error(SKY_ERR_STATE, "this is a pixel bank: use sky_score_pixels");
    if (n_items == 0) return SKY_OK;
    DeviceGuard g(b->device);
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    SimtArgs a;
    a.bank = b->data; a.row0 = item0 * b->L;
    a.dtype = b->dtype; a.n_items = n_items; a.L = b->L; a.D = b->D; a.Dp = b->Dp;
    a.t = t; a.w = w; a.Q = Q; a.metric = metric; a.combine = combine; a.n_top = n_top_sims;
    a.emit = out_scores; a.item0 = 0; a.n = n_items;
    SearchState s{};   // unused in emit mode
    const int sqc = stream_pick_qc(Q);
    if (!env_knob("SKY_SCORE_GENERIC", 0) && stream_supported(b, sqc))
        return launch_stream_search(b, a, s, stream_grid(b, a.row0, a.row0 + n_items * b->L), sqc, st);
    const int qc = (Q == 1) ? 1 : simt_pick_qc(b->Dp);
    int grid = 1;
    size_t smem = 0;
    int rc = simt_grid(b, metric, b->L, n_items, qc, n_top_sims, &grid, &smem);
    if (rc) return rc;
    return launch_simt_search(b, a, s, grid, qc, smem, st);
}

/* ---- pixel-space bank (BASELINE config 5) ---------------------------------------------------- */
int sky_pixel_bank_create(sky_bank_t** out, int device, int64_t n_items, int C, int H, int W) {
    if (!out) return set_error(SKY_ERR_ARG, "bank out-pointer is NULL");
    *out = nullptr;
    if (n_items < 0 || C < 1 || H < 1 || W < 1) return set_error(SKY_ERR_ARG, "bad pixel bank shape");
    const int64_t D = static_cast<int64_t>(C) * H * W;
    if (D % 4 != 0 || D > (1 << 24)) return set_error(SKY_ERR_UNSUPPORTED, "C*H*W = %lld must be a multiple of 4 (and < 2^24)", (long long)D);
    if (n_items >= 0xFFFFFFFFll) return set_error(SKY_ERR_UNSUPPORTED, "a bank shard is limited to 2^32-2 rows");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return set_error(SKY_ERR_CUDA, "no CUDA device: this engine has no CPU fallback");
    if (device < 0 || device >= ndev) return set_error(SKY_ERR_ARG, "device %d out of range (%d devices)", device, ndev);
    DeviceGuard g(device);
    if (!g.ok) return set_error(SKY_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return set_error(SKY_ERR_CUDA, "cudaGetDeviceProperties failed");
    if (prop.major != 10) return set_error(SKY_ERR_UNSUPPORTED, "device is sm_%d%d; this library is built for sm_100a only", prop.major, prop.minor);
    sky_bank* b = new (std::nothrow) sky_bank();
    if (!b) return set_error(SKY_ERR_NOMEM, "host allocation failed");
    b->device = device; b->num_sms = prop.multiProcessorCount;
    b->pixel = true; b->n_items = n_items; b->capacity = n_items; b->L = 1; b->D = static_cast<int>(D); b->Dp = b->D;
    b->dtype = SKY_F32; b->rows = n_items; b->rows_pad = n_items;
    const size_t data_b = static_cast<size_t>(n_items > 0 ? n_items : 1) * D * sizeof(float);
    cudaError_t e = cudaMalloc(&b->data, data_b);
    if (e != cudaSuccess) {
        int rc = set_error(SKY_ERR_NOMEM, "pixel bank allocation of %zu B failed: %s", data_b, cudaGetErrorString(e));
        sky_bank_destroy(b);
        return rc;
    }
    b->finalized = true;
    *out = b;
    return SKY_OK;
}

int sky_pixel_bank_upload(sky_bank_t* b, const float* src, int64_t item0, int64_t n_items, void* stream) {
    if (!b || (!src && n_items > 0)) return set_error(SKY_ERR_ARG, "NULL argument");
    if (!b->pixel) return set_error(SKY_ERR_STATE, "not a pixel bank");
    if (item0 < 0 || n_items < 0 || item0 + n_items > b->n_items) return set_error(SKY_ERR_ARG, "item range outside the bank");
    DeviceGuard g(b->device);
    // src may be a device or a (pinned / pageable) host pointer: the copy kind is resolved by the runtime
    SKY_CUDA(cudaMemcpyAsync(reinterpret_cast<float*>(b->data) + static_cast<size_t>(item0) * b->D, src,
                             static_cast<size_t>(n_items) * b->D * sizeof(float), cudaMemcpyDefault,
                             static_cast<cudaStream_t>(stream)));
    return SKY_OK;
}

static int pixel_impl(sky_bank_t* b, const float* q, const unsigned char* qmask, int Q, int k, int64_t idx_offset,
                      float* out_scores, int64_t* out_idx, int64_t item0, int64_t n_items, float* emit, cudaStream_t st) {
    if (!b || !q) return set_error(SKY_ERR_ARG, "NULL argument");
    if (!b->pixel) return set_error(SKY_ERR_STATE, "not a pixel bank");
    if (Q < 1) return set_error(SKY_ERR_ARG, "Q must be positive");
    const int64_t nq = static_cast<int64_t>(Q) * b->D;
    const size_t qp_bytes = static_cast<size_t>(round_up(nq * static_cast<int64_t>(sizeof(float)), 256));
    int rc = ensure_ws2(b, qp_bytes + static_cast<size_t>(Q) * sizeof(int) + 256);
    if (rc) return rc;
    float* qp = reinterpret_cast<float*>(b->ws2);
    int* excl = reinterpret_cast<int*>(reinterpret_cast<unsigned char*>(b->ws2) + qp_bytes);   // per query: pixels excluded?
    rc = launch_pixel_fold(q, qmask, nq, b->D, qp, excl, st);
    if (rc) return rc;
    const int qc = pixel_pick_qc(Q);
    SearchState s{};
    if (emit) {
        if (n_items == 0) return SKY_OK;
        return launch_pixel_search(b, qp, excl, Q, item0, item0 + n_items, s, pixel_grid(b, n_items, qc), qc, emit, st);
    }
    const int grid = pixel_grid(b, b->rows, qc);
    rc = plan_state(b, grid, Q, k, grid, &s);
    if (rc) return rc;
    rc = launch_init_state(s, grid, st);
    if (rc) return rc;
    if (b->rows > 0) {
        rc = launch_pixel_search(b, qp, excl, Q, 0, b->rows, s, grid, qc, nullptr, st);
        if (rc) return rc;
    }
    return launch_merge_lists(s, SKY_MSE, idx_offset, out_scores, out_idx, st);
}

int sky_search_pixels(sky_bank_t* b, const float* q, const unsigned char* qmask, int Q, int k, int64_t idx_offset,
                      float* out_scores, int64_t* out_idx, void* stream) {
    if (!b || !out_scores || !out_idx) return set_error(SKY_ERR_ARG, "NULL argument");
    if (k < 1) return set_error(SKY_ERR_ARG, "k must be positive");
    DeviceGuard g(b->device);
    return pixel_impl(b, q, qmask, Q, k, idx_offset, out_scores, out_idx, 0, 0, nullptr, static_cast<cudaStream_t>(stream));
}

int sky_score_pixels(sky_bank_t* b, const float* q, const unsigned char* qmask, int Q, int64_t item0, int64_t n_items,
                     float* out_scores, void* stream) {
    if (!b || !out_scores) return set_error(SKY_ERR_ARG, "NULL argument");
    if (item0 < 0 || n_items < 0 || item0 + n_items > b->n_items) return set_error(SKY_ERR_ARG, "item range outside the bank");
    DeviceGuard g(b->device);
    return pixel_impl(b, q, qmask, Q, 1, 0, nullptr, nullptr, item0, n_items, out_scores, static_cast<cudaStream_t>(stream));
}

static int device_of(const void* ptr, int device, const char* what) {
    if (device >= 0) return device;
    cudaPointerAttributes pa;
    if (cudaPointerGetAttributes(&pa, ptr) != cudaSuccess || pa.type != cudaMemoryTypeDevice) {
        cudaGetLastError();
        return set_error(SKY_ERR_ARG, "%s is not a device pointer (there is no CPU fallback)", what);
    }
    return pa.device;
}

int sky_pixel_snr(const float* cutouts, int64_t n_items, int C, int H, int W, int n_central_pix, int n_min_channels,
                  float* out_snr, float* out_min, int device, void* stream) {
    if (!cutouts || (!out_snr && !out_min)) return set_error(SKY_ERR_ARG, "NULL argument");
    if (n_items < 0 || C < 1 || H < 1 || W != H) return set_error(SKY_ERR_ARG, "bad cutout shape [%lld, %d, %d, %d] (square images)", (long long)n_items, C, H, W);
    if (n_central_pix < 1 || n_central_pix >= H) return set_error(SKY_ERR_ARG, "n_central_pix=%d must be in [1, %d)", n_central_pix, H);
    if (n_items > 0x7FFFFFFFll) return set_error(SKY_ERR_UNSUPPORTED, "at most 2^31-1 cutouts per call");
    const int dev = device_of(cutouts, device, "cutouts");
    if (dev < 0) return dev;
    DeviceGuard g(dev);
    if (!g.ok) return set_error(SKY_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    count_launch(n_items > 0 ? 1 : 0);
    return launch_snr(cutouts, n_items, C, H, W, n_central_pix, n_min_channels < C ? n_min_channels : C, out_snr, out_min,
                      static_cast<cudaStream_t>(stream));
}

int sky_pixel_bank_snr(const sky_bank_t* b, int C, int H, int W, int64_t item0, int64_t n_items, int n_central_pix,
                       int n_min_channels, float* out_snr, float* out_min, void* stream) {
    if (!b) return set_error(SKY_ERR_ARG, "bank is NULL");
    if (!b->pixel) return set_error(SKY_ERR_STATE, "not a pixel bank");
    if (static_cast<int64_t>(C) * H * W != b->D) return set_error(SKY_ERR_ARG, "C*H*W=%lld differs from the bank's %d pixels per cutout", (long long)C * H * W, b->D);
    if (item0 < 0 || n_items < 0 || item0 + n_items > b->n_items) return set_error(SKY_ERR_ARG, "item range outside the bank");
    return sky_pixel_snr(reinterpret_cast<const float*>(b->data) + static_cast<size_t>(item0) * b->D, n_items, C, H, W,
                         n_central_pix, n_min_channels, out_snr, out_min, b->device, stream);
}

int sky_tile_cutouts(const float* tile, int C, int H, int W, const int32_t* coords, int64_t n, int size, float pixel_min,
                     float pixel_max, float* out, int device, void* stream) {
    if (!tile || (!coords && n > 0) || (!out && n > 0)) return set_error(SKY_ERR_ARG, "NULL argument");
    if (C < 1 || size < 1 || H < size || W < size || n < 0) return set_error(SKY_ERR_ARG, "bad tile / cutout shape");
    const int dev = device_of(tile, device, "tile");
    if (dev < 0) return dev;
    DeviceGuard g(dev);
    if (!g.ok) return set_error(SKY_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    count_launch(n > 0 ? 1 : 0);
    return launch_tile_cutouts(tile, C, H, W, coords, n, size, pixel_min, pixel_max, out, static_cast<cudaStream_t>(stream));
}

int sky_center_clip(const float* src, int64_t n, int C, int Hs, int Ws, int size, float pixel_min, float pixel_max,
                    float* out, int device, void* stream) {
    if ((!src || !out) && n > 0) return set_error(SKY_ERR_ARG, "NULL argument");
    if (C < 1 || size < 1 || Hs < size || Ws < size || n < 0) return set_error(SKY_ERR_ARG, "bad cutout shape");
    if (n == 0) return SKY_OK;
    const int dev = device_of(src, device, "src");
    if (dev < 0) return dev;
    DeviceGuard g(dev);
    if (!g.ok) return set_error(SKY_ERR_CUDA, "cudaSetDevice(%d) failed", dev);
    count_launch(1);
    return launch_center_clip(src, n, C, Hs, Ws, size, pixel_min, pixel_max, out, static_cast<cudaStream_t>(stream));
}

int sky_merge_candidates(const float* scores, const int64_t* idx, int R, int Q, int k_in, int k_out, int metric,
                         float* out_scores, int64_t* out_idx, int device, void* stream) {
    if (!scores || !idx || !out_scores || !out_idx) return set_error(SKY_ERR_ARG, "NULL argument");
    if (!valid_metric(metric)) return set_error(SKY_ERR_ARG, "unknown metric %d", metric);
    if (R < 1 || Q < 0 || k_in < 1 || k_out < 1) return set_error(SKY_ERR_ARG, "bad merge shape");
    DeviceGuard g(device);
    if (!g.ok) return set_error(SKY_ERR_CUDA, "cudaSetDevice(%d) failed: no CUDA device, and there is no CPU fallback", device);
    return launch_merge_candidates(scores, idx, R, Q, k_in, 0, 0, k_out, metric, out_scores, out_idx, static_cast<cudaStream_t>(stream));
}

int sky_merge_candidates_strided(const float* scores, const int64_t* idx, int R, int Q, int k_in, int64_t stride_scores,
                                 int64_t stride_idx, int k_out, int metric, float* out_scores, int64_t* out_idx, int device,
                                 void* stream) {
    if (!scores || !idx || !out_scores || !out_idx) return set_error(SKY_ERR_ARG, "NULL argument");
    if (!valid_metric(metric)) return set_error(SKY_ERR_ARG, "unknown metric %d", metric);
    if (R < 1 || Q < 0 || k_in < 1 || k_out < 1) return set_error(SKY_ERR_ARG, "bad merge shape");
    if (stride_scores < static_cast<int64_t>(Q) * k_in || stride_idx < static_cast<int64_t>(Q) * k_in)
        return set_error(SKY_ERR_ARG, "rank strides must cover Q * k_in elements");
    DeviceGuard g(device);
    if (!g.ok) return set_error(SKY_ERR_CUDA, "cudaSetDevice(%d) failed: no CUDA device, and there is no CPU fallback", device);
    return launch_merge_candidates(scores, idx, R, Q, k_in, stride_scores, stride_idx, k_out, metric, out_scores, out_idx,
                                   static_cast<cudaStream_t>(stream));
}

}  // extern "C"
