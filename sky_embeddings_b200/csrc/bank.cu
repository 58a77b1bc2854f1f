// bank.cu -- K0: building the device-resident, normalised embedding bank, and query preparation.
//
// Reference (utils/similarity.py): token select / max-pool :87-95, first-batch mean / unbiased
// std :98-100, (x - mean) / (std + 1e-8) :101-102 -- done there once per batch per search, here
// once per bank.  determine_target_features :134-147 for the query side.
#include "bank.cuh"

namespace sky {

__device__ __forceinline__ float load_src(const void* src, int dtype, size_t i) {
    return dtype == SKY_F32 ? reinterpret_cast<const float*>(src)[i]
                            : __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(src)[i]);
}

// NaN-propagating max, like torch.max (utils/similarity.py:95)
__device__ __forceinline__ float max_nan(float a, float b) {
    return (a != a || b != b) ? __uint_as_float(0x7FC00000u) : fmaxf(a, b);
}

// One warp per output row, written in the tile-major layout; columns [D, Dp) are zero-filled.
template <typename BankT>
__global__ void ingest_kernel(const void* __restrict__ src, int src_dtype, int64_t n_items, int src_tokens,
                              int token_mode, int num_extra, int L, int D, int Dp,
                              const float* __restrict__ mu, const float* __restrict__ sp,
                              BankT* __restrict__ dst, float* __restrict__ rownorm, int64_t dst_row0) {
    const int lane = threadIdx.x & 31;
    const int64_t warp = (static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x) >> 5;
    const int64_t nrows = n_items * L;
    if (warp >= nrows) return;
    const int64_t item = warp / L;
    const int l = static_cast<int>(warp - item * L);
    const size_t item_base = static_cast<size_t>(item) * src_tokens * D;
    int tok0 = l, ntok = 1;
    if (token_mode == SKY_TOK_CLS) tok0 = 0;
    else if (token_mode == SKY_TOK_PATCHES) tok0 = num_extra + l;
    else if (token_mode == SKY_TOK_MAXPOOL) { tok0 = num_extra; ntok = src_tokens - num_extra; }
    BankT* out = dst + tile_row_base(dst_row0 + warp, Dp / kKBlock);
    float ss = 0.f;
    for (int d = lane; d < Dp; d += 32) {
        float z = 0.f;
        if (d < D) {
            float x = load_src(src, src_dtype, item_base + static_cast<size_t>(tok0) * D + d);
            for (int t = 1; t < ntok; ++t)
                x = max_nan(x, load_src(src, src_dtype, item_base + static_cast<size_t>(tok0 + t) * D + d));
            z = mu ? (x - mu[d]) / sp[d] : x;
        }
        float stored;
        if constexpr (sizeof(BankT) == 2) {
            __nv_bfloat16 b = __float2bfloat16_rn(z);
            out[tile_col_off(d)] = b;
            stored = __bfloat162float(b);
        } else {
            out[tile_col_off(d)] = z;
            stored = z;
        }
        ss = fmaf(stored, stored, ss);
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if (lane == 0) rownorm[dst_row0 + warp] = ss;
}

// Column statistics over rows: mean and unbiased std (two passes, fp64 accumulation).
// Optional on-the-fly normalisation of the input with (mu_in, sp_in).
// grid = ceil(D/32), block = (32, 8).
__global__ void col_stats_kernel(const float* __restrict__ x, int64_t n_rows, int D, int ld, int tiled_kblocks,
                                 const float* __restrict__ mu_in, const float* __restrict__ sp_in,
                                 float* __restrict__ mean_out, float* __restrict__ std_out) {
    __shared__ double red[8][33];
    const int d = blockIdx.x * 32 + threadIdx.x;
    const bool ok = d < D;
    const float m_in = (ok && mu_in) ? mu_in[d] : 0.f;
    const float s_in = (ok && sp_in) ? sp_in[d] : 1.f;
    const bool nrm = mu_in != nullptr;
    double s = 0.0;
    if (ok)
        for (int64_t r = threadIdx.y; r < n_rows; r += 8) {
            float v = tiled_kblocks ? x[tile_offset(r, d, tiled_kblocks)] : x[static_cast<size_t>(r) * ld + d];
            if (nrm) v = (v - m_in) / s_in;
            s += static_cast<double>(v);
        }
    red[threadIdx.y][threadIdx.x] = s;
    __syncthreads();
    double mean = 0.0;
    for (int i = 0; i < 8; ++i) mean += red[i][threadIdx.x];
    mean /= static_cast<double>(n_rows);
    __syncthreads();
    double ss = 0.0;
    if (ok)
        for (int64_t r = threadIdx.y; r < n_rows; r += 8) {
            float v = tiled_kblocks ? x[tile_offset(r, d, tiled_kblocks)] : x[static_cast<size_t>(r) * ld + d];
            if (nrm) v = (v - m_in) / s_in;
            double dv = static_cast<double>(v) - mean;
            ss += dv * dv;
        }
    red[threadIdx.y][threadIdx.x] = ss;
    __syncthreads();
    if (threadIdx.y == 0 && ok) {
        double tot = 0.0;
        for (int i = 0; i < 8; ++i) tot += red[i][threadIdx.x];
        mean_out[d] = static_cast<float>(mean);
        // n_rows == 1 -> 0/0 = NaN, like torch.std(unbiased=True)
        std_out[d] = static_cast<float>(sqrt(tot / static_cast<double>(n_rows - 1)));
    }
}

__global__ void add_eps_kernel(const float* __restrict__ sigma, float* __restrict__ sp, int D) {
    int d = blockIdx.x * blockDim.x + threadIdx.x;
    if (d < D) sp[d] = sigma[d] + 1e-8f;   // utils/similarity.py:101-102, evaluated in fp32 like torch
}

// w = 1/std^2, normalised to sum 1 (utils/similarity.py:143-145); ones if !use_weights (:246-247).
// Single block.  std_in may alias w_out (in-place use by sky_query_from_targets): no __restrict__ here.
__global__ void finish_weights_kernel(const float* std_in, int D, int use_weights, float* w_out) {
    __shared__ double red[32];
    double s = 0.0;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float sd = std_in[d];
        float w = 1.0f / (sd * sd);
        w_out[d] = w;
        s += static_cast<double>(w);
    }
    for (int off = 16; off > 0; off >>= 1) s += __shfl_xor_sync(0xffffffffu, s, off);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < (blockDim.x >> 5); ++i) tot += red[i];
    const float totf = static_cast<float>(tot);
    for (int d = threadIdx.x; d < D; d += blockDim.x) w_out[d] = use_weights ? w_out[d] / totf : 1.0f;
}

template <typename BankT>
__global__ void download_kernel(const BankT* __restrict__ data, int64_t row0, int64_t nrows, int D, int Dp,
                                float* __restrict__ dst) {
    int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= nrows * D) return;
    int64_t r = i / D;
    int d = static_cast<int>(i - r * D);
    BankT v = data[tile_offset(row0 + r, d, Dp / kKBlock)];
    if constexpr (sizeof(BankT) == 2) dst[i] = __bfloat162float(v);
    else dst[i] = v;
}

// ---------------------------------------------------------------------------------------------
// host launchers
// ---------------------------------------------------------------------------------------------
int launch_ingest(const void* src, int src_dtype, int64_t n_items, int src_tokens, int token_mode,
                  int num_extra, int L, int D, int Dp, const float* mu, const float* sp, void* dst,
                  int dst_dtype, float* rownorm, int64_t dst_row0, cudaStream_t st) {
    const int64_t nrows = n_items * L;
    if (nrows == 0) return SKY_OK;
    const int threads = 256;
    const int64_t blocks = (nrows * 32 + threads - 1) / threads;
    if (dst_dtype == SKY_BF16)
        ingest_kernel<__nv_bfloat16><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
            src, src_dtype, n_items, src_tokens, token_mode, num_extra, L, D, Dp, mu, sp,
            reinterpret_cast<__nv_bfloat16*>(dst), rownorm, dst_row0);
    else
        ingest_kernel<float><<<static_cast<unsigned>(blocks), threads, 0, st>>>(
            src, src_dtype, n_items, src_tokens, token_mode, num_extra, L, D, Dp, mu, sp,
            reinterpret_cast<float*>(dst), rownorm, dst_row0);
    SKY_LAUNCH_CHECK("ingest_kernel");
    return SKY_OK;
}

int launch_col_stats(const float* x, int64_t n_rows, int D, int ld, int tiled_kblocks, const float* mu_in,
                     const float* sp_in, float* mean_out, float* std_out, cudaStream_t st) {
    dim3 block(32, 8);
    col_stats_kernel<<<(D + 31) / 32, block, 0, st>>>(x, n_rows, D, ld, tiled_kblocks, mu_in, sp_in, mean_out, std_out);
    SKY_LAUNCH_CHECK("col_stats_kernel");
    return SKY_OK;
}

int launch_add_eps(const float* sigma, float* sp, int D, cudaStream_t st) {
    add_eps_kernel<<<(D + 255) / 256, 256, 0, st>>>(sigma, sp, D);
    SKY_LAUNCH_CHECK("add_eps_kernel");
    return SKY_OK;
}

int launch_finish_weights(const float* std_in, int D, int use_weights, float* w_out, cudaStream_t st) {
    finish_weights_kernel<<<1, 256, 0, st>>>(std_in, D, use_weights, w_out);
    SKY_LAUNCH_CHECK("finish_weights_kernel");
    return SKY_OK;
}

int launch_download(const void* data, int dtype, int64_t row0, int64_t nrows, int D, int Dp, float* dst,
                    cudaStream_t st) {
    const int64_t n = nrows * D;
    if (n == 0) return SKY_OK;
    const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
    if (dtype == SKY_BF16)
        download_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(reinterpret_cast<const __nv_bfloat16*>(data),
                                                              row0, nrows, D, Dp, dst);
    else
        download_kernel<float><<<blocks, 256, 0, st>>>(reinterpret_cast<const float*>(data), row0, nrows, D, Dp, dst);
    SKY_LAUNCH_CHECK("download_kernel");
    return SKY_OK;
}

}  // namespace sky
