// pixel_prep.cu -- K6: the pixel-side steps in front of the bank.
//
//  * snr_kernel: the S/N pre-filter of the reference's h5 search driver.  calculate_snr (reference
//    utils/misc.py:119-163) takes, per image and channel, mean(central n x n pixels) / (std(all other pixels) + 1e-8);
//    similarity_search.py:124-133 then keeps the rows whose nanmin over the first five channels lies inside
//    (snr_lo, snr_hi).  One CTA per image walks its channels: two coalesced passes over a 16 KB plane (the second
//    hits L1/L2), double-precision accumulators; NaN pixels propagate exactly as in numpy (a NaN anywhere in a
//    region makes that channel's S/N NaN, which nanmin then skips).
//  * tile_cutouts_kernel: FITS-tile streaming (reference utils/dataloaders.py:511-536 overlapping_cutouts, :657-661
//    clipping): cutout i = tile[:, h0_i : h0_i + size, w0_i : w0_i + size], clipped; the (h0, w0) list is
//    generate_overlap_coords (:481-509), computed on the host (pure index arithmetic) by sky_embeddings_b200.ingest.
//  * center_clip_kernel: the per-item path of the h5 loader (reference utils/dataloaders.py:291-300): clip at
//    pixel_min / pixel_max (NaN stays NaN: `x < min` is false), central size x size crop when the stored cutout is larger
//    (extract_center, :685-700).
#include "bank.cuh"

namespace sky {

constexpr int kSnrThreads = 256;

__device__ __forceinline__ double block_sum(double v, double* red) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
    __syncthreads();                       // red[] may still be read from the previous call
    if (lane == 0) red[warp] = v;
    __syncthreads();
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < kSnrThreads / 32; ++i) t += red[i];
    return t;
}

__global__ void __launch_bounds__(kSnrThreads)
snr_kernel(const float* __restrict__ img, int64_t n_items, int C, int H, int W, int n_central, int n_min_channels,
           float* __restrict__ out_snr, float* __restrict__ out_min) {
    __shared__ double red[kSnrThreads / 32];
    const int64_t item = blockIdx.x;
    const int start = (H - n_central) / 2, end = start + n_central;      // utils/misc.py:145-146 (img_size = H = W)
    const int plane = H * W;
    const int n_c = n_central * n_central, n_s = plane - n_c;
    float best = __uint_as_float(0x7FC00000u);                             // nanmin of nothing / of all-NaN = NaN
    for (int c = 0; c < C; ++c) {
        const float* p = img + (item * C + c) * static_cast<int64_t>(plane);
        double sc = 0.0, ss = 0.0;
        for (int i = threadIdx.x; i < plane; i += kSnrThreads) {
            const int y = i / W, x = i - y * W;
            const double v = static_cast<double>(__ldg(p + i));
            const bool central = (y >= start) & (y < end) & (x >= start) & (x < end);
            if (central) sc += v; else ss += v;
        }
        sc = block_sum(sc, red);
        ss = block_sum(ss, red);
        const double mean_s = ss / n_s;
        double dev = 0.0;
        for (int i = threadIdx.x; i < plane; i += kSnrThreads) {
            const int y = i / W, x = i - y * W;
            const bool central = (y >= start) & (y < end) & (x >= start) & (x < end);
            if (!central) { const double d = static_cast<double>(__ldg(p + i)) - mean_s; dev += d * d; }
        }
        dev = block_sum(dev, red);
        const float mean_c = static_cast<float>(sc / n_c);
        const float std_s = static_cast<float>(sqrt(dev / n_s));           // np.std: population (ddof = 0)
        const float snr = mean_c / (std_s + 1e-8f);
        if (threadIdx.x == 0 && out_snr) out_snr[item * C + c] = snr;
        if (c < n_min_channels && !isnan(snr)) best = isnan(best) ? snr : fminf(best, snr);
    }
    if (threadIdx.x == 0 && out_min) out_min[item] = best;
}

__device__ __forceinline__ float clip_px(float v, float lo, float hi) {
    // `cutout[cutout < lo] = lo`: comparisons with NaN are false, so NaN pixels (and NaN bounds = no clipping) pass through
    if (v < lo) v = lo;
    if (v > hi) v = hi;
    return v;
}

__global__ void tile_cutouts_kernel(const float* __restrict__ tile, int C, int H, int W, const int* __restrict__ coords,
                                    int64_t n, int size, float lo, float hi, float* __restrict__ out) {
    const int64_t per = static_cast<int64_t>(C) * size * size;
    const int64_t total = n * per;
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < total;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = j / per;
        const int r = static_cast<int>(j - i * per);
        const int c = r / (size * size), yx = r - c * size * size, y = yx / size, x = yx - y * size;
        const int h0 = coords[2 * i], w0 = coords[2 * i + 1];
        out[j] = clip_px(__ldg(tile + (static_cast<int64_t>(c) * H + (h0 + y)) * W + (w0 + x)), lo, hi);
    }
}

__global__ void center_clip_kernel(const float* __restrict__ src, int64_t n, int C, int Hs, int Ws, int size, float lo,
                                   float hi, float* __restrict__ out) {
    const int64_t per = static_cast<int64_t>(C) * size * size;
    const int64_t total = n * per;
    const int r0 = Hs / 2 - size / 2, c0 = Ws / 2 - size / 2;           // extract_center, utils/dataloaders.py:694-696
    for (int64_t j = blockIdx.x * static_cast<int64_t>(blockDim.x) + threadIdx.x; j < total;
         j += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t i = j / per;
        const int r = static_cast<int>(j - i * per);
        const int c = r / (size * size), yx = r - c * size * size, y = yx / size, x = yx - y * size;
        out[j] = clip_px(__ldg(src + ((i * C + c) * Hs + (r0 + y)) * static_cast<int64_t>(Ws) + (c0 + x)), lo, hi);
    }
}

int launch_snr(const float* img, int64_t n_items, int C, int H, int W, int n_central, int n_min_channels, float* out_snr,
               float* out_min, cudaStream_t st) {
    if (n_items == 0) return SKY_OK;
    snr_kernel<<<static_cast<unsigned>(n_items), kSnrThreads, 0, st>>>(img, n_items, C, H, W, n_central, n_min_channels, out_snr, out_min);
    SKY_LAUNCH_CHECK("snr_kernel");
    return SKY_OK;
}

static unsigned flat_grid(int64_t total) {
    const int64_t blocks = (total + 255) / 256;
    return static_cast<unsigned>(blocks < 148 * 32 ? (blocks < 1 ? 1 : blocks) : 148 * 32);
}

int launch_tile_cutouts(const float* tile, int C, int H, int W, const int* coords, int64_t n, int size, float lo, float hi,
                        float* out, cudaStream_t st) {
    if (n == 0) return SKY_OK;
    tile_cutouts_kernel<<<flat_grid(n * C * size * size), 256, 0, st>>>(tile, C, H, W, coords, n, size, lo, hi, out);
    SKY_LAUNCH_CHECK("tile_cutouts_kernel");
    return SKY_OK;
}

int launch_center_clip(const float* src, int64_t n, int C, int Hs, int Ws, int size, float lo, float hi, float* out,
                       cudaStream_t st) {
    if (n == 0) return SKY_OK;
    center_clip_kernel<<<flat_grid(n * C * size * size), 256, 0, st>>>(src, n, C, Hs, Ws, size, lo, hi, out);
    SKY_LAUNCH_CHECK("center_clip_kernel");
    return SKY_OK;
}

}  // namespace sky
