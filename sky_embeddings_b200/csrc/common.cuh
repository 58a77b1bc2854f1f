// common.cuh -- shared host/device definitions of libskysearch (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/sky_search.h"

namespace sky {

constexpr int kPruneSlack = 128;   // max candidate inserts per query between two prune points
constexpr int kTileRows = 128;     // bank rows per tensor tile (UMMA M)
constexpr int kKBlock = 64;        // bf16 elements per 128-byte swizzle row (one TMA box / K step)

// ---------------------------------------------------------------------------------------------
// Order-preserving score keys.  A candidate is one 64-bit "composite":
//     high 32 bits = key   (larger key = better score, NaN handled as torch.argsort does:
//                           NaN is the LARGEST value -- first for cosine, last for MSE/MAE,
//                           reference utils/similarity.py:24,29)
//     low  32 bits = ~idx  (so that among equal keys the LOWER bank index wins)
// Composite 0 never occurs for a real candidate (idx < 2^32-1), so 0 means "empty".
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t float_to_ordered(float f) {
#ifdef __CUDA_ARCH__
    uint32_t b = __float_as_uint(f);
#else
    union { float f; uint32_t u; } c; c.f = f; uint32_t b = c.u;
#endif
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_float(uint32_t o) {
    uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#ifdef __CUDA_ARCH__
    return __uint_as_float(b);
#else
    union { float f; uint32_t u; } c; c.u = b; return c.f;
#endif
}
__host__ __device__ __forceinline__ uint32_t score_to_key(float s, bool largest) {
    uint32_t o = (s != s) ? 0xFFFFFFFFu : float_to_ordered(s);
    return largest ? o : ~o;
}
__host__ __device__ __forceinline__ float key_to_score(uint32_t key, bool largest) {
    uint32_t o = largest ? key : ~key;
    if (o == 0xFFFFFFFFu) {
#ifdef __CUDA_ARCH__
        return __uint_as_float(0x7FC00000u);
#else
        union { float f; uint32_t u; } c; c.u = 0x7FC00000u; return c.f;
#endif
    }
    return ordered_to_float(o);
}
__host__ __device__ __forceinline__ uint64_t make_composite(uint32_t key, uint32_t idx) {
    return (static_cast<uint64_t>(key) << 32) | static_cast<uint64_t>(0xFFFFFFFFu - idx);
}
__host__ __device__ __forceinline__ uint32_t composite_key(uint64_t c) { return static_cast<uint32_t>(c >> 32); }
__host__ __device__ __forceinline__ uint32_t composite_idx(uint64_t c) { return 0xFFFFFFFFu - static_cast<uint32_t>(c); }

__host__ __device__ __forceinline__ bool metric_largest(int metric) { return metric == SKY_COSINE; }

__host__ __device__ __forceinline__ int64_t round_up(int64_t a, int64_t b) { return (a + b - 1) / b * b; }

// ---------------------------------------------------------------------------------------------
// Bank layout in HBM: TILE-MAJOR.  Rows are grouped into tiles of 128, columns into k-blocks of
// 64 elements; each (tile, k-block) is one contiguous [128 rows][64 elements] block, so that a
// TMA box of the tensor path is a single contiguous 16 KB (bf16) read instead of 128 scattered
// 128-byte bursts:
//     element (row, d)  ->  ((row / 128) * KB + d / 64) * 8192 + (row % 128) * 64 + d % 64
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t tile_row_base(int64_t row, int kblocks) {
    return static_cast<size_t>(row >> 7) * kblocks * 8192 + static_cast<size_t>(row & 127) * 64;
}
__host__ __device__ __forceinline__ size_t tile_col_off(int d) { return static_cast<size_t>(d >> 6) * 8192 + (d & 63); }
__host__ __device__ __forceinline__ size_t tile_offset(int64_t row, int d, int kblocks) {
    return tile_row_base(row, kblocks) + tile_col_off(d);
}

// ---------------------------------------------------------------------------------------------
// Experiment knobs (SKY_*_DEBUG / _STAGES / _POLICY ... environment variables, kernel timelines).  They exist only
// in a build made with SKY_NVCC_DEFS=-DSKY_EXPERIMENTS: the release library never reads the environment, so a stray
// variable cannot change a result (several of the debug bits skip parts of the math), and the debug branches and
// trace buffers are compiled out of the kernels.
// ---------------------------------------------------------------------------------------------
#ifdef SKY_EXPERIMENTS
#define SKY_DBG(p) ((p).debug)
int env_knob(const char* name, int dflt);     // api.cu
#else
#define SKY_DBG(p) 0
static inline int env_knob(const char*, int dflt) { return dflt; }
#endif

// ---------------------------------------------------------------------------------------------
// host-side error plumbing (api.cu owns the storage)
// ---------------------------------------------------------------------------------------------
int set_error(int code, const char* fmt, ...);
void count_launch(int n = 1);

#define SKY_CUDA(expr)                                                                        \
    do {                                                                                      \
        cudaError_t _e = (expr);                                                              \
        if (_e != cudaSuccess)                                                                \
            return ::sky::set_error(SKY_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,             \
                                    cudaGetErrorString(_e), __FILE__, __LINE__);              \
    } while (0)

#define SKY_LAUNCH_CHECK(name)                                                                \
    do {                                                                                      \
        ::sky::count_launch();                                                                \
        cudaError_t _e = cudaGetLastError();                                                  \
        if (_e != cudaSuccess)                                                                \
            return ::sky::set_error(SKY_ERR_CUDA, "launch of %s failed: %s", name,            \
                                    cudaGetErrorString(_e));                                  \
    } while (0)

// Launch `kernel` so that it may start while the previous kernel of the stream drains (see ptx::griddep_wait: the
// kernel MUST call it before touching anything its predecessor wrote).  pdl = false: an ordinary launch.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_maybe_pdl(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
constexpr int kUsePdl = 1;      // experiments: SKY_PDL=0 launches everything the ordinary way

}  // namespace sky
