// simt_search.cu -- K1: HBM-streaming warp-reduction scorer with fused top-k (CUDA cores).
//
// Direct-form evaluation of the reference metrics, any bank dtype, per-query weights, any L:
//   cosine  utils/similarity.py:163-170   sum(w t x) / (sqrt(sum w t^2) sqrt(sum w x^2) + 1e-6)
//   MSE     utils/similarity.py:188-192   sum(w (t-x)^2) / (sum w) / D
//   MAE     utils/similarity.py:208-212   sum(w |t-x|)   / (sum w) / D
//   n_top_sims :257-259, combine over the L tokens of an item :262-267, running top-k :18-35.
// The regime for this kernel is small query batches (the reference itself always has Q = 1):
// each warp streams 4 bank rows at a time with 128-bit loads, the (<= QC) query vectors live in
// shared memory, and candidates go straight into the CTA's sink -- no [Q, N] score matrix.
#include "bank.cuh"
#include "topk.cuh"

namespace sky {

constexpr int kSimtWarps = 8;
constexpr int kSimtThreads = kSimtWarps * 32;
constexpr int kRowsPerIter = 4;

struct SimtParams {
    const void* bank;
    int64_t row0;       // first bank row of item 0 (sky_score works on a sub-range)
    int64_t n_items;
    int L, D, Dp;
    const float* t;
    const float* w;     // may be null (ones)
    int q0, nq;         // queries [q0, q0+nq) handled by this launch
    int metric, combine, n_top;
    // sink
    uint64_t* lists; int* counts; uint32_t* gtop;
    int p_stride, Qtot, cap, k, use_gtau;
    // emit mode
    float* emit; int64_t item0; int64_t emit_n;
};

__device__ __forceinline__ float nan_f() { return __uint_as_float(0x7FC00000u); }

template <typename BankT>
__device__ __forceinline__ void load8(const BankT* p, float (&x)[8]);

template <>
__device__ __forceinline__ void load8<float>(const float* p, float (&x)[8]) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(p));
    const float4 b = __ldg(reinterpret_cast<const float4*>(p) + 1);
    x[0] = a.x; x[1] = a.y; x[2] = a.z; x[3] = a.w; x[4] = b.x; x[5] = b.y; x[6] = b.z; x[7] = b.w;
}
template <>
__device__ __forceinline__ void load8<__nv_bfloat16>(const __nv_bfloat16* p, float (&x)[8]) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(p));
    const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        x[2 * i] = __uint_as_float(u[i] << 16);
        x[2 * i + 1] = __uint_as_float(u[i] & 0xFFFF0000u);
    }
}

template <typename BankT, int METRIC, int QC>
__global__ void __launch_bounds__(kSimtThreads) simt_search_kernel(const SimtParams p) {
    constexpr int NC = (METRIC == SKY_COSINE) ? 2 : 1;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int Dp = p.Dp;
    float* sA = reinterpret_cast<float*>(smem_raw);            // [QC][Dp] cosine: w*t, else t
    float* sW = sA + QC * Dp;                                  // [QC][Dp]
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>(sW + QC * Dp);   // [QC]
    float* sQc = reinterpret_cast<float*>(sThr + QC);          // [QC] cosine: |t|_w, else sum(w)
    float* sThrF = sQc + QC;                                   // [QC] (unused by this kernel's exact test)
    int* sCnt = reinterpret_cast<int*>(sThrF + QC);            // [QC]
    uint32_t* sLmax = reinterpret_cast<uint32_t*>(sCnt + QC);  // [QC]
    uint32_t* sHist = sLmax + QC;                              // [warps][256]
    float* sTok = reinterpret_cast<float*>(sHist + kSimtWarps * 256);   // [warps][QC][L] iff n_top

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool largest = (METRIC == SKY_COSINE);
    const BankT* bank = reinterpret_cast<const BankT*>(p.bank);

    // ---- stage the query operands -----------------------------------------------------------
    for (int i = tid; i < QC * Dp; i += kSimtThreads) {
        const int q = i / Dp, d = i - q * Dp;
        float tv = 0.f, wv = 0.f;
        if (q < p.nq && d < p.D) {
            tv = p.t[static_cast<size_t>(p.q0 + q) * p.D + d];
            wv = p.w ? p.w[static_cast<size_t>(p.q0 + q) * p.D + d] : 1.0f;
        }
        sA[i] = (METRIC == SKY_COSINE) ? wv * tv : tv;
        sW[i] = wv;
    }
    if (tid < QC) { sThr[tid] = (tid < p.nq) ? 0ull : ~0ull; sThrF[tid] = 0.f; sCnt[tid] = 0; sLmax[tid] = 0; }
    __syncthreads();
    if (warp < QC) {
        double acc = 0.0;
        for (int d = lane; d < p.D; d += 32) {
            if (METRIC == SKY_COSINE) {
                // w t^2 (the reference squares t after the product with w: weights * target ** 2)
                const float wv = sW[warp * Dp + d];
                const float tv = (warp < p.nq) ? p.t[static_cast<size_t>(p.q0 + warp) * p.D + d] : 0.f;
                acc += static_cast<double>(wv * (tv * tv));
            } else {
                acc += static_cast<double>(sW[warp * Dp + d]);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) sQc[warp] = (METRIC == SKY_COSINE) ? sqrtf(static_cast<float>(acc)) : static_cast<float>(acc);
    }
    __syncthreads();

    Sink sink;
    sink.lists = p.lists ? p.lists + (static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0) * p.cap : nullptr;
    sink.thr = smem_addr(sThr); sink.thr_f = smem_addr(sThrF); sink.cnt = smem_addr(sCnt); sink.lmax = smem_addr(sLmax);
    sink.cap = p.cap; sink.k = p.k; sink.largest = largest;
    const bool emit = p.emit != nullptr;

    const int L = p.L;
    const bool flat = (L == 1);
    const int64_t total_rows = p.n_items * L;
    const int64_t n_groups = flat ? (p.n_items + kRowsPerIter - 1) / kRowsPerIter : p.n_items;
    const int64_t groups_per_round = static_cast<int64_t>(gridDim.x) * kSimtWarps;
    const int64_t n_rounds = (n_groups + groups_per_round - 1) / groups_per_round;
    const int inserts_per_round = kSimtWarps * (flat ? kRowsPerIter : 1);
    const int check_every = kPruneSlack / inserts_per_round;
    const float invD = 1.0f / static_cast<float>(p.D);
    int check_idx = 0;

    for (int64_t round = 0; round < n_rounds; ++round) {
        const int64_t group = (round * gridDim.x + blockIdx.x) * kSimtWarps + warp;
        if (group < n_groups) {
            const int64_t item_first = flat ? group * kRowsPerIter : group;
            const int n_tok_iters = flat ? 1 : (L + kRowsPerIter - 1) / kRowsPerIter;
            float cmb[QC];      // combine state over tokens (L > 1)
            bool cnan[QC];
#pragma unroll
            for (int q = 0; q < QC; ++q) { cmb[q] = (p.combine == SKY_MIN) ? INFINITY : (p.combine == SKY_MAX ? -INFINITY : 0.f); cnan[q] = false; }

            float s[kRowsPerIter][QC];
            for (int it = 0; it < n_tok_iters; ++it) {
                const int64_t row0 = flat ? item_first : item_first * L + static_cast<int64_t>(it) * kRowsPerIter;
                const BankT* rp[kRowsPerIter];     // tile-major: row base, then + tile_col_off(d)
#pragma unroll
                for (int r = 0; r < kRowsPerIter; ++r) {
                    int64_t row = row0 + r;
                    if (row >= total_rows) row = total_rows - 1;   // clamp; result masked below
                    rp[r] = bank + tile_row_base(p.row0 + row, Dp / kKBlock);
                }
                float acc[kRowsPerIter][QC][NC];
#pragma unroll
                for (int r = 0; r < kRowsPerIter; ++r)
#pragma unroll
                    for (int q = 0; q < QC; ++q)
#pragma unroll
                        for (int c = 0; c < NC; ++c) acc[r][q][c] = 0.f;

#pragma unroll 2
                for (int d0 = lane * 8; d0 < Dp; d0 += 256) {
                    float x[kRowsPerIter][8];
#pragma unroll
                    for (int r = 0; r < kRowsPerIter; ++r) load8<BankT>(rp[r] + tile_col_off(d0), x[r]);
#pragma unroll
                    for (int q = 0; q < QC; ++q) {
                        float a[8], wv[8];
                        *reinterpret_cast<float4*>(&a[0]) = *reinterpret_cast<const float4*>(&sA[q * Dp + d0]);
                        *reinterpret_cast<float4*>(&a[4]) = *reinterpret_cast<const float4*>(&sA[q * Dp + d0 + 4]);
                        *reinterpret_cast<float4*>(&wv[0]) = *reinterpret_cast<const float4*>(&sW[q * Dp + d0]);
                        *reinterpret_cast<float4*>(&wv[4]) = *reinterpret_cast<const float4*>(&sW[q * Dp + d0 + 4]);
#pragma unroll
                        for (int r = 0; r < kRowsPerIter; ++r)
#pragma unroll
                            for (int v = 0; v < 8; ++v) {
                                const float xv = x[r][v];
                                if (METRIC == SKY_COSINE) {
                                    acc[r][q][0] = fmaf(a[v], xv, acc[r][q][0]);
                                    acc[r][q][1] = fmaf(wv[v] * xv, xv, acc[r][q][1]);
                                } else if (METRIC == SKY_MSE) {
                                    const float dv = a[v] - xv;
                                    acc[r][q][0] = fmaf(wv[v] * dv, dv, acc[r][q][0]);
                                } else {
                                    acc[r][q][0] = fmaf(wv[v], fabsf(a[v] - xv), acc[r][q][0]);
                                }
                            }
                    }
                }
#pragma unroll
                for (int r = 0; r < kRowsPerIter; ++r)
#pragma unroll
                    for (int q = 0; q < QC; ++q) {
#pragma unroll
                        for (int c = 0; c < NC; ++c) {
                            float v = acc[r][q][c];
#pragma unroll
                            for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                            acc[r][q][c] = v;
                        }
                        if (METRIC == SKY_COSINE)
                            s[r][q] = acc[r][q][0] / (sQc[q] * sqrtf(acc[r][q][1]) + 1e-6f);
                        else
                            s[r][q] = acc[r][q][0] / sQc[q] * invD;
                    }
                if (!flat) {
#pragma unroll
                    for (int r = 0; r < kRowsPerIter; ++r) {
                        const int tok = it * kRowsPerIter + r;
                        if (tok < L) {
#pragma unroll
                            for (int q = 0; q < QC; ++q) {
                                const float v = s[r][q];
                                if (p.n_top > 0) {
                                    if (lane == 0) sTok[(warp * QC + q) * L + tok] = v;
                                } else {
                                    cnan[q] |= (v != v);
                                    if (p.combine == SKY_MIN) cmb[q] = fminf(cmb[q], v);
                                    else if (p.combine == SKY_MAX) cmb[q] = fmaxf(cmb[q], v);
                                    else cmb[q] += v;
                                }
                            }
                        }
                    }
                }
            }

            if (flat) {
                // lane j owns (row r = j / QC, query q = j % QC)
                float mine = 0.f;
#pragma unroll
                for (int r = 0; r < kRowsPerIter; ++r)
#pragma unroll
                    for (int q = 0; q < QC; ++q)
                        if (lane == r * QC + q) mine = s[r][q];
                const int r = lane / QC, q = lane % QC;
                const int64_t item = item_first + r;
                if (lane < kRowsPerIter * QC && q < p.nq && item < p.n_items) {
                    if (emit) {
                        if (item >= p.item0 && item < p.item0 + p.emit_n)
                            p.emit[static_cast<size_t>(p.q0 + q) * p.emit_n + (item - p.item0)] = mine;
                    } else {
                        sink_insert_one(sink, q, make_composite(score_to_key(mine, largest), static_cast<uint32_t>(item)));
                    }
                }
            } else {
                if (p.n_top > 0) {
                    // best-n_top token scores per query (torch.topk, utils/similarity.py:259), then combine
                    __syncwarp();
#pragma unroll
                    for (int q = 0; q < QC; ++q) {
                        const float* tk = sTok + (warp * QC + q) * L;
                        float lsum = 0.f, lmin = INFINITY, lmax = -INFINITY;
                        bool lnan = false;
                        for (int i = lane; i < L; i += 32) {
                            const float vi = tk[i];
                            const uint32_t ki = score_to_key(vi, largest);
                            int rank = 0;
                            for (int j = 0; j < L; ++j) {
                                const uint32_t kj = score_to_key(tk[j], largest);
                                rank += (kj > ki) || (kj == ki && j < i);
                            }
                            if (rank < p.n_top) {
                                lnan |= (vi != vi);
                                lsum += vi; lmin = fminf(lmin, vi); lmax = fmaxf(lmax, vi);
                            }
                        }
#pragma unroll
                        for (int off = 16; off > 0; off >>= 1) {
                            lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
                            lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, off));
                            lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
                        }
                        cnan[q] = __any_sync(0xffffffffu, lnan);
                        cmb[q] = (p.combine == SKY_MIN) ? lmin : (p.combine == SKY_MAX ? lmax : lsum / static_cast<float>(p.n_top));
                    }
                    __syncwarp();
                } else if (p.combine == SKY_MEAN) {
#pragma unroll
                    for (int q = 0; q < QC; ++q) cmb[q] = cmb[q] / static_cast<float>(L);
                }
                float mine = 0.f;
#pragma unroll
                for (int q = 0; q < QC; ++q) {
                    const float v = cnan[q] ? nan_f() : cmb[q];
                    if (lane == q) mine = v;
                }
                const int64_t item = item_first;
                if (lane < p.nq) {
                    if (emit) {
                        if (item >= p.item0 && item < p.item0 + p.emit_n)
                            p.emit[static_cast<size_t>(p.q0 + lane) * p.emit_n + (item - p.item0)] = mine;
                    } else {
                        sink_insert_one(sink, lane, make_composite(score_to_key(mine, largest), static_cast<uint32_t>(item)));
                    }
                }
            }
        }

        if (!emit && ((round + 1) % check_every == 0)) {
            __syncthreads();
            sink_prune_if_full(sink, p.nq, warp, kSimtWarps, smem_addr(sHist + warp * 256));
            // exchange the grid-wide bound often early on, then at a decaying rate
            if (p.use_gtau && (check_idx < 8 || (check_idx & (check_idx - 1)) == 0 || (check_idx & 15) == 0))
            {
                // gtop is [cta][query]: publish this CTA's row, then min over CTAs per query column
                if (warp == kSimtWarps - 1) exchange_publish(sink, p.nq, p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0);
                for (int q = warp; q < p.nq; q += kSimtWarps) {
                    const uint32_t lo = exchange_reduce(p.gtop + p.q0 + q, p.p_stride, p.Qtot, p.k);
                    if (lane == 0) exchange_apply(sink, q, lo);
                }
            }
            ++check_idx;
            __syncthreads();
        }
    }

    if (!emit) {
        __syncthreads();
        for (int q = warp; q < p.nq; q += kSimtWarps) {
            if (lane == 0) {
                p.counts[static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q] = sCnt[q];
                if (p.use_gtau && sLmax[q]) st_cg_u32(p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q, sLmax[q]);
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static size_t simt_smem_bytes(int Dp, int qc, int L, int n_top) {
    size_t b = static_cast<size_t>(2) * qc * Dp * sizeof(float);
    b += qc * (sizeof(unsigned long long) + 2 * sizeof(float) + sizeof(int) + sizeof(uint32_t));
    b += kSimtWarps * 256 * sizeof(uint32_t);
    if (n_top > 0) b += static_cast<size_t>(kSimtWarps) * qc * L * sizeof(float);
    return (b + 15) / 16 * 16;
}

int simt_pick_qc(int Dp) {
    // two fp32 operand vectors per query must fit in shared memory next to the sink state
    return (static_cast<size_t>(2) * 4 * Dp * sizeof(float) <= 96 * 1024) ? 4 : 1;
}

template <typename BankT, int METRIC, int QC>
static int simt_config(int Dp, int L, int n_top, int* blocks_per_sm, size_t* smem) {
    *smem = simt_smem_bytes(Dp, QC, L, n_top);
    if (*smem > 227 * 1024) return set_error(SKY_ERR_UNSUPPORTED, "SIMT scorer: D=%d needs %zu B of shared memory", Dp, *smem);
    SKY_CUDA(cudaFuncSetAttribute(simt_search_kernel<BankT, METRIC, QC>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                  static_cast<int>(*smem)));
    SKY_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(blocks_per_sm, simt_search_kernel<BankT, METRIC, QC>,
                                                           kSimtThreads, *smem));
    if (*blocks_per_sm < 1) *blocks_per_sm = 1;
    return SKY_OK;
}

template <typename BankT, int METRIC, int QC>
static int simt_launch_one(const SimtParams& p, int grid, size_t smem, cudaStream_t st) {
    simt_search_kernel<BankT, METRIC, QC><<<grid, kSimtThreads, smem, st>>>(p);
    SKY_LAUNCH_CHECK("simt_search_kernel");
    return SKY_OK;
}

#define SIMT_DISPATCH(FN, dtype, metric, qc, ...)                                                      \
    do {                                                                                               \
        if ((dtype) == SKY_BF16) {                                                                     \
            if ((metric) == SKY_COSINE) { if ((qc) == 4) return FN<__nv_bfloat16, SKY_COSINE, 4>(__VA_ARGS__); else return FN<__nv_bfloat16, SKY_COSINE, 1>(__VA_ARGS__); } \
            if ((metric) == SKY_MSE)    { if ((qc) == 4) return FN<__nv_bfloat16, SKY_MSE, 4>(__VA_ARGS__);    else return FN<__nv_bfloat16, SKY_MSE, 1>(__VA_ARGS__); }    \
            if ((qc) == 4) return FN<__nv_bfloat16, SKY_MAE, 4>(__VA_ARGS__); else return FN<__nv_bfloat16, SKY_MAE, 1>(__VA_ARGS__);                                    \
        } else {                                                                                       \
            if ((metric) == SKY_COSINE) { if ((qc) == 4) return FN<float, SKY_COSINE, 4>(__VA_ARGS__); else return FN<float, SKY_COSINE, 1>(__VA_ARGS__); }               \
            if ((metric) == SKY_MSE)    { if ((qc) == 4) return FN<float, SKY_MSE, 4>(__VA_ARGS__);    else return FN<float, SKY_MSE, 1>(__VA_ARGS__); }                  \
            if ((qc) == 4) return FN<float, SKY_MAE, 4>(__VA_ARGS__); else return FN<float, SKY_MAE, 1>(__VA_ARGS__);                                                  \
        }                                                                                              \
    } while (0)

static int simt_config_dispatch(int dtype, int metric, int qc, int Dp, int L, int n_top, int* bps, size_t* smem) {
    SIMT_DISPATCH(simt_config, dtype, metric, qc, Dp, L, n_top, bps, smem);
}
static int simt_launch_dispatch(int dtype, int metric, int qc, const SimtParams& p, int grid, size_t smem, cudaStream_t st) {
    SIMT_DISPATCH(simt_launch_one, dtype, metric, qc, p, grid, smem, st);
}

// Grid for a search over n_items: persistent CTAs, every CTA owns >= 1 work group.
int simt_grid(const sky_bank* b, int metric, int L, int64_t n_items, int qc, int n_top, int* grid, size_t* smem) {
    int bps = 1;
    int rc = simt_config_dispatch(b->dtype, metric, qc, b->Dp, L, n_top, &bps, smem);
    if (rc) return rc;
    const int64_t n_groups = (L == 1) ? (n_items + kRowsPerIter - 1) / kRowsPerIter : n_items;
    int64_t g = (n_groups + kSimtWarps - 1) / kSimtWarps;
    const int64_t cap = static_cast<int64_t>(b->num_sms) * bps;
    if (g > cap) g = cap;
    if (g < 1) g = 1;
    *grid = static_cast<int>(g);
    return SKY_OK;
}

int launch_simt_search(const sky_bank* b, const SimtArgs& a, const SearchState& s, int grid, int qc, size_t smem,
                       cudaStream_t st) {
    for (int q0 = 0; q0 < a.Q; q0 += qc) {
        SimtParams p;
        p.bank = a.bank; p.row0 = a.row0; p.n_items = a.n_items; p.L = a.L; p.D = a.D; p.Dp = a.Dp;
        p.t = a.t; p.w = a.w; p.q0 = q0; p.nq = (a.Q - q0 < qc) ? (a.Q - q0) : qc;
        p.metric = a.metric; p.combine = a.combine; p.n_top = a.n_top;
        p.lists = s.lists; p.counts = s.counts; p.gtop = s.gtop; p.p_stride = s.p_stride; p.Qtot = s.Qtot;
        p.cap = s.cap; p.k = s.k; p.use_gtau = s.use_gtau;
        p.emit = a.emit; p.item0 = a.item0; p.emit_n = a.n;
        if (!a.emit) prof_mark(b, st);
        int rc = simt_launch_dispatch(a.dtype, a.metric, qc, p, grid, smem, st);
        if (!a.emit) prof_mark(b, st);
        if (rc) return rc;
    }
    return SKY_OK;
}

}  // namespace sky
