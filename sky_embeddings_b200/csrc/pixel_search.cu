// pixel_search.cu -- K5: pixel-space masked-MSE search over raw cutouts (BASELINE config 5).
//
// The reference has no pixel-space search; SURVEY.md section 8(d) defines it so that the
// reference's own code stays the oracle: weighted_MSE (utils/similarity.py:174-192) with the
// weights replaced by a validity mask, and the NaN handling / normaliser of the MAE loss
// (utils/mim_vit.py:482-486, :509-519):
//     valid = ~isnan(q) & ~isnan(x);  m = valid * qmask;   score = sum m (q - x)^2 / (sum m + 1e-5)
// lower is better.  A cutout is one bank row of D = C*H*W fp32 values (5*64*64 = 20480 -> 80 KB),
// stored row-major WITH its NaNs (they are the validity mask).  The query is folded once into
// q' = q where it takes part, NaN elsewhere: then q' - x is NaN exactly where the pixel is excluded,
// so one compare yields the mask for free.
//
// HBM-bound for a handful of queries.  Same machinery as the streaming scorer (stream_search.cu):
// one producer thread bulk-copies 16 KB pieces of rows into a shared-memory ring, 8 consumer
// warps split every piece, a ninth warp trades grid-wide k-th-best bounds.  Rows are taken in
// groups of 8 and walked piece-major (piece c of all 8 rows, then piece c+1), so each lane keeps
// its slice of q' in registers for 8 rows: q' is re-read from L2 once per 8 rows (1/8 of the bank
// traffic) instead of once per row.
#include <cstdlib>

#include "bank.cuh"
#include "ptx.cuh"
#include "topk.cuh"

namespace sky {

constexpr int kPxChunk = 16384;                    // bytes per stage
constexpr int kPxChunkElems = kPxChunk / 4;
constexpr int kPxMaxStages = 13;
// Consumer shape (template parameters WARPS, ROWS of the kernel): 8 warps x 8 rows per group -- a lane keeps 16 pixels of
// q' per query in registers for 8 rows.  One or two queries per pass run at the HBM roofline.  Four queries per pass
// need 16 instructions per pixel and the SM issues ~0.47 warp instructions per scheduler and cycle whatever the mix
// (measured: integer counts 24.7 ms per 1 M cutouts, float counts 21.3, the PTX form below 21.8; 16 warps x 4 rows with
// half the pixels per lane 22.3 -- more warps do not help, it is not latency): 0.58 of the HBM roofline is the issue limit.

struct PixelParams {
    const unsigned char* bank;   // [rows][D] fp32, row-major
    const float* qp;             // [Q][D] folded queries (NaN = pixel excluded)
    const int* excl;             // [Q] != 0: the folded query excludes pixels (mask or NaN); 0 = every pixel takes part
    int64_t row_lo, row_hi;
    int D, nch;                  // pieces per row
    int q0, nq;
    int stages;
    unsigned long long policy;
    uint64_t* lists; int* counts; uint32_t* gtop;
    int p_stride, Qtot, cap, k, use_gtau;
    float* emit; int64_t emit_item0; int64_t emit_n;
};

// q' = q where (mask != 0 and q is not NaN), NaN elsewhere
__global__ void pixel_fold_query_kernel(const float* __restrict__ q, const unsigned char* __restrict__ mask, int64_t n, int D,
                                        float* __restrict__ qp, int* __restrict__ excl) {
    const int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float v = q[i];
    const float f = (mask == nullptr || mask[i] != 0) ? v : __uint_as_float(0x7FC00000u);
    qp[i] = f;
    if (f != f) atomicOr(excl + i / D, 1);      // rare: one flag per query, read by the search kernel
}

__device__ __forceinline__ void px_bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar)), "l"(policy)
        : "memory");
}

// One pixel of one query on the general path: acc += d * d and n += 1 unless d is NaN.  Written as PTX so that it stays
// FSETP + predicated FFMA + predicated integer add (with the subtraction in front: four instructions per pixel and query,
// two on the fp32 pipe and two on the ALU pipe).  Left to the compiler, an integer count under `if` became add + select
// (five instructions, three of them on the ALU pipe: 0.51 of the HBM roofline at four queries per pass against 0.58).
__device__ __forceinline__ void px_accumulate(float d, float& acc, int& n) {
    asm("{\n\t.reg .pred p;\n\t"
        "setp.eq.f32 p, %2, %2;\n\t"
        "@p fma.rn.f32 %0, %2, %2, %0;\n\t"
        "@p add.s32 %1, %1, 1;\n\t}"
        : "+f"(acc), "+r"(n) : "f"(d));
}

template <int QC, int WARPS, int ROWS>
__global__ void __launch_bounds__(WARPS * 32 + 64, 1) pixel_search_kernel(const PixelParams p) {
    constexpr int kPxWarps = WARPS, kPxRows = ROWS;                 // rows per group <= consumer warps: warp w finishes row w
    constexpr int kPxConsumers = WARPS * 32;
    constexpr int kPxProducerWarp = WARPS, kPxXchgWarp = WARPS + 1; // + producer warp + exchange warp
    constexpr int NJ = kPxChunkElems / (WARPS * 128);               // float4 per lane and piece (4 or 2)
    static_assert(ROWS <= WARPS && NJ * WARPS * 128 == kPxChunkElems, "consumer shape");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const int S = p.stages;
    unsigned char* sStage = base;                                                     // [S][16 KB]
    float* sPart = reinterpret_cast<float*>(sStage + static_cast<size_t>(S) * kPxChunk);   // [warps][rows][QC][2]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sPart + kPxWarps * kPxRows * QC * 2);  // [kPxMaxStages]
    uint64_t* empty_bar = full_bar + kPxMaxStages;
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>(empty_bar + kPxMaxStages);   // [QC]
    float* sThrF = reinterpret_cast<float*>(sThr + QC);
    int* sCnt = reinterpret_cast<int*>(sThrF + QC);
    uint32_t* sLmax = reinterpret_cast<uint32_t*>(sCnt + QC);
    uint32_t* sHist = sLmax + QC;                                                     // [warps][256]
    volatile int* sGroupsDone = reinterpret_cast<volatile int*>(sHist + kPxWarps * 256);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t n_rows = p.row_hi - p.row_lo;
    const int64_t n_groups = (n_rows + kPxRows - 1) / kPxRows;
    const int my_groups = (n_groups > static_cast<int64_t>(blockIdx.x))
                              ? static_cast<int>((n_groups - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;
    const int last_len = p.D - (p.nch - 1) * kPxChunkElems;      // elements in the last piece of a row
    // every query of this pass compares all its pixels (no mask, no NaN): the pixel's validity then depends on the bank
    // alone, one test and one count per pixel serve all queries of the pass
    bool clean = true;
    for (int q = 0; q < p.nq; ++q) clean = clean && (__ldg(p.excl + p.q0 + q) == 0);

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], kPxWarps); }
        ptx::fence_barrier_init();
        *sGroupsDone = 0;
    }
    if (tid < QC) { sThr[tid] = (tid < p.nq) ? 0ull : ~0ull; sThrF[tid] = __uint_as_float(0x7FC00000u); sCnt[tid] = 0; sLmax[tid] = 0; }
    __syncthreads();

    Sink sink;
    sink.lists = p.lists ? p.lists + (static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0) * p.cap : nullptr;
    sink.thr = smem_addr(sThr); sink.thr_f = smem_addr(sThrF); sink.cnt = smem_addr(sCnt); sink.lmax = smem_addr(sLmax);
    sink.cap = p.cap; sink.k = p.k; sink.largest = false;
    const bool emit = p.emit != nullptr;

    if (warp == kPxProducerWarp) {
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t stage0 = ptx::smem_u32(sStage);
            for (int g = 0; g < my_groups; ++g) {
                const int64_t r0 = p.row_lo + (blockIdx.x + static_cast<int64_t>(g) * gridDim.x) * kPxRows;
                const int rows = static_cast<int>(p.row_hi - r0 < kPxRows ? p.row_hi - r0 : kPxRows);
                for (int c = 0; c < p.nch; ++c) {
                    const uint32_t bytes = static_cast<uint32_t>((c == p.nch - 1 ? last_len : kPxChunkElems) * 4);
                    for (int r = 0; r < rows; ++r) {
                        ptx::mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, 32);
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], bytes);
                        px_bulk_load(stage0 + static_cast<uint32_t>(stage) * kPxChunk,
                                     p.bank + (static_cast<size_t>(r0 + r) * p.D + static_cast<size_t>(c) * kPxChunkElems) * 4,
                                     bytes, &full_bar[stage], p.policy);
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    } else if (warp == kPxXchgWarp) {
        if (!emit && p.use_gtau && my_groups > 0) {
            uint32_t* my_row = p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0;
            int round = 0;
            uint32_t published = 0;
            while (*sGroupsDone < my_groups) {
                if (lane < p.nq) {
                    const uint32_t mine = lds_u32(sink.lmax + lane * 4);
                    if (mine != published) { st_cg_u32(my_row + lane, mine); published = mine; }
                }
                for (int q = 0; q < p.nq; ++q) {
                    const uint32_t lo = exchange_reduce(p.gtop + p.q0 + q, p.p_stride, p.Qtot, p.k);
                    if (lane == 0) exchange_apply(sink, q, lo);
                }
                ++round;
                __nanosleep(round < 32 ? 500 : 8000);
            }
        }
    } else if (warp < kPxWarps) {
        const uint32_t stage0 = ptx::smem_u32(sStage);
        // this lane's elements of a piece: NJ x float4 at element warp * NJ * 128 + j * 128 + lane * 4
        const int e0 = warp * (NJ * 128) + lane * 4;
        int stage = 0;
        uint32_t phase = 0;
        for (int g = 0; g < my_groups; ++g) {
            const int64_t r0 = p.row_lo + (blockIdx.x + static_cast<int64_t>(g) * gridDim.x) * kPxRows;
            const int rows = static_cast<int>(p.row_hi - r0 < kPxRows ? p.row_hi - r0 : kPxRows);
            float acc[kPxRows][QC];
            int cntg[kPxRows][QC];       // pixels that took part, per query (clean path: slot 0 serves every query)
#pragma unroll
            for (int r = 0; r < kPxRows; ++r)
#pragma unroll
                for (int q = 0; q < QC; ++q) { acc[r][q] = 0.f; cntg[r][q] = 0; }

            for (int c = 0; c < p.nch; ++c) {
                const int len = (c == p.nch - 1) ? last_len : kPxChunkElems;
                // slice of the folded queries for this piece (L2 / L1 resident), excluded = NaN
                float4 qv[QC][NJ];
#pragma unroll
                for (int q = 0; q < QC; ++q)
#pragma unroll
                    for (int j = 0; j < NJ; ++j) {
                        const int e = e0 + j * 128;
                        if (q < p.nq && e < len)
                            qv[q][j] = __ldg(reinterpret_cast<const float4*>(p.qp + static_cast<size_t>(p.q0 + q) * p.D +
                                                                              static_cast<size_t>(c) * kPxChunkElems + e));
                        else
                            qv[q][j] = make_float4(__uint_as_float(0x7FC00000u), __uint_as_float(0x7FC00000u),
                                                   __uint_as_float(0x7FC00000u), __uint_as_float(0x7FC00000u));
                    }
#pragma unroll
                for (int r = 0; r < kPxRows; ++r) {
                    if (r < rows) {
                        ptx::mbar_wait_relaxed(&full_bar[stage], phase, 20);
                        const uint32_t src = stage0 + static_cast<uint32_t>(stage) * kPxChunk + e0 * 4;
                        float4 xv[NJ];
#pragma unroll
                        for (int j = 0; j < NJ; ++j) {
                            // beyond the end of a short last piece the ring holds stale data: q' is NaN there
                            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                         : "=f"(xv[j].x), "=f"(xv[j].y), "=f"(xv[j].z), "=f"(xv[j].w) : "r"(src + j * 512));
                        }
                        if (clean) {
                            // 2 instructions per pixel and query (subtract, predicated multiply-add) + a test and a
                            // count per pixel, instead of 4 per pixel and query
#pragma unroll
                            for (int j = 0; j < NJ; ++j) {
                                const bool in = e0 + j * 128 < len;      // a short last piece: stale ring data beyond it
                                const bool p0 = in && xv[j].x == xv[j].x, p1 = in && xv[j].y == xv[j].y;
                                const bool p2 = in && xv[j].z == xv[j].z, p3 = in && xv[j].w == xv[j].w;
                                cntg[r][0] += (p0 ? 1 : 0) + (p1 ? 1 : 0) + (p2 ? 1 : 0) + (p3 ? 1 : 0);
#pragma unroll
                                for (int q = 0; q < QC; ++q) {
                                    const float d0 = qv[q][j].x - xv[j].x, d1 = qv[q][j].y - xv[j].y;
                                    const float d2 = qv[q][j].z - xv[j].z, d3 = qv[q][j].w - xv[j].w;
                                    if (p0) acc[r][q] = fmaf(d0, d0, acc[r][q]);
                                    if (p1) acc[r][q] = fmaf(d1, d1, acc[r][q]);
                                    if (p2) acc[r][q] = fmaf(d2, d2, acc[r][q]);
                                    if (p3) acc[r][q] = fmaf(d3, d3, acc[r][q]);
                                }
                            }
                        } else {
#pragma unroll
                            for (int q = 0; q < QC; ++q)
#pragma unroll
                                for (int j = 0; j < NJ; ++j) {
                                    const float d0 = qv[q][j].x - xv[j].x, d1 = qv[q][j].y - xv[j].y;
                                    const float d2 = qv[q][j].z - xv[j].z, d3 = qv[q][j].w - xv[j].w;
                                    // NaN (either side missing or masked out) drops out of both sums
                                    px_accumulate(d0, acc[r][q], cntg[r][q]);
                                    px_accumulate(d1, acc[r][q], cntg[r][q]);
                                    px_accumulate(d2, acc[r][q], cntg[r][q]);
                                    px_accumulate(d3, acc[r][q], cntg[r][q]);
                                }
                        }
                        __syncwarp();
                        if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                }
            }

            // ---- reduce: lanes -> warp partials -> warp w finishes row w --------------------------------
#pragma unroll
            for (int r = 0; r < kPxRows; ++r)
#pragma unroll
                for (int q = 0; q < QC; ++q) {
                    float a = acc[r][q];
                    int n = clean ? cntg[r][0] : cntg[r][q];
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        a += __shfl_xor_sync(0xffffffffu, a, off);
                        n += __shfl_xor_sync(0xffffffffu, n, off);
                    }
                    if (lane == 0) {
                        sPart[((warp * kPxRows + r) * QC + q) * 2] = a;
                        sPart[((warp * kPxRows + r) * QC + q) * 2 + 1] = static_cast<float>(n);
                    }
                }
            ptx::named_bar_sync(1, kPxConsumers);
            if (warp < rows && lane < p.nq) {
                float a = 0.f, n = 0.f;
#pragma unroll
                for (int w = 0; w < kPxWarps; ++w) {
                    a += sPart[((w * kPxRows + warp) * QC + lane) * 2];
                    n += sPart[((w * kPxRows + warp) * QC + lane) * 2 + 1];
                }
                const float score = a / (n + 1e-5f);
                const int64_t row = r0 + warp;
                if (emit) {
                    if (row >= p.emit_item0 && row < p.emit_item0 + p.emit_n)
                        p.emit[static_cast<size_t>(p.q0 + lane) * p.emit_n + (row - p.emit_item0)] = score;
                } else {
                    sink_insert_one(sink, lane, make_composite(score_to_key(score, false), static_cast<uint32_t>(row)));
                }
            }
            ptx::named_bar_sync(1, kPxConsumers);
            if (!emit) {
                sink_prune_if_full(sink, p.nq, warp, kPxWarps, smem_addr(sHist + warp * 256));
                ptx::named_bar_sync(1, kPxConsumers);
            }
            if (tid == 0) *sGroupsDone = g + 1;
        }
        if (!emit) {
            ptx::named_bar_sync(1, kPxConsumers);
            if (tid < p.nq && p.use_gtau && sLmax[tid])
                st_cg_u32(p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + tid, sLmax[tid]);
            ptx::named_bar_sync(1, kPxConsumers);
            // shrink the lists to what can still be in the top-k under the grid-wide bound known now (see stream_search.cu)
            if (warp < p.nq) {
                const int q = warp;
                int n = sCnt[q];
                if (p.use_gtau) {
                    const uint32_t lo = exchange_reduce(p.gtop + p.q0 + q, p.p_stride, p.Qtot, p.k);
                    if (lo != 0u) n = warp_compact_ge(sink.lists + static_cast<size_t>(q) * p.cap, n, static_cast<uint64_t>(lo) << 32);
                }
                if (lane == 0) p.counts[static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q] = n;
            }
        }
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
static int px_warps(int) { return 8; }
static int px_rows(int) { return 8; }

static size_t px_fixed_bytes(int qc) {
    const int warps = px_warps(qc), rows = px_rows(qc);
    size_t b = static_cast<size_t>(warps) * rows * qc * 2 * sizeof(float);
    b += 2 * kPxMaxStages * sizeof(uint64_t);
    b += qc * (sizeof(unsigned long long) + sizeof(float) + sizeof(int) + sizeof(uint32_t));
    b += warps * 256 * sizeof(uint32_t) + 16;
    return b + 256;
}

int pixel_pick_qc(int Q) {
    const int e = env_knob("SKY_PX_QC", 0);      // experiments
    if (e == 1 || e == 2 || e == 4) return e;
    return Q == 1 ? 1 : (Q == 2 ? 2 : 4);
}

int pixel_grid(const sky_bank* b, int64_t n_rows, int qc) {
    const int rows = px_rows(qc);
    const int64_t groups = (n_rows + rows - 1) / rows;
    int64_t g = groups < b->num_sms ? groups : b->num_sms;
    return static_cast<int>(g < 1 ? 1 : g);
}

int launch_pixel_fold(const float* q, const unsigned char* mask, int64_t n, int D, float* qp, int* excl, cudaStream_t st) {
    SKY_CUDA(cudaMemsetAsync(excl, 0, static_cast<size_t>((n + D - 1) / D) * sizeof(int), st));
    pixel_fold_query_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(q, mask, n, D, qp, excl);
    SKY_LAUNCH_CHECK("pixel_fold_query_kernel");
    return SKY_OK;
}

template <int QC, int WARPS, int ROWS>
static int pixel_launch_one(const PixelParams& p, int grid, size_t smem, cudaStream_t st) {
    SKY_CUDA(cudaFuncSetAttribute(pixel_search_kernel<QC, WARPS, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    pixel_search_kernel<QC, WARPS, ROWS><<<grid, WARPS * 32 + 64, smem, st>>>(p);
    SKY_LAUNCH_CHECK("pixel_search_kernel");
    return SKY_OK;
}

// qp: folded queries [Q][D]; rows [row_lo, row_hi) of the bank; emit != null -> scores only
int launch_pixel_search(const sky_bank* b, const float* qp, const int* excl, int Q, int64_t row_lo, int64_t row_hi,
                        const SearchState& s, int grid, int qc, float* emit, cudaStream_t st) {
    int stages = static_cast<int>((227 * 1024 - px_fixed_bytes(qc)) / kPxChunk);
    if (stages > kPxMaxStages) stages = kPxMaxStages;
    { const int e = env_knob("SKY_PX_STAGES", 0); if (e >= 2 && e < stages) stages = e; }
    const size_t smem = static_cast<size_t>(stages) * kPxChunk + px_fixed_bytes(qc);
    for (int q0 = 0; q0 < Q; q0 += qc) {
        PixelParams p;
        p.bank = reinterpret_cast<const unsigned char*>(b->data);
        p.qp = qp;
        p.excl = excl;
        p.row_lo = row_lo; p.row_hi = row_hi;
        p.D = b->D; p.nch = (b->D + kPxChunkElems - 1) / kPxChunkElems;
        p.q0 = q0; p.nq = (Q - q0 < qc) ? (Q - q0) : qc;
        p.stages = stages;
        p.policy = ptx::kEvictFirst;
        p.lists = s.lists; p.counts = s.counts; p.gtop = s.gtop;
        p.p_stride = s.p_stride; p.Qtot = s.Qtot; p.cap = s.cap; p.k = s.k; p.use_gtau = s.use_gtau;
        p.emit = emit; p.emit_item0 = row_lo; p.emit_n = row_hi - row_lo;
        if (!emit) prof_mark(b, st);
        int rc = qc == 1 ? pixel_launch_one<1, 8, 8>(p, grid, smem, st)
                         : (qc == 2 ? pixel_launch_one<2, 8, 8>(p, grid, smem, st) : pixel_launch_one<4, 8, 8>(p, grid, smem, st));
        if (!emit) prof_mark(b, st);
        if (rc) return rc;
    }
    return SKY_OK;
}

}  // namespace sky
