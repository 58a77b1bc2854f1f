// stream_search.cu -- K1: HBM-streaming scorer with bulk-copy (TMA engine) staging and fused top-k.
//
// The regime of the reference itself: ONE query (utils/similarity.py has Q = 1) or a handful,
// per-feature weights, fp32 or bf16 bank, any of the three metrics, L tokens per item:
//   cosine  utils/similarity.py:163-170   sum(w t x) / (sqrt(sum w t^2) sqrt(sum w x^2) + 1e-6)
//   MSE     utils/similarity.py:188-192   sum(w (t-x)^2) / (sum w) / D
//   MAE     utils/similarity.py:208-212   sum(w |t-x|)   / (sum w) / D
//   n_top_sims :257-259, combine over the L tokens of an item :262-267, running top-k :18-35.
// Arithmetic intensity is <= 4 queries x 3 flop per element, so the kernel must run at HBM speed.
// The bank is tile-major (common.cuh): every (tile, k-block) is one contiguous 16 KB piece, which
// one elected producer thread streams into a ring of shared-memory stages with cp.async.bulk
// (UBLKCP; completion on an mbarrier), so ~150 KB per SM are in flight with no register cost.
// Eight consumer warps read the stages with conflict-free 128-bit shared loads, keep the (<= 4)
// query vectors in shared memory, accumulate with packed fp32 FMAs (FFMA2) and push scores that
// beat the running threshold straight into the CTA's candidate sink -- no [Q, N] matrix in HBM.
// One more warp trades grid-wide k-th-best bounds while the stream runs (topk.cuh).
#include <cstdlib>

#include "bank.cuh"
#include "ptx.cuh"
#include "topk.cuh"

namespace sky {

constexpr int kStWarps = 8;                        // consumer warps
constexpr int kStConsumers = kStWarps * 32;
constexpr int kStProducerWarps = 2;                // cp.async producers (bulk mode: one elected thread)
constexpr int kStProducers = kStProducerWarps * 32;
constexpr int kStThreads = kStConsumers + kStProducers + 32;      // + exchange warp
constexpr int kStProducerWarp = kStWarps;
constexpr int kStXchgWarp = kStWarps + kStProducerWarps;
constexpr int kStChunk = 16384;                    // bytes per stage
constexpr int kStMaxStages = 12;

struct StreamParams {
    const unsigned char* bank;
    const float* rownorm;
    int64_t row_lo, row_hi;     // valid bank rows [row_lo, row_hi) (item aligned)
    int L, D, Dp, KB;
    const float* t;
    const float* w;             // may be null (ones)
    int q0, nq;                 // queries [q0, q0 + nq) of this launch
    int combine, n_top;
    int stages;
    int debug;                  // experiments: bit0 skip the stage math, bit1 skip the row-block epilogue
    int spin;                   // experiments: 1 = consumers spin on try_wait instead of the suspending wait
    int split;                  // > 0: bulk copies (UBLKCP) per stage; 0: cp.async (LDGSTS) producers
    unsigned long long policy;  // L2 hint of the bank stream
    // sink
    uint64_t* lists; int* counts; uint32_t* gtop;
    int p_stride, Qtot, cap, k, use_gtau;
    // emit mode (sky_score): scores of items [emit_item0, emit_item0 + emit_n) -> emit[q * emit_n + i]
    float* emit; int64_t emit_item0; int64_t emit_n;
};

#ifdef SKY_EXPERIMENTS
__device__ unsigned long long g_st_stats[8];   // debug (SKY_ST_DEBUG bit 4): [0] inserts tried, [1] passed exact test, [2] prunes
#endif

// ---- packed fp32 pairs (FFMA2 / FMUL2 / FADD2: two fp32 lanes per issue slot) ------------------
__device__ __forceinline__ uint64_t f2_pack(float lo, float hi) {
    return (static_cast<uint64_t>(__float_as_uint(hi)) << 32) | static_cast<uint64_t>(__float_as_uint(lo));
}
__device__ __forceinline__ float f2_lo(uint64_t v) { return __uint_as_float(static_cast<uint32_t>(v)); }
__device__ __forceinline__ float f2_hi(uint64_t v) { return __uint_as_float(static_cast<uint32_t>(v >> 32)); }
#ifdef SKY_SCALAR_F2      // experiment: scalar FFMA / FMUL / FADD instead of the packed forms
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    return f2_pack(fmaf(f2_lo(a), f2_lo(b), f2_lo(c)), fmaf(f2_hi(a), f2_hi(b), f2_hi(c)));
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) { return f2_pack(f2_lo(a) * f2_lo(b), f2_hi(a) * f2_hi(b)); }
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) { return f2_pack(f2_lo(a) - f2_lo(b), f2_hi(a) - f2_hi(b)); }
#else
__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t r; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c)); return r;
}
__device__ __forceinline__ uint64_t mul2(uint64_t a, uint64_t b) {
    uint64_t r; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
__device__ __forceinline__ uint64_t sub2(uint64_t a, uint64_t b) {
    uint64_t r; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b)); return r;
}
#endif

__device__ __forceinline__ void bulk_load(uint32_t smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
        ::"r"(smem_dst), "l"(gsrc), "r"(bytes), "r"(ptx::smem_u32(bar)), "l"(policy)
        : "memory");
}
__device__ __forceinline__ uint4 lds128(uint32_t a) {
    uint4 v;
    asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(a));
    return v;
}

// Geometry of one 16 KB stage: R bank rows x 64 elements.
template <typename BankT> struct StGeo;
template <> struct StGeo<__nv_bfloat16> {
    static constexpr int R = 128, LPR = 8, CPL = 8, ROWB = 128;     // rows, lanes per row, columns per lane, row bytes
};
template <> struct StGeo<float> {
    static constexpr int R = 64, LPR = 16, CPL = 4, ROWB = 256;
};

// byte offset of stage (row block rb, k-block kb) in the tile-major bank
template <typename BankT>
__device__ __forceinline__ size_t chunk_offset(int64_t rb, int kb, int KB) {
    if constexpr (sizeof(BankT) == 2) return (static_cast<size_t>(rb) * KB + kb) * kStChunk;
    else return (static_cast<size_t>(rb >> 1) * KB + kb) * (2 * kStChunk) + static_cast<size_t>(rb & 1) * kStChunk;
}

template <typename BankT, int METRIC, int QC, bool WEIGHTED>
__global__ void __launch_bounds__(kStThreads, 1) stream_search_kernel(const StreamParams p) {
    using G = StGeo<BankT>;
    constexpr int R = G::R, LPR = G::LPR, CPL = G::CPL, ROWB = G::ROWB;
    constexpr int RPG = 32 / LPR;            // rows per 128-bit warp load
    constexpr int RPW = R / kStWarps;        // rows per warp
    constexpr int NRG = RPW / RPG;           // row groups per warp (4 for both element types)
    constexpr int NP = CPL / 2;              // fp32 pairs per lane per row
    constexpr bool COS = (METRIC == SKY_COSINE);
    constexpr int NC = (COS && WEIGHTED) ? 2 : 1;
    constexpr bool largest = COS;

    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const int S = p.stages, Dp = p.Dp, KB = p.KB;
    unsigned char* sStage = base;                                             // [S][16 KB]
    float* sA = reinterpret_cast<float*>(sStage + static_cast<size_t>(S) * kStChunk);   // [QC][Dp] cosine: w*t, else t
    float* sW = sA + QC * Dp;                                                 // [QC][Dp]
    float* sTok = sW + QC * Dp;                                               // [QC][R] token scores (L > 1)
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sTok + QC * R);          // [kStMaxStages]
    uint64_t* empty_bar = full_bar + kStMaxStages;                            // [kStMaxStages]
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>(empty_bar + kStMaxStages);   // [QC]
    float* sQc = reinterpret_cast<float*>(sThr + QC);                         // [QC] cosine: |t|_w, else sum(w)
    float* sThrF = sQc + QC;                                                  // [QC]
    int* sCnt = reinterpret_cast<int*>(sThrF + QC);                           // [QC]
    uint32_t* sLmax = reinterpret_cast<uint32_t*>(sCnt + QC);                 // [QC]
    uint32_t* sHist = sLmax + QC;                                             // [kStWarps][256]
    volatile int* sBlocksDone = reinterpret_cast<volatile int*>(sHist + kStWarps * 256);   // [1]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // Programmatic dependent launch (search mode): this grid may start while the kernel that zeroes the exchange state
    // (init_state_kernel) still runs.  The query staging, the bank stream and the scoring do not depend on it; the two
    // places that touch the exchange state (exchange warp, final publish + counts) wait for it first.
    ptx::griddep_launch_dependents();

    // row blocks of this CTA: rb_lo + blockIdx.x, + gridDim.x, ...
    const int64_t rb_lo = p.row_lo / R, rb_hi = (p.row_hi + R - 1) / R;
    const int64_t nblk = rb_hi - rb_lo;
    const int my_blocks = (nblk > static_cast<int64_t>(blockIdx.x))
                              ? static_cast<int>((nblk - blockIdx.x + gridDim.x - 1) / gridDim.x) : 0;

    if (tid == 0) {
        for (int s = 0; s < S; ++s) { ptx::mbar_init(&full_bar[s], p.split > 0 ? 1 : kStProducers); ptx::mbar_init(&empty_bar[s], kStWarps); }
        ptx::fence_barrier_init();
        *sBlocksDone = 0;
    }
    // ---- stage the query operands ---------------------------------------------------------------
    for (int i = tid; i < QC * Dp; i += kStThreads) {
        const int q = i / Dp, d = i - q * Dp;
        float tv = 0.f, wv = 0.f;
        if (q < p.nq && d < p.D) {
            tv = p.t[static_cast<size_t>(p.q0 + q) * p.D + d];
            wv = p.w ? p.w[static_cast<size_t>(p.q0 + q) * p.D + d] : 1.0f;
        }
        // Inside a 64-element k-block a lane owns CPL consecutive elements and reads them as CPL / 4 128-bit loads.  With
        // CPL = 8 (bf16 bank) the natural order puts the lanes' loads 32 bytes apart: lanes lc and lc + 4 hit the same
        // banks (2-way conflict on every query load, 12 M conflict cycles per 1 M-row search in ncu).  Store half h of
        // every lane contiguously instead: element lc*8 + h*4 + e -> position h*32 + lc*4 + e.
        int pi = i;
        if constexpr (CPL == 8) {
            const int e64 = d & 63;
            pi = q * Dp + (d & ~63) + ((e64 >> 2) & 1) * 32 + (e64 >> 3) * 4 + (e64 & 3);
        }
        sA[pi] = (COS && WEIGHTED) ? wv * tv : tv;
        sW[pi] = wv;
    }
    // the float mirror of the threshold starts as NaN: every comparison fails, so everything passes the pre-filter
    if (tid < QC) { sThr[tid] = (tid < p.nq) ? 0ull : ~0ull; sThrF[tid] = __uint_as_float(0x7FC00000u); sCnt[tid] = 0; sLmax[tid] = 0; }
    __syncthreads();
    if (warp < QC) {
        double acc = 0.0;
        for (int d = lane; d < p.D; d += 32) {
            // (from global memory: the shared-memory copies may be stored in a permuted order)
            const float wv = (warp < p.nq) ? (p.w ? p.w[static_cast<size_t>(p.q0 + warp) * p.D + d] : 1.0f) : 0.f;
            if (COS) {
                // w t^2 (the reference squares t after the product with w: weights * target ** 2)
                const float tv = (warp < p.nq) ? p.t[static_cast<size_t>(p.q0 + warp) * p.D + d] : 0.f;
                acc += static_cast<double>(wv * (tv * tv));
            } else {
                acc += static_cast<double>(wv);
            }
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
        if (lane == 0) sQc[warp] = COS ? sqrtf(static_cast<float>(acc)) : static_cast<float>(acc);
    }
    __syncthreads();

    Sink sink;
    sink.lists = p.lists ? p.lists + (static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0) * p.cap : nullptr;
    sink.thr = smem_addr(sThr); sink.thr_f = smem_addr(sThrF); sink.cnt = smem_addr(sCnt); sink.lmax = smem_addr(sLmax);
    sink.cap = p.cap; sink.k = p.k; sink.largest = largest;
    const bool emit = p.emit != nullptr;

    if (warp >= kStProducerWarp && warp < kStXchgWarp && p.split == 0) {
        // ===================== producers: 64 threads feed the ring with 16-byte cp.async =====================
        const int pt = tid - kStConsumers;
        int stage = 0;
        uint32_t phase = 0;
        const uint32_t stage0 = ptx::smem_u32(sStage) + pt * 16;
        for (int i = 0; i < my_blocks; ++i) {
            const int64_t rb = rb_lo + blockIdx.x + static_cast<int64_t>(i) * gridDim.x;
            for (int kb = 0; kb < KB; ++kb) {
                if (lane == 0) ptx::mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, 32);
                __syncwarp();
                const unsigned char* src = p.bank + chunk_offset<BankT>(rb, kb, KB) + pt * 16;
                const uint32_t dst = stage0 + static_cast<uint32_t>(stage) * kStChunk;
#pragma unroll
                for (int j = 0; j < kStChunk / (16 * kStProducers); ++j)
                    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(dst + j * kStProducers * 16),
                                 "l"(src + j * kStProducers * 16), "l"(p.policy) : "memory");
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ptx::smem_u32(&full_bar[stage])) : "memory");
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == kStProducerWarp) {
        // ===================== producer: one thread feeds the ring with bulk copies =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            const uint32_t stage0 = ptx::smem_u32(sStage);
            for (int i = 0; i < my_blocks; ++i) {
                const int64_t rb = rb_lo + blockIdx.x + static_cast<int64_t>(i) * gridDim.x;
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, 32);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], kStChunk);
                    const unsigned char* src = p.bank + chunk_offset<BankT>(rb, kb, KB);
                    const uint32_t dst = stage0 + static_cast<uint32_t>(stage) * kStChunk;
                    const uint32_t piece = kStChunk / p.split;
                    for (int s = 0; s < p.split; ++s)
                        bulk_load(dst + s * piece, src + s * piece, piece, &full_bar[stage], p.policy);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == kStXchgWarp) {
        // ===================== grid-wide bound exchange =====================
        if (!emit && p.use_gtau && my_blocks > 0) {
            ptx::griddep_wait();
            uint32_t* my_row = p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0;
            int round = 0;
            uint32_t published = 0;      // lane q: the key this CTA last published for query q
            while (*sBlocksDone < my_blocks) {
                // publish only what changed: all CTAs' entries of a query share a few L2 lines, and rewriting them
                // every round turns those lines into a hot spot that delays the bank stream behind it
                if (lane < p.nq) {
                    const uint32_t mine = lds_u32(sink.lmax + lane * 4);
                    if (mine != published) { st_cg_u32(my_row + lane, mine); published = mine; }
                }
                for (int q = 0; q < p.nq; ++q) {
                    const uint32_t lo = exchange_reduce(p.gtop + p.q0 + q, p.p_stride, p.Qtot, p.k);
                    if (lane == 0) exchange_apply(sink, q, lo);
                }
                ++round;
                __nanosleep((SKY_DBG(p) & 128) ? 4000 : (round < 32 ? 200 : 4000));
            }
        }
    } else if (warp < kStWarps) {
        // ===================== consumers =====================
        const int lr = lane / LPR, lc = lane % LPR;
        const uint32_t stage0 = ptx::smem_u32(sStage);
        const uint32_t my_off = static_cast<uint32_t>((warp * RPW + lr) * ROWB + lc * 16);
        // lane lc's h-th 128-bit piece of a query k-block: at lc*16 + h*HSTEP (see the staging loop)
        constexpr int HSTEP = (CPL == 8) ? 128 : 16;
        const uint32_t sA_addr = ptx::smem_u32(sA) + lc * 16, sW_addr = ptx::smem_u32(sW) + lc * 16;
        const float invD = 1.0f / static_cast<float>(p.D);
        const int L = p.L;
        int stage = 0;
        uint32_t phase = 0;
        // a list takes at most R (L = 1) or R / L inserts per row block; with cap well above k the CTA-wide
        // barrier pair of the overflow check is only needed every few blocks
        const int per_block = (L == 1) ? R : R / L;
        int check_every = emit ? 1 : (p.cap - p.k) / (2 * per_block);
        check_every = check_every < 1 ? 1 : (check_every > 16 ? 16 : check_every);
        const int prune_slack = check_every * per_block;

        for (int i = 0; i < my_blocks; ++i) {
            const int64_t rb = rb_lo + blockIdx.x + static_cast<int64_t>(i) * gridDim.x;
            const int64_t row_base = rb * R + warp * RPW + lr;          // + rg * RPG
            float rn[NRG];
            if (COS && !WEIGHTED) {
#pragma unroll
                for (int rg = 0; rg < NRG; ++rg)     // volatile: issued here, a whole row block ahead of its use
                    asm volatile("ld.global.nc.f32 %0, [%1];" : "=f"(rn[rg]) : "l"(p.rownorm + row_base + rg * RPG));
            }
            uint64_t acc[NRG][QC][NC];
#pragma unroll
            for (int rg = 0; rg < NRG; ++rg)
#pragma unroll
                for (int q = 0; q < QC; ++q)
#pragma unroll
                    for (int c = 0; c < NC; ++c) acc[rg][q][c] = 0ull;

            for (int kb = 0; kb < KB; ++kb) {
                if (p.spin) ptx::mbar_wait(&full_bar[stage], phase);
                else ptx::mbar_wait_relaxed(&full_bar[stage], phase, 20);
                const uint32_t src = stage0 + static_cast<uint32_t>(stage) * kStChunk + my_off;
                if (SKY_DBG(p) & 1) {
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                    continue;
                }
                uint64_t x2[NRG][NP];
#pragma unroll
                for (int rg = 0; rg < NRG; ++rg) {
                    const uint4 v = lds128(src + rg * RPG * ROWB);
                    if constexpr (sizeof(BankT) == 2) {
                        const uint32_t u[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            x2[rg][j] = (static_cast<uint64_t>(u[j] & 0xFFFF0000u) << 32) | static_cast<uint64_t>(u[j] << 16);
                    } else {
                        x2[rg][0] = (static_cast<uint64_t>(v.y) << 32) | v.x;
                        x2[rg][1] = (static_cast<uint64_t>(v.w) << 32) | v.z;
                    }
                }
                uint64_t xx2[NRG][NP];
                if (COS && WEIGHTED) {
#pragma unroll
                    for (int rg = 0; rg < NRG; ++rg)
#pragma unroll
                        for (int j = 0; j < NP; ++j) xx2[rg][j] = mul2(x2[rg][j], x2[rg][j]);
                }
                const uint32_t qoff = static_cast<uint32_t>(kb * kKBlock * 4);
#pragma unroll
                for (int q = 0; q < QC; ++q) {
                    uint64_t a2[NP], w2[NP];
#pragma unroll
                    for (int h = 0; h < NP / 2; ++h) {
                        const uint4 av = lds128(sA_addr + q * Dp * 4 + qoff + h * HSTEP);
                        a2[2 * h] = (static_cast<uint64_t>(av.y) << 32) | av.x;
                        a2[2 * h + 1] = (static_cast<uint64_t>(av.w) << 32) | av.z;
                        if (WEIGHTED) {
                            const uint4 wv = lds128(sW_addr + q * Dp * 4 + qoff + h * HSTEP);
                            w2[2 * h] = (static_cast<uint64_t>(wv.y) << 32) | wv.x;
                            w2[2 * h + 1] = (static_cast<uint64_t>(wv.w) << 32) | wv.z;
                        }
                    }
#pragma unroll
                    for (int rg = 0; rg < NRG; ++rg)
#pragma unroll
                        for (int j = 0; j < NP; ++j) {
                            if (COS) {
                                acc[rg][q][0] = fma2(a2[j], x2[rg][j], acc[rg][q][0]);
                                if (WEIGHTED) acc[rg][q][NC - 1] = fma2(w2[j], xx2[rg][j], acc[rg][q][NC - 1]);
                            } else if (METRIC == SKY_MSE) {
                                const uint64_t d2 = sub2(a2[j], x2[rg][j]);
                                if (WEIGHTED) acc[rg][q][0] = fma2(w2[j], mul2(d2, d2), acc[rg][q][0]);
                                else acc[rg][q][0] = fma2(d2, d2, acc[rg][q][0]);
                            } else {
                                const float d0 = fabsf(f2_lo(a2[j]) - f2_lo(x2[rg][j]));
                                const float d1 = fabsf(f2_hi(a2[j]) - f2_hi(x2[rg][j]));
                                float s0 = f2_lo(acc[rg][q][0]), s1 = f2_hi(acc[rg][q][0]);
                                if (WEIGHTED) { s0 = fmaf(f2_lo(w2[j]), d0, s0); s1 = fmaf(f2_hi(w2[j]), d1, s1); }
                                else { s0 += d0; s1 += d1; }
                                acc[rg][q][0] = f2_pack(s0, s1);
                            }
                        }
                }
                __syncwarp();
                if (lane == 0) ptx::mbar_arrive(&empty_bar[stage]);
                if (++stage == S) { stage = 0; phase ^= 1; }
            }

            if (SKY_DBG(p) & 2) { if (tid == 0) *sBlocksDone = i + 1; continue; }
            // ---- row sums: fold the pair, then across the LPR lanes that share a row ---------------
            float mine[NRG];     // score of (row group rg, row lr, query lc) on lanes with lc < QC
            const float qc = sQc[lc < QC ? lc : 0];
            const float mse_scale = __fdividef(invD, qc);
            // threshold as a score, read once per row block; it may be stale (it only tightens) and the exact
            // composite compare inside sink_insert_one decides
            const float thrf = lds_f32(sink.thr_f + (lc < QC ? lc : 0) * 4);
#pragma unroll
            for (int rg = 0; rg < NRG; ++rg) {
                float sel0 = 0.f, sel1 = 0.f;
#pragma unroll
                for (int q = 0; q < QC; ++q) {
#pragma unroll
                    for (int c = 0; c < NC; ++c) {
                        float v = f2_lo(acc[rg][q][c]) + f2_hi(acc[rg][q][c]);
#pragma unroll
                        for (int off = LPR / 2; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
                        if (lc == q) { if (c == 0) sel0 = v; else sel1 = v; }
                    }
                }
                if (COS) {
                    const float xx = WEIGHTED ? sel1 : rn[rg];
                    mine[rg] = __fdividef(sel0, fmaf(qc, sqrtf(xx), 1e-6f));
                } else {
                    mine[rg] = sel0 * mse_scale;
                }
            }

            if (L == 1) {
#pragma unroll
                for (int rg = 0; rg < NRG; ++rg) {
                    const int64_t row = row_base + rg * RPG;
                    if (lc < p.nq && row >= p.row_lo && row < p.row_hi) {
                        if (emit) {
                            if (row >= p.emit_item0 && row < p.emit_item0 + p.emit_n)
                                p.emit[static_cast<size_t>(p.q0 + lc) * p.emit_n + (row - p.emit_item0)] = mine[rg];
                        } else if ((SKY_DBG(p) & 8) || ((SKY_DBG(p) & 32) && i >= 2) || ((SKY_DBG(p) & 64) && i < 2)) {
                            if (mine[rg] == 123.456f) p.counts[0] = 1;      // keep the score alive
                        } else if (largest ? !(mine[rg] < thrf) : !(mine[rg] > thrf)) {      // NaN passes
#ifdef SKY_EXPERIMENTS
                            if (SKY_DBG(p) & 16) {
                                atomicAdd(&g_st_stats[0], 1ull);
                                if (make_composite(score_to_key(mine[rg], largest), static_cast<uint32_t>(row)) > sink_thr(sink, lc)) atomicAdd(&g_st_stats[1], 1ull);
                                if (i < 8) atomicAdd(&g_st_stats[3], 1ull);
                                if (i < 2) atomicAdd(&g_st_stats[4], 1ull);
                            }
#endif
                            sink_insert_one(sink, lc, make_composite(score_to_key(mine[rg], largest), static_cast<uint32_t>(row)));
                        }
                    }
                }
            } else {
                // token scores -> shared memory; items never straddle a row block (R % L == 0)
                if (lc < QC) {
#pragma unroll
                    for (int rg = 0; rg < NRG; ++rg) sTok[lc * R + warp * RPW + rg * RPG + lr] = mine[rg];
                }
                ptx::named_bar_sync(1, kStConsumers);
                const int items = R / L;
                for (int pr = warp; pr < items * p.nq; pr += kStWarps) {
                    const int it = pr / p.nq, q = pr - it * p.nq;
                    const float* tk = sTok + q * R + it * L;
                    const int64_t item = (rb * R) / L + it;
                    float lsum = 0.f, lmin = INFINITY, lmax = -INFINITY;
                    bool lnan = false;
                    for (int ii = lane; ii < L; ii += 32) {
                        const float vi = tk[ii];
                        bool take = true;
                        if (p.n_top > 0) {
                            // best-n_top token scores (torch.topk, utils/similarity.py:259)
                            const uint32_t ki = score_to_key(vi, largest);
                            int rank = 0;
                            for (int j = 0; j < L; ++j) {
                                const uint32_t kj = score_to_key(tk[j], largest);
                                rank += (kj > ki) || (kj == ki && j < ii);
                            }
                            take = rank < p.n_top;
                        }
                        if (take) { lnan |= (vi != vi); lsum += vi; lmin = fminf(lmin, vi); lmax = fmaxf(lmax, vi); }
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        lsum += __shfl_xor_sync(0xffffffffu, lsum, off);
                        lmin = fminf(lmin, __shfl_xor_sync(0xffffffffu, lmin, off));
                        lmax = fmaxf(lmax, __shfl_xor_sync(0xffffffffu, lmax, off));
                    }
                    const bool anynan = __any_sync(0xffffffffu, lnan);
                    const int cnt = p.n_top > 0 ? p.n_top : L;
                    float v = (p.combine == SKY_MIN) ? lmin : (p.combine == SKY_MAX ? lmax : lsum / static_cast<float>(cnt));
                    if (anynan) v = __uint_as_float(0x7FC00000u);
                    const int64_t row0 = item * L;
                    if (lane == 0 && row0 >= p.row_lo && row0 < p.row_hi) {
                        if (emit) {
                            if (item >= p.emit_item0 && item < p.emit_item0 + p.emit_n)
                                p.emit[static_cast<size_t>(p.q0 + q) * p.emit_n + (item - p.emit_item0)] = v;
                        } else {
                            sink_insert_one(sink, q, make_composite(score_to_key(v, largest), static_cast<uint32_t>(item)));
                        }
                    }
                }
            }
            if (SKY_DBG(p) & 4) {
            } else if (!emit && (i + 1) % check_every == 0) {
                // overflow check, every check_every row blocks: at most prune_slack inserts per query in between
                ptx::named_bar_sync(1, kStConsumers);
                sink_prune_if_full(sink, p.nq, warp, kStWarps, smem_addr(sHist + warp * 256), prune_slack);
                ptx::named_bar_sync(1, kStConsumers);
            } else if (L != 1) {
                ptx::named_bar_sync(1, kStConsumers);       // token scores of this block are consumed
            }
            if (tid == 0) *sBlocksDone = i + 1;
        }
        if (!emit) {
            ptx::griddep_wait();
            ptx::named_bar_sync(1, kStConsumers);
            if (tid < p.nq && p.use_gtau && sLmax[tid])
                st_cg_u32(p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + tid, sLmax[tid]);
            ptx::named_bar_sync(1, kStConsumers);
            // Shrink this CTA's lists before handing them to the merge: the first row blocks went in unfiltered (no
            // bound existed yet); what is below the grid-wide bound known NOW can never be in the top-k.  One warp
            // per query compacts in place, so the merge kernel reads a handful of entries per CTA instead of hundreds.
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
static int st_rows_per_block(int dtype) { return dtype == SKY_BF16 ? 128 : 64; }

static size_t st_fixed_bytes(int Dp, int qc, int R) {
    size_t b = static_cast<size_t>(2) * qc * Dp * sizeof(float);        // sA, sW
    b += static_cast<size_t>(qc) * R * sizeof(float);                   // sTok
    b += 2 * kStMaxStages * sizeof(uint64_t);                           // barriers
    b += qc * (sizeof(unsigned long long) + 2 * sizeof(float) + sizeof(int) + sizeof(uint32_t));
    b += kStWarps * 256 * sizeof(uint32_t) + 16;                        // histograms, progress flag
    return b + 256;                                                     // alignment slack
}

static int st_stages(int Dp, int qc, int R) {
    const size_t budget = 227 * 1024;
    const size_t fixed = st_fixed_bytes(Dp, qc, R);
    if (fixed + 3 * kStChunk > budget) return 0;
    size_t s = (budget - fixed) / kStChunk;
    return static_cast<int>(s > kStMaxStages ? kStMaxStages : s);
}

#ifdef SKY_EXPERIMENTS
int debug_stream_stats(unsigned long long* h_out, int reset) {
    SKY_CUDA(cudaDeviceSynchronize());
    SKY_CUDA(cudaMemcpyFromSymbol(h_out, g_st_stats, sizeof(g_st_stats)));
    if (reset) { unsigned long long z[8] = {0}; SKY_CUDA(cudaMemcpyToSymbol(g_st_stats, z, sizeof(z))); }
    return SKY_OK;
}
#endif

int stream_pick_qc(int Q) { return Q == 1 ? 1 : 4; }

bool stream_supported(const sky_bank* b, int qc) {
    const int R = st_rows_per_block(b->dtype);
    if (b->L != 1 && (R % b->L) != 0) return false;
    return st_stages(b->Dp, qc, R) >= 3;
}

int stream_grid(const sky_bank* b, int64_t row_lo, int64_t row_hi) {
    const int R = st_rows_per_block(b->dtype);
    const int64_t nblk = (row_hi + R - 1) / R - row_lo / R;
    int64_t g = nblk < b->num_sms ? nblk : b->num_sms;
    return static_cast<int>(g < 1 ? 1 : g);
}

template <typename BankT, int METRIC, int QC, bool WEIGHTED>
static int stream_launch_one(const StreamParams& p, int grid, size_t smem, cudaStream_t st) {
    SKY_CUDA(cudaFuncSetAttribute(stream_search_kernel<BankT, METRIC, QC, WEIGHTED>,
                                  cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    // search mode follows init_state_kernel (or the previous pass of this kernel): start under its tail.  Emit mode
    // (sky_score) may follow the caller's own kernels, which can have produced the query: ordinary launch.
    const bool pdl = p.emit == nullptr && env_knob("SKY_PDL", kUsePdl) != 0;
    SKY_CUDA(launch_maybe_pdl(pdl, stream_search_kernel<BankT, METRIC, QC, WEIGHTED>, dim3(grid), dim3(kStThreads), smem, st, p));
    SKY_LAUNCH_CHECK("stream_search_kernel");
    return SKY_OK;
}

template <typename BankT, int METRIC>
static int stream_launch_qw(int qc, bool weighted, const StreamParams& p, int grid, size_t smem, cudaStream_t st) {
    if (qc == 1) return weighted ? stream_launch_one<BankT, METRIC, 1, true>(p, grid, smem, st)
                                 : stream_launch_one<BankT, METRIC, 1, false>(p, grid, smem, st);
    return weighted ? stream_launch_one<BankT, METRIC, 4, true>(p, grid, smem, st)
                    : stream_launch_one<BankT, METRIC, 4, false>(p, grid, smem, st);
}

template <typename BankT>
static int stream_launch_m(int metric, int qc, bool weighted, const StreamParams& p, int grid, size_t smem, cudaStream_t st) {
    if (metric == SKY_COSINE) return stream_launch_qw<BankT, SKY_COSINE>(qc, weighted, p, grid, smem, st);
    if (metric == SKY_MSE) return stream_launch_qw<BankT, SKY_MSE>(qc, weighted, p, grid, smem, st);
    return stream_launch_qw<BankT, SKY_MAE>(qc, weighted, p, grid, smem, st);
}

// a: the same argument block as the generic SIMT scorer (bank.cuh); s: candidate state (unused in emit mode)
int launch_stream_search(const sky_bank* b, const SimtArgs& a, const SearchState& s, int grid, int qc, cudaStream_t st) {
    const int R = st_rows_per_block(a.dtype);
    int stages = st_stages(a.Dp, qc, R);
    { const int e = env_knob("SKY_ST_STAGES", 0); if (e >= 2 && e < stages) stages = e; }
    int split = 1;
    { const int e = env_knob("SKY_ST_SPLIT", -1); if (e >= 0 && e <= 16) split = e; }
    unsigned long long policy = ptx::kEvictFirst;
    { const int pv = env_knob("SKY_ST_POLICY", 0);
      policy = pv == 1 ? 0x1000000000000000ull : (pv == 2 ? ptx::kEvictLast : ptx::kEvictFirst); }
    if (stages < 2) return set_error(SKY_ERR_UNSUPPORTED, "streaming scorer: D=%d does not fit in shared memory", a.Dp);
    const size_t smem = static_cast<size_t>(stages) * kStChunk + st_fixed_bytes(a.Dp, qc, R);
    for (int q0 = 0; q0 < a.Q; q0 += qc) {
        StreamParams p;
        p.bank = reinterpret_cast<const unsigned char*>(a.bank);
        p.rownorm = b->rownorm;
        p.row_lo = a.row0; p.row_hi = a.row0 + a.n_items * a.L;
        p.L = a.L; p.D = a.D; p.Dp = a.Dp; p.KB = a.Dp / kKBlock;
        p.t = a.t; p.w = a.w; p.q0 = q0; p.nq = (a.Q - q0 < qc) ? (a.Q - q0) : qc;
        p.combine = a.combine; p.n_top = a.n_top;
        p.stages = stages;
        p.split = split;
        p.spin = env_knob("SKY_ST_SPIN", 0);
        p.debug = env_knob("SKY_ST_DEBUG", 0);
        p.policy = policy;
        p.lists = s.lists; p.counts = s.counts; p.gtop = s.gtop;
        p.p_stride = s.p_stride; p.Qtot = s.Qtot; p.cap = s.cap; p.k = s.k; p.use_gtau = s.use_gtau;
        p.emit = a.emit; p.emit_item0 = a.row0 / a.L + a.item0; p.emit_n = a.n;
        if (!a.emit) prof_mark(b, st);
        int rc = (a.dtype == SKY_BF16) ? stream_launch_m<__nv_bfloat16>(a.metric, qc, a.w != nullptr, p, grid, smem, st)
                                       : stream_launch_m<float>(a.metric, qc, a.w != nullptr, p, grid, smem, st);
        if (!a.emit) prof_mark(b, st);
        if (rc) return rc;
    }
    return SKY_OK;
}

}  // namespace sky
