// tc_weighted.cu -- K2w: tcgen05 scorer for queries WITH per-feature weights (the reference's default,
// use_weights=True in both drivers: similarity_search.py:170, sky_sim_search.py:163).
//
// Contraction forms of the weighted metrics, a = w o t:
//   cosine  utils/similarity.py:163-170   s = a.z / ( sqrt(sum w t^2) sqrt(w.(z o z)) + 1e-6 )
//   MSE     utils/similarity.py:188-192   s = ( sum w t^2 - 2 a.z + w.(z o z) ) / (D sum w)
// Two contractions per bank tile: D1 = Z A^T (A = [a_q]) and D2 = (Z o Z) W^T (W = [w_q]).  The bank holds
// only Z; the squared tile is made ON CHIP: four "squarer" warps read every 16 KB stage the producers
// land (thread = bank row), square it with packed bf16 multiplies and store it to TENSOR MEMORY with
// tcgen05.st; the second MMA takes its A operand from there (no shared memory for Z o Z, so the ring stays
// 6 stages deep).  HBM traffic is one pass over the bank for up to 64 weighted queries -- HBM bound like K2.  Candidate sink, grid-wide bound exchange and merge are shared with K1/K2.
#include <cuda_bf16.h>

#include <cstdio>
#include <cstdlib>

#include "bank.cuh"
#include "ptx.cuh"
#include "topk.cuh"

namespace sky {

constexpr int kTwBN = 64;                        // queries per launch
constexpr int kTwProducerWarps = 4;              // warps 0-3: cp.async bank stream (+ TMA of the query k-blocks)
constexpr int kTwSquareWarp0 = 4;                // warps 4-7: z -> z o z
constexpr int kTwEpiWarp0 = 8;                   // warps 8-11: epilogue (warp & 3 = TMEM lane quarter)
constexpr int kTwMmaWarp = 12;
constexpr int kTwXchgWarp = 13;
constexpr int kTwThreads = 14 * 32;
constexpr int kTwProducers = kTwProducerWarps * 32;
constexpr int kTwStageA = kTileRows * 128;       // 16 KB
constexpr int kTwStageB = kTwBN * 128;           // 8 KB
constexpr int kTwStage = kTwStageA + 2 * kTwStageB;       // A | Ba | Bw = 32 KB
constexpr int kTwStages = 6;
constexpr int kTwSqCols = 32;                    // TMEM columns of one squared k-block: 64 bf16 per row, two per column
constexpr int kTwSqBase = 4 * kTwBN;             // behind the 2 x (D1 | D2) accumulators

struct TwParams {
    const unsigned char* bank;
    const float* qc1;          // [64] cosine: sqrt(sum w t^2); MSE: sum w t^2
    const float* qc2;          // [64] MSE: 1 / (D sum w); cosine: unused
    uint64_t* lists; int* counts; uint32_t* gtop; uint32_t* gtau;
    int p_stride, Qtot, q0, nq, cap, k, use_gtau;
    int64_t rows;
    int num_tiles, kblocks;
    unsigned long long bank_policy;
    int debug;                 // experiments: bit0 squarers only signal, bit1 no second MMA, bit2 epilogue only drains TMEM
};

__device__ __forceinline__ void tw_cp_async_16(uint32_t smem_dst, const void* gsrc, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "l"(policy) : "memory");
}

// Per-query pre-filter coefficients from the current thresholds (one thread per query; called by the epilogue warps
// between the two barriers that end a tile, and once before the first tile).
template <bool COS>
__device__ __forceinline__ void tw_prefilter_coeffs(int q, int nq, const Sink& sink, const float* sQ1, const float* sQ2, float* sC) {
    if (q >= nq) {                       // padding query: nothing passes
        sC[q] = COS ? INFINITY : -INFINITY;
        return;
    }
    const float th = lds_f32(sink.thr_f + q * 4);        // NaN = no threshold yet -> NaN coefficient -> everything passes
    if (COS) {
        const float t2 = th * sQ1[q];
        const float c = t2 * fabsf(t2);
        sC[q] = fmaf(-fabsf(c), 8e-6f, c);
    } else {
        const float b = th / sQ2[q] - sQ1[q];            // q2 = 1 / (D sum w) > 0
        sC[q] = fmaf(fabsf(b) + fabsf(sQ1[q]), 8e-6f, b);
    }
}

// Epilogue of one 128-row tile for one warp (thread = bank row = TMEM lane): D1 | D2 accumulators of 64 queries at
// `taddr` -> conservative pre-filter bitmask -> exact score and insert of the survivors.  `release` runs once the
// last accumulator chunk is in registers (the accumulator stage may then be overwritten by the next tile's MMAs).
template <bool COS, typename Release>
__device__ __forceinline__ void tw_score_tile(uint32_t taddr, uint32_t row, bool valid, int nq, int debug, const Sink& sink,
                                              const float* sQ1, const float* sQ2, const float* sC, Release release) {
    constexpr bool largest = COS;
#pragma unroll 1
    for (int c = 0; c < kTwBN / 32; ++c) {
        uint32_t v1[32], v2[32];
        ptx::tmem_ld_32x32b_x32(taddr + c * 32, v1);
        ptx::tmem_ld_32x32b_x32(taddr + kTwBN + c * 32, v2);
        ptx::tmem_ld_wait();
        if (c == kTwBN / 32 - 1) {
            ptx::tc_fence_before();
            __syncwarp();
            release();
        }
        if (debug & 4) continue;
        auto score = [&](float d1, float d2, float q1, float q2) -> float {
            // cosine: the accumulated w.(z o z) can round slightly below zero for a near-null row
            return COS ? __fdividef(d1, fmaf(q1, sqrtf(fmaxf(d2, 0.f)), 1e-6f)) : (q1 - 2.0f * d1 + d2) * q2;
        };
        // FAST PATH (kept short: three warp roles share each scheduler's instruction cache, and a long unrolled
        // epilogue slows the MMA issuer and the squarers down): a conservative pre-filter against per-query
        // coefficients sC[q] that tw_prefilter_coeffs derives from the thresholds once per tile
        //   cosine: s >= th  <=>  d1 >= th (q1 sqrt(d2) + 1e-6); with f(x) = x |x| (monotone) this is implied by
        //           f(d1 + 1e-6) >= f(th q1) d2;   sC = f(th q1), lowered by a few 1e-6 relative
        //   MSE:    s <= th  <=>  (q1 - 2 d1 + d2) q2 <= th  <=>  d2 - 2 d1 <= th / q2 - q1 = sC (raised a little)
        // sC may be a tile stale (thresholds only tighten); NaN (no threshold yet) passes everything on to the exact
        // test; padding queries carry +-inf and never pass.
        uint32_t mbits = 0;
#pragma unroll
        for (int g4 = 0; g4 < 8; ++g4) {
            const float4 cq = *reinterpret_cast<const float4*>(sC + c * 32 + g4 * 4);
            const float cc[4] = {cq.x, cq.y, cq.z, cq.w};
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int j = g4 * 4 + u;
                const float d1 = __uint_as_float(v1[j]), d2 = __uint_as_float(v2[j]);
                bool pass;
                if (COS) {
                    const float x = d1 + 1e-6f;
                    pass = !(x * fabsf(x) < cc[u] * fmaxf(d2, 0.f));
                } else {
                    pass = !(fmaf(-2.0f, d1, d2) > cc[u]);
                }
                mbits |= (pass ? 1u : 0u) << j;
            }
        }
        mbits = valid ? mbits : 0u;
#pragma unroll 1
        while (mbits) {
            const int j = __ffs(mbits) - 1;
            mbits &= mbits - 1;
            const int q = c * 32 + j;
            if (q >= nq) continue;
            uint32_t a16[16], a8[8], a4[4], a2[2], b16[16], b8[8], b4[4], b2[2];
#pragma unroll
            for (int i = 0; i < 16; ++i) { a16[i] = (j & 1) ? v1[2 * i + 1] : v1[2 * i]; b16[i] = (j & 1) ? v2[2 * i + 1] : v2[2 * i]; }
#pragma unroll
            for (int i = 0; i < 8; ++i) { a8[i] = (j & 2) ? a16[2 * i + 1] : a16[2 * i]; b8[i] = (j & 2) ? b16[2 * i + 1] : b16[2 * i]; }
#pragma unroll
            for (int i = 0; i < 4; ++i) { a4[i] = (j & 4) ? a8[2 * i + 1] : a8[2 * i]; b4[i] = (j & 4) ? b8[2 * i + 1] : b8[2 * i]; }
#pragma unroll
            for (int i = 0; i < 2; ++i) { a2[i] = (j & 8) ? a4[2 * i + 1] : a4[2 * i]; b2[i] = (j & 8) ? b4[2 * i + 1] : b4[2 * i]; }
            const float d1 = __uint_as_float((j & 16) ? a2[1] : a2[0]);
            const float d2 = __uint_as_float((j & 16) ? b2[1] : b2[0]);
            const float sv = score(d1, d2, sQ1[q], sQ2[q]);
            sink_insert_one(sink, q, make_composite(score_to_key(sv, largest), row));
        }
    }
}

template <bool COS>
__global__ void __launch_bounds__(kTwThreads, 1)
tc_weighted_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_w, const TwParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sStage = base;                                                   // [stages][A | Ba | Bw]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStage + kTwStages * kTwStage);   // producers -> squarers / MMA
    uint64_t* sq_bar = full_bar + kTwStages;                                        // squarers -> MMA
    uint64_t* empty_bar = sq_bar + kTwStages;                                       // MMA -> producers
    uint64_t* tmem_full = empty_bar + kTwStages;                                    // [2]
    uint64_t* tmem_empty = tmem_full + 2;                                           // [2]
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>(tmem_empty + 2);   // [64]
    float* sQ1 = reinterpret_cast<float*>(sThr + kTwBN);
    float* sQ2 = sQ1 + kTwBN;
    float* sThrF = sQ2 + kTwBN;
    float* sC = sThrF + kTwBN;                                                      // [64] pre-filter coefficients
    int* sCnt = reinterpret_cast<int*>(sC + kTwBN);
    uint32_t* sLmax = reinterpret_cast<uint32_t*>(sCnt + kTwBN);
    uint32_t* sHist = sLmax + kTwBN;                                                // [4][256]
    uint32_t* sTmemBase = sHist + 4 * 256;
    volatile int* sTilesDone = reinterpret_cast<volatile int*>(sTmemBase + 1);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool largest = COS;
    const int KB = p.kblocks;
    const int my_tiles = (p.num_tiles > static_cast<int>(blockIdx.x))
                             ? (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
    auto tile_of = [&](int it) -> int { return static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x); };

    if (tid == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_w);
        for (int s = 0; s < kTwStages; ++s) {
            ptx::mbar_init(&full_bar[s], kTwProducers + 1);      // 128 cp.async arrivals + the TMA issuer's expect_tx
            ptx::mbar_init(&sq_bar[s], 128);
            ptx::mbar_init(&empty_bar[s], 1);
        }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 4); }
        ptx::fence_barrier_init();
        *sTilesDone = 0;
    }
    if (warp == kTwMmaWarp) {
        ptx::tmem_alloc(sTmemBase, 512);                         // 2 x (D1 | D2) accumulators + 6 squared k-blocks
        ptx::tmem_relinquish();
    }
    for (int q = tid; q < kTwBN; q += kTwThreads) {
        sThr[q] = (q < p.nq) ? 0ull : ~0ull;
        sThrF[q] = (q < p.nq) ? __uint_as_float(0x7FC00000u) : (largest ? INFINITY : -INFINITY);
        sQ1[q] = p.qc1[q];
        sQ2[q] = p.qc2[q];
        sC[q] = (q < p.nq) ? __uint_as_float(0x7FC00000u) : (largest ? INFINITY : -INFINITY);
        sCnt[q] = 0;
        sLmax[q] = 0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *sTmemBase;

    Sink sink;
    sink.lists = p.lists + (static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0) * p.cap;
    sink.thr = smem_addr(sThr); sink.thr_f = smem_addr(sThrF); sink.cnt = smem_addr(sCnt); sink.lmax = smem_addr(sLmax);
    sink.cap = p.cap; sink.k = p.k; sink.largest = largest;

    if (warp < kTwProducerWarps) {
        // ===================== producers: bank stage by cp.async, query k-blocks by TMA =====================
        uint32_t dst_off[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = j * kTwProducers + tid;                // 16-byte chunk of the [128][64] block
            const int row = g >> 3, c = g & 7;
            dst_off[j] = static_cast<uint32_t>(row * 128 + ((c ^ (row & 7)) << 4));   // 128-byte swizzle
        }
        const uint32_t s0 = ptx::smem_u32(sStage);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const unsigned char* src = p.bank + (static_cast<size_t>(tile_of(it)) * KB) * kTwStageA + static_cast<size_t>(tid) * 16;
            for (int kb = 0; kb < KB; ++kb) {
                if (lane == 0) ptx::mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
                __syncwarp();
                const uint32_t dst = s0 + static_cast<uint32_t>(stage) * kTwStage;
                if (tid == 0) {
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], 2 * kTwStageB);
                    unsigned char* sb = sStage + static_cast<size_t>(stage) * kTwStage + kTwStageA;
                    ptx::tma_load_2d(&tmap_a, sb, &full_bar[stage], kb * kKBlock, 0, ptx::kEvictLast);
                    ptx::tma_load_2d(&tmap_w, sb + kTwStageB, &full_bar[stage], kb * kKBlock, 0, ptx::kEvictLast);
                }
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    tw_cp_async_16(dst + dst_off[j], src + static_cast<size_t>(j) * kTwProducers * 16, p.bank_policy);
                asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ptx::smem_u32(&full_bar[stage])) : "memory");
                src += kTwStageA;
                if (++stage == kTwStages) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp < kTwEpiWarp0) {
        // ===================== squarers: (Z o Z) k-block -> tensor memory =====================
        // thread = bank row of the tile (TMEM lane); its 64 bf16 sit in 8 swizzled 16-byte chunks of the stage
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t s0 = ptx::smem_u32(sStage) + row * 128;
        const uint32_t sw = static_cast<uint32_t>(row & 7);
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kTwSqBase;
        int stage = 0;
        uint32_t phase = 0;
        const int total = my_tiles * KB;
        for (int i = 0; i < total; ++i) {
            ptx::mbar_wait_relaxed(&full_bar[stage], phase, 20);
            // the previous MMAs that read this TMEM buffer retired before the producers refilled the stage (empty_bar)
            if (SKY_DBG(p) & 1) { ptx::mbar_arrive(&sq_bar[stage]); if (++stage == kTwStages) { stage = 0; phase ^= 1; } continue; }
            const uint32_t a = s0 + static_cast<uint32_t>(stage) * kTwStage;
            uint32_t v[32];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v[4 * c]), "=r"(v[4 * c + 1]), "=r"(v[4 * c + 2]), "=r"(v[4 * c + 3]) : "r"(a + ((c ^ sw) << 4)));
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&v[u]);
                x = __hmul2(x, x);
                v[u] = *reinterpret_cast<uint32_t*>(&x);
            }
            ptx::tmem_st_32x32b_x32(t0 + stage * kTwSqCols, v);
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            ptx::mbar_arrive(&sq_bar[stage]);
            if (++stage == kTwStages) { stage = 0; phase ^= 1; }
        }
    } else if (warp == kTwMmaWarp) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileRows, kTwBN);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait_relaxed(&tmem_empty[acc], acc_phase ^ 1, 32);
                ptx::tc_fence_after();
                const uint32_t d1 = tmem_base + static_cast<uint32_t>(acc * 2 * kTwBN);
                const uint32_t d2 = d1 + kTwBN;
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait_relaxed(&full_bar[stage], phase, 20);    // bank stage + query k-blocks have landed
                    ptx::mbar_wait_relaxed(&sq_bar[stage], phase, 20);      // and the squared copy is written
                    ptx::fence_proxy_async();                               // cp.async (generic proxy) writes of A
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(sStage + static_cast<size_t>(stage) * kTwStage);
                    const uint32_t ba_addr = a_addr + kTwStageA;
                    const uint32_t bw_addr = ba_addr + kTwStageB;
                    const uint32_t a2_tmem = tmem_base + kTwSqBase + stage * kTwSqCols;
#pragma unroll
                    for (int k = 0; k < kKBlock / 16; ++k) {
                        ptx::umma_bf16(d1, ptx::make_sw128_kmajor_desc(a_addr + k * 32), ptx::make_sw128_kmajor_desc(ba_addr + k * 32),
                                       idesc, (kb | k) != 0 ? 1u : 0u);
                        if (!(SKY_DBG(p) & 2)) ptx::umma_bf16_ts(d2, a2_tmem + k * 8, ptx::make_sw128_kmajor_desc(bw_addr + k * 32), idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);
                    if (++stage == kTwStages) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full[acc]);
            }
        }
    } else if (warp >= kTwEpiWarp0 && warp < kTwEpiWarp0 + 4) {
        // ===================== epilogue =====================
        const int e = warp - kTwEpiWarp0;
        const int quarter = warp & 3;
        const uint32_t hist = smem_addr(sHist + e * 256);
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = tile_of(it);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int64_t row = static_cast<int64_t>(tile) * kTileRows + quarter * 32 + lane;
            const bool valid = row < p.rows;
            if (lane == 0) ptx::mbar_wait_relaxed(&tmem_full[acc], acc_phase, 128);
            __syncwarp();
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 2 * kTwBN);
            tw_score_tile<COS>(taddr, static_cast<uint32_t>(row), valid, p.nq, SKY_DBG(p), sink, sQ1, sQ2, sC,
                               [&]() { if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]); });
            ptx::named_bar_sync(1, 128);
            sink_prune_if_full(sink, p.nq, e, 4, hist);
            if (e * 32 + lane < kTwBN) tw_prefilter_coeffs<COS>(e * 32 + lane, p.nq, sink, sQ1, sQ2, sC);
            ptx::named_bar_sync(1, 128);
            if (e == 0 && lane == 0) *sTilesDone = it + 1;
        }
        ptx::named_bar_sync(1, 128);
        for (int q = e * 32 + lane; q < p.nq; q += 128) {
            p.counts[static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q] = static_cast<int>(lds_u32(sink.cnt + q * 4));
            const uint32_t mine = lds_u32(sink.lmax + q * 4);
            if (p.use_gtau && mine) st_cg_u32(p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q, mine);
        }
    } else if (warp == kTwXchgWarp) {
        // ===================== grid-wide bound exchange (as K2) =====================
        if (p.use_gtau && my_tiles > 0) {
            uint32_t* my_row = p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0;
            const int rq = static_cast<int>(blockIdx.x) % p.nq;
            int round = 0;
            uint32_t last_pub[kTwBN / 32] = {};
            uint32_t last_lo = 0;
            while (*sTilesDone < my_tiles) {
                exchange_publish_changed(sink, p.nq, my_row, last_pub);
                const uint32_t lo = exchange_reduce(p.gtop + p.q0 + rq, p.p_stride, p.Qtot, p.k);
                if (lane == 0 && lo > last_lo) atomicMax(p.gtau + p.q0 + rq, lo);
                last_lo = lo > last_lo ? lo : last_lo;
                for (int q = lane; q < p.nq; q += 32) exchange_apply(sink, q, ld_cg_u32(p.gtau + p.q0 + q));
                ++round;
                __nanosleep(round < 24 ? 200 : 3000);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kTwMmaWarp) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 512);
    }
}

// ---------------------------------------------------------------------------------------------
// K2w2: the same search on a CTA PAIR (cta_group::2) with the query operands RESIDENT in shared memory.
//
// K2w above re-streams the 2 x 8 KB query k-blocks from L2 with every 16 KB bank stage (they do not fit beside a
// ring: 64 queries x Dp x 2 matrices = 192 KB at Dp = 768), which doubles the shared-memory fill traffic and halves
// the ring.  One tcgen05.mma.cta_group::2 of M = 256 takes 128 bank rows from EACH CTA of the pair but only HALF of
// the B rows from each: every CTA keeps just 32 of the 64 queries (both matrices: 96 KB at Dp = 768) resident for
// the whole kernel and the ring carries nothing but bank tiles, landed by TMA (7 x 16 KB at Dp = 768).
//   warp 0      bank stream (one thread): TMA of the CTA's own tiles, contiguous 16 KB (tile, k-block) boxes
//   warp 1      MMA issuer (one thread of the EVEN CTA issues for the pair); owns the pair's TMEM allocation
//   warp 2      grid-wide bound exchange (per CTA, as K2)
//   warps 4-7   squarers (per CTA): z -> z o z into the CTA's own TMEM; one arrival per warp on the even CTA's barrier
//   warps 8-11  epilogue (per CTA): own 128 rows x 64 queries, own candidate lists
// Both CTAs walk the same number of tiles in lock-step (tile = blockIdx.x + it * gridDim.x); the odd CTA's last tile
// may lie past the end of the bank, it then re-reads the last real tile and its rows are masked as invalid.
// ---------------------------------------------------------------------------------------------
constexpr int kT2Threads = 12 * 32;
constexpr int kT2ProdWarp = 0, kT2MmaWarp = 1, kT2XchgWarp = 2, kT2SquareWarp0 = 4, kT2EpiWarp0 = 8;
constexpr int kT2HalfN = kTwBN / 2;              // query rows resident per CTA
constexpr int kT2BBlock = kT2HalfN * 128;        // 4 KB: one k-block of one operand half
constexpr int kT2MaxStages = 8;                  // TMEM: 256 accumulator columns + 32 per stage <= 512

// experiment builds: timeline of cluster 0 (rows: 0 TMA issue, 1 squarer saw full (warp 0), 2 squarer arrived, 3 MMA saw
// sq_bar, 4 MMA committed, 5 MMA got tmem_empty (per tile), 6 epilogue saw tmem_full, 7 epilogue released, 8 epilogue tile
// done; the odd CTA writes rows 9.. the same way), SKY_TW_DEBUG bit 5
#ifdef SKY_EXPERIMENTS
constexpr int kTwTraceLen = 1024;
constexpr int kTwTraceRows = 13;   // + 9..11 squarer warps 1..3 arrived, 12 MMA starts waiting for sq_bar
__device__ unsigned long long g_tw_trace[2 * kTwTraceRows * kTwTraceLen];
#define TW_TRACE(row, i) do { if ((SKY_DBG(p) & 32) && blockIdx.x < 2 && (i) < kTwTraceLen && (threadIdx.x & 31) == 0) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); g_tw_trace[(blockIdx.x * kTwTraceRows + (row)) * kTwTraceLen + (i)] = t_; } } while (0)
#else
#define TW_TRACE(row, i) do { } while (0)
#endif

struct Tw2Params {
    TwParams w;
    int stages;
    int spin;       // barriers signalled from the OTHER SM are polled (no suspend hint, no back-off)
};

__device__ __forceinline__ void t2_wait_remote(uint64_t* bar, uint32_t parity, int spin) {
    if (spin) ptx::mbar_wait(bar, parity);
    else ptx::mbar_wait_relaxed(bar, parity, 32);
}

template <bool COS>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kT2Threads, 1)
tc_weighted2_kernel(const __grid_constant__ CUtensorMap tmap_bank, const __grid_constant__ CUtensorMap tmap_a,
                    const __grid_constant__ CUtensorMap tmap_w, const Tw2Params pp) {
    const TwParams& p = pp.w;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KB = p.kblocks, S = pp.stages;
    unsigned char* sBa = base;                                                      // [KB][32 rows][128 B]
    unsigned char* sBw = sBa + static_cast<size_t>(KB) * kT2BBlock;
    unsigned char* sA = sBw + static_cast<size_t>(KB) * kT2BBlock;                  // [S][128 rows][128 B]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sA + static_cast<size_t>(S) * kTwStageA);   // TMA -> squarers (per CTA)
    uint64_t* sq_bar = full_bar + kT2MaxStages;                                     // squarers of BOTH CTAs -> MMA (even CTA)
    uint64_t* empty_bar = sq_bar + kT2MaxStages;                                    // MMA -> producer (multicast commit)
    uint64_t* b_full = empty_bar + kT2MaxStages;                                    // [1] resident operands landed
    uint64_t* tmem_full = b_full + 1;                                               // [2] multicast commit
    uint64_t* tmem_empty = tmem_full + 2;                                           // [2] epilogue warps of BOTH CTAs -> MMA
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(tmem_empty + 2) + 15) & ~uintptr_t(15));
    float* sQ1 = reinterpret_cast<float*>(sThr + kTwBN);
    float* sQ2 = sQ1 + kTwBN;
    float* sThrF = sQ2 + kTwBN;
    float* sC = sThrF + kTwBN;                                                      // [64] pre-filter coefficients
    int* sCnt = reinterpret_cast<int*>(sC + kTwBN);
    uint32_t* sLmax = reinterpret_cast<uint32_t*>(sCnt + kTwBN);
    uint32_t* sHist = sLmax + kTwBN;                                                // [4][256]
    uint32_t* sTmemBase = sHist + 4 * 256;
    volatile int* sTilesDone = reinterpret_cast<volatile int*>(sTmemBase + 1);
    volatile int* sBoot = sTilesDone + 1;   // 0 = not started, 1 = tile-0 maxima ready, 2 = bounds applied

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool largest = COS;
    const uint32_t rank = ptx::cluster_ctarank();           // 0 = even CTA: issues the pair's MMAs
    const int lead = static_cast<int>(blockIdx.x) - static_cast<int>(rank);
    // lock-step: both CTAs run as many tiles as the even CTA has
    const int my_tiles = (p.num_tiles > lead) ? (p.num_tiles - lead + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
    auto tile_of = [&](int it) -> int { return static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x); };

    if (tid == 0) {
        ptx::prefetch_tmap(&tmap_bank);
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_w);
        for (int s = 0; s < S; ++s) {
            ptx::mbar_init(&full_bar[s], 1);
            ptx::mbar_init(&sq_bar[s], 8);          // 4 squarer warps x 2 CTAs
            ptx::mbar_init(&empty_bar[s], 1);
        }
        ptx::mbar_init(b_full, 1);
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 8); }
        ptx::fence_barrier_init();
        *sTilesDone = 0;
        *sBoot = 0;
    }
    if (warp == kT2MmaWarp) {
        ptx::tmem_alloc2(sTmemBase, 512);           // 2 x (D1 | D2) accumulators + S squared k-blocks, in both CTAs
        ptx::tmem_relinquish2();
    }
    // programmatic dependent launch: the set-up above ran under the tail of the packing kernel; its outputs (query
    // constants, operand matrices, zeroed exchange state) are read from here on
    ptx::griddep_wait();
    ptx::griddep_launch_dependents();
    for (int q = tid; q < kTwBN; q += kT2Threads) {
        sThr[q] = (q < p.nq) ? 0ull : ~0ull;
        sThrF[q] = (q < p.nq) ? __uint_as_float(0x7FC00000u) : (largest ? INFINITY : -INFINITY);
        sQ1[q] = p.qc1[q];
        sQ2[q] = p.qc2[q];
        sC[q] = (q < p.nq) ? __uint_as_float(0x7FC00000u) : (largest ? INFINITY : -INFINITY);
        sCnt[q] = 0;
        sLmax[q] = 0;
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();                        // the peer's barriers are initialised before anyone signals them
    ptx::tc_fence_after();
    const uint32_t tmem_base = *sTmemBase;

    Sink sink;
    sink.lists = p.lists + (static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0) * p.cap;
    sink.thr = smem_addr(sThr); sink.thr_f = smem_addr(sThrF); sink.cnt = smem_addr(sCnt); sink.lmax = smem_addr(sLmax);
    sink.cap = p.cap; sink.k = p.k; sink.largest = largest;

    if (warp == kT2ProdWarp) {
        // ===================== producer: resident query halves once, then the bank stream, all by TMA =====================
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(b_full, static_cast<uint32_t>(KB) * 2 * kT2BBlock);
            for (int kb = 0; kb < KB; ++kb) {
                ptx::tma_load_2d(&tmap_a, sBa + static_cast<size_t>(kb) * kT2BBlock, b_full, kb * kKBlock, static_cast<int>(rank) * kT2HalfN, ptx::kEvictLast);
                ptx::tma_load_2d(&tmap_w, sBw + static_cast<size_t>(kb) * kT2BBlock, b_full, kb * kKBlock, static_cast<int>(rank) * kT2HalfN, ptx::kEvictLast);
            }
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int tile = min(tile_of(it), p.num_tiles - 1);
                for (int kb = 0; kb < KB; ++kb) {
                    t2_wait_remote(&empty_bar[stage], phase ^ 1, pp.spin);
                    TW_TRACE(0, it * KB + kb);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], kTwStageA);
                    ptx::tma_load_2d(&tmap_bank, sA + static_cast<size_t>(stage) * kTwStageA, &full_bar[stage], 0,
                                     (tile * KB + kb) * kTileRows, p.bank_policy);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp >= kT2SquareWarp0 && warp < kT2SquareWarp0 + 4) {
        // ===================== squarers: (Z o Z) k-block -> this CTA's tensor memory =====================
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        const uint32_t s0 = ptx::smem_u32(sA) + row * 128;
        const uint32_t sw = static_cast<uint32_t>(row & 7);
        const uint32_t t0 = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + kTwSqBase;
        int stage = 0;
        uint32_t phase = 0;
        const int total = my_tiles * KB;
        if (total > 0) ptx::mbar_wait_relaxed(b_full, 0, 64);    // the first sq_bar arrival also vouches for the resident operands
        for (int i = 0; i < total; ++i) {
            ptx::mbar_wait_relaxed(&full_bar[stage], phase, 20);
            if (quarter == 0 && lane == 0) TW_TRACE(1, i);
            if (SKY_DBG(p) & 1) {      // experiment: no squaring, only the hand-off
                __syncwarp();
                if (lane == 0) { if (rank == 0) ptx::mbar_arrive(&sq_bar[stage]); else ptx::mbar_arrive_cluster(&sq_bar[stage], 0); }
                if (++stage == S) { stage = 0; phase ^= 1; }
                continue;
            }
            const uint32_t a = s0 + static_cast<uint32_t>(stage) * kTwStageA;
            uint32_t v[32];
#pragma unroll
            for (int c = 0; c < 8; ++c)
                asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];"
                             : "=r"(v[4 * c]), "=r"(v[4 * c + 1]), "=r"(v[4 * c + 2]), "=r"(v[4 * c + 3]) : "r"(a + ((c ^ sw) << 4)));
#pragma unroll
            for (int u = 0; u < 32; ++u) {
                __nv_bfloat162 x = *reinterpret_cast<__nv_bfloat162*>(&v[u]);
                x = __hmul2(x, x);
                v[u] = *reinterpret_cast<uint32_t*>(&x);
            }
            ptx::tmem_st_32x32b_x32(t0 + stage * kTwSqCols, v);
            ptx::tmem_st_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (rank == 0) ptx::mbar_arrive(&sq_bar[stage]);
                else ptx::mbar_arrive_cluster(&sq_bar[stage], 0);
                if (quarter == 0) TW_TRACE(2, i); else TW_TRACE(8 + quarter, i);
            }
            if (++stage == S) { stage = 0; phase ^= 1; }
        }
    } else if (warp == kT2MmaWarp) {
        // ===================== MMA issuer: the even CTA's warp drives both SMs =====================
        // The warp stays converged: every lane polls the barriers and carries the (warp-uniform) descriptors, one
        // elected lane issues.  The descriptors are loop-carried and only ever incremented (a stage is 16 KB = 1024 in
        // the descriptor's address >> 4 field, a resident k-block 4 KB = 256, a k-step 32 B = 2): rebuilding them from
        // addresses every k-block costs ~50 dependent uniform-datapath instructions, several hundred cycles per
        // stage on the one thread everything else waits for.
        if (rank == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(2 * kTileRows, kTwBN);
            const uint32_t issuer = ptx::elect_one();
            const uint64_t a_desc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sA));
            const uint64_t ba_desc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sBa));
            const uint64_t bw_desc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sBw));
            const uint32_t sq0 = tmem_base + kTwSqBase;
            int stage = 0;
            uint32_t phase = 0;
            uint64_t a_desc = a_desc0;
            uint32_t a2_tmem = sq0;
            for (int it = 0; it < my_tiles; ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                t2_wait_remote(&tmem_empty[acc], acc_phase ^ 1, pp.spin);
                TW_TRACE(5, it);
                ptx::tc_fence_after();
                const uint32_t d1 = tmem_base + static_cast<uint32_t>(acc * 2 * kTwBN);
                const uint32_t d2 = d1 + kTwBN;
                uint64_t ba_desc = ba_desc0, bw_desc = bw_desc0;
                for (int kb = 0; kb < KB; ++kb) {
                    // 8 arrivals: each squarer warp of either CTA saw its own stage land (and, first time, its resident
                    // operands) and finished writing the squared copy
                    TW_TRACE(12, it * KB + kb);
                    t2_wait_remote(&sq_bar[stage], phase, pp.spin);
                    TW_TRACE(3, it * KB + kb);
                    ptx::tc_fence_after();
                    const uint32_t acc0 = kb != 0 ? 1u : 0u;
                    if (issuer) {
                        if (!(SKY_DBG(p) & 8)) {
#pragma unroll
                            for (int k = 0; k < kKBlock / 16; ++k) ptx::umma2_bf16(d1, a_desc + 2 * k, ba_desc + 2 * k, idesc, k ? 1u : acc0);
                        }
                        if (!(SKY_DBG(p) & 2)) {
#pragma unroll
                            for (int k = 0; k < kKBlock / 16; ++k) ptx::umma2_bf16_ts(d2, a2_tmem + k * 8, bw_desc + 2 * k, idesc, k ? 1u : acc0);
                        }
                        ptx::umma2_commit_mc(&empty_bar[stage], 0b11);      // both producers may refill the stage
                        if (kb == KB - 1) ptx::umma2_commit_mc(&tmem_full[acc], 0b11);   // both epilogues may read
                        TW_TRACE(4, it * KB + kb);
                    }
                    ba_desc += kT2BBlock >> 4;
                    bw_desc += kT2BBlock >> 4;
                    a_desc += kTwStageA >> 4;
                    a2_tmem += kTwSqCols;
                    if (++stage == S) { stage = 0; phase ^= 1; a_desc = a_desc0; a2_tmem = sq0; }
                }
            }
        }
    } else if (warp >= kT2EpiWarp0 && warp < kT2EpiWarp0 + 4) {
        // ===================== epilogue (own 128 rows, all 64 queries) =====================
        const int e = warp - kT2EpiWarp0;
        const int quarter = warp & 3;
        const uint32_t hist = smem_addr(sHist + e * 256);
        for (int it = 0; it < my_tiles; ++it) {
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int64_t row = static_cast<int64_t>(tile_of(it)) * kTileRows + quarter * 32 + lane;
            const bool valid = row < p.rows;
            if (lane == 0) { if (pp.spin) ptx::mbar_wait(&tmem_full[acc], acc_phase); else ptx::mbar_wait_relaxed(&tmem_full[acc], acc_phase, 128); if (e == 0) TW_TRACE(6, it); }
            __syncwarp();
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * 2 * kTwBN);
            if (it == 0 && p.use_gtau && !(SKY_DBG(p) & 4)) {
                // BOOTSTRAP (once per launch, as K2): with no bound yet every row of the first tile would be a candidate
                // for every query (8192 inserts per CTA, and lists that the merge then has to wade through).  Read the
                // accumulators twice: this first pass only takes the per-query maximum of the tile, the exchange warp
                // trades maxima with the other CTAs, and the normal pass below runs with a grid-wide bound in place.
#pragma unroll 1
                for (int c = 0; c < kTwBN / 32; ++c) {
                    uint32_t v1[32], v2[32];
                    ptx::tmem_ld_32x32b_x32(taddr + c * 32, v1);
                    ptx::tmem_ld_32x32b_x32(taddr + kTwBN + c * 32, v2);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int q = c * 32 + j;
                        const float d1 = __uint_as_float(v1[j]), d2 = __uint_as_float(v2[j]);
                        const float sv = COS ? __fdividef(d1, fmaf(sQ1[q], sqrtf(fmaxf(d2, 0.f)), 1e-6f)) : (sQ1[q] - 2.0f * d1 + d2) * sQ2[q];
                        const uint32_t key = (valid && q < p.nq) ? score_to_key(sv, largest) : 0u;
                        const uint32_t best = __reduce_max_sync(0xffffffffu, key);
                        if (lane == 0 && best) reds_max_u32(sink.lmax + q * 4, best);
                    }
                }
                ptx::named_bar_sync(1, 128);
                if (e == 0 && lane == 0) *sBoot = 1;
                if (lane == 0) {
                    const long long t_end = clock64() + 60000;         // ~30 us: never wait for a bound forever
                    while (*sBoot != 2 && clock64() < t_end) __nanosleep(100);
                }
                __syncwarp();
                if (e * 32 + lane < kTwBN) tw_prefilter_coeffs<COS>(e * 32 + lane, p.nq, sink, sQ1, sQ2, sC);
                ptx::named_bar_sync(1, 128);
            }
            tw_score_tile<COS>(taddr, static_cast<uint32_t>(row), valid, p.nq, SKY_DBG(p), sink, sQ1, sQ2, sC, [&]() {
                if (lane == 0) {
                    if (rank == 0) ptx::mbar_arrive(&tmem_empty[acc]);
                    else ptx::mbar_arrive_cluster(&tmem_empty[acc], 0);
                    if (e == 0) TW_TRACE(7, it);
                }
            });
            ptx::named_bar_sync(1, 128);
            sink_prune_if_full(sink, p.nq, e, 4, hist);
            if (e * 32 + lane < kTwBN) tw_prefilter_coeffs<COS>(e * 32 + lane, p.nq, sink, sQ1, sQ2, sC);
            ptx::named_bar_sync(1, 128);
            if (e == 0 && lane == 0) { *sTilesDone = it + 1; TW_TRACE(8, it); }
        }
        ptx::named_bar_sync(1, 128);
        for (int q = e * 32 + lane; q < p.nq; q += 128) {
            p.counts[static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q] = static_cast<int>(lds_u32(sink.cnt + q * 4));
            const uint32_t mine = lds_u32(sink.lmax + q * 4);
            if (p.use_gtau && mine) st_cg_u32(p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q, mine);
        }
    } else if (warp == kT2XchgWarp) {
        // ===================== grid-wide bound exchange (as K2) =====================
        if (p.use_gtau && my_tiles > 0) {
            uint32_t* my_row = p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0;
            const int rq = static_cast<int>(blockIdx.x) % p.nq;
            int round = 0;
            uint32_t last_pub[kTwBN / 32] = {};
            uint32_t last_lo = 0;
            // bootstrap: wait for the tile-0 maxima, then trade them until every query has a bound
            while (*sBoot == 0 && *sTilesDone < my_tiles) __nanosleep(50);
            {
                const long long t_end = clock64() + 40000;
                bool all = false;
                while (!all && clock64() < t_end) {
                    exchange_publish_changed(sink, p.nq, my_row, last_pub);
                    const uint32_t lo = exchange_reduce(p.gtop + p.q0 + rq, p.p_stride, p.Qtot, p.k);
                    if (lane == 0 && lo > last_lo) atomicMax(p.gtau + p.q0 + rq, lo);
                    last_lo = lo > last_lo ? lo : last_lo;
                    bool mine_ok = true;
                    for (int q = lane; q < p.nq; q += 32) {
                        const uint32_t g = ld_cg_u32(p.gtau + p.q0 + q);
                        exchange_apply(sink, q, g);
                        mine_ok = mine_ok && (g != 0u);
                    }
                    all = __all_sync(0xffffffffu, mine_ok);
                }
                __syncwarp();
                if (lane == 0) *sBoot = 2;
            }
            while (*sTilesDone < my_tiles) {
                exchange_publish_changed(sink, p.nq, my_row, last_pub);
                const uint32_t lo = exchange_reduce(p.gtop + p.q0 + rq, p.p_stride, p.Qtot, p.k);
                if (lane == 0 && lo > last_lo) atomicMax(p.gtau + p.q0 + rq, lo);
                last_lo = lo > last_lo ? lo : last_lo;
                for (int q = lane; q < p.nq; q += 32) exchange_apply(sink, q, ld_cg_u32(p.gtau + p.q0 + q));
                ++round;
                __nanosleep(round < 24 ? 200 : 3000);
            }
        }
    }

    // neither CTA may leave (or free tensor memory) while the pair's MMAs still read its shared / tensor memory
    ptx::tc_fence_before();
    __syncthreads();
    ptx::cluster_sync_all();
    if (warp == kT2MmaWarp) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc2(tmem_base, 512);
    }
}

#ifdef SKY_EXPERIMENTS
int debug_read_tw_trace(unsigned long long* h_out, int n) {
    if (n > 2 * kTwTraceRows * kTwTraceLen) n = 2 * kTwTraceRows * kTwTraceLen;
    SKY_CUDA(cudaDeviceSynchronize());
    SKY_CUDA(cudaMemcpyFromSymbol(h_out, g_tw_trace, sizeof(unsigned long long) * n));
    return SKY_OK;
}
#endif

// a = w o t and w as bf16 operand matrices [64, Dp] (zero padded) + per-query constants
__global__ void pack_weighted_kernel(const float* __restrict__ t, const float* __restrict__ w, int nq, int D, int Dp, int metric,
                                     __nv_bfloat16* __restrict__ ba, __nv_bfloat16* __restrict__ bw,
                                     float* __restrict__ qc1, float* __restrict__ qc2, const StateInit si) {
    ptx::griddep_launch_dependents();      // the scorer behind this kernel may set itself up now (it waits before reading)
    state_init_gridwide(si);
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ double red[2][8];
    double s_wtt = 0.0, s_w = 0.0;
    for (int d = threadIdx.x; d < Dp; d += blockDim.x) {
        float tv = 0.f, wv = 0.f;
        if (q < nq && d < D) { tv = t[static_cast<size_t>(q) * D + d]; wv = w[static_cast<size_t>(q) * D + d]; }
        const __nv_bfloat16 ar = __float2bfloat16_rn(wv * tv), wr = __float2bfloat16_rn(wv);
        ba[static_cast<size_t>(q) * Dp + d] = ar;
        bw[static_cast<size_t>(q) * Dp + d] = wr;
        // sum w t^2 of the ROUNDED operands a~ = bf16(w t), w~ = bf16(w): sum a~^2 / w~, so that
        // sum w t^2 - 2 a.z + w.z^2 = sum w~ (a~/w~ - z)^2 stays a true weighted squared distance
        const float af = __bfloat162float(ar), wf = __bfloat162float(wr);
        s_wtt += (wf > 0.f) ? static_cast<double>(af) * af / wf : static_cast<double>(wv * (tv * tv));
        s_w += static_cast<double>(wv);
    }
    for (int off = 16; off > 0; off >>= 1) { s_wtt += __shfl_xor_sync(0xffffffffu, s_wtt, off); s_w += __shfl_xor_sync(0xffffffffu, s_w, off); }
    if (lane == 0) { red[0][warp] = s_wtt; red[1][warp] = s_w; }
    __syncthreads();
    if (threadIdx.x == 0) {
        double a = 0.0, b = 0.0;
        for (int i = 0; i < nw; ++i) { a += red[0][i]; b += red[1][i]; }
        const float wtt = static_cast<float>(a), sw = static_cast<float>(b);
        qc1[q] = (metric == SKY_COSINE) ? sqrtf(wtt) : wtt;
        qc2[q] = (metric == SKY_COSINE) ? 0.f : 1.0f / (sw * static_cast<float>(D));
    }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int make_tmap_2d(CUtensorMap* m, const void* ptr, int64_t rows, int Dp, int box_rows);     // tc_search.cu

bool tc_weighted_supported(const sky_bank* b, int metric, bool weighted, int n_top) {
    return b->dtype == SKY_BF16 && b->L == 1 && weighted && n_top == 0 && (metric == SKY_COSINE || metric == SKY_MSE) && b->rows > 0;
}

size_t tc_weighted_scratch_bytes(const sky_bank* b) {
    return 2 * static_cast<size_t>(kTwBN) * b->Dp * 2 + 2 * kTwBN * sizeof(float) + 512;
}

static size_t tw2_tail_bytes() {
    return (3 * kT2MaxStages + 1 + 4) * sizeof(uint64_t) + 16 + kTwBN * (8 + 6 * 4) + 4 * 256 * 4 + 32;
}

// ring depth of the CTA-pair kernel (0 = the resident query halves do not leave room for a useful ring)
static int tw2_stages(int Dp) {
    const size_t budget = 227 * 1024 - 1024 /*alignment slack*/ - tw2_tail_bytes();
    const size_t resident = static_cast<size_t>(Dp / kKBlock) * 2 * kT2BBlock;
    if (resident + 4 * kTwStageA > budget) return 0;
    size_t s = (budget - resident) / kTwStageA;
    if (s > kT2MaxStages) s = kT2MaxStages;
    return static_cast<int>(s);
}

static bool tw2_usable(const sky_bank* b) {
    return b->tmap_ready && b->num_sms >= 2 && tw2_stages(b->Dp) >= 4 && env_knob("SKY_TW_PAIR", 1) != 0;
}

// CTAs of the weighted tensor scorer: the pair kernel needs an even grid
int tc_weighted_grid(const sky_bank* b) {
    if (!tw2_usable(b)) return tc_grid(b);
    const int64_t tiles = (b->rows + kTileRows - 1) / kTileRows;
    const int64_t want = (tiles + 1) & ~static_cast<int64_t>(1);
    const int cap = b->num_sms & ~1;
    return static_cast<int>(want < cap ? want : cap);
}

// scratch (bank->ws2): Ba [64, Dp] bf16 | Bw [64, Dp] bf16 | qc1 [64] | qc2 [64]; one launch per 64 queries
int launch_tc_weighted(sky_bank* b, const float* t, const float* w, int Q, int metric, const SearchState& s, cudaStream_t st) {
    unsigned char* ws = reinterpret_cast<unsigned char*>(b->ws2);
    __nv_bfloat16* ba = reinterpret_cast<__nv_bfloat16*>(ws);
    __nv_bfloat16* bw = ba + static_cast<size_t>(kTwBN) * b->Dp;
    float* qc1 = reinterpret_cast<float*>(ws + round_up(2 * static_cast<int64_t>(kTwBN) * b->Dp * 2, 256));
    float* qc2 = qc1 + kTwBN;
    const bool pair = tw2_usable(b) && (s.P % 2) == 0;
    const int stages2 = pair ? tw2_stages(b->Dp) : 0;
    const size_t smem = pair ? 1024 + static_cast<size_t>(b->Dp / kKBlock) * 2 * kT2BBlock + static_cast<size_t>(stages2) * kTwStageA + tw2_tail_bytes()
                             : 1024 + static_cast<size_t>(kTwStages) * kTwStage + (3 * kTwStages + 4) * 8 + kTwBN * (8 + 6 * 4) + 4 * 256 * 4 + 64;
    if (pair) {
        if (metric == SKY_COSINE) SKY_CUDA(cudaFuncSetAttribute(tc_weighted2_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        else SKY_CUDA(cudaFuncSetAttribute(tc_weighted2_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    } else {
        if (metric == SKY_COSINE) SKY_CUDA(cudaFuncSetAttribute(tc_weighted_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
        else SKY_CUDA(cudaFuncSetAttribute(tc_weighted_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    }
    for (int q0 = 0; q0 < Q; q0 += kTwBN) {
        const int nq = (Q - q0 < kTwBN) ? (Q - q0) : kTwBN;
        pack_weighted_kernel<<<kTwBN, 256, 0, st>>>(t + static_cast<size_t>(q0) * b->D, w + static_cast<size_t>(q0) * b->D, nq, b->D, b->Dp,
                                                   metric, ba, bw, qc1, qc2,
                                                   q0 == 0 ? StateInit{s.gtop, s.gtau, s.counts, s.Qtot, s.p_stride, s.P} : StateInit{});
        SKY_LAUNCH_CHECK("pack_weighted_kernel");
        CUtensorMap tma, tmw;
        // box = the rows one CTA keeps: all 64 (streamed per stage) or its resident half of 32 (pair kernel)
        int rc = make_tmap_2d(&tma, ba, kTwBN, b->Dp, pair ? kT2HalfN : kTwBN);
        if (rc) return rc;
        rc = make_tmap_2d(&tmw, bw, kTwBN, b->Dp, pair ? kT2HalfN : kTwBN);
        if (rc) return rc;
        TwParams p;
        p.bank = reinterpret_cast<const unsigned char*>(b->data);
        p.qc1 = qc1; p.qc2 = qc2;
        p.lists = s.lists; p.counts = s.counts; p.gtop = s.gtop; p.gtau = s.gtau;
        p.p_stride = s.p_stride; p.Qtot = s.Qtot; p.q0 = q0; p.nq = nq; p.cap = s.cap; p.k = s.k; p.use_gtau = s.use_gtau;
        p.rows = b->rows;
        p.num_tiles = static_cast<int>((b->rows + kTileRows - 1) / kTileRows);
        p.kblocks = b->Dp / kKBlock;
        p.bank_policy = ptx::kEvictFirst;
        p.debug = env_knob("SKY_TW_DEBUG", 0);
        prof_mark(b, st);
        if (pair) {
            Tw2Params pp;
            pp.w = p;
            pp.w.bank_policy = env_knob("SKY_TW2_POLICY", 0) == 1 ? ptx::kEvictFirst : 0x1000000000000000ull;   // evict-normal, as K2b's bank boxes
            pp.stages = stages2;
            { const int e = env_knob("SKY_TW2_STAGES", 0); if (e >= 2 && e < stages2) pp.stages = e; }
            pp.spin = env_knob("SKY_TW2_SPIN", 0);      // measured: polling with back-off beats spinning (0.346 vs 0.374 ms)
            if (env_knob("SKY_TW2_OCC", 0)) {
                int ncl = -1;
                cudaLaunchConfig_t cfg = {};
                cfg.gridDim = dim3(s.P); cfg.blockDim = dim3(kT2Threads); cfg.dynamicSmemBytes = smem;
                cudaLaunchAttribute at[1];
                at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
                cfg.attrs = at; cfg.numAttrs = 1;
                cudaOccupancyMaxActiveClusters(&ncl, tc_weighted2_kernel<true>, &cfg);
                fprintf(stderr, "[sky] K2w2 grid %d, smem %zu, stages %d, max active clusters %d\n", s.P, smem, pp.stages, ncl);
            }
            const bool pdl = env_knob("SKY_PDL", kUsePdl) != 0;
            if (metric == SKY_COSINE) SKY_CUDA(launch_maybe_pdl(pdl, tc_weighted2_kernel<true>, dim3(s.P), dim3(kT2Threads), smem, st, b->tmap_bank, tma, tmw, pp));
            else SKY_CUDA(launch_maybe_pdl(pdl, tc_weighted2_kernel<false>, dim3(s.P), dim3(kT2Threads), smem, st, b->tmap_bank, tma, tmw, pp));
        } else if (metric == SKY_COSINE) tc_weighted_kernel<true><<<s.P, kTwThreads, smem, st>>>(tma, tmw, p);
        else tc_weighted_kernel<false><<<s.P, kTwThreads, smem, st>>>(tma, tmw, p);
        prof_mark(b, st);
        SKY_LAUNCH_CHECK("tc_weighted_kernel");
    }
    return SKY_OK;
}

}  // namespace sky
