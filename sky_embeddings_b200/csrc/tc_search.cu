// tc_search.cu -- K2: tcgen05 tensor-core scorer with fused top-k (bf16 bank, sm_100a).
//
// The contraction form of the reference metrics for queries without
// per-feature weights (use_weights=False, reference utils/similarity.py:246-247):
//   cosine :163-170   s = t.z / (|t| |z| + 1e-6)            |z|^2 = row norm stored with the bank
//   MSE    :188-192   s = (|t|^2 - 2 t.z + |z|^2) / D^2     (rank-equivalent to L2)
// One persistent CTA per SM.  Warp roles:
//   warps 0-3   bank stream: thread 0 lands each contiguous 16 KB (tile, k-block) piece with ONE TMA box load
//               (128-byte swizzle applied by the TMA unit) and loads the query matrix B[BN, Dp] once by TMA
//               (resident in shared memory for the whole kernel); the other producer threads idle.  The round-1
//               stream (128 threads x 8 16-byte cp.async per stage) is kept behind use_tma = 0: it measured
//               3 % slower on BASELINE config 2 (0.2518 vs 0.2447 ms)
//   warps 4-7   epilogue: tcgen05.ld accumulators -> conservative pre-filter bitmask -> exact insert of survivors
//   warp 8      MMA issuer (one thread): tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM, two accumulator
//               stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warp 9      grid-wide threshold exchange (publishes / refreshes the top-k lower bound)
// The bank streams from HBM exactly once per launch; no score ever goes to HBM.
#include <cstdlib>

#include "bank.cuh"
#include "ptx.cuh"
#include "topk.cuh"

namespace sky {

constexpr int kProducerWarps = 4;               // warps 0-3: cp.async bank stream (warp 0 also loads B by TMA)
constexpr int kEpiWarp0 = 4;                    // warps 4-7: epilogue (warp & 3 = TMEM lane quarter)
constexpr int kMmaWarp = 8;                     // warp 8: MMA issuer, owns TMEM
constexpr int kXchgWarp = 9;                    // warp 9: grid-wide threshold exchange
constexpr int kTcThreads = 10 * 32;
constexpr int kProducerThreads = kProducerWarps * 32;
constexpr int kEpiThreads = 128;
constexpr int kStageBytes = kTileRows * 128;   // 128 rows x 64 bf16 = one contiguous (tile, k-block) of the bank
constexpr int kMaxStages = 8;

struct TcParams {
    const unsigned char* bank;   // tile-major bf16 bank
    const float* rownorm;   // [rows_pad]
    const float* qconst;    // [BN] cosine: |t| ; MSE: |t|^2
    uint64_t* lists; int* counts; uint32_t* gtop; uint32_t* gtau;
    int p_stride, Qtot, q0, nq, cap, k, use_gtau;
    int64_t rows;           // valid bank rows
    int num_tiles, kblocks, stages, metric;
    unsigned long long bank_policy;   // L2 cache hint of the bank stream
    int debug;              // bit0: epilogue skips scoring, bit1: MMA skipped, bit2: contiguous tiles (experiments)
    int use_tma;            // bank stages by one TMA box load each (else 1024 16-byte cp.async by 128 threads)
    float inv_dd;           // 1 / D^2
};

// 16-byte async copy global -> shared with an L2 eviction hint (the bank is streamed once)
__device__ __forceinline__ void cp_async_16(uint32_t smem_dst, const void* gsrc, uint64_t policy) {
    asm volatile("cp.async.cg.shared.global.L2::cache_hint [%0], [%1], 16, %2;" ::"r"(smem_dst), "l"(gsrc), "l"(policy)
                 : "memory");
}
// the mbarrier receives one arrival from this thread once all its earlier cp.async have landed
__device__ __forceinline__ void cp_async_arrive_noinc(uint64_t* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(ptx::smem_u32(bar)) : "memory");
}

// debug timeline (SKY_TC_DEBUG bit 3): clock64 stamps of CTA 0, rows: 0 producer issue, 1 MMA saw full,
// 2 MMA committed, 3 epilogue saw tmem_full, 4 epilogue tile done
#ifdef SKY_EXPERIMENTS
constexpr int kTraceLen = 1024;
__device__ unsigned long long g_trace[6 * kTraceLen];   // row 5: per-tile counters of CTA 0 (slow groups of warp e=0)
__device__ unsigned long long g_epi[64 * 12];   // per-tile stamps inside the epilogue of CTA 0, warp e=0
#define SKY_EPI(it, slot) do { if ((SKY_DBG(p) & 8) && blockIdx.x == 0 && e == 0 && lane == 0 && (it) < 64) g_epi[(it) * 12 + (slot)] = clock64(); } while (0)
#define SKY_TRACE(row, i) do { if ((SKY_DBG(p) & 8) && blockIdx.x == 0 && (i) < kTraceLen) g_trace[(row) * kTraceLen + (i)] = clock64(); } while (0)
#else
#define SKY_EPI(it, slot) do { } while (0)
#define SKY_TRACE(row, i) do { } while (0)
#endif

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_search_kernel(const __grid_constant__ CUtensorMap tmap_q, const __grid_constant__ CUtensorMap tmap_bank, const TcParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // 1024-byte alignment is required by the 128-byte swizzle atoms
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const int KB = p.kblocks, S = p.stages;
    unsigned char* sB = base;                                   // [KB][BN rows][128 B]
    unsigned char* sA = sB + static_cast<size_t>(KB) * BN * 128;   // [S][128 rows][128 B]
    unsigned char* tail = sA + static_cast<size_t>(S) * kStageBytes;
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(tail);     // [kMaxStages]
    uint64_t* empty_bar = full_bar + kMaxStages;                // [kMaxStages]
    uint64_t* b_full = empty_bar + kMaxStages;                  // [1]
    uint64_t* tmem_full = b_full + 1;                           // [2]
    uint64_t* tmem_empty = tmem_full + 2;                       // [2]
    // 16-byte aligned: the threshold / constant arrays are read with 128-bit shared loads
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>((reinterpret_cast<uintptr_t>(tmem_empty + 2) + 15) & ~uintptr_t(15));   // [BN]
    float* sQc = reinterpret_cast<float*>(sThr + BN);           // [BN]
    float* sThrF = sQc + BN;                                    // [BN] threshold as a score (pre-filter)
    int* sCnt = reinterpret_cast<int*>(sThrF + BN);             // [BN]
    uint32_t* sLmax = reinterpret_cast<uint32_t*>(sCnt + BN);   // [BN]
    uint32_t* sHist = sLmax + BN;                               // [4][256]
    uint32_t* sTmemBase = sHist + 4 * 256;                      // [1]
    volatile int* sTilesDone = reinterpret_cast<volatile int*>(sTmemBase + 1);   // [1]
    volatile int* sBoot = sTilesDone + 1;   // [1] 0 = not started, 1 = tile-0 maxima ready, 2 = bounds applied

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool largest = (p.metric == SKY_COSINE);

    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int tiles_per_cta = (p.num_tiles + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x);
    int my_tiles = (p.num_tiles > static_cast<int>(blockIdx.x))
                       ? (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                       : 0;
    if (SKY_DBG(p) & 4) {   // experiment: contiguous tile range per CTA
        const int lo = static_cast<int>(blockIdx.x) * tiles_per_cta;
        my_tiles = p.num_tiles > lo ? min(tiles_per_cta, p.num_tiles - lo) : 0;
    }
    auto tile_of = [&](int it) -> int {
        return (SKY_DBG(p) & 4) ? static_cast<int>(blockIdx.x) * tiles_per_cta + it : static_cast<int>(blockIdx.x + it * gridDim.x);
    };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_q);
        for (int s = 0; s < S; ++s) { ptx::mbar_init(&full_bar[s], p.use_tma ? 1 : kProducerThreads); ptx::mbar_init(&empty_bar[s], 1); }
        if (p.use_tma) ptx::prefetch_tmap(&tmap_bank);
        ptx::mbar_init(b_full, 1);
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 4); }
        ptx::fence_barrier_init();
        *sTilesDone = 0;
        *sBoot = 0;
    }
    if (warp == kMmaWarp) {
        ptx::tmem_alloc(sTmemBase, 2 * BN);
        ptx::tmem_relinquish();
    }
    // everything above overlaps the tail of the packing kernel (programmatic dependent launch); from here on the
    // kernel reads what that kernel wrote: query constants, the bf16 query matrix, the zeroed exchange state
    ptx::griddep_wait();
    ptx::griddep_launch_dependents();      // the merge kernel may be scheduled as soon as SMs free up
    for (int q = tid; q < BN; q += kTcThreads) {
        sThr[q] = (q < p.nq) ? 0ull : ~0ull;
        // NaN lets everything through the pre-filter; padding queries are rejected by it
        sThrF[q] = (q < p.nq) ? __uint_as_float(0x7FC00000u) : (largest ? INFINITY : -INFINITY);
        sQc[q] = p.qconst[q];
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

    if (warp < kProducerWarps) {
        // ===================== bank stream: cp.async producers =====================
        // Each (tile, k-block) is one contiguous 16 KB block in HBM (tile-major bank).  128 threads copy
        // it as 1024 16-byte chunks into the 128-byte-swizzled layout the UMMA descriptors expect:
        // chunk (row, c) -> row * 128 + ((c ^ (row & 7)) << 4).  The stage's full barrier gets one
        // arrival per producer thread when that thread's copies have landed.
        if (tid == 0) {
            ptx::mbar_arrive_expect_tx(b_full, static_cast<uint32_t>(KB) * BN * 128);
            for (int kb = 0; kb < KB; ++kb)
                ptx::tma_load_2d(&tmap_q, sB + static_cast<size_t>(kb) * BN * 128, b_full, kb * kKBlock, 0, ptx::kEvictLast);
        }
        if (p.use_tma) {
            // one thread, one TMA box per stage: the tile-major bank is a [rows * KB, 64] tensor whose 128-row boxes are
            // the contiguous 16 KB (tile, k-block) pieces; the TMA unit applies the 128-byte swizzle on the way in
            if (tid == 0) {
                int stage = 0;
                uint32_t phase = 0;
                for (int it = 0; it < my_tiles; ++it) {
                    const int tile = tile_of(it);
                    for (int kb = 0; kb < KB; ++kb) {
                        ptx::mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
                        ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                        ptx::tma_load_2d(&tmap_bank, sA + static_cast<size_t>(stage) * kStageBytes, &full_bar[stage], 0,
                                         (tile * KB + kb) * kTileRows, p.bank_policy);
                        if (++stage == S) { stage = 0; phase ^= 1; }
                    }
                }
            }
        } else {
        uint32_t dst_off[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int g = j * kProducerThreads + tid;           // chunk index in the block
            const int row = g >> 3, c = g & 7;
            dst_off[j] = static_cast<uint32_t>(row * 128 + ((c ^ (row & 7)) << 4));
        }
        const uint32_t sA_addr = ptx::smem_u32(sA);
        int stage = 0;
        uint32_t phase = 0;
        for (int it = 0; it < my_tiles; ++it) {
            const unsigned char* src = p.bank + (static_cast<size_t>(tile_of(it)) * KB) * kStageBytes + static_cast<size_t>(tid) * 16;
            for (int kb = 0; kb < KB; ++kb) {
                if (lane == 0) ptx::mbar_wait_relaxed(&empty_bar[stage], phase ^ 1);
                __syncwarp();
                if (tid == 0) SKY_TRACE(0, it * KB + kb);
                const uint32_t dst = sA_addr + static_cast<uint32_t>(stage) * kStageBytes;
#pragma unroll
                for (int j = 0; j < 8; ++j)
                    cp_async_16(dst + dst_off[j], src + static_cast<size_t>(j) * kProducerThreads * 16, p.bank_policy);
                cp_async_arrive_noinc(&full_bar[stage]);
                src += kStageBytes;
                if (++stage == S) { stage = 0; phase ^= 1; }
            }
        }
        }
    } else if (warp == kMmaWarp) {
        // ===================== MMA issuer =====================
        if (lane == 0) {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileRows, BN);
            ptx::mbar_wait(b_full, 0);
            ptx::tc_fence_after();
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int acc = it & 1;
                const uint32_t acc_phase = (it >> 1) & 1;
                ptx::mbar_wait_relaxed(&tmem_empty[acc], acc_phase ^ 1, 32);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait_relaxed(&full_bar[stage], phase, 20);
                    SKY_TRACE(1, it * KB + kb);
                    ptx::fence_proxy_async();                 // cp.async (generic proxy) writes -> UMMA (async proxy) reads
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(sA + static_cast<size_t>(stage) * kStageBytes);
                    const uint32_t b_addr = ptx::smem_u32(sB + static_cast<size_t>(kb) * BN * 128);
#pragma unroll
                    for (int k = 0; k < kKBlock / 16; ++k) {
                        const uint64_t a_desc = ptx::make_sw128_kmajor_desc(a_addr + k * 32);
                        const uint64_t b_desc = ptx::make_sw128_kmajor_desc(b_addr + k * 32);
                        if (!(SKY_DBG(p) & 2)) ptx::umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);      // stage reusable once these MMAs retire
                    SKY_TRACE(2, it * KB + kb);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full[acc]);            // accumulator ready for the epilogue
            }
        }
    } else if (warp >= kEpiWarp0 && warp < kEpiWarp0 + 4) {
        // ===================== epilogue =====================
        const int e = warp - kEpiWarp0;
        const int quarter = warp & 3;                          // TMEM lanes this warp may access
        const uint32_t hist = smem_addr(sHist + e * 256);
        const uint32_t thrf_addr = smem_addr(sThrF);
        const uint32_t qc_addr = smem_addr(sQc);
        // row norms are prefetched two tiles ahead: a cold load under a saturated HBM stream costs
        // thousands of cycles, which would otherwise be exposed once per tile
        auto load_rn = [&](int it) -> float {
            if (it >= my_tiles) return 0.f;
            const int64_t r = static_cast<int64_t>(tile_of(it)) * kTileRows + quarter * 32 + lane;
            return r < p.rows ? __ldg(p.rownorm + r) : 0.f;
        };
        float rn_cur = load_rn(0), rn_nxt = load_rn(1);
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = tile_of(it);
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int64_t row = static_cast<int64_t>(tile) * kTileRows + quarter * 32 + lane;
            const bool valid = row < p.rows;
            const float rn = rn_cur;
            rn_cur = rn_nxt;
            rn_nxt = load_rn(it + 2);
            const float mx = sqrtf(rn);
            const uint32_t ridx = static_cast<uint32_t>(row);
            if (lane == 0) ptx::mbar_wait_relaxed(&tmem_full[acc], acc_phase, 128);
            __syncwarp();
            if (e == 0 && lane == 0) SKY_TRACE(3, it);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
            if (it == 0 && p.use_gtau && !(SKY_DBG(p) & 1)) {
                // BOOTSTRAP (once per launch): with no bound yet, every row of the first tile would be a
                // candidate for every query.  Instead read the accumulator twice: this first pass only takes
                // the per-query maximum of the tile, the exchange warp trades maxima with the other CTAs,
                // and the normal pass below then runs with a grid-wide bound already in place.
#pragma unroll 1
                for (int c = 0; c < BN / 32; ++c) {
                    uint32_t v[32];
                    ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
                    ptx::tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 32; ++j) {
                        const int q = c * 32 + j;
                        const float dot = __uint_as_float(v[j]);
                        const float qcv = sQc[q];
                        const float sv = largest ? dot / fmaf(qcv, mx, 1e-6f) : (qcv - 2.0f * dot + rn) * p.inv_dd;
                        const uint32_t key = (valid && q < p.nq) ? score_to_key(sv, largest) : 0u;
                        const uint32_t best = __reduce_max_sync(0xffffffffu, key);
                        if (lane == 0 && best) reds_max_u32(sink.lmax + q * 4, best);
                    }
                }
                ptx::named_bar_sync(1, kEpiThreads);
                if (e == 0 && lane == 0) *sBoot = 1;
                if (lane == 0) {
                    const long long t_end = clock64() + 60000;         // ~30 us: never wait for a bound forever
                    while (*sBoot != 2 && clock64() < t_end) __nanosleep(100);
                }
                __syncwarp();
            }
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t v[32];
                SKY_EPI(it, c * 4 + 0);
                ptx::tmem_ld_32x32b_x32(taddr + c * 32, v);
                ptx::tmem_ld_wait();
                SKY_EPI(it, c * 4 + 1);
                if (c == BN / 32 - 1) {                        // last chunk is in registers: free the accumulator
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
                }
                if (SKY_DBG(p) & 1) continue;
                // FAST PATH, branch-free and compact (instruction-cache friendly): 32 conservative
                // pre-filters in the space of the accumulator -- no division, no key -- into one bitmask.
                // Thresholds are read with vector loads and may be slightly stale (they only tighten);
                // padding queries carry a rejecting threshold; NaN always goes on to the exact test
                // (NaN ranks first for cosine).
                uint32_t mbits = 0;
#pragma unroll
                for (int g = 0; g < 8; ++g) {
                    const int q0 = c * 32 + g * 4;
                    float th[4], qc[4];
                    asm volatile("ld.volatile.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(th[0]), "=f"(th[1]), "=f"(th[2]), "=f"(th[3]) : "r"(thrf_addr + q0 * 4) : "memory");
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                                 : "=f"(qc[0]), "=f"(qc[1]), "=f"(qc[2]), "=f"(qc[3]) : "r"(qc_addr + q0 * 4));
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float dot = __uint_as_float(v[g * 4 + u]);
                        bool m;
                        if (largest) {
                            float bound = th[u] * fmaf(qc[u], mx, 1e-6f);
                            bound = fmaf(-fabsf(bound), 1e-6f, bound);
                            m = !(dot < bound);
                        } else {
                            m = !((qc[u] - 2.0f * dot + rn) * p.inv_dd > th[u]);
                        }
                        mbits |= (m ? 1u : 0u) << (g * 4 + u);
                    }
                }
                mbits = valid ? mbits : 0u;
                SKY_EPI(it, c * 4 + 2);
                // EXACT PATH, rare after warm-up: every lane walks its own survivors (the warp runs as many
                // rounds as the busiest lane has bits); inserts are per-lane shared-memory atomics
#pragma unroll 1
                while (mbits) {
                    const int j = __ffs(mbits) - 1;
                    mbits &= mbits - 1;
                    const int q = c * 32 + j;
                    if (q >= p.nq) continue;
                    float dot = 0.f;
#pragma unroll
                    for (int jj = 0; jj < 32; ++jj) dot = (jj == j) ? __uint_as_float(v[jj]) : dot;
                    const float qcv = lds_f32(qc_addr + q * 4);
                    const float sv = largest ? dot / fmaf(qcv, mx, 1e-6f) : (qcv - 2.0f * dot + rn) * p.inv_dd;
                    sink_insert_one(sink, q, make_composite(score_to_key(sv, largest), ridx));
                }
                SKY_EPI(it, c * 4 + 3);
            }
            ptx::named_bar_sync(1, kEpiThreads);
            SKY_EPI(it, 8);
            sink_prune_if_full(sink, p.nq, e, 4, hist);
            SKY_EPI(it, 9);
            ptx::named_bar_sync(1, kEpiThreads);
            SKY_EPI(it, 10);
            if (e == 0 && lane == 0) { *sTilesDone = it + 1; SKY_TRACE(4, it); }
        }
        // final: counts and the last published bound
        ptx::named_bar_sync(1, kEpiThreads);
        for (int q = e * 32 + lane; q < p.nq; q += 128) {
            p.counts[static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q] = static_cast<int>(lds_u32(sink.cnt + q * 4));
            const uint32_t mine = lds_u32(sink.lmax + q * 4);
            if (p.use_gtau && mine) st_cg_u32(p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q, mine);
        }
    } else if (warp == kXchgWarp) {
        // ===================== threshold exchange =====================
        // publish this CTA's best keys, reduce one query column for everybody, apply all bounds;
        // runs free of the epilogue (all state is monotone), fast at first, then at a trickle
        if (p.use_gtau && my_tiles > 0) {
            uint32_t* my_row = p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0;
            const int rq = static_cast<int>(blockIdx.x) % p.nq;
            int round = 0;
            uint32_t last_pub[BN / 32] = {};     // keys already published (only changes are rewritten)
            uint32_t last_lo = 0;                // bound this CTA last contributed for its column
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
                __nanosleep(round < 24 ? 100 : 3000);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == kMmaWarp) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 2 * BN);
    }
}

#ifdef SKY_EXPERIMENTS
int debug_read_trace(unsigned long long* h_out, int n) {
    if (n > 6 * kTraceLen) n = 6 * kTraceLen;
    SKY_CUDA(cudaDeviceSynchronize());
    SKY_CUDA(cudaMemcpyFromSymbol(h_out, g_trace, sizeof(unsigned long long) * n));
    return SKY_OK;
}
int debug_read_epi(unsigned long long* h_out) {
    SKY_CUDA(cudaDeviceSynchronize());
    SKY_CUDA(cudaMemcpyFromSymbol(h_out, g_epi, sizeof(g_epi)));
    return SKY_OK;
}
#endif

// Queries -> bf16 operand matrix [q_pad, Dp] (zero padded) + per-query constants.
__global__ void pack_queries_kernel(const float* __restrict__ t, int Q, int D, int Dp, int q_pad, int metric,
                                    __nv_bfloat16* __restrict__ bq, float* __restrict__ qconst, const StateInit si) {
    ptx::griddep_launch_dependents();      // the scorer behind this kernel may set itself up now (it waits before reading)
    state_init_gridwide(si);
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ double red[8];
    double acc = 0.0;
    for (int d = threadIdx.x; d < Dp; d += blockDim.x) {
        const float v = (q < Q && d < D) ? t[static_cast<size_t>(q) * D + d] : 0.f;
        const __nv_bfloat16 r = __float2bfloat16_rn(v);
        bq[static_cast<size_t>(q) * Dp + d] = r;
        // |t|^2 of the ROUNDED query: the contraction sees bf16(t), so cosine and MSE stay a true cosine / squared
        // distance of that vector (never negative for a near-duplicate row)
        const float vr = __bfloat162float(r);
        acc += static_cast<double>(vr * vr);
    }
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tot = 0.0;
        for (int i = 0; i < nw; ++i) tot += red[i];
        const float tt = static_cast<float>(tot);
        qconst[q] = (metric == SKY_COSINE) ? sqrtf(tt) : tt;
    }
    (void)q_pad;
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// 2-D bf16 [rows, Dp] tensor (row pitch Dp elements), box = 64 columns x box_rows rows, 128-byte swizzle.
int make_tmap_2d(CUtensorMap* m, const void* ptr, int64_t rows, int Dp, int box_rows) {
    EncodeTiledFn fn = get_encode_fn();
    if (!fn) return set_error(SKY_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    cuuint64_t dims[2] = {static_cast<cuuint64_t>(Dp), static_cast<cuuint64_t>(rows)};
    cuuint64_t strides[1] = {static_cast<cuuint64_t>(Dp) * 2};
    cuuint32_t box[2] = {static_cast<cuuint32_t>(kKBlock), static_cast<cuuint32_t>(box_rows)};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return set_error(SKY_ERR_CUDA, "cuTensorMapEncodeTiled failed with %d", static_cast<int>(r));
    return SKY_OK;
}

int launch_pack_queries(const float* t, int Q, int D, int Dp, int q_pad, int metric, void* bq, float* qconst, cudaStream_t st) {
    pack_queries_kernel<<<q_pad, 256, 0, st>>>(t, Q, D, Dp, q_pad, metric, reinterpret_cast<__nv_bfloat16*>(bq), qconst, StateInit{});
    SKY_LAUNCH_CHECK("pack_queries_kernel");
    return SKY_OK;
}

constexpr int kTcBN = 64;
constexpr int kTcUseTma = 1;      // bank stream of K2: 1 = one TMA box per stage (measured 0.2447 ms on C2), 0 = 1024 cp.async per stage (0.2518 ms)

static size_t tc_tail_bytes(int BN) {
    return (2 * kMaxStages + 1 + 4) * sizeof(uint64_t) + 16 + BN * (8 + 4 + 4 + 4 + 4) + 4 * 256 * 4 + 16;
}

static int tc_stages(int Dp, int BN) {
    const size_t budget = 227 * 1024 - 1024 /*alignment slack*/ - tc_tail_bytes(BN);
    const size_t bq = static_cast<size_t>(Dp) * BN * 2;
    if (bq + 2 * kStageBytes > budget) return 0;
    size_t s = (budget - bq) / kStageBytes;
    if (s > kMaxStages) s = kMaxStages;
    return static_cast<int>(s);
}

bool tc_supported(const sky_bank* b, int metric, bool weighted, int n_top) {
    return b->dtype == SKY_BF16 && b->L == 1 && !weighted && n_top == 0 &&
           (metric == SKY_COSINE || metric == SKY_MSE) && tc_stages(b->Dp, kTcBN) >= 2 && b->rows > 0;
}

int tc_grid(const sky_bank* b) {
    const int64_t tiles = (b->rows + kTileRows - 1) / kTileRows;
    return static_cast<int>(tiles < b->num_sms ? tiles : b->num_sms);
}

int tc_make_bank_tmap(sky_bank* b) {
    // tile-major bank: a [rows_pad * KB, 64] tensor whose 128-row boxes are the contiguous (tile, k-block) blocks
    int rc = make_tmap_2d(&b->tmap_bank, b->data, b->rows_pad * (b->Dp / kKBlock), kKBlock, kTileRows);
    if (rc) return rc;
    b->tmap_ready = true;
    return SKY_OK;
}

size_t tc_scratch_bytes(const sky_bank* b, int Q) {
    const int64_t q_pad = round_up(Q, kTcBN);
    return static_cast<size_t>(q_pad) * b->Dp * 2 + static_cast<size_t>(q_pad) * sizeof(float) + 256;
}

// scratch (bank->ws2): [q_pad, Dp] bf16 | [q_pad] f32
int launch_tc_search(sky_bank* b, const float* t, int Q, int metric, const SearchState& s, cudaStream_t st) {
    constexpr int BN = kTcBN;
    const int q_pad = static_cast<int>(round_up(Q, BN));
    __nv_bfloat16* bq = reinterpret_cast<__nv_bfloat16*>(b->ws2);
    float* qconst = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(b->ws2) +
                                             round_up(static_cast<int64_t>(q_pad) * b->Dp * 2, 256));
    // the packing grid also zeroes the exchange state of this search (no separate init_state launch)
    const StateInit si{s.gtop, s.gtau, s.counts, s.Qtot, s.p_stride, s.P};
    pack_queries_kernel<<<q_pad, 256, 0, st>>>(t, Q, b->D, b->Dp, q_pad, metric, bq, qconst, si);
    SKY_LAUNCH_CHECK("pack_queries_kernel");

    int stages = tc_stages(b->Dp, BN);
    { const int e = env_knob("SKY_TC_STAGES", 0); if (e >= 2 && e < stages) stages = e; }
    const size_t smem = 1024 + static_cast<size_t>(b->Dp) * BN * 2 + static_cast<size_t>(stages) * kStageBytes + tc_tail_bytes(BN);
    SKY_CUDA(cudaFuncSetAttribute(tc_search_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    int grid = s.P;
    { const int e = env_knob("SKY_TC_GRID", 0); if (e >= 1 && e < grid) grid = e; }
    for (int q0 = 0; q0 < Q; q0 += BN) {
        CUtensorMap tmq;
        int rc = make_tmap_2d(&tmq, bq + static_cast<size_t>(q0) * b->Dp, BN, b->Dp, BN);
        if (rc) return rc;
        TcParams p;
        p.bank = reinterpret_cast<const unsigned char*>(b->data);
        p.rownorm = b->rownorm;
        p.qconst = qconst + q0;
        p.lists = s.lists; p.counts = s.counts; p.gtop = s.gtop; p.gtau = s.gtau;
        p.p_stride = s.p_stride; p.Qtot = s.Qtot; p.q0 = q0; p.nq = (Q - q0 < BN) ? (Q - q0) : BN;
        p.cap = s.cap; p.k = s.k; p.use_gtau = s.use_gtau;
        p.rows = b->rows;
        p.num_tiles = static_cast<int>((b->rows + kTileRows - 1) / kTileRows);
        p.kblocks = b->Dp / kKBlock;
        p.stages = stages;
        p.metric = metric;
        p.debug = env_knob("SKY_TC_DEBUG", 0);
        { const int pv = env_knob("SKY_TC_POLICY", 0);
          p.bank_policy = pv == 1 ? 0x1000000000000000ull /* evict normal */ : (pv == 2 ? ptx::kEvictLast : ptx::kEvictFirst); }
        p.inv_dd = 1.0f / (static_cast<float>(b->D) * static_cast<float>(b->D));
        prof_mark(b, st);
        p.use_tma = (b->tmap_ready && env_knob("SKY_TC_TMA", kTcUseTma) != 0) ? 1 : 0;
        SKY_CUDA(launch_maybe_pdl(env_knob("SKY_PDL", kUsePdl) != 0, tc_search_kernel<BN>, dim3(grid), dim3(kTcThreads), smem, st,
                                  tmq, b->tmap_bank, p));
        prof_mark(b, st);
        SKY_LAUNCH_CHECK("tc_search_kernel");
    }
    return SKY_OK;
}

}  // namespace sky
