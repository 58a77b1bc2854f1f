// tc_batch.cu -- K2b: batched tcgen05 scorer for LARGE query batches (tensor-pipe bound regime).
//
// Same contraction forms as tc_search.cu (unweighted cosine / MSE, reference
// utils/similarity.py:163-170, :188-192, :246-247), but shaped as a GEMM: an output tile is
// 128 bank rows x 256 queries, both operands stream through a TMA ring (the query matrix no longer
// fits in shared memory), fp32 accumulators live in two 256-column TMEM stages so that the epilogue
// of one tile overlaps the MMAs of the next.  A bank tile is visited by all query groups back to
// back, so it is read from HBM once and from L2 afterwards; the query matrix stays in L2.
//
// Top-k without a score matrix, for thousands of queries and k up to 4096: the bank is walked in
// PHASES of geometrically growing size.  Inside a phase the kernel only filters: a score that
// beats the query's current bound (the exact k-th best of everything scanned in earlier phases)
// is appended to the CTA's candidate list.  Between phases a merge kernel folds the lists into the
// running top-k ("carry") of every query and refreshes the bounds -- the running merge of
// update_best_scores (utils/similarity.py:18-35), done once per phase instead of once per batch.
// The bound is fixed inside a phase, so a phase lets k * (rows of the phase / rows before it) rows per query through
// (at most ~3.2 k with phase sizes 1, 1, 4, 16, ... tiles per CTA), so lists stay short; an adversarially ordered bank
// is still exact: a list that fills up is pruned in place to its k best.  Rows that pass the pre-filter are queued per
// epilogue warp and scored exactly / inserted 32 at a time (see the epilogue).
#include <cstdlib>

#include "bank.cuh"
#include "ptx.cuh"
#include "topk.cuh"

namespace sky {

constexpr int kTbBN = 256;                     // queries per tile (UMMA N)
// Warp 0 TMA, warp 1 MMA, warps 2-9 epilogue.  EIGHT epilogue warps: a warp may only read the tensor-memory lanes of
// its quarter (warp % 4), so two warps share each quarter and split the accumulator's 256 columns.  With four warps the
// filter of a visit took ~4.7 k cycles + ~1 k of write-back against the 6.1 k cycles of its MMAs (%clock timeline of
// CTA 0, tools/trace_tb_phase.py): the kernel was epilogue-bound even without survivors.
constexpr int kTbEpiWarps = 8;
constexpr int kTbThreads = (2 + kTbEpiWarps) * 32;
constexpr int kTbEpiThreads = kTbEpiWarps * 32;
constexpr int kTbQPT = kTbBN / kTbEpiThreads;                 // queries of a group staged per epilogue thread (1)
constexpr int kTbWarpChunks = (kTbBN / 32) / (kTbEpiWarps / 4);   // 32-column chunks per warp and visit (4)
static_assert(kTbQPT * kTbEpiThreads == kTbBN && kTbWarpChunks % 2 == 0 && kTbBN / kTbEpiWarps == 32, "epilogue shape");
constexpr int kTbStageA = kTileRows * 128;     // 16 KB: 128 rows x 64 bf16
constexpr int kTbStageB = kTbBN * 128;         // 32 KB: 256 queries x 64 bf16
constexpr int kTbStage = kTbStageA + kTbStageB;
constexpr int kTbStages = 4;
constexpr int kTbQueue = 192;                  // survivor queue entries per epilogue warp (8 x 1.5 KB of shared memory)

struct TbParams {
    const float* rownorm;      // [rows_pad]
    const float* qconst;       // [Qp] cosine |t|, MSE |t|^2
    const float* bound1;       // [Qp] pre-filter coefficients (see batch_bounds_kernel)
    const float* bound2;       // [Qp]
    const uint64_t* tauc;      // [Qp] exact bound: composite of the current k-th best (0 = none)
    uint64_t* lthr;            // [P][Qp] CTA-local bound after an overflow prune (0 = none)
    uint64_t* lists;           // [P][Qp][cap]
    int* counts;               // [P][Qp]
    int Qp, nq, cap, k, groups, metric, kblocks;
    int64_t rows;              // valid bank rows
    int tile0, tile1;          // tiles [tile0, tile1) of this phase
    int dense;                 // first phase: no bound exists, every row is a candidate -> lists are written densely
    int debug;                 // experiments: bit0 epilogue only drains TMEM, bit1 no MMA issue
    float inv_dd;
};

__device__ __forceinline__ int ld_cg_i32(const int* p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg_i32(int* p, int v) { __stcg(p, v); }
// barrier over `nthreads` threads that also ORs a predicate
__device__ __forceinline__ bool named_bar_or(int id, int nthreads, bool pred) {
    uint32_t r;
    asm volatile(
        "{\n\t.reg .pred p, q;\n\t"
        "setp.ne.u32 q, %3, 0;\n\t"
        "bar.red.or.pred p, %1, %2, q;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(r) : "r"(id), "r"(nthreads), "r"(static_cast<uint32_t>(pred)) : "memory");
    return r != 0;
}

// debug timeline (SKY_TB_DEBUG bit 5 = 32): clock64 stamps of CTA 0, epilogue warp e = 0, per visit of the launch:
// [0] visit start, [1] accumulator ready, [2] chunks done, [3] write-back done, [4] = 1 if anything was inserted
#ifdef SKY_EXPERIMENTS
constexpr int kTbTrace = 4096;
__device__ long long g_tb_trace[kTbTrace * 4 * 5];
#define TB_TRACE(slot, val) do { if ((SKY_DBG(p) & 32) && blockIdx.x == 0 && lane == 0 && v < kTbTrace && e < 4) g_tb_trace[(v * 4 + e) * 5 + (slot)] = (val); } while (0)
#else
#define TB_TRACE(slot, val) do { } while (0)
#endif

// tcgen05.wait::ld that also "touches" the destination registers of an earlier tcgen05.ld, so that the compiler
// cannot move their first use above the wait when other work sits between the load and the wait
__device__ __forceinline__ void tmem_ld_wait_touch(uint32_t (&r)[32]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15]),
                   "+r"(r[16]), "+r"(r[17]), "+r"(r[18]), "+r"(r[19]), "+r"(r[20]), "+r"(r[21]), "+r"(r[22]), "+r"(r[23]),
                   "+r"(r[24]), "+r"(r[25]), "+r"(r[26]), "+r"(r[27]), "+r"(r[28]), "+r"(r[29]), "+r"(r[30]), "+r"(r[31])
                 :: "memory");
}

template <bool COS>
__global__ void __launch_bounds__(kTbThreads, 1)
tc_batch_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b, const TbParams p) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* base = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    unsigned char* sStage = base;                                                   // [stages][A 16 KB | B 32 KB]
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(sStage + kTbStages * kTbStage);   // [stages]
    uint64_t* empty_bar = full_bar + kTbStages;
    uint64_t* tmem_full = empty_bar + kTbStages;                                    // [2]
    uint64_t* tmem_empty = tmem_full + 2;                                           // [2]
    unsigned long long* sTau = reinterpret_cast<unsigned long long*>(tmem_empty + 2);   // [2][256] exact bound (composite)
    float* sB1 = reinterpret_cast<float*>(sTau + 2 * kTbBN);                        // [2][256] pre-filter coefficients
    float* sB2 = sB1 + 2 * kTbBN;                                                   // [2][256]
    float* sQc = sB2 + 2 * kTbBN;                                                   // [2][256] |t| or |t|^2
    int* sCnt = reinterpret_cast<int*>(sQc + 2 * kTbBN);                            // [2][256] list fill of this CTA
    uint32_t* sHist = reinterpret_cast<uint32_t*>(sCnt + 2 * kTbBN);                // [epilogue warps][256]
    uint2* sQueue = reinterpret_cast<uint2*>(sHist + kTbEpiWarps * 256);            // [epilogue warps][kTbQueue] survivors of a visit
    uint32_t* sTmemBase = reinterpret_cast<uint32_t*>(sQueue + kTbEpiWarps * kTbQueue);

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr bool largest = COS;
    const int KB = p.kblocks, G = p.groups;
    const int ntiles = p.tile1 - p.tile0;
    const int my_tiles = (ntiles > static_cast<int>(blockIdx.x))
                             ? (ntiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x) : 0;
    const int visits = my_tiles * G;
    auto tile_of = [&](int it) -> int { return p.tile0 + static_cast<int>(blockIdx.x) + it * static_cast<int>(gridDim.x); };

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_a);
        ptx::prefetch_tmap(&tmap_b);
        for (int s = 0; s < kTbStages; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], kTbEpiWarps); }
        ptx::fence_barrier_init();
    }
    if (warp == 1) {
        ptx::tmem_alloc(sTmemBase, 2 * kTbBN);
        ptx::tmem_relinquish();
    }
    ptx::tc_fence_before();
    __syncthreads();
    ptx::tc_fence_after();
    const uint32_t tmem_base = *sTmemBase;

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int v = 0; v < visits; ++v) {
                const int tile = tile_of(v / G), g = v % G;
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait_relaxed(&empty_bar[stage], phase ^ 1, 32);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], kTbStage);
                    unsigned char* dst = sStage + static_cast<size_t>(stage) * kTbStage;
                    // bank: tile-major, (tile, k-block) = 128 consecutive rows of a [rows_pad * KB, 64] tensor
                    ptx::tma_load_2d(&tmap_a, dst, &full_bar[stage], 0, (tile * KB + kb) * kTileRows, 0x1000000000000000ull);
                    ptx::tma_load_2d(&tmap_b, dst + kTbStageA, &full_bar[stage], kb * kKBlock, g * kTbBN, ptx::kEvictLast);
                    if (++stage == kTbStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer =====================
        // The warp stays converged (every lane polls the barriers and carries the warp-uniform descriptors), one elected
        // lane issues; the descriptors are loop-carried and only incremented (stage = 48 KB = 3072 in the descriptor's
        // address >> 4 field, k-step = 32 B = 2).  Inside an `if (lane == 0)` region ptxas wraps every UTCHMMA in an
        // ELECT / R2UR loop and rebuilds both descriptors with ~50 dependent uniform-datapath instructions: ~120 cycles
        // per MMA on the issuing thread, about as long as the 128 cycles a 128 x 256 x 16 MMA executes -- the tensor
        // pipe then idles ~10 % of the time waiting for its next instruction.
        {
            constexpr uint32_t idesc = ptx::make_idesc_bf16(kTileRows, kTbBN);
            const uint32_t issuer = ptx::elect_one();
            const uint64_t a_desc0 = ptx::make_sw128_kmajor_desc(ptx::smem_u32(sStage));
            uint64_t a_desc = a_desc0;
            int stage = 0;
            uint32_t phase = 0;
            for (int v = 0; v < visits; ++v) {
                const int acc = v & 1;
                const uint32_t acc_phase = (v >> 1) & 1;
                ptx::mbar_wait_relaxed(&tmem_empty[acc], acc_phase ^ 1, 32);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * kTbBN);
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait_relaxed(&full_bar[stage], phase, 20);
                    ptx::tc_fence_after();
                    const uint64_t b_desc = a_desc + (kTbStageA >> 4);
                    const uint32_t acc0 = kb != 0 ? 1u : 0u;
                    if (issuer) {
                        if (!(SKY_DBG(p) & 2)) {
#pragma unroll
                            for (int k = 0; k < kKBlock / 16; ++k) ptx::umma_bf16(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, k ? 1u : acc0);
                        }
                        ptx::umma_commit(&empty_bar[stage]);
                        if (kb == KB - 1) ptx::umma_commit(&tmem_full[acc]);
                    }
                    a_desc += kTbStage >> 4;
                    if (++stage == kTbStages) { stage = 0; phase ^= 1; a_desc = a_desc0; }
                }
            }
        }
    } else {
        // ===================== epilogue: 8 warps, warp & 3 = TMEM lane quarter, e / 4 = column half =====================
        const int e = warp - 2;                   // 0..7: histogram slot, queue, prune ownership
        const int quarter = warp & 3;
        const int c_lo = (e >> 2) * kTbWarpChunks;   // first 32-column chunk of this warp
        const int et = tid - 64;                  // 0..255
        const uint32_t hist = smem_addr(sHist + e * 256);
        uint64_t* my_lists = p.lists + static_cast<size_t>(blockIdx.x) * p.Qp * p.cap;
        int* my_counts = p.counts + static_cast<size_t>(blockIdx.x) * p.Qp;
        uint64_t* my_lthr = p.lthr + static_cast<size_t>(blockIdx.x) * p.Qp;

        // per-query state of a visit's query group is staged in shared memory: pre-filter coefficients, |t|,
        // the exact bound max(global k-th best, CTA-local bound) and the fill of this CTA's list.  It is
        // prefetched into registers one visit ahead (thread et owns query et of the group).
        float nb1[kTbQPT], nb2[kTbQPT], nqc[kTbQPT];
        unsigned long long ntau[kTbQPT];
        int ncnt[kTbQPT];
        auto prefetch_group = [&](int gq) {
#pragma unroll
            for (int j = 0; j < kTbQPT; ++j) {
                const int q = gq * kTbBN + et + j * kTbEpiThreads;
                nb1[j] = __ldg(p.bound1 + q); nb2[j] = __ldg(p.bound2 + q); nqc[j] = __ldg(p.qconst + q);
                const unsigned long long tg = __ldg(reinterpret_cast<const unsigned long long*>(p.tauc + q));
                const unsigned long long tl = ld_cg_u64(my_lthr + q);
                ntau[j] = tl > tg ? tl : tg;
                ncnt[j] = ld_cg_i32(my_counts + q);
            }
        };
        prefetch_group(0);
        uint2* myq = sQueue + e * kTbQueue;
        const uint32_t lt_mask = (1u << lane) - 1u;
        int qn = 0;                               // queue fill (warp-uniform); empty at the end of every visit

        for (int v = 0; v < visits; ++v) {
            const int it = v / G, g = v - it * G;
            const int tile = tile_of(it);
            const int acc = v & 1;
            const uint32_t acc_phase = (v >> 1) & 1;
            const int buf = v & 1;
#pragma unroll
            for (int j = 0; j < kTbQPT; ++j) {
                const int i = buf * kTbBN + et + j * kTbEpiThreads;
                sB1[i] = nb1[j]; sB2[i] = nb2[j]; sQc[i] = nqc[j]; sTau[i] = ntau[j]; sCnt[i] = ncnt[j];
            }
            // the next group is a different set of queries, except when there is a single group: then its
            // counts are re-read after this visit's write-back (see the end of the loop)
            if (G > 1) prefetch_group((g + 1 == G) ? 0 : g + 1);
            TB_TRACE(0, clock64());
            const int64_t row = static_cast<int64_t>(tile) * kTileRows + quarter * 32 + lane;
            const bool valid = row < p.rows;
            const float rn = valid ? __ldg(p.rownorm + row) : 0.f;
            const float mx = sqrtf(rn);
            // per-row term of the pre-filter (see batch_bounds_kernel): cosine bound = b1*|z| + b2, MSE bound = b1 + rn/2 (lowered)
            const float rterm = largest ? mx : 0.5f * rn * (1.0f - 4e-6f);
            ptx::named_bar_sync(1, kTbEpiThreads);           // bounds of this visit are in shared memory
            if (lane == 0) ptx::mbar_wait_relaxed(&tmem_full[acc], acc_phase, 64);
            __syncwarp();
            TB_TRACE(1, clock64());
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * kTbBN);
            bool inserted = false;
            // DRAIN the survivor queue, 32 entries per step, one per lane: exact score, composite, test against the
            // query's bound, append to the CTA's list.  The dependent chain of a survivor (shared loads, division, shared
            // atomic, global store: ~350 cycles) used to run once per survivor on the warp that found it, with nothing
            // to overlap -- 43 ns of epilogue per survivor and CTA, 4.8 of C4's 17.4 ms per GPU (k = 1000 lets ~16 k rows
            // per query through); here it runs for 32 survivors at a time.  Row terms come from the lane that owns the row.
            auto drain = [&]() {
                __syncwarp();
                for (int i0 = 0; i0 < qn; i0 += 32) {
                    const bool has = i0 + lane < qn;
                    const uint2 en = has ? myq[i0 + lane] : make_uint2(0u, static_cast<uint32_t>(lane) << 8);
                    const int sl = static_cast<int>(en.y >> 8) & 31, col = static_cast<int>(en.y & 255u);
                    const float rn_s = __shfl_sync(0xffffffffu, rn, sl);
                    if (has) {
                        const float dot = __uint_as_float(en.x);
                        const int qi = buf * kTbBN + col;
                        const float qcv = sQc[qi];
                        const float sv = largest ? dot / fmaf(qcv, sqrtf(rn_s), 1e-6f) : (qcv - 2.0f * dot + rn_s) * p.inv_dd;
                        const uint32_t row_s = static_cast<uint32_t>(static_cast<int64_t>(tile) * kTileRows + quarter * 32 + sl);
                        const uint64_t comp = make_composite(score_to_key(sv, largest), row_s);
                        if (comp > sTau[qi] && !(SKY_DBG(p) & 16)) {
                            // cap >= k + 256 and at most 128 rows per visit: the list cannot overflow before the check below
                            const uint32_t pos = atoms_add_u32(smem_addr(&sCnt[qi]), 1u);
                            st_cg_u64(my_lists + static_cast<size_t>(g * kTbBN + col) * p.cap + pos, comp);
                            inserted = true;
                        }
                    }
                }
                __syncwarp();
                qn = 0;
            };
            // one 32-column chunk: conservative pre-filter in the space of the accumulator (one FMA / ADD and one
            // compare per score, coefficients read with 128-bit shared loads issued up front), then the rare exact path
            auto process = [&](const uint32_t (&vv)[32], int c) {
                if (p.dense) {
                    // FIRST PHASE: everything is a candidate.  No filter, no atomics: row r of the tile goes to slot r
                    // of the (empty) list of every query, 32 lanes = 256 contiguous bytes per store.
                    const float4* pqc = reinterpret_cast<const float4*>(sQc + buf * kTbBN + c * 32);
                    float4 qv[8];
#pragma unroll
                    for (int i = 0; i < 8; ++i) qv[i] = pqc[i];
                    uint64_t* dst = my_lists + static_cast<size_t>(g * kTbBN + c * 32) * p.cap + quarter * 32 + lane;
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float qc4[4] = {qv[i].x, qv[i].y, qv[i].z, qv[i].w};
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const float dot = __uint_as_float(vv[i * 4 + u]);
                            const float sv = largest ? dot / fmaf(qc4[u], mx, 1e-6f) : (qc4[u] - 2.0f * dot + rn) * p.inv_dd;
                            const uint64_t comp = valid ? make_composite(score_to_key(sv, largest), static_cast<uint32_t>(row)) : 0ull;
                            st_cg_u64(dst + static_cast<size_t>(i * 4 + u) * p.cap, comp);
                        }
                    }
                    return;
                }
                const float4* pb1 = reinterpret_cast<const float4*>(sB1 + buf * kTbBN + c * 32);
                const float4* pb2 = reinterpret_cast<const float4*>(sB2 + buf * kTbBN + c * 32);
                float4 c1[8], c2[8];
#pragma unroll
                for (int i = 0; i < 8; ++i) { c1[i] = pb1[i]; if (COS) c2[i] = pb2[i]; }
                uint32_t mbits = 0;
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    const float b1[4] = {c1[i].x, c1[i].y, c1[i].z, c1[i].w};
                    const float b2[4] = {COS ? c2[i].x : 0.f, COS ? c2[i].y : 0.f, COS ? c2[i].z : 0.f, COS ? c2[i].w : 0.f};
#pragma unroll
                    for (int u = 0; u < 4; ++u) {
                        const float dot = __uint_as_float(vv[i * 4 + u]);
                        const float bound = COS ? fmaf(b1[u], rterm, b2[u]) : b1[u] + rterm;
                        mbits |= (!(dot < bound) ? 1u : 0u) << (i * 4 + u);       // NaN passes
                    }
                }
                if (SKY_DBG(p) & 4) mbits = 0;
                mbits = valid ? mbits : 0u;
                // SURVIVORS go to this warp's queue as (accumulator bits, column, lane); the exact test and the insert
                // happen in drain().  Every lane takes its lowest set bit per round (the warp runs as many rounds as its
                // busiest lane has bits); a round is ballot + select tree + one shared store, nothing to wait for.
                // columns beyond the real queries carry a +inf bound and never pass
#pragma unroll 1
                for (;;) {
                    const uint32_t act = __ballot_sync(0xffffffffu, mbits != 0u);
                    if (act == 0u) break;
                    if (qn > kTbQueue - 32) drain();
                    if (mbits) {
                        const int j = __ffs(mbits) - 1;
                        mbits &= mbits - 1;
                        // vv[j] by a 5-level select tree (depth 5 instead of a 32-deep chain)
                        uint32_t s16[16], s8[8], s4[4], s2[2];
#pragma unroll
                        for (int i = 0; i < 16; ++i) s16[i] = (j & 1) ? vv[2 * i + 1] : vv[2 * i];
#pragma unroll
                        for (int i = 0; i < 8; ++i) s8[i] = (j & 2) ? s16[2 * i + 1] : s16[2 * i];
#pragma unroll
                        for (int i = 0; i < 4; ++i) s4[i] = (j & 4) ? s8[2 * i + 1] : s8[2 * i];
#pragma unroll
                        for (int i = 0; i < 2; ++i) s2[i] = (j & 8) ? s4[2 * i + 1] : s4[2 * i];
                        const uint32_t dotbits = (j & 16) ? s2[1] : s2[0];
                        myq[qn + __popc(act & lt_mask)] = make_uint2(dotbits, static_cast<uint32_t>(c * 32 + j) | (static_cast<uint32_t>(lane) << 8));
                    }
                    qn += __popc(act);
                }
            };
            // software pipeline over this warp's chunks: the tcgen05.ld of chunk c+1 is in flight while chunk c is filtered
            uint32_t va[32], vb[32];
            ptx::tmem_ld_32x32b_x32(taddr + c_lo * 32, va);
#pragma unroll 1
            for (int c2 = 0; c2 < kTbWarpChunks / 2; ++c2) {
                const int c = c_lo + 2 * c2;
                tmem_ld_wait_touch(va);
                ptx::tmem_ld_32x32b_x32(taddr + (c + 1) * 32, vb);
                if (!(SKY_DBG(p) & 1)) process(va, c);
                tmem_ld_wait_touch(vb);
                if (c2 + 1 < kTbWarpChunks / 2) {
                    ptx::tmem_ld_32x32b_x32(taddr + (c + 2) * 32, va);
                } else {                                      // this warp's share of the accumulator is in registers: hand it back
                    ptx::tc_fence_before();
                    __syncwarp();
                    if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);
                }
                if (!(SKY_DBG(p) & 1)) process(vb, c + 1);
            }
            if (!p.dense) drain();
            TB_TRACE(2, clock64());
            TB_TRACE(4, inserted ? 1 : 0);
            // write the fills back; lists of this group that could not take another 128 rows are pruned in place
            // to their k best (rare once bounds exist), which also yields a CTA-local bound
            if (p.dense) {
#pragma unroll
                for (int j = 0; j < kTbQPT; ++j) st_cg_i32(my_counts + g * kTbBN + et + j * kTbEpiThreads, kTileRows);
            } else if (!(SKY_DBG(p) & 8) && named_bar_or(3, kTbEpiThreads, inserted)) {
                // warp e owns queries [32 e, 32 e + 32) of the group, one per lane
                {
                    const int qq = e * 32 + lane;
                    const int n = sCnt[buf * kTbBN + qq];
                    uint32_t full = __ballot_sync(0xffffffffu, n > p.cap - kTileRows);
                    while (full) {                                  // rare: prune one list per iteration, whole warp
                        const int l = __ffs(full) - 1;
                        full &= full - 1;
                        const int qf = e * 32 + l;
                        const int nf = __shfl_sync(0xffffffffu, n, l);
                        uint64_t* lst = my_lists + static_cast<size_t>(g * kTbBN + qf) * p.cap;
                        const uint64_t kth = warp_select_kth(lst, nf, p.k, hist);
                        const int m = warp_compact_ge(lst, nf, kth);
                        if (lane == 0) { st_cg_u64(my_lthr + g * kTbBN + qf, kth); st_cg_i32(my_counts + g * kTbBN + qf, m); }
                        __syncwarp();
                    }
                    if (!(n > p.cap - kTileRows)) st_cg_i32(my_counts + g * kTbBN + qq, n);
                }
                ptx::named_bar_sync(2, kTbEpiThreads);
            }
            if (G == 1) prefetch_group(0);
            TB_TRACE(3, clock64());
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 2 * kTbBN);
    }
}

// ---------------------------------------------------------------------------------------------
// between phases
// ---------------------------------------------------------------------------------------------
// Pre-filter coefficients of every query from its exact bound tauc (composite of the k-th best so far):
//   cosine:  s = dot / (qc |z| + 1e-6) >= tau   <=>  dot >= tau qc |z| + tau 1e-6      -> b1 = tau qc, b2 = tau 1e-6
//   MSE:     s = (qc - 2 dot + rn) / D^2 <= tau <=>  dot >= (qc - tau D^2) / 2 + rn/2  -> b1 = (qc - tau D^2) / 2
// both moved a few 1e-6 relative towards "pass" so that fp32 rounding can never reject what the exact
// composite test would accept.  No bound yet: everything passes.  Padding queries: nothing passes.
__global__ void batch_bounds_kernel(const uint64_t* __restrict__ tauc, const float* __restrict__ qconst, int nq, int Qp,
                                    int metric, float dd, float* __restrict__ b1, float* __restrict__ b2) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Qp) return;
    const bool largest = metric == SKY_COSINE;
    float o1, o2;
    if (q >= nq) { o1 = largest ? 0.f : INFINITY; o2 = INFINITY; }
    else if (tauc[q] == 0ull) { o1 = largest ? 0.f : -INFINITY; o2 = -INFINITY; }
    else {
        const float tau = key_to_score(composite_key(tauc[q]), largest);
        const float qc = qconst[q];
        if (largest) {
            const float f = tau >= 0.f ? (1.0f - 4e-6f) : (1.0f + 4e-6f);
            o1 = tau * qc * f;
            o2 = tau * 1e-6f * f;
        } else {
            const float td = tau * dd;
            o1 = 0.5f * (qc - td) - 2e-6f * (fabsf(qc) + fabsf(td));
            o2 = 0.f;
        }
    }
    b1[q] = o1;
    b2[q] = o2;
}

// carry[q][kpad] (running top-k as composites, 0 = empty) + the P lists of this phase -> new carry, tauc;
// resets the lists.  last != 0: also writes the final (score, index) rows.
constexpr int kMpThreads = 256;
constexpr int kMpSample = 4096;      // first (dense) phase; later phases run with a small pool and sample (more CTAs per SM)
__global__ void __launch_bounds__(kMpThreads)
merge_phase_kernel(uint64_t* __restrict__ carry, uint64_t* __restrict__ tauc, const uint64_t* __restrict__ lists,
                   int* __restrict__ counts, uint64_t* __restrict__ lthr, int P, int Qp, int cap, int k, int kpad, int kMpPool, int nsample, int largest,
                   int last, int nq, int64_t idx_offset, float* __restrict__ out_scores, int64_t* __restrict__ out_idx,
                   const unsigned char* __restrict__ only, int keep_sorted) {
    if (only && !only[blockIdx.x]) return;        // this query was merged by the warp-per-query kernel
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sel = reinterpret_cast<uint64_t*>(smem_raw);    // [kpad]
    uint64_t* pool = sel + kpad;                              // [kMpPool]
    uint64_t* extra = pool + kMpPool;                         // [nsample] survivors of the sampled pre-select
    __shared__ uint32_t hist[256];
    __shared__ uint32_t scratch[4];
    __shared__ uint32_t npool;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kMpThreads / 32;
    if (tid == 0) npool = 0;
    __syncthreads();
    uint64_t* cq = carry + static_cast<size_t>(q) * kpad;
    // list entries were admitted with comp > tauc[q]: every one of them counts
    for (int j = tid; j < k; j += kMpThreads) {
        const uint64_t v = cq[j];
        if (v) { const uint32_t pos = atomicAdd(&npool, 1u); if (pos < kMpPool) pool[pos] = v; }
    }
    // all list fills first (one L2 round trip instead of one per list), an exclusive scan gives every list its
    // place in the pool, then the gather runs with independent loads (no atomics on the way)
    __shared__ int s_n[1024];
    __shared__ int s_off[1024];
    for (int p = tid; p < P; p += kMpThreads) s_n[p] = counts[static_cast<size_t>(p) * Qp + q];
    __syncthreads();
    if (warp == 0) {
        int run = static_cast<int>(npool);           // carry entries already in the pool
        for (int p0 = 0; p0 < P; p0 += 32) {
            const int n = (p0 + lane < P) ? s_n[p0 + lane] : 0;
            int incl = n;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
            if (p0 + lane < P) s_off[p0 + lane] = run + incl - n;
            run += __shfl_sync(0xffffffffu, incl, 31);
        }
        if (lane == 0) npool = static_cast<uint32_t>(run);
    }
    __syncthreads();
    const int np = static_cast<int>(npool);          // all candidates: carry + every list
    // gather: everything if it fits the pool, else the first kMpPool candidates (a sample).  Four independent loads
    // per lane and step: a dependent one-load-per-lane loop leaves ~6 KB in flight per SM and the merge of the dense first
    // phase (19 k candidates per query, 620 MB over C3's queries) crawled at 0.7 TB/s.
    if (P <= kMpThreads && np <= 32 * P) {
        // later phases: a handful of survivors per list.  One THREAD per list, four loads in flight each: ~n / 4 L2 round
        // trips for the whole query instead of one per list and warp (P / 8 of them, one after the other)
        if (tid < P) {
            const int n = s_n[tid], o = s_off[tid];
            const uint64_t* e = lists + (static_cast<size_t>(tid) * Qp + q) * cap;
            for (int i0 = 0; i0 < n && o + i0 < kMpPool; i0 += 4) {
                uint64_t v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) v[u] = (i0 + u < n && o + i0 + u < kMpPool) ? ld_cg_u64(e + i0 + u) : 0ull;
#pragma unroll
                for (int u = 0; u < 4; ++u) if (i0 + u < n && o + i0 + u < kMpPool) pool[o + i0 + u] = v[u];
            }
        }
    } else {
        for (int p = warp; p < P; p += nwarps) {
            const int n = s_n[p], o = s_off[p];
            if (o >= kMpPool) continue;
            const uint64_t* e = lists + (static_cast<size_t>(p) * Qp + q) * cap;
            for (int i0 = 0; i0 < n && o + i0 < kMpPool; i0 += 128) {
                uint64_t v[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 32 + lane;
                    v[u] = (i < n && o + i < kMpPool) ? ld_cg_u64(e + i) : 0ull;
                }
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int i = i0 + u * 32 + lane;
                    if (i < n && o + i < kMpPool) pool[o + i] = v[u];
                }
            }
        }
    }
    __syncthreads();
    bool done = false;
    // the carry between two phases only has to be the right SET (this kernel gathers it into the pool again); sorted
    // order is needed by the last phase (the result) and by the warp-per-query merge (binary searches over the carry)
    const bool sorted = last || keep_sorted;
    uint64_t kth = 0;
    bool have_kth = false;
    if (np <= kRankSortMax) {
        block_rank_topk(pool, np, k, kpad, sel);
        done = true;
    } else if (np <= kMpPool) {
        kth = block_select_sort([&](int j) { return pool[j]; }, np, k, kpad, sel, hist, scratch, sorted);
        have_kth = true;
        done = true;
    } else {
        // big candidate set (first phase: every row of every CTA): the k-th best of the sample is a lower bound of
        // the real one -- stream all candidates once more, keep what beats it, finish on the few survivors
        block_select_sort([&](int j) { return pool[j]; }, kMpPool, k, kpad, sel, hist, scratch);
        __syncthreads();
        const uint64_t tau0 = sel[k - 1];
        __shared__ uint32_t nkeep;
        if (tid == 0) nkeep = 0;
        __syncthreads();
        if (tau0 != 0) {
            for (int j = tid; j < k; j += kMpThreads) {
                const uint64_t v = cq[j];
                if (v >= tau0) { const uint32_t pos = atomicAdd(&nkeep, 1u); if (pos < static_cast<uint32_t>(nsample)) extra[pos] = v; }
            }
            // two lists per step, four entries per lane and list: eight independent loads in flight per lane
            for (int p0 = warp; p0 < P; p0 += 2 * nwarps) {
                const int pb[2] = {p0, p0 + nwarps};
                const int nb[2] = {s_n[p0], p0 + nwarps < P ? s_n[p0 + nwarps] : 0};
                const uint64_t* eb[2] = {lists + (static_cast<size_t>(pb[0]) * Qp + q) * cap,
                                         lists + (static_cast<size_t>(pb[1] < P ? pb[1] : pb[0]) * Qp + q) * cap};
                const int nmax = nb[0] > nb[1] ? nb[0] : nb[1];
                for (int i0 = 0; i0 < nmax; i0 += 128) {
                    uint64_t v[2][4];
#pragma unroll
                    for (int l = 0; l < 2; ++l)
#pragma unroll
                        for (int u = 0; u < 4; ++u) {
                            const int i = i0 + u * 32 + lane;
                            v[l][u] = i < nb[l] ? ld_cg_u64(eb[l] + i) : 0ull;
                        }
#pragma unroll
                    for (int l = 0; l < 2; ++l)
#pragma unroll
                        for (int u = 0; u < 4; ++u)
                            if (v[l][u] >= tau0) {      // tau0 != 0: empty entries (0) never pass
                                const uint32_t pos = atomicAdd(&nkeep, 1u);
                                if (pos < static_cast<uint32_t>(nsample)) extra[pos] = v[l][u];
                            }
                }
            }
            __syncthreads();
            if (nkeep <= kRankSortMax) {
                block_rank_topk(extra, static_cast<int>(nkeep), k, kpad, sel);
                done = true;
            } else if (nkeep <= static_cast<uint32_t>(nsample)) {
                block_select_sort([&](int j) { return extra[j]; }, static_cast<int>(nkeep), k, kpad, sel, hist, scratch);
                done = true;
            }
        }
        __syncthreads();
    }
    if (!done) {
        // adversarial order / huge k: more candidates than the pool holds -> select straight from L2
        auto fetch = [&](int j) -> uint64_t {
            if (j < k) return cq[j];
            const int jj = j - k;
            const int p = jj / cap, i = jj - p * cap;
            if (i >= counts[static_cast<size_t>(p) * Qp + q]) return 0ull;
            return lists[(static_cast<size_t>(p) * Qp + q) * cap + i];
        };
        block_select_sort(fetch, k + P * cap, k, kpad, sel, hist, scratch);
    }
    __syncthreads();
    for (int j = tid; j < kpad; j += kMpThreads) cq[j] = (j < k) ? sel[j] : 0ull;
    if (tid == 0) tauc[q] = have_kth ? kth : sel[k - 1];        // 0 while fewer than k rows have been seen
    for (int p = tid; p < P; p += kMpThreads) { counts[static_cast<size_t>(p) * Qp + q] = 0; lthr[static_cast<size_t>(p) * Qp + q] = 0ull; }
    if (last && q < nq) {
        for (int j = tid; j < k; j += kMpThreads) {
            const uint64_t c = sel[j];
            float* so = out_scores + static_cast<size_t>(q) * k + j;
            int64_t* io = out_idx + static_cast<size_t>(q) * k + j;
            if (c == 0) { *so = largest ? -INFINITY : INFINITY; *io = -1; }
            else { *so = key_to_score(composite_key(c), largest != 0); *io = static_cast<int64_t>(composite_idx(c)) + idx_offset; }
        }
    }
}

// Warp-per-query merge of one LATER phase (k <= 128): the carry is sorted, the phase's survivors are few (~k ln 5 per
// query), so one warp ranks them with binary searches instead of a 256-thread CTA paying six dependent L2 round trips
// per query -- four queries per CTA.  A query with more survivors than a warp holds is flagged
// in `slow` and left (untouched) to merge_phase_kernel.
constexpr int kMwWarps = 4;
constexpr int kMwMaxM = 512;         // a phase lets k * (rows of the phase / rows before it) <= ~3.2 k survivors per query through: 320 at k = 100
constexpr int kMwMaxK = 128;
constexpr int kMwMaxPL = 8;          // lists per lane: P <= 256
__global__ void __launch_bounds__(kMwWarps * 32)
merge_phase_warp_kernel(uint64_t* __restrict__ carry, uint64_t* __restrict__ tauc, const uint64_t* __restrict__ lists,
                        int* __restrict__ counts, uint64_t* __restrict__ lthr, unsigned char* __restrict__ slow, int P, int Qp, int cap,
                        int k, int kpad, int largest, int last, int nq, int64_t idx_offset, float* __restrict__ out_scores,
                        int64_t* __restrict__ out_idx) {
    __shared__ uint64_t sS[kMwWarps][kMwMaxM];      // survivors as gathered
    __shared__ uint64_t sT[kMwWarps][kMwMaxM];      // survivors, best first
    __shared__ uint64_t sC[kMwWarps][kMwMaxK];      // carry of the earlier phases (best first, zeros behind)
    __shared__ uint64_t sN[kMwWarps][kMwMaxK];      // new carry
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int q = blockIdx.x * kMwWarps + w;
    if (q >= Qp) return;
    // fills of this query's P lists: lane owns lists lane, lane + 32, ...
    int cnt[kMwMaxPL];
    int mine = 0;
#pragma unroll
    for (int i = 0; i < kMwMaxPL; ++i) {
        const int p = lane + 32 * i;
        cnt[i] = p < P ? ld_cg_i32(counts + static_cast<size_t>(p) * Qp + q) : 0;
        mine += cnt[i];
    }
    int incl = mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) { const int o = __shfl_up_sync(0xffffffffu, incl, off); if (lane >= off) incl += o; }
    const int m = __shfl_sync(0xffffffffu, incl, 31);
    if (m > kMwMaxM) {
        if (lane == 0) slow[q] = 1;
        return;
    }
    if (lane == 0) slow[q] = 0;
    uint64_t* cq = carry + static_cast<size_t>(q) * kpad;
    for (int j = lane; j < kpad; j += 32) { sC[w][j] = cq[j]; sN[w][j] = 0ull; }
    int off0 = incl - mine;
    // a list holds one or two survivors as a rule: their loads are issued for all lists of the lane before the first
    // use (a load-then-store loop per list would pay one L2 round trip per list, one after the other)
    uint64_t first[kMwMaxPL][2];
#pragma unroll
    for (int i = 0; i < kMwMaxPL; ++i) {
        const int p = lane + 32 * i;
        const uint64_t* e = lists + (static_cast<size_t>(p < P ? p : 0) * Qp + q) * cap;
        first[i][0] = cnt[i] > 0 ? ld_cg_u64(e) : 0ull;
        first[i][1] = cnt[i] > 1 ? ld_cg_u64(e + 1) : 0ull;
    }
#pragma unroll
    for (int i = 0; i < kMwMaxPL; ++i) {
        const int p = lane + 32 * i;
        if (p < P) {
            const uint64_t* e = lists + (static_cast<size_t>(p) * Qp + q) * cap;
            if (cnt[i] > 0) sS[w][off0] = first[i][0];
            if (cnt[i] > 1) sS[w][off0 + 1] = first[i][1];
            for (int j = 2; j < cnt[i]; ++j) sS[w][off0 + j] = ld_cg_u64(e + j);
            off0 += cnt[i];
        }
    }
    __syncwarp();
    // carry entries in use (sorted, zeros behind)
    int c = 0;
    for (int j0 = 0; j0 < kpad; j0 += 32) c += __popc(__ballot_sync(0xffffffffu, sC[w][j0 + lane] != 0ull));
    // survivors best first: rank among themselves (composites are distinct)
    for (int i = lane; i < m; i += 32) {
        const uint64_t v = sS[w][i];
        int r = 0;
        for (int j = 0; j < m; ++j) r += (sS[w][j] > v) ? 1 : 0;
        sT[w][r] = v;
    }
    __syncwarp();
    // merged rank = own position + number of elements of the other sorted run that beat the element
    for (int i = lane; i < m; i += 32) {
        const uint64_t v = sT[w][i];
        int lo = 0, hi = c;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sC[w][mid] > v) lo = mid + 1; else hi = mid; }
        if (i + lo < k) sN[w][i + lo] = v;
    }
    for (int j = lane; j < c; j += 32) {
        const uint64_t v = sC[w][j];
        int lo = 0, hi = m;
        while (lo < hi) { const int mid = (lo + hi) >> 1; if (sT[w][mid] > v) lo = mid + 1; else hi = mid; }
        if (j + lo < k) sN[w][j + lo] = v;
    }
    __syncwarp();
    for (int j = lane; j < kpad; j += 32) cq[j] = (j < k) ? sN[w][j] : 0ull;
    if (lane == 0) tauc[q] = sN[w][k - 1];        // 0 while fewer than k rows have been seen
#pragma unroll
    for (int i = 0; i < kMwMaxPL; ++i) {
        const int p = lane + 32 * i;
        if (p < P) { counts[static_cast<size_t>(p) * Qp + q] = 0; lthr[static_cast<size_t>(p) * Qp + q] = 0ull; }
    }
    if (last && q < nq) {
        for (int j = lane; j < k; j += 32) {
            const uint64_t cv = sN[w][j];
            float* so = out_scores + static_cast<size_t>(q) * k + j;
            int64_t* io = out_idx + static_cast<size_t>(q) * k + j;
            if (cv == 0) { *so = largest ? -INFINITY : INFINITY; *io = -1; }
            else { *so = key_to_score(composite_key(cv), largest != 0); *io = static_cast<int64_t>(composite_idx(cv)) + idx_offset; }
        }
    }
}

__global__ void batch_init_kernel(uint64_t* carry, uint64_t* tauc, int* counts, uint64_t* lthr, size_t n_carry, size_t n_q, size_t n_pq) {
    const size_t i = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i < n_carry) carry[i] = 0ull;
    if (i < n_q) tauc[i] = 0ull;
    if (i < n_pq) { counts[i] = 0; lthr[i] = 0ull; }
}

// ---------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------
int make_tmap_2d(CUtensorMap* m, const void* ptr, int64_t rows, int Dp, int box_rows);                 // tc_search.cu
int launch_pack_queries(const float* t, int Q, int D, int Dp, int q_pad, int metric, void* bq, float* qconst, cudaStream_t st);

#ifdef SKY_EXPERIMENTS
int debug_read_tb_trace(long long* h_out, int n) {
    if (n > kTbTrace * 20) n = kTbTrace * 20;
    SKY_CUDA(cudaDeviceSynchronize());
    SKY_CUDA(cudaMemcpyFromSymbol(h_out, g_tb_trace, sizeof(long long) * n));
    return SKY_OK;
}
#endif

static int tb_next_pow2(int v) { int p = 32; while (p < v) p <<= 1; return p; }

bool tc_batch_supported(const sky_bank* b, int metric, bool weighted, int n_top, int k) {
    return b->dtype == SKY_BF16 && b->L == 1 && !weighted && n_top == 0 && (metric == SKY_COSINE || metric == SKY_MSE) &&
           b->rows > 0 && k <= 4096 && b->tmap_ready;
}

int launch_tc_batch(sky_bank* b, const float* t, int Q, int metric, int k, int64_t idx_offset, float* out_scores,
                    int64_t* out_idx, cudaStream_t st) {
    const int Qp = static_cast<int>(round_up(Q, kTbBN));
    const int G = Qp / kTbBN;
    const int KB = b->Dp / kKBlock;
    const int num_tiles = static_cast<int>((b->rows + kTileRows - 1) / kTileRows);
    const int P = num_tiles < b->num_sms ? num_tiles : b->num_sms;
    const int kpad = tb_next_pow2(k);
    const int cap = k + 2 * kTileRows;
    // workspace: lists | carry | tauc | lthr | counts | b1 | b2 | qconst | bq
    size_t off = 0;
    auto take = [&](size_t bytes) { const size_t o = off; off += static_cast<size_t>(round_up(static_cast<int64_t>(bytes), 256)); return o; };
    const size_t o_lists = take(static_cast<size_t>(P) * Qp * cap * 8);
    const size_t o_carry = take(static_cast<size_t>(Qp) * kpad * 8);
    const size_t o_tauc = take(static_cast<size_t>(Qp) * 8);
    const size_t o_lthr = take(static_cast<size_t>(P) * Qp * 8);
    const size_t o_counts = take(static_cast<size_t>(P) * Qp * 4);
    const size_t o_b1 = take(static_cast<size_t>(Qp) * 4);
    const size_t o_b2 = take(static_cast<size_t>(Qp) * 4);
    const size_t o_qc = take(static_cast<size_t>(Qp) * 4);
    const size_t o_bq = take(static_cast<size_t>(Qp) * b->Dp * 2);
    const size_t o_slow = take(static_cast<size_t>(Qp));
    int rc = ensure_ws(b, off);
    if (rc) return rc;
    unsigned char* ws = reinterpret_cast<unsigned char*>(b->ws);
    uint64_t* lists = reinterpret_cast<uint64_t*>(ws + o_lists);
    uint64_t* carry = reinterpret_cast<uint64_t*>(ws + o_carry);
    uint64_t* tauc = reinterpret_cast<uint64_t*>(ws + o_tauc);
    uint64_t* lthr = reinterpret_cast<uint64_t*>(ws + o_lthr);
    int* counts = reinterpret_cast<int*>(ws + o_counts);
    float* b1 = reinterpret_cast<float*>(ws + o_b1);
    float* b2 = reinterpret_cast<float*>(ws + o_b2);
    float* qconst = reinterpret_cast<float*>(ws + o_qc);
    void* bq = ws + o_bq;
    unsigned char* slow = ws + o_slow;

    rc = launch_pack_queries(t, Q, b->D, b->Dp, Qp, metric, bq, qconst, st);
    if (rc) return rc;
    {
        const size_t n = static_cast<size_t>(P) * Qp > static_cast<size_t>(Qp) * kpad ? static_cast<size_t>(P) * Qp : static_cast<size_t>(Qp) * kpad;
        batch_init_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, st>>>(carry, tauc, counts, lthr, static_cast<size_t>(Qp) * kpad,
                                                                               static_cast<size_t>(Qp), static_cast<size_t>(P) * Qp);
        SKY_LAUNCH_CHECK("batch_init_kernel");
    }
    CUtensorMap tmq;
    rc = make_tmap_2d(&tmq, bq, Qp, b->Dp, kTbBN);
    if (rc) return rc;

    const size_t smem = 1024 + static_cast<size_t>(kTbStages) * kTbStage + (2 * kTbStages + 4) * 8 + 2 * kTbBN * 8 + 8 * kTbBN * 4 + kTbEpiWarps * 256 * 4 + kTbEpiWarps * kTbQueue * 8 + 64;
    SKY_CUDA(cudaFuncSetAttribute(tc_batch_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    SKY_CUDA(cudaFuncSetAttribute(tc_batch_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    SKY_CUDA(cudaFuncSetAttribute(merge_phase_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));

    TbParams p;
    p.rownorm = b->rownorm; p.qconst = qconst; p.bound1 = b1; p.bound2 = b2; p.tauc = tauc; p.lthr = lthr;
    p.lists = lists; p.counts = counts;
    p.Qp = Qp; p.nq = Q; p.cap = cap; p.k = k; p.groups = G; p.metric = metric; p.kblocks = KB;
    p.rows = b->rows;
    p.debug = env_knob("SKY_TB_DEBUG", 0);
    p.inv_dd = 1.0f / (static_cast<float>(b->D) * static_cast<float>(b->D));
    const float dd = static_cast<float>(b->D) * static_cast<float>(b->D);

    // phases: 1, 1, 2, 4, ... tiles per CTA
    int t0 = 0;
    int per_cta = 1;
    bool first = true;
    int phase_idx = 0;
    int growth = 4;       // phase sizes 1, 1, 4, 16, ... tiles per CTA (measured: growth 16 saves two merges on a small shard but its
                          // k ln 17 survivors per query overflow lists and pools at k = 1000: C4's share 20 -> 29 ms)
    { const int e = env_knob("SKY_TB_PHASE0", 0); if (e >= 1) per_cta = e; }
    { const int e = env_knob("SKY_TB_GROWTH", 0); if (e >= 2) growth = e; }
    // Dense first phase = one tile per CTA.  Measured on C3's 8-GPU shard (1.25 M rows, Q = 4096, k = 100): cutting it to
    // 31 tiles (so that the first merge reads 4 k instead of 19 k candidates per query) LOSES 0.7 ms -- the weaker first
    // bound lets 4.8 k instead of 1 k rows per query through the next phase, and survivors are what costs (see the queue).
    int dense_tiles = P;
    { const int e = env_knob("SKY_TB_DENSE0", 0); if (e >= 1) dense_tiles = e < P ? e : P; }
    while (t0 < num_tiles) {
        int t1 = t0 + ((first && per_cta == 1) ? dense_tiles : per_cta * P);
        if (t1 > num_tiles || num_tiles - t1 < P) t1 = num_tiles;      // fold a short tail into this phase
        batch_bounds_kernel<<<(Qp + 255) / 256, 256, 0, st>>>(tauc, qconst, Q, Qp, metric, dd, b1, b2);
        SKY_LAUNCH_CHECK("batch_bounds_kernel");
        p.tile0 = t0; p.tile1 = t1;
        { const int tp = env_knob("SKY_TB_TRACE_PHASE", -1);      // experiments: timeline (debug bit 5) of one phase only
          if (tp >= 0) p.debug = (env_knob("SKY_TB_DEBUG", 0) & ~32) | (phase_idx == tp ? 32 : 0); }
        ++phase_idx;
        // dense first phase: one tile per CTA into empty lists (a longer first phase keeps the filtered path)
        p.dense = (first && per_cta == 1 && t1 - t0 <= P) ? 1 : 0;
        const int grid = (t1 - t0) < P ? (t1 - t0) : P;
        prof_mark(b, st);
        if (metric == SKY_COSINE) tc_batch_kernel<true><<<grid, kTbThreads, smem, st>>>(b->tmap_bank, tmq, p);
        else tc_batch_kernel<false><<<grid, kTbThreads, smem, st>>>(b->tmap_bank, tmq, p);
        prof_mark(b, st);
        SKY_LAUNCH_CHECK("tc_batch_kernel");
        const int last = t1 == num_tiles ? 1 : 0;
        // sample pool of the merge: large enough that the k-th best of the sample keeps fewer than kMpSample of
        // all candidates (first phase: every row of the phase's tiles is one)
        // later phases see k carried entries plus ~k ln 5 survivors: a pool of max(2048, 4 kpad) holds them (a phase
        // that overflows it falls back to the exact select over L2), and the small footprint lets 8 merge CTAs share an
        // SM instead of 3 -- on small shards the per-phase merges are a fifth of the search
        int pool = 4 * kpad > 2048 ? 4 * kpad : 2048, nsample = 2 * kpad > 1024 ? 2 * kpad : 1024;
        if (first) {
            const int64_t np_max = static_cast<int64_t>(grid) * kTileRows + k;
            // the k-th best of a sample of `pool` candidates keeps ~k np_max / pool of all of them: 1.5 x headroom under
            // kMpSample (the count scatters by ~1 / sqrt(k)); a smaller pool lets two merge CTAs share an SM at k = 1000
            const int64_t want = round_up(3 * np_max * k / (2 * kMpSample), 1024);
            pool = static_cast<int>(want < 4096 ? 4096 : (want > 16384 ? 16384 : want));
            nsample = kMpSample;
        }
        const size_t msmem = static_cast<size_t>(kpad + pool + nsample) * 8;
        const bool warp_merge = !first && kpad <= kMwMaxK && P <= 32 * kMwMaxPL;
        if (warp_merge) {
            merge_phase_warp_kernel<<<(Qp + kMwWarps - 1) / kMwWarps, kMwWarps * 32, 0, st>>>(
                carry, tauc, lists, counts, lthr, slow, P, Qp, cap, k, kpad, metric_largest(metric) ? 1 : 0, last, Q, idx_offset,
                out_scores, out_idx);
            SKY_LAUNCH_CHECK("merge_phase_warp_kernel");
        }
        merge_phase_kernel<<<Qp, kMpThreads, msmem, st>>>(carry, tauc, lists, counts, lthr, P, Qp, cap, k, kpad, pool, nsample,
                                                         metric_largest(metric) ? 1 : 0, last, Q, idx_offset, out_scores, out_idx,
                                                         warp_merge ? slow : nullptr, (kpad <= kMwMaxK || env_knob("SKY_TB_SORTED", 0)) ? 1 : 0);
        SKY_LAUNCH_CHECK("merge_phase_kernel");
        t0 = t1;
        if (!first) per_cta *= growth;
        first = false;
    }
    return SKY_OK;
}

}  // namespace sky
