// tc_search.cu -- K2: tcgen05 tensor-core scorer with fused top-k (bf16 bank, sm_100a).
//
// The contraction form of the reference metrics for queries without
// per-feature weights (use_weights=False, reference utils/similarity.py:246-247):
//   cosine :163-170   s = t.z / (|t| |z| + 1e-6)            |z|^2 = row norm stored with the bank
//   MSE    :188-192   s = (|t|^2 - 2 t.z + |z|^2) / D^2     (rank-equivalent to L2)
// One persistent CTA per SM.  Warp roles:
//   warp 0      TMA producer: the query matrix B[BN, Dp] once (resident in shared memory for the
//               whole kernel), then 128-row x 64-col bank boxes through a ring of stages
//   warp 1      MMA issuer (one thread): tcgen05.mma 128 x BN x 16, fp32 accumulators in TMEM,
//               two accumulator stages so the epilogue of tile i overlaps the MMAs of tile i+1
//   warps 2-5   epilogue: tcgen05.ld accumulators -> score -> threshold filter -> candidate sink
//   warp 6      grid-wide threshold exchange (publishes / refreshes the top-k lower bound)
// The bank streams from HBM exactly once per launch; no score ever goes to HBM.
#include "bank.cuh"
#include "ptx.cuh"
#include "topk.cuh"

namespace sky {

constexpr int kTcThreads = 7 * 32;
constexpr int kEpiWarp0 = 2;
constexpr int kEpiThreads = 128;
constexpr int kStageBytes = kTileRows * 128;   // 128 rows x 64 bf16
constexpr int kMaxStages = 8;

struct TcParams {
    const float* rownorm;   // [rows_pad]
    const float* qconst;    // [BN] cosine: |t| ; MSE: |t|^2
    uint64_t* lists; int* counts; uint32_t* gtop; uint32_t* gtau;
    int p_stride, Qtot, q0, nq, cap, k, use_gtau;
    int64_t rows;           // valid bank rows
    int num_tiles, kblocks, stages, metric;
    float inv_dd;           // 1 / D^2
};

template <int BN>
__global__ void __launch_bounds__(kTcThreads, 1)
tc_search_kernel(const __grid_constant__ CUtensorMap tmap_bank, const __grid_constant__ CUtensorMap tmap_q,
                 const TcParams p) {
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
    unsigned long long* sThr = reinterpret_cast<unsigned long long*>(tmem_empty + 2);   // [BN]
    float* sQc = reinterpret_cast<float*>(sThr + BN);           // [BN]
    float* sThrF = sQc + BN;                                    // [BN] threshold as a score (pre-filter)
    int* sCnt = reinterpret_cast<int*>(sThrF + BN);             // [BN]
    uint32_t* sLmax = reinterpret_cast<uint32_t*>(sCnt + BN);   // [BN]
    uint32_t* sHist = sLmax + BN;                               // [4][256]
    uint32_t* sTmemBase = sHist + 4 * 256;                      // [1]
    volatile int* sTilesDone = reinterpret_cast<volatile int*>(sTmemBase + 1);   // [1]

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const bool largest = (p.metric == SKY_COSINE);

    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const int my_tiles = (p.num_tiles > static_cast<int>(blockIdx.x))
                             ? (p.num_tiles - static_cast<int>(blockIdx.x) + static_cast<int>(gridDim.x) - 1) / static_cast<int>(gridDim.x)
                             : 0;

    if (warp == 0 && lane == 0) {
        ptx::prefetch_tmap(&tmap_bank);
        ptx::prefetch_tmap(&tmap_q);
        for (int s = 0; s < S; ++s) { ptx::mbar_init(&full_bar[s], 1); ptx::mbar_init(&empty_bar[s], 1); }
        ptx::mbar_init(b_full, 1);
        for (int a = 0; a < 2; ++a) { ptx::mbar_init(&tmem_full[a], 1); ptx::mbar_init(&tmem_empty[a], 4); }
        ptx::fence_barrier_init();
        *sTilesDone = 0;
    }
    if (warp == 1) {
        ptx::tmem_alloc(sTmemBase, 2 * BN);
        ptx::tmem_relinquish();
    }
    for (int q = tid; q < BN; q += kTcThreads) {
        sThr[q] = (q < p.nq) ? 0ull : ~0ull;
        sThrF[q] = __uint_as_float(0x7FC00000u);   // NaN: the pre-filter lets everything through
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

    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            ptx::mbar_arrive_expect_tx(b_full, static_cast<uint32_t>(KB) * BN * 128);
            for (int kb = 0; kb < KB; ++kb)
                ptx::tma_load_2d(&tmap_q, sB + static_cast<size_t>(kb) * BN * 128, b_full, kb * kKBlock, 0, ptx::kEvictLast);
            int stage = 0;
            uint32_t phase = 0;
            for (int it = 0; it < my_tiles; ++it) {
                const int tile = blockIdx.x + it * gridDim.x;
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(&empty_bar[stage], phase ^ 1);
                    ptx::mbar_arrive_expect_tx(&full_bar[stage], kStageBytes);
                    ptx::tma_load_2d(&tmap_bank, sA + static_cast<size_t>(stage) * kStageBytes, &full_bar[stage],
                                     kb * kKBlock, tile * kTileRows, ptx::kEvictFirst);
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
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
                ptx::mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
                ptx::tc_fence_after();
                const uint32_t d_tmem = tmem_base + static_cast<uint32_t>(acc * BN);
                for (int kb = 0; kb < KB; ++kb) {
                    ptx::mbar_wait(&full_bar[stage], phase);
                    ptx::tc_fence_after();
                    const uint32_t a_addr = ptx::smem_u32(sA + static_cast<size_t>(stage) * kStageBytes);
                    const uint32_t b_addr = ptx::smem_u32(sB + static_cast<size_t>(kb) * BN * 128);
#pragma unroll
                    for (int k = 0; k < kKBlock / 16; ++k) {
                        const uint64_t a_desc = ptx::make_sw128_kmajor_desc(a_addr + k * 32);
                        const uint64_t b_desc = ptx::make_sw128_kmajor_desc(b_addr + k * 32);
                        ptx::umma_bf16(d_tmem, a_desc, b_desc, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    ptx::umma_commit(&empty_bar[stage]);      // stage reusable once these MMAs retire
                    if (++stage == S) { stage = 0; phase ^= 1; }
                }
                ptx::umma_commit(&tmem_full[acc]);            // accumulator ready for the epilogue
            }
        }
    } else if (warp < kEpiWarp0 + 4) {
        // ===================== epilogue =====================
        const int e = warp - kEpiWarp0;
        const int quarter = warp & 3;                          // TMEM lanes this warp may access
        const uint32_t hist = smem_addr(sHist + e * 256);
        for (int it = 0; it < my_tiles; ++it) {
            const int tile = blockIdx.x + it * gridDim.x;
            const int acc = it & 1;
            const uint32_t acc_phase = (it >> 1) & 1;
            const int64_t row = static_cast<int64_t>(tile) * kTileRows + quarter * 32 + lane;
            const bool valid = row < p.rows;
            const float rn = valid ? __ldg(p.rownorm + row) : 0.f;
            ptx::mbar_wait(&tmem_full[acc], acc_phase);
            ptx::tc_fence_after();
            const uint32_t taddr = tmem_base + (static_cast<uint32_t>(quarter * 32) << 16) + static_cast<uint32_t>(acc * BN);
            uint32_t v[BN];
#pragma unroll
            for (int c = 0; c < BN / 32; ++c) ptx::tmem_ld_32x32b_x32(taddr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
            ptx::tmem_ld_wait();
            ptx::tc_fence_before();
            __syncwarp();
            if (lane == 0) ptx::mbar_arrive(&tmem_empty[acc]);   // accumulator stage free again

            const float mx = sqrtf(rn);
            const uint32_t ridx = static_cast<uint32_t>(row);
#pragma unroll
            for (int q = 0; q < BN; ++q) {
                if (q < p.nq) {                                  // warp-uniform
                    const float dot = __uint_as_float(v[q]);
                    const float thf = sink_thr_score(sink, q);
                    // cheap, conservative pre-filter in the space of the accumulator: no division,
                    // no key; NaN always goes on to the exact test (NaN ranks first for cosine)
                    float den = 0.f, s = 0.f;
                    bool maybe;
                    if (largest) {
                        den = fmaf(sQc[q], mx, 1e-6f);
                        float bound = thf * den;
                        bound = fmaf(-fabsf(bound), 1e-6f, bound);
                        maybe = !(dot < bound);
                    } else {
                        s = (sQc[q] - 2.0f * dot + rn) * p.inv_dd;
                        maybe = !(s > thf);
                    }
                    maybe = maybe && valid;
                    if (__any_sync(0xffffffffu, maybe)) {
                        if (largest) s = dot / den;
                        const uint64_t comp = make_composite(score_to_key(s, largest), ridx);
                        const bool pass = maybe && (comp > sink_thr(sink, q));
                        sink_insert_rows(sink, q, pass, comp);
                    }
                }
            }
            ptx::named_bar_sync(1, kEpiThreads);
            sink_prune_if_full(sink, p.nq, e, 4, hist);
            ptx::named_bar_sync(1, kEpiThreads);
            if (e == 0 && lane == 0) *sTilesDone = it + 1;
        }
        // final: counts and the last published bound
        ptx::named_bar_sync(1, kEpiThreads);
        for (int q = e * 32 + lane; q < p.nq; q += 128) {
            p.counts[static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q] = static_cast<int>(lds_u32(sink.cnt + q * 4));
            const uint32_t mine = lds_u32(sink.lmax + q * 4);
            if (p.use_gtau && mine) st_cg_u32(p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0 + q, mine);
        }
    } else {
        // ===================== threshold exchange =====================
        // publish this CTA's best keys, reduce one query column for everybody, apply all bounds;
        // runs free of the epilogue (all state is monotone), fast at first, then at a trickle
        if (p.use_gtau && my_tiles > 0) {
            uint32_t* my_row = p.gtop + static_cast<size_t>(blockIdx.x) * p.Qtot + p.q0;
            const int rq = static_cast<int>(blockIdx.x) % p.nq;
            int round = 0;
            while (*sTilesDone < my_tiles) {
                exchange_publish(sink, p.nq, my_row);
                const uint32_t lo = exchange_reduce(p.gtop + p.q0 + rq, p.p_stride, p.Qtot);
                if (lane == 0 && lo != 0u) atomicMax(p.gtau + p.q0 + rq, lo);
                for (int q = lane; q < p.nq; q += 32) exchange_apply(sink, q, ld_cg_u32(p.gtau + p.q0 + q));
                ++round;
                __nanosleep(round < 24 ? 100 : 3000);
            }
        }
    }

    ptx::tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        ptx::tc_fence_after();
        ptx::tmem_dealloc(tmem_base, 2 * BN);
    }
}

// Queries -> bf16 operand matrix [q_pad, Dp] (zero padded) + per-query constants.
__global__ void pack_queries_kernel(const float* __restrict__ t, int Q, int D, int Dp, int q_pad, int metric,
                                    __nv_bfloat16* __restrict__ bq, float* __restrict__ qconst) {
    const int q = blockIdx.x;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __shared__ double red[8];
    double acc = 0.0;
    for (int d = threadIdx.x; d < Dp; d += blockDim.x) {
        float v = (q < Q && d < D) ? t[static_cast<size_t>(q) * D + d] : 0.f;
        bq[static_cast<size_t>(q) * Dp + d] = __float2bfloat16_rn(v);
        acc += static_cast<double>(v * v);
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

// 2-D bf16 row-major [rows, Dp] tensor, box = 64 columns x box_rows rows, 128-byte swizzle.
static int make_tmap_2d(CUtensorMap* m, const void* ptr, int64_t rows, int Dp, int box_rows) {
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

constexpr int kTcBN = 64;

static size_t tc_tail_bytes(int BN) {
    return (2 * kMaxStages + 1 + 4) * sizeof(uint64_t) + BN * (8 + 4 + 4 + 4 + 4) + 4 * 256 * 4 + 16;
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
    int rc = make_tmap_2d(&b->tmap_bank, b->data, b->rows_pad, b->Dp, kTileRows);
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
    if (!b->tmap_ready) return set_error(SKY_ERR_STATE, "bank has no TMA descriptor (finalize first)");
    constexpr int BN = kTcBN;
    const int q_pad = static_cast<int>(round_up(Q, BN));
    __nv_bfloat16* bq = reinterpret_cast<__nv_bfloat16*>(b->ws2);
    float* qconst = reinterpret_cast<float*>(reinterpret_cast<unsigned char*>(b->ws2) +
                                             round_up(static_cast<int64_t>(q_pad) * b->Dp * 2, 256));
    pack_queries_kernel<<<q_pad, 256, 0, st>>>(t, Q, b->D, b->Dp, q_pad, metric, bq, qconst);
    SKY_LAUNCH_CHECK("pack_queries_kernel");

    const int stages = tc_stages(b->Dp, BN);
    const size_t smem = 1024 + static_cast<size_t>(b->Dp) * BN * 2 + static_cast<size_t>(stages) * kStageBytes + tc_tail_bytes(BN);
    SKY_CUDA(cudaFuncSetAttribute(tc_search_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    const int grid = s.P;
    for (int q0 = 0; q0 < Q; q0 += BN) {
        CUtensorMap tmq;
        int rc = make_tmap_2d(&tmq, bq + static_cast<size_t>(q0) * b->Dp, BN, b->Dp, BN);
        if (rc) return rc;
        TcParams p;
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
        p.inv_dd = 1.0f / (static_cast<float>(b->D) * static_cast<float>(b->D));
        prof_mark(b, st);
        tc_search_kernel<BN><<<grid, kTcThreads, smem, st>>>(b->tmap_bank, tmq, p);
        prof_mark(b, st);
        SKY_LAUNCH_CHECK("tc_search_kernel");
    }
    return SKY_OK;
}

}  // namespace sky
