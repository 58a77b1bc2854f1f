// merge.cu -- K3/K4: merging candidate lists into the final, best-first top-k.
//
//  * merge_lists_kernel: per-CTA candidate lists of one scorer launch -> top-k of the shard.
//  * merge_candidates_kernel: R (score, index) lists per query -> top-k; this is the device merge
//    after the NCCL all-gather of per-shard results, and the running merge that replaces
//    update_best_scores (reference utils/similarity.py:18-35: cat + argsort + [:n_save]).
// Output padding when fewer than k candidates exist follows the reference's initial fill
// (utils/similarity.py:66): -inf for cosine, +inf otherwise; index -1.
#include "bank.cuh"
#include "ptx.cuh"
#include "topk.cuh"

namespace sky {

constexpr int kMergeThreads = 256;
constexpr int kPool = 4096;      // smem pool of pre-filtered candidates per query
constexpr int kMaxK = 4096;

__global__ void init_state_kernel(uint32_t* gtop, uint32_t* gtau, int* counts, int Qtot, int p_stride, int p_active, int P) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < Qtot * p_stride) gtop[i] = 0u;   // [p][q]; 0 = nothing published
    if (i < Qtot * P) counts[i] = 0;
    if (i < Qtot) gtau[i] = 0u;
}

__device__ __forceinline__ void write_result(uint64_t c, bool largest, int64_t idx_offset, float* score, int64_t* idx) {
    if (c == 0) {
        *score = largest ? -INFINITY : INFINITY;
        *idx = -1;
    } else {
        *score = key_to_score(composite_key(c), largest);
        *idx = static_cast<int64_t>(composite_idx(c)) + idx_offset;
    }
}

__global__ void __launch_bounds__(kMergeThreads)
merge_lists_kernel(const uint64_t* __restrict__ lists, const int* __restrict__ counts,
                   const uint32_t* __restrict__ gtop, int P, int p_stride, int Qtot, int cap, int k, int kpad,
                   int use_gtau, int largest, int64_t idx_offset, float* __restrict__ out_scores,
                   int64_t* __restrict__ out_idx, const XchgTarget xt) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sel = reinterpret_cast<uint64_t*>(smem_raw);    // [kpad]
    uint64_t* pool = sel + kpad;                              // [kPool]
    __shared__ uint32_t hist[256];
    __shared__ uint32_t scratch[4];
    __shared__ uint32_t npool;
    __shared__ uint32_t tau_s;
    const int q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    constexpr int nwarps = kMergeThreads / 32;

    if (tid == 0) { npool = 0; tau_s = 0u; }
    ptx::griddep_wait();      // launched with programmatic serialization: the scorer's lists are complete from here on
    __syncthreads();
    // list fills: issued before the bound reduction so both L2 round trips overlap
    constexpr int kCountCache = 2048;
    const int per = (P + kMergeThreads - 1) / kMergeThreads;
    int mine[kCountCache / kMergeThreads];
    int tot = 0;
    if (P <= kCountCache) {
#pragma unroll
        for (int u = 0; u < kCountCache / kMergeThreads; ++u) {
            const int pi = tid * per + u;
            mine[u] = (u < per && pi < P) ? min(counts[static_cast<size_t>(pi) * Qtot + q], cap) : 0;
            tot += mine[u];
        }
    }
    // final grid-wide bound: the k-th largest of the per-CTA best keys (see topk.cuh)
    if (use_gtau && warp == 0) {
        const uint32_t lo = exchange_reduce(gtop + q, p_stride, Qtot, k);
        if (lane == 0) tau_s = lo;
    }
    __syncthreads();
    // every composite with key >= tau must be kept; tau == 0 / disabled keeps everything valid
    const uint64_t keep_ge = (use_gtau && tau_s != 0u) ? (static_cast<uint64_t>(tau_s) << 32) : 1ull;

    // all list fills first (one L2 round trip), then ONE flat pass over the concatenation of the lists: entry j of
    // the concatenation belongs to the list whose offset range holds j (binary search in shared memory), so every
    // thread's loads are independent of each other -- a list-per-warp loop would pay one dependent L2 round trip
    // per list (P / 8 of them in a row)
    __shared__ int s_off[kCountCache + 1];
    __shared__ int s_wsum[kMergeThreads / 32];
    if (P <= kCountCache) {
        // exclusive scan of the fills (loaded above): thread t owns lists [t * per, (t + 1) * per)
        int incl = tot;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
            const int o = __shfl_up_sync(0xffffffffu, incl, off);
            if (lane >= off) incl += o;
        }
        if (lane == 31) s_wsum[warp] = incl;
        __syncthreads();
        int wbase = 0;
        for (int i = 0; i < warp; ++i) wbase += s_wsum[i];
        int run = wbase + incl - tot;
#pragma unroll
        for (int u = 0; u < kCountCache / kMergeThreads; ++u) {
            const int pi = tid * per + u;
            if (u < per && pi < P) s_off[pi] = run;
            run += mine[u];
        }
        if (tid == kMergeThreads - 1) s_off[P] = run;
        __syncthreads();
        const int total = s_off[P];
        for (int j0 = tid; j0 < total; j0 += 4 * kMergeThreads) {
            uint64_t v[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {           // four independent loads in flight per thread
                const int j = j0 + u * kMergeThreads;
                v[u] = 0ull;
                if (j < total) {
                    int lo = 0, hi = P;             // largest lo with s_off[lo] <= j
                    while (hi - lo > 1) {
                        const int mid = (lo + hi) >> 1;
                        if (s_off[mid] <= j) lo = mid; else hi = mid;
                    }
                    v[u] = __ldcg(reinterpret_cast<const unsigned long long*>(lists + (static_cast<size_t>(lo) * Qtot + q) * cap + (j - s_off[lo])));
                }
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if (v[u] >= keep_ge) {
                    const uint32_t pos = atomicAdd(&npool, 1u);
                    if (pos < kPool) pool[pos] = v[u];
                }
            }
        }
    } else {
        for (int p = warp; p < P; p += nwarps) {
            const int n = counts[static_cast<size_t>(p) * Qtot + q];
            const uint64_t* e = lists + (static_cast<size_t>(p) * Qtot + q) * cap;
            for (int i = lane; i < n; i += 32) {
                const uint64_t v = e[i];
                if (v >= keep_ge) {
                    const uint32_t pos = atomicAdd(&npool, 1u);
                    if (pos < kPool) pool[pos] = v;
                }
            }
        }
    }
    __syncthreads();
    const int np = static_cast<int>(npool);
    if (np <= kRankSortMax) {
        block_rank_topk(pool, np, k, kpad, sel);
    } else if (np <= kPool) {
        block_select_sort([&](int j) { return pool[j]; }, np, k, kpad, sel, hist, scratch);
    } else {
        // rare: too many survivors (adversarial order / massive ties) -> select straight from L2
        auto fetch = [&](int j) -> uint64_t {
            const int p = j / cap, i = j - p * cap;
            if (i >= counts[static_cast<size_t>(p) * Qtot + q]) return 0ull;
            const uint64_t v = lists[(static_cast<size_t>(p) * Qtot + q) * cap + i];
            return v >= keep_ge ? v : 0ull;
        };
        block_select_sort(fetch, P * cap, k, kpad, sel, hist, scratch);
    }
    __syncthreads();
    if (xt.world > 0) {
        // sharded search: the shard's top-k of query q goes straight into this rank's slot of EVERY peer's exchange
        // buffer (plain stores over NVLink), then one flag per peer says "query q of search `seq` has landed"
        const size_t nq = static_cast<size_t>(Qtot) * k;
        for (int p = 0; p < xt.world; ++p) {
            unsigned char* slot = xt.base[p] + xchg_slot_off(xt.slot_units, xt.world, xt.parity, xt.rank);
            int64_t* di = reinterpret_cast<int64_t*>(slot) + static_cast<size_t>(q) * k;
            float* ds = reinterpret_cast<float*>(slot + nq * 8) + static_cast<size_t>(q) * k;
            for (int j = tid; j < k; j += kMergeThreads) write_result(sel[j], largest != 0, idx_offset, ds + j, di + j);
        }
        __threadfence_system();
        __syncthreads();
        if (tid < xt.world)
            st_release_sys_u32(reinterpret_cast<unsigned*>(xt.base[tid] + xchg_flag_off(xt.slot_units, xt.world, xt.max_Q, xt.parity, xt.rank)) + q, xt.seq);
        return;
    }
    for (int j = tid; j < k; j += kMergeThreads)
        write_result(sel[j], largest != 0, idx_offset, out_scores + static_cast<size_t>(q) * k + j,
                     out_idx + static_cast<size_t>(q) * k + j);
}

__global__ void __launch_bounds__(kMergeThreads)
merge_candidates_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, int R, int Q, int k_in,
                        int64_t stride_s, int64_t stride_i, int k_out, int kpad, int largest,
                        float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sel = reinterpret_cast<uint64_t*>(smem_raw);    // [kpad]
    __shared__ uint32_t hist[256];
    __shared__ uint32_t scratch[4];
    const int q = blockIdx.x, tid = threadIdx.x;
    const int n = R * k_in;
    // composite low word = position in the (rank-major) concatenation: lists are best-first with
    // ties by lower index and shards hold increasing row ranges, so position order == index order
    // list r of this query starts at r * stride (elements): dense [R, Q, k_in] or one gathered buffer per rank
    auto src_s = [&](int j) -> size_t {
        const int r = j / k_in, i = j - r * k_in;
        return static_cast<size_t>(r) * stride_s + static_cast<size_t>(q) * k_in + i;
    };
    auto src_i = [&](int j) -> size_t {
        const int r = j / k_in, i = j - r * k_in;
        return static_cast<size_t>(r) * stride_i + static_cast<size_t>(q) * k_in + i;
    };
    auto fetch = [&](int j) -> uint64_t {
        if (idx[src_i(j)] < 0) return 0ull;
        return make_composite(score_to_key(scores[src_s(j)], largest != 0), static_cast<uint32_t>(j));
    };
    if (n <= kRankSortMax) {
        uint64_t* cand = sel + kpad;                          // [n] (dynamic smem sized by the launcher)
        for (int j = tid; j < n; j += kMergeThreads) cand[j] = fetch(j);
        __syncthreads();
        block_rank_topk(cand, n, k_out, kpad, sel);
    } else {
        block_select_sort(fetch, n, k_out, kpad, sel, hist, scratch);
    }
    __syncthreads();
    for (int j = tid; j < k_out; j += kMergeThreads) {
        const uint64_t c = sel[j];
        float* so = out_scores + static_cast<size_t>(q) * k_out + j;
        int64_t* io = out_idx + static_cast<size_t>(q) * k_out + j;
        if (c == 0) {
            *so = largest ? -INFINITY : INFINITY;
            *io = -1;
        } else {
            const int j0 = static_cast<int>(composite_idx(c));
            *so = scores[src_s(j0)];
            *io = idx[src_i(j0)];
        }
    }
}

static int next_pow2(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}

int launch_init_state(const SearchState& s, int p_active, cudaStream_t st) {
    const int n = s.Qtot * (s.p_stride > s.P ? s.p_stride : s.P);
    init_state_kernel<<<(n + 255) / 256, 256, 0, st>>>(s.gtop, s.gtau, s.counts, s.Qtot, s.p_stride, p_active, s.P);
    SKY_LAUNCH_CHECK("init_state_kernel");
    return SKY_OK;
}

int launch_merge_lists(const SearchState& s, int metric, int64_t idx_offset, float* out_scores, int64_t* out_idx,
                       cudaStream_t st, XchgTarget* xt) {
    if (s.k > kMaxK) return set_error(SKY_ERR_UNSUPPORTED, "k=%d exceeds the merge limit %d", s.k, kMaxK);
    const int kpad = next_pow2(s.k);
    const size_t smem = static_cast<size_t>(kpad + kPool) * sizeof(uint64_t);
    SKY_CUDA(cudaFuncSetAttribute(merge_lists_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    SKY_CUDA(launch_maybe_pdl(env_knob("SKY_PDL", kUsePdl) != 0, merge_lists_kernel, dim3(s.Qtot), dim3(kMergeThreads), smem, st,
                              s.lists, s.counts, s.gtop, s.P, s.p_stride, s.Qtot, s.cap, s.k, kpad, s.use_gtau,
                              metric_largest(metric) ? 1 : 0, idx_offset, out_scores, out_idx, xt ? *xt : XchgTarget{}));
    SKY_LAUNCH_CHECK("merge_lists_kernel");
    if (xt) xt->fused = true;
    return SKY_OK;
}

int launch_merge_candidates(const float* scores, const int64_t* idx, int R, int Q, int k_in, int64_t stride_s,
                            int64_t stride_i, int k_out, int metric, float* out_scores, int64_t* out_idx, cudaStream_t st) {
    if (k_out > kMaxK) return set_error(SKY_ERR_UNSUPPORTED, "k=%d exceeds the merge limit %d", k_out, kMaxK);
    if (Q == 0) return SKY_OK;
    const int kpad = next_pow2(k_out);
    const size_t smem = static_cast<size_t>(kpad + (R * k_in <= kRankSortMax ? R * k_in : 0)) * sizeof(uint64_t);
    SKY_CUDA(cudaFuncSetAttribute(merge_candidates_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    if (stride_s <= 0) stride_s = static_cast<int64_t>(Q) * k_in;
    if (stride_i <= 0) stride_i = static_cast<int64_t>(Q) * k_in;
    merge_candidates_kernel<<<Q, kMergeThreads, smem, st>>>(scores, idx, R, Q, k_in, stride_s, stride_i, k_out, kpad,
                                                           metric_largest(metric) ? 1 : 0, out_scores, out_idx);
    SKY_LAUNCH_CHECK("merge_candidates_kernel");
    return SKY_OK;
}

}  // namespace sky
