// topk.cuh -- fused top-k machinery shared by the SIMT and tcgen05 scorers and the merge kernels.
//
// Replaces update_best_scores (reference utils/similarity.py:18-35: cat + argsort + gather per
// batch).  Scores never form a [Q, N] matrix in HBM:
//   * each CTA keeps, per query, a private candidate list in L2-resident global memory, gated by a
//     threshold held in shared memory;
//   * the threshold is the max of (a) the CTA's own k-th best after an in-place warp radix-select
//     prune (overflow protection) and (b) a grid-wide lower bound: every CTA publishes the best
//     key it has seen per query; if k <= #CTAs, at least k bank rows score >= the minimum of the
//     published keys, so anything below it can never be in the global top-k;
//   * a final kernel merges the per-CTA lists (filter by the final bound, select, sort).
#pragma once
#include "common.cuh"

namespace sky {

__device__ __forceinline__ uint64_t ld_cg_u64(const uint64_t* p) {
    return __ldcg(reinterpret_cast<const unsigned long long*>(p));
}
__device__ __forceinline__ void st_cg_u64(uint64_t* p, uint64_t v) {
    __stcg(reinterpret_cast<unsigned long long*>(p), static_cast<unsigned long long>(v));
}
__device__ __forceinline__ uint32_t ld_cg_u32(const uint32_t* p) { return __ldcg(p); }
__device__ __forceinline__ void st_cg_u32(uint32_t* p, uint32_t v) { __stcg(p, v); }

// ---------------------------------------------------------------------------------------------
// explicit shared-memory accessors (32-bit shared-window addresses): the sink state is reached
// through structs, where the compiler would otherwise fall back to slow generic LD/ST/ATOM
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_addr(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }
__device__ __forceinline__ uint32_t lds_u32(uint32_t a) {
    uint32_t v; asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ uint64_t lds_u64(uint32_t a) {
    uint64_t v; asm volatile("ld.volatile.shared.u64 %0, [%1];" : "=l"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ float lds_f32(uint32_t a) {
    float v; asm volatile("ld.volatile.shared.f32 %0, [%1];" : "=f"(v) : "r"(a) : "memory"); return v;
}
__device__ __forceinline__ void sts_u32(uint32_t a, uint32_t v) { asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void sts_f32(uint32_t a, float v) { asm volatile("st.volatile.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory"); }
__device__ __forceinline__ uint32_t atoms_add_u32(uint32_t a, uint32_t v) {
    uint32_t o; asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(o) : "r"(a), "r"(v) : "memory"); return o;
}
__device__ __forceinline__ void reds_add_u32(uint32_t a, uint32_t v) { asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_max_u32(uint32_t a, uint32_t v) { asm volatile("red.shared.max.u32 [%0], %1;" ::"r"(a), "r"(v) : "memory"); }
__device__ __forceinline__ void reds_max_u64(uint32_t a, uint64_t v) { asm volatile("red.shared.max.u64 [%0], %1;" ::"r"(a), "l"(v) : "memory"); }

// ---------------------------------------------------------------------------------------------
// Warp-level exact selection: the kth-largest (1-based) of n distinct 64-bit composites that
// live in global memory (L2).  8 passes of 8 bits, MSB first; hist = 256 u32 of warp-private smem
// (shared-window address).  All 32 lanes must call; all get the result.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void radix_pick_bucket(uint32_t hist, int lane, int remaining,
                                                  uint32_t& digit, uint32_t& cnt_above) {
    // lane L owns bins [8L, 8L+8); higher lanes own higher digits
    uint32_t c[8];
    uint32_t mine = 0;
#pragma unroll
    for (int b = 0; b < 8; ++b) { c[b] = lds_u32(hist + (lane * 8 + b) * 4); mine += c[b]; }
    uint32_t incl = mine;   // inclusive suffix sum over lanes
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        uint32_t o = __shfl_down_sync(0xffffffffu, incl, off);
        if (lane + off < 32) incl += o;
    }
    const uint32_t above = incl - mine;
    const uint32_t rem = static_cast<uint32_t>(remaining);
    const bool owner = (above < rem) && (rem <= above + mine);
    uint32_t d = 0, ca = 0;
    if (owner) {
        uint32_t run = above;
#pragma unroll
        for (int b = 7; b >= 0; --b) {
            if (run < rem && rem <= run + c[b]) { d = lane * 8 + b; ca = run; }
            run += c[b];
        }
    }
    const uint32_t ballot = __ballot_sync(0xffffffffu, owner);
    const int src = __ffs(ballot) - 1;
    digit = __shfl_sync(0xffffffffu, d, src);
    cnt_above = __shfl_sync(0xffffffffu, ca, src);
}

__device__ __forceinline__ uint64_t warp_select_kth(const uint64_t* e, int n, int kth, uint32_t hist) {
    const int lane = threadIdx.x & 31;
    uint64_t prefix = 0, mask = 0;
    int remaining = kth;
#pragma unroll 1
    for (int pass = 7; pass >= 0; --pass) {
        const int shift = pass * 8;
#pragma unroll
        for (int b = 0; b < 8; ++b) sts_u32(hist + (lane * 8 + b) * 4, 0u);
        __syncwarp();
        for (int i = lane; i < n; i += 32) {
            uint64_t v = ld_cg_u64(e + i);
            if ((v & mask) == prefix) reds_add_u32(hist + static_cast<uint32_t>((v >> shift) & 0xFF) * 4, 1u);
        }
        __syncwarp();
        uint32_t digit, cnt_above;
        radix_pick_bucket(hist, lane, remaining, digit, cnt_above);
        remaining -= static_cast<int>(cnt_above);
        prefix |= static_cast<uint64_t>(digit) << shift;
        mask |= static_cast<uint64_t>(0xFF) << shift;
        __syncwarp();
    }
    return prefix;
}

// Keep the entries >= thr at the front of e[0..n).  Returns the new count.
// In-place and safe: writes always trail reads.
__device__ __forceinline__ int warp_compact_ge(uint64_t* e, int n, uint64_t thr) {
    const int lane = threadIdx.x & 31;
    int out = 0;
    for (int base = 0; base < n; base += 32) {
        int i = base + lane;
        uint64_t v = (i < n) ? ld_cg_u64(e + i) : 0;
        bool keep = (i < n) && (v >= thr);
        uint32_t m = __ballot_sync(0xffffffffu, keep);
        __syncwarp();
        if (keep) st_cg_u64(e + out + __popc(m & ((1u << lane) - 1u)), v);
        out += __popc(m);
        __syncwarp();
    }
    return out;
}

// ---------------------------------------------------------------------------------------------
// CTA-private candidate sink.  Shared-memory state is addressed through the shared window.
// Invariant: at most kPruneSlack inserts per query happen between two calls of
// sink_prune_if_full, and cap >= k + kPruneSlack, so a list can never overflow.
// ---------------------------------------------------------------------------------------------
struct Sink {
    uint64_t* lists;   // global: [nq][cap] for this CTA
    uint32_t thr;      // smem u64[nq]: a candidate passes iff composite > thr
    uint32_t thr_f;    // smem f32[nq]: the threshold as a score, for the cheap pre-filter
    uint32_t cnt;      // smem u32[nq]
    uint32_t lmax;     // smem u32[nq]: best key inserted so far (published grid-wide)
    int cap;
    int k;
    bool largest;
};

__device__ __forceinline__ uint64_t sink_thr(const Sink& s, int q) { return lds_u64(s.thr + q * 8); }
__device__ __forceinline__ float sink_thr_score(const Sink& s, int q) { return lds_f32(s.thr_f + q * 4); }

// Raise the threshold of query q to composite `c` (monotone; callable concurrently).
__device__ __forceinline__ void sink_raise(const Sink& s, int q, uint64_t c) {
    reds_max_u64(s.thr + q * 8, c);
    // the float mirror may lag or be slightly stale: it only gates the pre-filter, which is
    // conservative, and the exact composite compare decides.  Benign race: values only tighten.
    const uint64_t now = lds_u64(s.thr + q * 8);
    sts_f32(s.thr_f + q * 4, key_to_score(composite_key(now), s.largest));
}

// Tensor-path pattern: the 32 lanes of a warp hold 32 different bank rows for the SAME query q.
__device__ __forceinline__ void sink_insert_rows(const Sink& s, int q, bool pass, uint64_t comp) {
    const uint32_t m = __ballot_sync(0xffffffffu, pass);
    if (m == 0) return;
    const int lane = threadIdx.x & 31;
    const int leader = __ffs(m) - 1;
    const uint32_t kmax = __reduce_max_sync(0xffffffffu, pass ? composite_key(comp) : 0u);
    uint32_t base = 0;
    if (lane == leader) {
        base = atoms_add_u32(s.cnt + q * 4, __popc(m));
        reds_max_u32(s.lmax + q * 4, kmax);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    if (pass) st_cg_u64(s.lists + static_cast<size_t>(q) * s.cap + base + __popc(m & ((1u << lane) - 1u)), comp);
}

// SIMT-path pattern: each calling lane owns one (item, query) pair.
__device__ __forceinline__ void sink_insert_one(const Sink& s, int q, uint64_t comp) {
    if (comp > sink_thr(s, q)) {
        const uint32_t pos = atoms_add_u32(s.cnt + q * 4, 1u);
        reds_max_u32(s.lmax + q * 4, composite_key(comp));
        st_cg_u64(s.lists + static_cast<size_t>(q) * s.cap + pos, comp);
    }
}

// One warp prunes query q down to its k best and raises the threshold.  All lanes call.
__device__ __forceinline__ void sink_prune(const Sink& s, int q, uint32_t hist) {
    const int n = static_cast<int>(lds_u32(s.cnt + q * 4));
    if (n <= s.k) return;
    uint64_t* e = s.lists + static_cast<size_t>(q) * s.cap;
    const uint64_t kth = warp_select_kth(e, n, s.k, hist);
    const int m = warp_compact_ge(e, n, kth);
    if ((threadIdx.x & 31) == 0) {
        sts_u32(s.cnt + q * 4, static_cast<uint32_t>(m));   // == k (composites are distinct)
        sink_raise(s, q, kth);
    }
    __syncwarp();
}

// Called by `nwarps` warps (warp_in_group = 0..nwarps-1) between two barriers.
// slack: the most inserts per query that can happen before the next call (cap >= k + slack).
__device__ __forceinline__ void sink_prune_if_full(const Sink& s, int nq, int warp_in_group, int nwarps,
                                                   uint32_t hist, int slack = kPruneSlack) {
    // one vector pass over the counts instead of a serial chain of dependent shared loads
    const int lane = threadIdx.x & 31;
    for (int base = 0; base < nq; base += 32 * nwarps) {
        const int q = base + lane * nwarps + warp_in_group;      // this warp owns q % nwarps == warp_in_group
        const bool full = q < nq && static_cast<int>(lds_u32(s.cnt + q * 4)) > s.cap - slack;
        uint32_t todo = __ballot_sync(0xffffffffu, full);
        while (todo) {
            const int l = __ffs(todo) - 1;
            todo &= todo - 1;
            sink_prune(s, base + l * nwarps + warp_in_group, hist);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Grid-wide bound exchange.
//   gtop [p_stride][Qtot] u32: row p = best key CTA p has inserted per query (0 = nothing yet)
//   gtau [Qtot] u32:           the k-th largest entry of column q, maintained by reducer CTAs
// Every published key belongs to a DIFFERENT bank row (CTAs own disjoint rows) that sits in its
// CTA's candidate list, so at least k rows have key >= (k-th largest published key): a candidate
// with a smaller key can never be in the global top-k.  Fewer than k published keys give 0 = no
// information.  All values only grow; every race is benign.
// ---------------------------------------------------------------------------------------------
// publish this CTA's best keys for queries [0, nq) (one warp)
__device__ __forceinline__ void exchange_publish(const Sink& s, int nq, uint32_t* gtop_row) {
    for (int q = threadIdx.x & 31; q < nq; q += 32) {
        const uint32_t mine = lds_u32(s.lmax + q * 4);
        if (mine) st_cg_u32(gtop_row + q, mine);
    }
}
// same, but only keys that changed since the last call are rewritten (last[i] <-> query lane + 32 i): the
// entries of all CTAs share a few L2 lines, and rewriting them every round makes those lines a hot spot
// that delays the bank stream queued behind it
template <int N>
__device__ __forceinline__ void exchange_publish_changed(const Sink& s, int nq, uint32_t* gtop_row, uint32_t (&last)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) {
        const int q = (threadIdx.x & 31) + 32 * i;
        if (q < nq) {
            const uint32_t mine = lds_u32(s.lmax + q * 4);
            if (mine != last[i]) { st_cg_u32(gtop_row + q, mine); last[i] = mine; }
        }
    }
}
// one warp: k-th largest (1-based) of one query column of gtop; p_stride <= 32 * kMaxPerLane
constexpr int kXchgPerLane = 32;   // up to 1024 CTAs
__device__ __forceinline__ uint32_t exchange_reduce(const uint32_t* gtop_col, int p_stride, int q_stride, int k) {
    const int lane = threadIdx.x & 31;
    // bitwise binary search for the largest x with #{v >= x} >= k, values re-read from L2 each round
    // would be slow: keep them in registers (fixed upper bound, predicated)
    uint32_t v[8];
    uint32_t prefix = 0;
    if (p_stride <= 256) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int p = lane + 32 * i;
            v[i] = p < p_stride ? ld_cg_u32(gtop_col + static_cast<size_t>(p) * q_stride) : 0u;
        }
#pragma unroll 1
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = prefix | (1u << bit);
            int c = 0;
#pragma unroll
            for (int i = 0; i < 8; ++i) c += (v[i] >= cand);
            if (__reduce_add_sync(0xffffffffu, c) >= k) prefix = cand;
        }
    } else {
#pragma unroll 1
        for (int bit = 31; bit >= 0; --bit) {
            const uint32_t cand = prefix | (1u << bit);
            int c = 0;
            for (int p = lane; p < p_stride; p += 32) c += (ld_cg_u32(gtop_col + static_cast<size_t>(p) * q_stride) >= cand);
            if (__reduce_add_sync(0xffffffffu, c) >= k) prefix = cand;
        }
    }
    return prefix;
}
// raise the local threshold of query q to the grid-wide bound `lo` (a key; 0 = no information)
__device__ __forceinline__ void exchange_apply(const Sink& s, int q, uint32_t lo) {
    // every composite with key >= lo must be kept: pass iff comp > (lo << 32) - 1
    if (lo != 0u) sink_raise(s, q, (static_cast<uint64_t>(lo) << 32) - 1ull);
}

// ---------------------------------------------------------------------------------------------
// Peer-memory exchange layout (csrc/exchange.cu): a rank's buffer = slots[2][world][slot_units] int64 | flags[2][world][max_Q] u32
// ---------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ size_t xchg_slot_off(size_t slot_units, int world, int parity, int r) {
    return (static_cast<size_t>(parity) * world + r) * slot_units * 8;
}
__host__ __device__ __forceinline__ size_t xchg_flag_off(size_t slot_units, int world, int max_Q, int parity, int r) {
    return 2 * static_cast<size_t>(world) * slot_units * 8 + (static_cast<size_t>(parity) * world + r) * static_cast<size_t>(max_Q) * 4;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}

// Zero the grid-wide exchange state of a search (gtop / gtau / counts) from inside another kernel's grid: the query
// packing kernels of the tensor paths do this on the side, which saves the separate init_state launch.
struct StateInit {
    uint32_t* gtop; uint32_t* gtau; int* counts;
    int Qtot, p_stride, P;
};
__device__ __forceinline__ void state_init_gridwide(const StateInit& si) {
    if (!si.gtop) return;
    const int n = si.Qtot * (si.p_stride > si.P ? si.p_stride : si.P);
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        if (i < si.Qtot * si.p_stride) si.gtop[i] = 0u;
        if (i < si.Qtot * si.P) si.counts[i] = 0;
        if (i < si.Qtot) si.gtau[i] = 0u;
    }
}

// ---------------------------------------------------------------------------------------------
// Block-level exact top-k over a virtual array fetch(j), j in [0, n): 0 = empty entry.
// Radix-selects the k-th largest, gathers the winners into sel[kpad] (smem, kpad = pow2 >= k),
// bitonic-sorts them descending.  Afterwards sel[0..k) holds the result (0 = padding).
//   hist: 256 u32 smem; scratch: 4 u32 smem.
// sorted = false: a caller that only needs the SET of winners and the k-th best (a running top-k between two phases)
// skips the bitonic sort -- more than half of the barriers of this routine at kpad = 1024 -- whenever more than k
// candidates exist; the return value is the k-th best composite (0 while fewer than k candidates exist).
// ---------------------------------------------------------------------------------------------
template <typename Fetch>
__device__ uint64_t block_select_sort(Fetch fetch, int n, int k, int kpad, uint64_t* sel, uint32_t* hist,
                                      uint32_t* scratch, bool sorted = true) {
    const int tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { scratch[0] = 0; scratch[3] = 0; }
    __syncthreads();
    {
        uint32_t local = 0;
        for (int j = tid; j < n; j += nt) local += (fetch(j) != 0);
        if (local) atomicAdd(&scratch[0], local);
    }
    __syncthreads();
    const int total = static_cast<int>(scratch[0]);
    uint64_t thr = 1;   // keep everything valid
    if (total > k) {
        uint64_t prefix = 0, mask = 0;
        int remaining = k;
        for (int pass = 7; pass >= 0; --pass) {
            const int shift = pass * 8;
            for (int b = tid; b < 256; b += nt) hist[b] = 0;
            __syncthreads();
            for (int j = tid; j < n; j += nt) {
                uint64_t v = fetch(j);
                if (v != 0 && (v & mask) == prefix) atomicAdd(&hist[(v >> shift) & 0xFF], 1u);
            }
            __syncthreads();
            if (tid < 32) {
                uint32_t digit, cnt_above;
                radix_pick_bucket(smem_addr(hist), tid, remaining, digit, cnt_above);
                if (tid == 0) { scratch[1] = digit; scratch[2] = cnt_above; }
            }
            __syncthreads();
            remaining -= static_cast<int>(scratch[2]);
            prefix |= static_cast<uint64_t>(scratch[1]) << shift;
            mask |= static_cast<uint64_t>(0xFF) << shift;
            __syncthreads();
        }
        thr = prefix;
    }
    for (int j = tid; j < kpad; j += nt) sel[j] = 0;
    __syncthreads();
    for (int j = tid; j < n; j += nt) {
        uint64_t v = fetch(j);
        if (v != 0 && v >= thr) {
            uint32_t pos = atomicAdd(&scratch[3], 1u);
            if (pos < static_cast<uint32_t>(kpad)) sel[pos] = v;
        }
    }
    __syncthreads();
    if (!sorted && total > k) return thr;      // sel[0..k) = the k winners in arbitrary order (composites are distinct)
    // bitonic sort, descending
    for (int size = 2; size <= kpad; size <<= 1) {
        for (int stride = size >> 1; stride > 0; stride >>= 1) {
            for (int i = tid; i < (kpad >> 1); i += nt) {
                int lo = ((i / stride) * (stride << 1)) + (i % stride);
                int hi = lo + stride;
                bool desc = ((lo & size) == 0);
                uint64_t a = sel[lo], b = sel[hi];
                if ((a < b) == desc) { sel[lo] = b; sel[hi] = a; }
            }
            __syncthreads();
        }
    }
    return sel[k - 1];
}

// ---------------------------------------------------------------------------------------------
// Block-level exact top-k for SMALL candidate sets held in shared memory (n <= ~1024): every thread ranks
// its candidates by counting the larger ones (all threads read the same element -> broadcast, no bank
// conflicts) and drops them at sel[rank].  No radix passes, two barriers instead of ~80: this is the common
// case of every merge once bounds exist (k plus a few hundred survivors).  Composites are distinct (0 = empty).
// ---------------------------------------------------------------------------------------------
constexpr int kRankSortMax = 1024;
__device__ __forceinline__ void block_rank_topk(const uint64_t* cand, int n, int k, int kpad, uint64_t* sel) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = tid; j < kpad; j += nt) sel[j] = 0;
    __syncthreads();
    for (int i = tid; i < n; i += nt) {
        const uint64_t v = cand[i];
        if (v == 0) continue;
        int rank = 0;
#pragma unroll 4
        for (int j = 0; j < n; ++j) rank += (cand[j] > v) ? 1 : 0;
        if (rank < k) sel[rank] = v;
    }
    __syncthreads();
}

// ---------------------------------------------------------------------------------------------
// Block-level merge of R lists that are each sorted best-first (composites strictly decreasing, 0 = empty tail), held
// in shared memory as cand[R][k_in]: the global rank of an element is its position in its own list plus, for every
// other list, the number of elements there that beat it -- R - 1 binary searches instead of the R * k_in comparisons
// of the rank sort.  This is the shard merge of a sharded search (every shard's result is sorted by construction).
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void block_merge_sorted(const uint64_t* cand, int R, int k_in, int k_out, int kpad, uint64_t* sel) {
    const int tid = threadIdx.x, nt = blockDim.x;
    for (int j = tid; j < kpad; j += nt) sel[j] = 0;
    __syncthreads();
    const int n = R * k_in;
    for (int idx = tid; idx < n; idx += nt) {
        const uint64_t v = cand[idx];
        if (v == 0) continue;
        const int r = idx / k_in;
        int rank = idx - r * k_in;
        for (int o = 0; o < R && rank < k_out; ++o) {
            if (o == r) continue;
            const uint64_t* lst = cand + o * k_in;
            int lo = 0, hi = k_in;                 // first position whose element does not beat v
            while (lo < hi) {
                const int mid = (lo + hi) >> 1;
                if (lst[mid] > v) lo = mid + 1; else hi = mid;
            }
            rank += lo;
        }
        if (rank < k_out) sel[rank] = v;
    }
    __syncthreads();
}

}  // namespace sky
