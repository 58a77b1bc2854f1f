// exchange.cu -- K7: the candidate exchange of a row-sharded search over peer memory (NVLink / NVSwitch).
//
// The reference's running top-k (utils/similarity.py:18-35) is a chunk-wise merge, so a bank sharded by rows needs
// one exchange per search: every rank's [Q, k] (score, global index) candidates to every rank, then a merge.  With
// NCCL that is an all-gather launch (tens of microseconds of launch + protocol latency for 77 KB at BASELINE config
// 2) in front of the merge kernel.  Here every rank owns a buffer that all peers map (cudaIpc handles, one-time
// setup over torch.distributed), and the exchange is two of this library's kernels on the search stream:
//   push   plain 16-byte stores of the local block into slot[parity][my_rank] of EVERY peer's buffer, a system-scope
//          fence, then one flag store per (peer, query block): flag = sequence number of this search;
//   merge  one CTA per query: spin (ld.acquire.sys) until all R flags of its query block carry this search's
//          sequence number, then merge the R best-first lists that now sit in LOCAL memory (read with ld.cg: the
//          peers' stores land in this GPU's L2, a stale L1 line from two searches ago must not be used).
// No collective launch, no host synchronisation; two slot parities make the reuse safe (a rank can only be one
// search ahead of a peer: it needs that peer's push of search s+1 to finish s+1, and the peer issues that push
// after its own merge of search s in stream order).
#include <cstring>
#include <new>

#include "bank.cuh"
#include "topk.cuh"

struct sky_exchange {
    int device = 0, rank = 0, world = 1;
    int max_Q = 0, max_k = 0;
    size_t slot_units = 0;          // int64 units of one rank's block: idx [Q*k] | scores [Q*k] f32
    size_t bytes = 0;
    unsigned char* local = nullptr; // this rank's buffer: slots[2][world][slot_units] int64 | flags[2][world][max_Q] u32
    unsigned char* peer[16] = {};   // mapped base of every rank's buffer (peer[rank] == local)
    bool opened[16] = {};
    bool ready = false;
    unsigned seq = 0;
};

namespace sky {

constexpr int kXchgThreads = 256;
constexpr int kXchgSortedMax = 20480;     // candidates of one query held in shared memory (160 KB) for the sorted merge

struct PeerPtrs { unsigned char* base[kMaxPeers]; };

struct DeviceGuardX {
    int prev = -1;
    explicit DeviceGuardX(int dev) { cudaGetDevice(&prev); cudaSetDevice(dev); }
    ~DeviceGuardX() { if (prev >= 0) cudaSetDevice(prev); }
};

#define slot_off xchg_slot_off
#define flag_off xchg_flag_off

// grid = queries; CTA q copies row q of the local (scores, idx) into every peer, then raises that query's flag there
__global__ void __launch_bounds__(kXchgThreads)
xchg_push_kernel(const float* __restrict__ scores, const int64_t* __restrict__ idx, int Q, int k, PeerPtrs peers, int world,
                 int rank, size_t slot_units, int max_Q, int parity, unsigned seq, int skip_self) {
    const int q = blockIdx.x;
    const size_t n = static_cast<size_t>(Q) * k;
    for (int p = 0; p < world; ++p) {
        if (skip_self && p == rank) continue;           // the source IS this rank's own slot
        unsigned char* slot = peers.base[p] + slot_off(slot_units, world, parity, rank);
        int64_t* di = reinterpret_cast<int64_t*>(slot) + static_cast<size_t>(q) * k;
        float* ds = reinterpret_cast<float*>(slot + n * 8) + static_cast<size_t>(q) * k;
        for (int j = threadIdx.x; j < k; j += kXchgThreads) {
            di[j] = idx[static_cast<size_t>(q) * k + j];
            ds[j] = scores[static_cast<size_t>(q) * k + j];
        }
    }
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x < world) {
        unsigned* f = reinterpret_cast<unsigned*>(peers.base[threadIdx.x] + flag_off(slot_units, world, max_Q, parity, rank)) + q;
        st_release_sys_u32(f, seq);
    }
}

// grid = queries; waits for the R pushes of query q, then merges the R best-first lists (lower index wins ties: lists
// hold increasing row ranges by rank, position order == index order)
__global__ void __launch_bounds__(kXchgThreads)
xchg_merge_kernel(const unsigned char* __restrict__ local, int world, size_t slot_units, int max_Q, int parity, unsigned seq,
                  int Q, int k, int k_out, int kpad, int largest, float* __restrict__ out_scores, int64_t* __restrict__ out_idx) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    uint64_t* sel = reinterpret_cast<uint64_t*>(smem_raw);     // [kpad]
    uint64_t* cand = sel + kpad;                               // [world * k] when it fits the rank sort
    __shared__ uint32_t hist[256];
    __shared__ uint32_t scratch[4];
    const int q = blockIdx.x, tid = threadIdx.x;
    if (tid < world) {
        const unsigned* f = reinterpret_cast<const unsigned*>(local + flag_off(slot_units, world, max_Q, parity, tid)) + q;
        // sequence numbers only grow; a peer that is one search ahead has already overwritten nothing of this parity.
        // A peer that never arrives (crashed rank, mismatched call counts) must not hang the GPU for ever: after ~10 s
        // the kernel traps, which surfaces as a CUDA error on the next call of this process.
        const long long t_end = clock64() + 20000000000ll;
        while (static_cast<int>(ld_acquire_sys_u32(f) - seq) < 0) {
            __nanosleep(64);
            if (clock64() > t_end) __trap();
        }
    }
    __syncthreads();
    const size_t nq = static_cast<size_t>(Q) * k;
    const int n = world * k;
    auto src_i = [&](int j) -> const int64_t* {
        const int r = j / k, i = j - r * k;
        return reinterpret_cast<const int64_t*>(local + slot_off(slot_units, world, parity, r)) + static_cast<size_t>(q) * k + i;
    };
    auto src_s = [&](int j) -> const float* {
        const int r = j / k, i = j - r * k;
        return reinterpret_cast<const float*>(local + slot_off(slot_units, world, parity, r) + nq * 8) + static_cast<size_t>(q) * k + i;
    };
    auto fetch = [&](int j) -> uint64_t {
        if (__ldcg(reinterpret_cast<const long long*>(src_i(j))) < 0) return 0ull;
        return make_composite(score_to_key(__ldcg(src_s(j)), largest != 0), static_cast<uint32_t>(j));
    };
    if (n <= kXchgSortedMax) {
        // every shard's list is sorted best-first (lower index first among equal scores, and shards hold increasing row
        // ranges): rank by binary search across the lists
        for (int j = tid; j < n; j += kXchgThreads) cand[j] = fetch(j);
        __syncthreads();
        block_merge_sorted(cand, world, k, k_out, kpad, sel);
    } else {
        block_select_sort(fetch, n, k_out, kpad, sel, hist, scratch);
    }
    __syncthreads();
    for (int j = tid; j < k_out; j += kXchgThreads) {
        const uint64_t c = sel[j];
        float* so = out_scores + static_cast<size_t>(q) * k_out + j;
        int64_t* io = out_idx + static_cast<size_t>(q) * k_out + j;
        if (c == 0) {
            *so = largest ? -INFINITY : INFINITY;
            *io = -1;
        } else {
            const int j0 = static_cast<int>(composite_idx(c));
            *so = __ldcg(src_s(j0));
            *io = static_cast<int64_t>(__ldcg(reinterpret_cast<const long long*>(src_i(j0))));
        }
    }
}

static int next_pow2_x(int v) {
    int p = 32;
    while (p < v) p <<= 1;
    return p;
}


// Start one exchange: next sequence number / slot parity, and where every peer keeps this rank's block.
int xchg_begin(sky_exchange* x, int Q, int k, XchgTarget* xt) {
    if (!x) return set_error(SKY_ERR_ARG, "exchange is NULL");
    if (!x->ready) return set_error(SKY_ERR_STATE, "exchange is not connected: call sky_exchange_open first");
    if (Q < 1 || Q > x->max_Q || k < 1 || k > x->max_k)
        return set_error(SKY_ERR_ARG, "Q=%d k=%d exceed the exchange capacity (%d, %d)", Q, k, x->max_Q, x->max_k);
    const unsigned seq = ++x->seq;
    for (int r = 0; r < kMaxPeers; ++r) xt->base[r] = r < x->world ? x->peer[r] : nullptr;
    xt->world = x->world; xt->rank = x->rank; xt->slot_units = x->slot_units; xt->max_Q = x->max_Q;
    xt->parity = static_cast<int>(seq & 1u); xt->seq = seq; xt->fused = false;
    return SKY_OK;
}

// this rank's own slot of the current exchange (a search can deliver its result there and push it afterwards)
void xchg_local_slot(const XchgTarget& xt, int Q, int k, float** scores, int64_t** idx) {
    unsigned char* slot = xt.base[xt.rank] + xchg_slot_off(xt.slot_units, xt.world, xt.parity, xt.rank);
    *idx = reinterpret_cast<int64_t*>(slot);
    *scores = reinterpret_cast<float*>(slot + static_cast<size_t>(Q) * k * 8);
}

int launch_xchg_push(const XchgTarget& xt, const float* scores, const int64_t* idx, int Q, int k, int skip_self, int device, cudaStream_t st) {
    DeviceGuardX g(device);
    PeerPtrs pp;
    for (int r = 0; r < kMaxPeers; ++r) pp.base[r] = xt.base[r];
    xchg_push_kernel<<<Q, kXchgThreads, 0, st>>>(scores, idx, Q, k, pp, xt.world, xt.rank, xt.slot_units, xt.max_Q, xt.parity, xt.seq, skip_self);
    SKY_LAUNCH_CHECK("xchg_push_kernel");
    return SKY_OK;
}

int launch_xchg_merge(const XchgTarget& xt, int Q, int k, int k_out, int metric, float* out_scores, int64_t* out_idx, int device, cudaStream_t st) {
    if (k_out < 1 || k_out > 4096) return set_error(SKY_ERR_ARG, "k_out=%d out of range", k_out);
    DeviceGuardX g(device);
    const int kpad = next_pow2_x(k_out);
    const int n = xt.world * k;
    const size_t smem = static_cast<size_t>(kpad + (n <= kXchgSortedMax ? n : 0)) * sizeof(uint64_t);
    SKY_CUDA(cudaFuncSetAttribute(xchg_merge_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem)));
    xchg_merge_kernel<<<Q, kXchgThreads, smem, st>>>(xt.base[xt.rank], xt.world, xt.slot_units, xt.max_Q, xt.parity, xt.seq, Q, k, k_out, kpad,
                                                    metric_largest(metric) ? 1 : 0, out_scores, out_idx);
    SKY_LAUNCH_CHECK("xchg_merge_kernel");
    return SKY_OK;
}

int xchg_device(const sky_exchange* x) { return x->device; }

}  // namespace sky

using namespace sky;

extern "C" {

int sky_exchange_create(sky_exchange_t** out, int device, int rank, int world, int max_Q, int max_k) {
    if (!out) return set_error(SKY_ERR_ARG, "exchange out-pointer is NULL");
    *out = nullptr;
    if (world < 1 || world > kMaxPeers || rank < 0 || rank >= world) return set_error(SKY_ERR_ARG, "bad rank %d / world %d (at most %d ranks)", rank, world, kMaxPeers);
    if (max_Q < 1 || max_k < 1 || max_k > 4096) return set_error(SKY_ERR_ARG, "bad exchange capacity Q=%d k=%d", max_Q, max_k);
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || device < 0 || device >= ndev)
        return set_error(SKY_ERR_CUDA, "no CUDA device %d: the exchange has no CPU fallback", device);
    int prev = -1;
    cudaGetDevice(&prev);
    if (cudaSetDevice(device) != cudaSuccess) return set_error(SKY_ERR_CUDA, "cudaSetDevice(%d) failed", device);
    sky_exchange* x = new (std::nothrow) sky_exchange();
    if (!x) { cudaSetDevice(prev); return set_error(SKY_ERR_NOMEM, "host allocation failed"); }
    x->device = device; x->rank = rank; x->world = world; x->max_Q = max_Q; x->max_k = max_k;
    const size_t n = static_cast<size_t>(max_Q) * max_k;
    x->slot_units = n + (n + 1) / 2;
    x->bytes = 2 * static_cast<size_t>(world) * x->slot_units * 8 + 2 * static_cast<size_t>(world) * max_Q * 4;
    cudaError_t e = cudaMalloc(reinterpret_cast<void**>(&x->local), x->bytes);
    if (e == cudaSuccess) e = cudaMemset(x->local, 0, x->bytes);
    if (e == cudaSuccess) e = cudaDeviceSynchronize();
    cudaSetDevice(prev);
    if (e != cudaSuccess) {
        if (x->local) cudaFree(x->local);
        delete x;
        return set_error(SKY_ERR_NOMEM, "exchange buffer of %zu B failed: %s", n, cudaGetErrorString(e));
    }
    x->peer[rank] = x->local;
    x->opened[rank] = false;
    x->ready = (world == 1);
    *out = x;
    return SKY_OK;
}

int sky_exchange_handle_bytes(void) { return static_cast<int>(sizeof(cudaIpcMemHandle_t)); }

int sky_exchange_handle(sky_exchange_t* x, void* h_handle) {
    if (!x || !h_handle) return set_error(SKY_ERR_ARG, "NULL argument");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(x->device);
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, x->local);
    cudaSetDevice(prev);
    if (e != cudaSuccess) return set_error(SKY_ERR_CUDA, "cudaIpcGetMemHandle failed: %s", cudaGetErrorString(e));
    memcpy(h_handle, &h, sizeof(h));
    return SKY_OK;
}

int sky_exchange_open(sky_exchange_t* x, const void* h_handles) {
    if (!x || !h_handles) return set_error(SKY_ERR_ARG, "NULL argument");
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(x->device);
    const unsigned char* hb = reinterpret_cast<const unsigned char*>(h_handles);
    for (int r = 0; r < x->world; ++r) {
        if (r == x->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hb + static_cast<size_t>(r) * sizeof(h), sizeof(h));
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaSetDevice(prev);
            return set_error(SKY_ERR_CUDA, "cudaIpcOpenMemHandle for rank %d failed: %s (no peer access between the GPUs?)", r, cudaGetErrorString(e));
        }
        x->peer[r] = reinterpret_cast<unsigned char*>(p);
        x->opened[r] = true;
    }
    cudaSetDevice(prev);
    x->ready = true;
    return SKY_OK;
}

/* ranks living in ONE process (tests; a single process driving several devices with peer access enabled): the peers'
 * buffers are passed as plain device pointers (sky_exchange_local_ptr of the other handles). */
int sky_exchange_open_local(sky_exchange_t* x, void* const* peer_ptrs) {
    if (!x || !peer_ptrs) return set_error(SKY_ERR_ARG, "NULL argument");
    for (int r = 0; r < x->world; ++r)
        if (r != x->rank) {
            if (!peer_ptrs[r]) return set_error(SKY_ERR_ARG, "peer pointer %d is NULL", r);
            x->peer[r] = reinterpret_cast<unsigned char*>(peer_ptrs[r]);
        }
    x->ready = true;
    return SKY_OK;
}

void* sky_exchange_local_ptr(sky_exchange_t* x) { return x ? x->local : nullptr; }

int sky_exchange_destroy(sky_exchange_t* x) {
    if (!x) return SKY_OK;
    int prev = -1;
    cudaGetDevice(&prev);
    cudaSetDevice(x->device);
    cudaDeviceSynchronize();
    for (int r = 0; r < x->world; ++r)
        if (x->opened[r] && x->peer[r]) cudaIpcCloseMemHandle(x->peer[r]);
    if (x->local) cudaFree(x->local);
    cudaSetDevice(prev);
    delete x;
    return SKY_OK;
}

int sky_exchange_merge(sky_exchange_t* x, const float* scores, const int64_t* idx, int Q, int k, int k_out, int metric,
                       float* out_scores, int64_t* out_idx, void* stream) {
    if (!x || !scores || !idx || !out_scores || !out_idx) return set_error(SKY_ERR_ARG, "NULL argument");
    if (metric != SKY_COSINE && metric != SKY_MSE && metric != SKY_MAE) return set_error(SKY_ERR_ARG, "unknown metric %d", metric);
    XchgTarget xt;
    int rc = xchg_begin(x, Q, k, &xt);
    if (rc) return rc;
    cudaStream_t st = static_cast<cudaStream_t>(stream);
    rc = launch_xchg_push(xt, scores, idx, Q, k, /*skip_self=*/0, x->device, st);
    if (rc) return rc;
    return launch_xchg_merge(xt, Q, k, k_out, metric, out_scores, out_idx, x->device, st);
}

}  // extern "C"
