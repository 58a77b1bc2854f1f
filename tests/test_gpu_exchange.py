"""Peer-memory candidate exchange (csrc/exchange.cu) on ONE GPU: two logical ranks in one process map each other's
buffers (sky_exchange_open_local) and run their push + flag-waiting merge on two streams.  The merged top-k must be
bit-identical to the merge of the concatenated candidates (utils/similarity.py:18-35 semantics: cat + sort + [:k]),
on every rank, search after search (both slot parities, growing sequence numbers).  The multi-process cudaIpc route
is covered by bench.py --gpus N (merge_bit_exact gate) under torchrun."""
import ctypes as C

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _mk(world, Q, k, dev):
    from sky_embeddings_b200 import _lib as L
    lib = L.load()
    hs = []
    for r in range(world):
        h = C.c_void_p()
        L.check(lib.sky_exchange_create(C.byref(h), dev.index or 0, r, world, Q, k))
        hs.append(h)
    ptrs = (C.c_void_p * world)(*[lib.sky_exchange_local_ptr(h) for h in hs])
    for h in hs:
        L.check(lib.sky_exchange_open_local(h, ptrs))
    return lib, hs


@pytest.mark.parametrize("metric", ["cosine", "MSE"])
@pytest.mark.parametrize("world,Q,k", [(2, 64, 100), (4, 7, 1000), (8, 3, 10)])
def test_peer_exchange_matches_concatenate_and_sort(metric, world, Q, k):
    from sky_embeddings_b200 import _lib as L
    dev = torch.device("cuda:0")
    lib, hs = _mk(world, Q, k, dev)
    streams = [torch.cuda.Stream(dev) for _ in range(world)]
    largest = metric == "cosine"
    g = torch.Generator(device="cpu").manual_seed(1234 + world)
    for it in range(5):                                   # both parities, sequence numbers 1..5
        sc, ix = [], []
        for r in range(world):
            s = torch.randn(Q, k, generator=g)
            s[:, ::7] = torch.round(s[:, ::7] * 4) / 4 + 0.0  # ties across ranks: the lower index must win (+0.0: no -0.0)
            s, _ = torch.sort(s, dim=1, descending=largest)
            i = (torch.arange(k)[None, :] * world + r + 1000 * r * 0).repeat(Q, 1).to(torch.int64) + r * 10_000_000
            if r == world - 1 and it == 2:                  # a short shard: empty slots (idx -1, +-inf) at the tail
                s[:, k // 2:] = float("-inf") if largest else float("inf")
                i[:, k // 2:] = -1
            sc.append(s.to(dev)); ix.append(i.to(dev))
        outs = [(torch.empty(Q, k, device=dev), torch.empty(Q, k, dtype=torch.int64, device=dev)) for _ in range(world)]
        torch.cuda.synchronize()
        for r in range(world):
            with torch.cuda.stream(streams[r]):
                L.check(lib.sky_exchange_merge(hs[r], sc[r].data_ptr(), ix[r].data_ptr(), Q, k, k, L.METRICS[metric],
                                               outs[r][0].data_ptr(), outs[r][1].data_ptr(), streams[r].cuda_stream))
        torch.cuda.synchronize()
        cat_s, cat_i = torch.cat(sc, 1), torch.cat(ix, 1)
        # stable sort of the rank-major concatenation: ties resolve to the lower rank = lower global index
        key = cat_s.clone()
        key[cat_i < 0] = float("-inf") if largest else float("inf")
        order = torch.sort(key, dim=1, descending=largest, stable=True)[1][:, :k]
        want_s, want_i = torch.gather(key, 1, order), torch.gather(cat_i, 1, order)
        for r in range(world):
            got_s, got_i = outs[r]
            assert torch.equal(got_i, want_i), f"iteration {it} rank {r}: indices differ"
            assert torch.equal(got_s.view(torch.int32), want_s.view(torch.int32)), f"iteration {it} rank {r}: scores differ"
    for h in hs:
        lib.sky_exchange_destroy(h)


def test_exchange_rejects_bad_use():
    from sky_embeddings_b200 import _lib as L
    lib = L.load()
    h = C.c_void_p()
    assert lib.sky_exchange_create(C.byref(h), 0, 3, 2, 8, 8) != 0          # rank outside the world
    L.check(lib.sky_exchange_create(C.byref(h), 0, 0, 2, 8, 8))
    s = torch.zeros(8, 8, device="cuda:0"); i = torch.zeros(8, 8, dtype=torch.int64, device="cuda:0")
    rc = lib.sky_exchange_merge(h, s.data_ptr(), i.data_ptr(), 8, 8, 8, 0, s.data_ptr(), i.data_ptr(), None)
    assert rc == -3 and b"not connected" in lib.sky_last_error()             # SKY_ERR_STATE before open
    lib.sky_exchange_destroy(h)


@pytest.mark.parametrize("metric,weighted,Q,k,path", [("cosine", False, 64, 100, "auto"), ("MSE", True, 16, 10, "auto"),
                                                      ("cosine", True, 3, 20, "auto"), ("MSE", False, 32, 50, "batch")])
def test_fused_sharded_search_equals_unsharded(metric, weighted, Q, k, path):
    """sky_search_sharded on two shards of one bank (two logical ranks on one GPU, two streams): the shard merge kernel
    delivers into both exchange buffers, the flag-waiting merge returns the global top-k -- identical, bit for bit, to
    the search over the unsharded bank; K1 / K2 / K2w2 deliver fused, K2b (path="batch") through the push kernel.
    Q stays small here: with both ranks on ONE GPU the first rank's waiting merge CTAs share the SMs with the second
    rank's scorer, and hundreds of them would leave it no room (one process per GPU has no such coupling)."""
    from sky_embeddings_b200 import Bank, synth
    from sky_embeddings_b200 import _lib as L
    dev = torch.device("cuda:0")
    n, D = 6000, 256
    lat = torch.from_numpy(synth.latents(n, 1, D, stream=901)).to(dev)
    full = Bank.from_latents(lat, norm_rows=64, dtype="bf16")
    mu, sigma = full.norm()
    cut = 3328                                            # shard boundary (a multiple of the 128-row tile is not required)
    shards = []
    for lo, hi in ((0, cut), (cut, n)):
        b = Bank(hi - lo, 1, D, "bf16", dev)
        b.set_norm(mu, sigma)
        b.upload(lat[lo:hi], 0)
        b.finalize()
        shards.append((b, lo))
    g = torch.Generator(device="cpu").manual_seed(7)
    z = full.download()[:, 0]
    t = (z[torch.randint(0, n, (Q,), generator=g)] + 0.1 * torch.randn(Q, D, generator=g).to(dev)).contiguous()
    w = None
    if weighted:
        w = (torch.rand(Q, D, generator=g) + 0.25).to(dev)
        w = (w / w.sum(1, keepdim=True)).contiguous()
    want_s, want_i = full.search(t, w, k=k, metric=metric, path=path)
    lib, hs = _mk(2, Q, k, dev)
    streams = [torch.cuda.Stream(dev) for _ in range(2)]
    # Two ranks in ONE process share one host thread and one CUDA context: rank 0's merge kernel spins on the GPU until
    # rank 1's kernels have run, so nothing on rank 1's launch path may synchronise the device (a first-use workspace
    # allocation or module load would).  One ordinary search per shard warms all of that up; outputs are preallocated.
    # (One process per GPU -- the real deployment -- has no such coupling.)
    outs = []
    for b, lo in shards:
        b.search(t, w, k=k, metric=metric, idx_offset=lo, path=path)
        outs.append((torch.empty(Q, k, device=dev), torch.empty(Q, k, dtype=torch.int64, device=dev)))
    for it in range(3):
        torch.cuda.synchronize()
        for r, (b, lo) in enumerate(shards):
            with torch.cuda.stream(streams[r]):
                b.search_sharded(hs[r], t, w, k=k, metric=metric, path=path, idx_offset=lo, out_scores=outs[r][0], out_idx=outs[r][1])
        torch.cuda.synchronize()
        for r in range(2):
            assert torch.equal(outs[r][1], want_i), f"iteration {it} rank {r}: indices differ from the unsharded search"
            assert torch.equal(outs[r][0].view(torch.int32), want_s.view(torch.int32)), f"iteration {it} rank {r}: scores differ"
    for h in hs:
        lib.sky_exchange_destroy(h)
    for b, _ in shards:
        b.close()
    full.close()
