"""CPU, world_size 2, gloo: the host-side logic of the sharded search (shard ranges, candidate
all-gather layout, merge to the global top-k).  The per-shard scorer and the merge are stood in by
the oracle here (the CUDA kernels need a GPU; tests/test_gpu_parity.py covers them, including an
emulated 4-shard run on one device)."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import sky_oracle as O
from sky_embeddings_b200 import synth
from sky_embeddings_b200.distributed import shard_range, sharded_search


def test_shard_range_covers_everything():
    for n, world, align in [(10, 3, 1), (1_000_000, 8, 65536), (5, 8, 1), (0, 2, 1), (100, 4, 16)]:
        spans = [shard_range(n, r, world, align) for r in range(world)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        for a, b in zip(spans, spans[1:]):
            assert a[1] == b[0]
        assert all(lo % align == 0 or lo == n for lo, _ in spans)


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, D, Q, k, metric, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    z = synth.latents(n, 1, D, stream=201).astype(np.float64)
    t = synth.latents(Q, 1, D, stream=202)[:, 0].astype(np.float64)
    lo, hi = shard_range(n, rank, world)

    def local_search():
        s, i = O.search(t, None, z[lo:hi], k, metric, "min")
        i = np.where(i >= 0, i + lo, -1)
        return torch.from_numpy(s.astype(np.float32)), torch.from_numpy(i)

    def merge(gs, gi, k_out, m):
        outs, outi = [], []
        for q in range(gs.shape[1]):
            s, i = O.merge_topk([gs[r, q].numpy() for r in range(gs.shape[0])],
                                [gi[r, q].numpy() for r in range(gi.shape[0])], k_out, m)
            outs.append(s)
            outi.append(i)
        return torch.from_numpy(np.stack(outs).astype(np.float32)), torch.from_numpy(np.stack(outi))

    s, i = sharded_search(local_search, k, metric, merge=merge)
    ref_s, ref_i = O.search(t, None, z, k, metric, "min")
    ok = np.array_equal(i.numpy(), ref_i) and np.allclose(s.numpy(), ref_s.astype(np.float32), rtol=0, atol=0)
    out[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_gloo_sharded_search_equals_global():
    world = 2
    port = _free_port()
    with mp.Manager() as mgr:
        out = mgr.dict()
        mp.spawn(_worker, args=(world, port, 301, 32, 5, 40, "cosine", out), nprocs=world, join=True)
        assert dict(out) == {0: True, 1: True}
        out2 = mgr.dict()
        # a shard shorter than k: padding (-1) must not leak into the merged result
        mp.spawn(_worker, args=(world, _free_port(), 50, 16, 3, 30, "MSE", out2), nprocs=world, join=True)
        assert dict(out2) == {0: True, 1: True}


def test_peer_exchange_needs_cuda():
    """The peer-memory exchange has no CPU path: constructing it on the CPU raises (the gloo tests above cover the
    collective-based host logic)."""
    import pytest
    import torch
    from sky_embeddings_b200.distributed import PeerExchange, make_exchange
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PeerExchange(4, 10, torch.device("cpu"))
    with pytest.raises(ValueError):
        make_exchange(4, 10, torch.device("cpu"), kind="smoke-signals")
