"""Dump the timeline of the CTA-pair weighted kernel (experiment build, SKY_TW_DEBUG=32) for one C2-sized search."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sky_embeddings_b200 import Bank, _lib
dev = torch.device("cuda:0")
n, D, Q, k = 1_000_000, 768, 64, 100
g = torch.Generator(device=dev); g.manual_seed(5)
bank = Bank(n, 1, D, "bf16", dev)
for r0 in range(0, n, 1 << 17):
    m = min(1 << 17, n - r0)
    bank.upload(torch.randn(m, 1, D, device=dev, generator=g), item0=r0)
bank.finalize()
t = torch.randn(Q, D, device=dev, generator=g)
w = torch.rand(Q, D, device=dev, generator=g) + 0.1
w = w / w.sum(1, keepdim=True)
for _ in range(3):
    bank.search(t, w, k=k, metric="cosine", path="tensor")
torch.cuda.synchronize()
lib = _lib.load()
L, R = 1024, 13
buf = (C.c_ulonglong * (2 * R * L))()
lib.sky_debug_tw_trace.argtypes = [C.c_void_p, C.c_int]
lib.sky_debug_tw_trace(buf, 2 * R * L)
a = np.array(buf[:], dtype=np.int64).reshape(2, R, L)
t0 = a[0, 0, 0]
KB = 12
print("per k-block, ns since first TMA issue (CTA 0): tma_issue sq_full | arrive w0 w1 w2 w3 | mma_wait_start mma_sq mma_commit || CTA1: tma_issue sq_full | arrive w0 w1 w2 w3")
for i in list(range(0, 40)) + list(range(300, 340)):
    print(i, *(int(a[0, r, i] - t0) for r in range(2)), '|', *(int(a[0, r, i] - t0) for r in (2, 9, 10, 11)), '|', int(a[0, 12, i] - t0), int(a[0, 3, i] - t0), int(a[0, 4, i] - t0),
          '||', *(int(a[1, r, i] - t0) for r in range(2)), '|', *(int(a[1, r, i] - t0) for r in (2, 9, 10, 11)))
print("per tile: mma_got_empty | epi_full epi_release epi_done (CTA0) || epi_full epi_release epi_done (CTA1) ; last mma commit of the tile")
for it in range(0, 40):
    print(it, int(a[0, 5, it] - t0), '|', *(int(a[0, r, it] - t0) for r in (6, 7, 8)), '||', *(int(a[1, r, it] - t0) for r in (6, 7, 8)), ';', int(a[0, 4, it * KB + KB - 1] - t0))
d = np.diff(a[0, 4, :600])
print("mean ns between MMA commits (k-blocks 24..600):", d[24:].mean())
