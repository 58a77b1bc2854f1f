"""Dump the tensor-kernel timeline of CTA 0 (SKY_TC_DEBUG=8|...) for one C2-sized search."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from sky_embeddings_b200 import Bank, _lib, synth
dev = torch.device("cuda:0")
n, D, Q, k = 400_000, 768, 64, 100
bank = Bank(n, 1, D, "bf16", dev)
bank.fit_norm(synth.device_bank_chunk(0, 512, D, dev))
for c in range((n + synth.CHUNK_ROWS - 1) // synth.CHUNK_ROWS):
    rows = min(synth.CHUNK_ROWS, n - c * synth.CHUNK_ROWS)
    bank.upload(synth.device_bank_chunk(c, rows, D, dev), c * synth.CHUNK_ROWS)
bank.finalize()
t = bank.download(0, Q)[:, 0] + 0.1
for _ in range(3):
    bank.search(t, None, k=k, metric="cosine", path="tensor")
torch.cuda.synchronize()
lib = _lib.load()
L = 1024
buf = (C.c_ulonglong * (6 * L))()
lib.sky_debug_trace.argtypes = [C.c_void_p, C.c_int]
lib.sky_debug_trace(buf, 6 * L)
a = np.array(buf[:], dtype=np.int64).reshape(6, L)
t0 = a[0, 0]
names = ["prod_issue", "mma_full", "mma_commit", "epi_full", "epi_done"]
nk = 12 * 6
print("k-block events (cycles since first producer issue), first", nk)
for i in range(nk):
    print(i, *(int(a[r, i] - t0) for r in range(3)))
print("tile events")
for i in range(22):
    print(i, int(a[3, i] - t0), int(a[4, i] - t0), 'epilogue cycles', int(a[4, i] - a[3, i]), 'slow groups (of 16, x3 searches)', int(a[5, i]))
eb = (C.c_ulonglong * (64 * 12))()
lib.sky_debug_epi.argtypes = [C.c_void_p]
lib.sky_debug_epi(eb)
ep = np.array(eb[:], dtype=np.int64).reshape(64, 12)
print("epilogue breakdown per tile (cycles): ld0 fast0 slow0 | ld1 fast1 slow1 | bar1 prune bar2")
for i in range(22):
    r = ep[i]
    print(i, r[1]-r[0], r[2]-r[1], r[3]-r[2], '|', r[5]-r[4], r[6]-r[5], r[7]-r[6], '|', r[8]-r[7], r[9]-r[8], r[10]-r[9])
d = np.diff(a[1, :200])
print("mean cycles between full-barrier completions (k-blocks 12..200):", d[12:].mean())
