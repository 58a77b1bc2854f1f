"""Debug (experiments build): timeline of ONE phase of the batched kernel's epilogue, CTA 0, all four epilogue warps.
usage: trace_tb_phase.py <phase> [n] [Q] [k] [metric]   stamps per visit: start, accumulator ready, chunks done, write-back done"""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
phase = int(sys.argv[1])
n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_250_000
Q = int(sys.argv[3]) if len(sys.argv) > 3 else 4096
k = int(sys.argv[4]) if len(sys.argv) > 4 else 100
metric = sys.argv[5] if len(sys.argv) > 5 else "MSE"
os.environ["SKY_TB_DEBUG"] = "32"
os.environ["SKY_TB_TRACE_PHASE"] = str(phase)
from sky_embeddings_b200 import _lib, Bank
lib = _lib.load()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(11)
D = 768
bank = Bank(n, 1, D, dtype="bf16", device=dev)
for r0 in range(0, n, 1 << 18):
    m = min(1 << 18, n - r0)
    bank.upload(torch.randn(m, 1, D, device=dev, generator=g), item0=r0)
bank.finalize()
t = torch.randn(Q, D, device=dev, generator=g)
for it in range(3):
    bank.search(t, None, k=k, metric=metric, path="batch")
torch.cuda.synchronize()
N = 4096
buf = (C.c_longlong * (N * 20))()
lib.sky_debug_tb_trace(buf, N * 20)
a = np.frombuffer(buf, dtype=np.int64).reshape(N, 4, 5)
G = (Q + 255) // 256
print(f"phase {phase}: n={n} Q={Q} k={k} {metric}; clock64 cycles, CTA 0")
t0 = a[0, :, 0].min()
nv = min(N, 4 * G)
for v in range(nv):
    if a[v, 0, 0] <= 0:
        break
    print("visit %3d " % v + " ".join("[e%d start=%7d acc+%6d chunks+%6d tail+%6d ins=%d]" % (
        e, a[v, e, 0] - t0, a[v, e, 1] - a[v, e, 0], a[v, e, 2] - a[v, e, 1], a[v, e, 3] - a[v, e, 2], a[v, e, 4]) for e in (0, 3)))
