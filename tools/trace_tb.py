"""Debug: per-visit timeline of the batched tensor kernel's epilogue (SKY_TB_DEBUG=32), last phase of a search."""
import ctypes as C, os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SKY_TB_DEBUG"] = os.environ.get("SKY_TB_DEBUG", "32")
import bench
from sky_embeddings_b200 import _lib
lib = _lib.load()
dev = torch.device("cuda:0")
n, D, Q, k = int(sys.argv[1]) if len(sys.argv) > 1 else 2_000_000, 768, 4096, 100
bank = bench.build_bank(n, D, dev)
t = bank.download(1000, Q)[:, 0] + 0.1 * torch.randn((Q, D), device=dev)
for it in range(2):
    s, i = bank.search(t, None, k=k, metric="MSE", path="batch")
torch.cuda.synchronize()
N = 4096
buf = (C.c_longlong * (N * 20))()
lib.sky_debug_tb_trace(buf, N * 20)
a = np.frombuffer(buf, dtype=np.int64).reshape(N, 4, 5)
a = a[a[:, 0, 0] > 0]
print("visits traced", len(a))
for e in range(4):
    b = a[:, e]
    per = np.diff(b[:, 0])
    print("warp e=%d: period med %.0f | wait-acc med %.0f mean %.0f | chunks med %.0f mean %.0f | tail med %.0f mean %.0f | ins %.3f" % (
        e, np.median(per), np.median(b[:, 1] - b[:, 0]), (b[:, 1] - b[:, 0]).mean(), np.median(b[:, 2] - b[:, 1]), (b[:, 2] - b[:, 1]).mean(),
        np.median(b[:, 3] - b[:, 2]), (b[:, 3] - b[:, 2]).mean(), b[:, 4].mean()))
t0 = a[:, :, 0].min()
for v in range(min(100, len(a) - 6), min(106, len(a))):
    print("visit", v, " ".join("[e%d s=%d r=%d c=%d t=%d]" % (e, a[v, e, 0] - t0, a[v, e, 1] - a[v, e, 0], a[v, e, 2] - a[v, e, 1], a[v, e, 3] - a[v, e, 2]) for e in range(4)))
