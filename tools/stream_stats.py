"""Debug: how many candidates reach the sink of the streaming scorer (SKY_ST_DEBUG=16)."""
import ctypes as C, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ["SKY_ST_DEBUG"] = "16"
import bench
from sky_embeddings_b200 import _lib, synth
lib = _lib.load()
dev = torch.device("cuda:0")
n, D, k = 1_000_000, 768, 100
bank = bench.build_bank(n, D, dev, dtype=sys.argv[1] if len(sys.argv) > 1 else "bf16")
Q = int(sys.argv[2]) if len(sys.argv) > 2 else 1
t = bank.download(500_000, Q)[:, 0] + 0.1 * torch.randn((Q, D), device=dev)
out = (C.c_ulonglong * 8)()
lib.sky_debug_stream_stats(out, 1)
for it in range(3):
    s, i = bank.search(t, None, k=k, metric="cosine", path="simt")
    torch.cuda.synchronize()
    lib.sky_debug_stream_stats(out, 1)
    print("iter", it, "prefilter-passed", out[0], "exact-passed", out[1], "in first 8 blocks", out[3], "in first 2 blocks", out[4], "top idx", i[:, 0].tolist())
