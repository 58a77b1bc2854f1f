"""Kernel timing of one search configuration on a C2-shaped random bank (no parity: for knob experiments).
usage: time_search.py [--n 1000000] [--q 64] [--k 100] [--weighted] [--metric cosine] [--path tensor] [--steps 20]"""
import argparse, sys
import torch
sys.path.insert(0, ".")
from sky_embeddings_b200 import Bank

ap = argparse.ArgumentParser()
ap.add_argument("--n", type=int, default=1000000)
ap.add_argument("--d", type=int, default=768)
ap.add_argument("--q", type=int, default=64)
ap.add_argument("--k", type=int, default=100)
ap.add_argument("--weighted", action="store_true")
ap.add_argument("--metric", default="cosine")
ap.add_argument("--path", default="tensor")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--tag", default="")
a = ap.parse_args()
dev = torch.device("cuda:0")
g = torch.Generator(device=dev); g.manual_seed(5)
bank = Bank(a.n, 1, a.d, dtype="bf16", device=dev)
chunk = 1 << 17
for r0 in range(0, a.n, chunk):
    m = min(chunk, a.n - r0)
    bank.upload(torch.randn(m, 1, a.d, device=dev, generator=g), item0=r0)
bank.finalize()
t = torch.randn(a.q, a.d, device=dev, generator=g)
w = torch.rand(a.q, a.d, device=dev, generator=g) + 0.1 if a.weighted else None
if w is not None:
    w = w / w.sum(1, keepdim=True)
for _ in range(3):
    bank.search(t, w, k=a.k, metric=a.metric, path=a.path)
torch.cuda.synchronize()
bank.profile(True)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(a.steps):
    bank.search(t, w, k=a.k, metric=a.metric, path=a.path)
e1.record()
torch.cuda.synchronize()
nl, ms = bank.profile_read()
step = e0.elapsed_time(e1) / a.steps
kern = ms / max(nl, 1) * (nl / a.steps)
gb = a.n * a.d * 2 / 1e9
print(f"{a.tag:16s} n={a.n} q={a.q} weighted={a.weighted} step={step:.4f} ms  scoring kernels/step={nl / a.steps:.0f}  kernel={kern:.4f} ms  "
      f"-> {gb / kern * 1e3:.0f} GB/s ({gb / kern * 1e3 / 6540.8:.3f} of HBM peak)")
bank.close()
