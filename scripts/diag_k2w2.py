"""GPU diagnostic for the CTA-pair weighted kernel (K2w2): which query column lands where, and how fast."""
import sys, time
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import sky_oracle as O
from sky_embeddings_b200 import Bank, synth

dev = torch.device("cuda:0")
n, Q, k, D = 20000, 64, 10, 768
lat = synth.latents(n, 1, D, stream=301)
bank = Bank.from_latents(torch.from_numpy(lat).to(dev), norm_rows=64, dtype="bf16")
z = bank.download().cpu().numpy()[:, 0]
ts, ws = [], []
for q in range(Q):
    grp = synth.target_group(z[:, None, :].astype(np.float32), [(37 * q + 3) % n, (91 * q + 5) % n], copies=6, noise=0.4, stream=310 + q)
    tq, wq = O.target_features(grp)
    ts.append(tq); ws.append(wq)
t = np.stack(ts).astype(np.float32); w = np.stack(ws).astype(np.float32)
for metric in ("cosine", "MSE"):
    sc, ix = bank.search(torch.from_numpy(t).to(dev), torch.from_numpy(w).to(dev), k=k, metric=metric, path="tensor")
    torch.cuda.synchronize()
    sc, ix = sc.cpu().numpy(), ix.cpu().numpy()
    ref_s, ref_i = O.search(t.astype(np.float64), w.astype(np.float64), z[:, None].astype(np.float64), k, metric, "min")
    top1 = {int(ref_i[q, 0]): q for q in range(Q)}
    where = [top1.get(int(ix[q, 0]), -1) for q in range(Q)]
    exact = sum(int(np.array_equal(ix[q], ref_i[q])) for q in range(Q))
    print(metric, "queries with identical top-k:", exact, "/", Q)
    print(" gpu column -> oracle query whose top-1 it returned:", where)
    print(" q0 gpu", ix[0, :5], sc[0, :5], "ref", ref_i[0, :5], ref_s[0, :5])
bank.close()
