"""Which error bounds the bf16 tensor paths actually meet on a D = 768 bank ("say which bound a D = 768
weighted-MSE top-100 actually meets").  For each path: the worst error of a returned top-100 score against the fp64 oracle on
the SAME stored (bf16-rounded) bank, relative to (a) that score itself, (b) a typical (median) score of the bank."""
import sys
import numpy as np
import torch
sys.path.insert(0, ".")
from oracle import sky_oracle as O
from sky_embeddings_b200 import Bank, synth

dev = torch.device("cuda:0")
n, D, Q, k = 200_000, 768, 64, 100
lat = synth.latents(n, 1, D, stream=777)
bank = Bank.from_latents(torch.from_numpy(lat).to(dev), norm_rows=64, dtype="bf16")
z = bank.download().cpu().numpy()[:, 0].astype(np.float64)
ts, ws = [], []
for q in range(Q):
    grp = synth.target_group(z[:, None, :].astype(np.float32), [(37 * q + 3) % n, (91 * q + 5) % n], copies=6, noise=0.4, stream=800 + q)
    tq, wq = O.target_features(grp)
    ts.append(tq); ws.append(wq)
t = np.stack(ts).astype(np.float32); w = np.stack(ws).astype(np.float32)
for metric in ("cosine", "MSE"):
    for weighted in (False, True):
        sc, ix = bank.search(torch.from_numpy(t).to(dev), torch.from_numpy(w).to(dev) if weighted else None, k=k, metric=metric, path="tensor")
        sc, ix = sc.cpu().numpy().astype(np.float64), ix.cpu().numpy()
        worst_own = worst_typ = 0.0
        same = 0
        for q in range(0, Q, 8):
            wq = w[q].astype(np.float64) if weighted else np.ones(D)
            allv = O.item_scores(t[q].astype(np.float64), wq, z[:, None], metric, "min")
            ref = allv[ix[q]]
            err = np.abs(sc[q] - ref)
            worst_own = max(worst_own, float((err / np.maximum(np.abs(ref), 1e-30)).max()))
            worst_typ = max(worst_typ, float(err.max() / np.median(np.abs(allv))))
            rs, ri = O.topk(allv, k, metric)
            same += int(np.array_equal(ri, ix[q]))
        print(f"{metric:6s} weighted={weighted!s:5s}: worst |err| / |own score| = {worst_own:.2e}   worst |err| / median |score| = {worst_typ:.2e}   "
              f"identical top-{k} index lists: {same}/8 sampled queries")
bank.close()
