"""CPU: the torch port used as the timed CPU baseline reproduces the reference's outputs."""
import numpy as np
import torch

from oracle import ref_port as R
from oracle import sky_oracle as O
from tests import golden_inputs as G


def test_port_matches_reference_fixtures():
    g = G.load("simsearch_small")
    bank, tgt, bs, k = G.simsearch_inputs(g)
    for name in g["names"]:
        kw = G.parse_simsearch_name(str(name))
        sc, ix = R.search_loop(torch.from_numpy(tgt), torch.from_numpy(bank), bs, k, **kw)
        ok, msg = O.check_topk_parity(sc.numpy(), ix.numpy(), g[f"scores.{name}"], g[f"idx.{name}"], 1e-5)
        assert ok, f"{name}: {msg}"


def test_multi_query_loop_matches_oracle():
    from sky_embeddings_b200 import synth
    z = synth.latents(500, 1, 64, stream=91)
    q = synth.latents(3, 1, 64, stream=92)[:, 0]
    for metric in ("cosine", "MSE"):
        got = R.multi_query_loop(torch.from_numpy(q), None, torch.from_numpy(z), 128, 7, metric)
        ref_s, ref_i = O.search(q.astype(np.float64), None, z.astype(np.float64), 7, metric, "min")
        for j, (s, i) in enumerate(got):
            ok, msg = O.check_topk_parity(s.numpy(), i.numpy(), ref_s[j], ref_i[j], 1e-5)
            assert ok, msg
