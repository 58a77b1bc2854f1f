"""INTEGRATION.md section 3 shows the ctypes stub a maintainer of the reference would add.  This test EXECUTES that
code block as printed (only the library path is made absolute) and checks its result against the CPU oracle."""
import os
import re

import numpy as np
import pytest
import torch

from oracle import sky_oracle as O

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_integration_md_ctypes_stub_runs_and_matches_the_oracle():
    from sky_embeddings_b200 import _lib, synth
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    sec = text[text.index("## 3."):text.index("## 4.")]
    code = re.search(r"```python\n(.*?)```", sec, re.S).group(1)
    assert 'C.CDLL("libskysearch.so")' in code
    code = code.replace('C.CDLL("libskysearch.so")', f'C.CDLL({_lib.LIB_PATH!r})')
    ns = {}
    exec(compile(code, "INTEGRATION.md#3", "exec"), ns)
    dev = torch.device("cuda:0")
    n, tokens, D, k, bs = 5000, 5, 768, 25, 64
    lat = synth.latents(n, tokens, D, stream=611)
    tgt = synth.target_group(lat, [40, 3210], copies=8, noise=0.3, stream=612)
    tsel = torch.from_numpy(O.token_select(tgt, 1, False, True)).reshape(-1, D).contiguous().to(dev)
    sc, ix = ns["search_latents"](torch.from_numpy(lat).to(dev), tsel, bs, k, ns["SKY_COSINE"], ns["SKY_MIN"], ns["SKY_TOK_MAXPOOL"])
    torch.cuda.synchronize()
    ref_s, ref_i, *_ = O.simsearch(tgt, lat, bs, k, metric="cosine", combine="min", use_weights=True, max_pool=True)
    # the stub stores the bank in bf16: compare against the oracle at the bf16 tolerance, anchors first
    assert set(ix[:2].cpu().tolist()) == {40, 3210}
    got = dict(zip(ix.cpu().tolist(), sc.cpu().tolist()))
    for i_, s_ in zip(ref_i.tolist()[:10], ref_s.tolist()[:10]):
        assert i_ in got and abs(got[i_] - s_) <= 2e-3 * max(abs(s_), 1e-3), (i_, s_, got.get(i_))
