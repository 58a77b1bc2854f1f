"""Property tests (hypothesis): random shapes, ties, +-inf, NaN, N < k -- SURVEY.md section 4.

CPU half: the numpy oracle against the reference's own functions executed on the same random inputs
(oracle/ref_harness.load_reference: /root/reference in the build container, oracle/_ref on the GPU box).
GPU half: the CUDA engine through the C ABI against the oracle.
"""
import numpy as np
import pytest
import torch
from hypothesis import HealthCheck, given, settings
from hypothesis import strategies as st

from oracle import ref_harness as H
from oracle import sky_oracle as O

SET = dict(deadline=None, suppress_health_check=[HealthCheck.too_slow, HealthCheck.function_scoped_fixture, HealthCheck.data_too_large])


def _inputs(seed, n, L, D, T, specials):
    rng = np.random.Generator(np.random.PCG64(seed))
    bank = rng.standard_normal((n, L, D)).astype(np.float32) * (1.0 + rng.random(D, dtype=np.float32)) + 0.3
    tgt = (bank[rng.integers(0, n, T)] + 0.2 * rng.standard_normal((T, L, D)).astype(np.float32))
    if specials and n >= 8:
        bank[3] = bank[1]                              # exact tie
        bank[5, 0, 0] = np.nan                          # NaN score
        bank[6, 0, 1] = np.inf                          # inf feature -> inf / NaN score
    return bank, tgt


# ------------------------------------------------------------------------------------------------
# CPU: oracle == reference function
# ------------------------------------------------------------------------------------------------
@settings(max_examples=40, **SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 60), L=st.sampled_from([1, 2, 5]), D=st.sampled_from([3, 16, 50]),
       T=st.integers(2, 9), metric=st.sampled_from(["cosine", "MSE", "MAE"]), combine=st.sampled_from(["mean", "min", "max"]),
       use_weights=st.booleans(), specials=st.booleans(), top=st.booleans())
def test_oracle_compute_similarity_equals_reference(seed, n, L, D, T, metric, combine, use_weights, specials, top):
    ref = H.load_reference()
    if ref is None:
        pytest.skip("neither /root/reference nor oracle/_ref is present")
    bank, tgt = _inputs(seed, n, L, D, T, specials)
    n_top = max(1, L - 1) if (top and L > 1) else None
    want = ref.compute_similarity(torch.from_numpy(tgt), torch.from_numpy(bank), metric=metric, combine=combine,
                                  use_weights=use_weights, n_top_sims=n_top).numpy()
    got = O.compute_similarity(tgt, bank, metric, combine, use_weights, n_top_sims=n_top)
    fin = np.isfinite(want)
    assert np.array_equal(np.isnan(want), np.isnan(got))
    assert np.array_equal(np.isinf(want), np.isinf(got))
    if fin.any():
        scale = max(float(np.abs(want[fin]).max()), 1e-30)
        assert np.all(np.abs(got[fin] - want[fin]) <= 2e-5 * scale + 1e-6 * np.abs(want[fin]))


@settings(max_examples=40, **SET)
@given(seed=st.integers(0, 2 ** 31), nb=st.integers(0, 30), nn=st.integers(1, 30), n_save=st.integers(1, 40),
       metric=st.sampled_from(["cosine", "MSE"]), specials=st.booleans())
def test_oracle_topk_equals_reference_update_best_scores(seed, nb, nn, n_save, metric, specials):
    """update_best_scores (utils/similarity.py:18-35) == order_best_first of the concatenation: NaN = largest, +-inf kept."""
    ref = H.load_reference()
    if ref is None:
        pytest.skip("neither /root/reference nor oracle/_ref is present")
    rng = np.random.Generator(np.random.PCG64(seed))
    a = rng.standard_normal(nb).astype(np.float32)
    b = rng.standard_normal(nn).astype(np.float32)
    if specials:
        if nn > 3:
            b[0], b[1], b[2] = np.nan, np.inf, -np.inf
        if nb > 1:
            a[0] = np.nan
    cat = np.concatenate([a, b])
    ids = np.arange(cat.shape[0], dtype=np.float32)
    bs, br, bsc = ref.update_best_scores(torch.from_numpy(ids[nb:, None]), torch.from_numpy(np.stack([ids[nb:], ids[nb:]], 1)),
                                         torch.from_numpy(b), torch.from_numpy(ids[:nb, None]),
                                         torch.from_numpy(np.stack([ids[:nb], ids[:nb]], 1)), torch.from_numpy(a), n_save, metric)
    order = O.order_best_first(cat, metric)[:n_save]
    want = cat[order]
    got = bsc.numpy()
    assert np.array_equal(np.isnan(got), np.isnan(want))
    assert np.array_equal(got[~np.isnan(got)], want[~np.isnan(want)])          # scores identical, order identical up to ties
    # payload rows follow their scores (ties may permute among equal scores: the reference's argsort is unstable)
    assert np.array_equal(np.nan_to_num(cat[br[:, 0].numpy().astype(int)], nan=7e33), np.nan_to_num(got, nan=7e33))


# ------------------------------------------------------------------------------------------------
# GPU: engine == oracle
# ------------------------------------------------------------------------------------------------
@pytest.fixture(scope="module")
def dev():
    assert torch.cuda.is_available(), "GPU tests need a CUDA device"
    from sky_embeddings_b200 import _lib
    _lib.load()
    return torch.device("cuda:0")


@pytest.mark.gpu
@settings(max_examples=60, **SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 700), L=st.sampled_from([1, 1, 2, 4, 5, 64]), D=st.sampled_from([8, 48, 64, 200, 768]),
       k=st.sampled_from([1, 7, 100, 300]), metric=st.sampled_from(["cosine", "MSE", "MAE"]), combine=st.sampled_from(["mean", "min", "max"]),
       use_weights=st.booleans(), specials=st.booleans(), Q=st.sampled_from([1, 1, 3, 6]), top=st.booleans())
def test_engine_search_equals_oracle_fp32(dev, seed, n, L, D, k, metric, combine, use_weights, specials, Q, top):
    """fp32 bank, every scorer the AUTO path picks (K1 stream / generic), any L / D / k, N < k, ties, NaN, inf."""
    from sky_embeddings_b200 import Bank
    if L == 64:
        n = min(n, 60)
    bank_np, _ = _inputs(seed, n, L, D, 4, specials)
    rng = np.random.Generator(np.random.PCG64(seed + 1))
    t = (bank_np[rng.integers(0, n, Q), rng.integers(0, L, Q)] + 0.1 * rng.standard_normal((Q, D))).astype(np.float32)
    w = (rng.random((Q, D), dtype=np.float32) + 0.25) if use_weights else None
    n_top = max(1, L // 2) if (top and L > 1) else None
    bank = Bank.from_latents(torch.from_numpy(bank_np).to(dev), norm_rows=None, dtype="fp32")
    sc, ix = bank.search(torch.from_numpy(t).to(dev), None if w is None else torch.from_numpy(w).to(dev), k=k, metric=metric,
                         combine=combine, n_top_sims=n_top)
    sc, ix = sc.cpu().numpy(), ix.cpu().numpy()
    bank.close()
    z = bank_np.astype(np.float64)
    for q in range(Q):
        wq = np.ones(D) if w is None else w[q].astype(np.float64)
        allv = O.item_scores(t[q].astype(np.float64), wq, z, metric, combine, n_top)
        ref_s, ref_i = O.topk(allv, k, metric)
        # rows with a non-finite feature: the reference's own arithmetic decides between inf and NaN; both rank at the
        # ends and are compared as a group through check_topk_parity (NaN == NaN, inf == inf)
        ok, msg = O.check_topk_parity(sc[q], ix[q], ref_s, ref_i, 2e-5, all_scores=allv)
        assert ok, f"n={n} L={L} D={D} k={k} {metric}/{combine} w={use_weights} Q={Q} top={n_top} q{q}: {msg}"
        kk = min(k, n)
        assert np.all(ix[q][kk:] == -1)


@pytest.mark.gpu
@settings(max_examples=40, **SET)
@given(seed=st.integers(0, 2 ** 31), n=st.integers(1, 3000), D=st.sampled_from([64, 96, 256, 768]), k=st.sampled_from([1, 10, 100, 1000]),
       metric=st.sampled_from(["cosine", "MSE"]), Q=st.sampled_from([5, 64, 70, 200]), weighted=st.booleans(), ties=st.booleans())
def test_engine_tensor_paths_equal_oracle_bf16(dev, seed, n, D, k, metric, Q, weighted, ties):
    """bf16 bank, the tcgen05 scorers (K2 / K2w / K2b picked by AUTO): random N / D / k / Q, N < k, exact ties."""
    from sky_embeddings_b200 import Bank
    rng = np.random.Generator(np.random.PCG64(seed))
    bank_np = (rng.standard_normal((n, 1, D)).astype(np.float32) * (1.0 + rng.random(D, dtype=np.float32)))
    if ties and n > 10:
        bank_np[7] = bank_np[2]
        bank_np[9] = bank_np[2]
    bank = Bank.from_latents(torch.from_numpy(bank_np).to(dev), norm_rows=None, dtype="bf16")
    z = bank.download().cpu().numpy().astype(np.float64)       # stored (rounded) rows: both sides see the same bank
    t = (z[rng.integers(0, n, Q), 0] + 0.2 * rng.standard_normal((Q, D))).astype(np.float32)
    if ties and n > 10:
        t[0] = z[2, 0].astype(np.float32)
    w = (rng.random((Q, D), dtype=np.float32) + 0.25) if weighted else None
    if w is not None:
        w /= w.sum(1, keepdims=True)
    sc, ix = bank.search(torch.from_numpy(t).to(dev), None if w is None else torch.from_numpy(w).to(dev), k=k, metric=metric)
    sc, ix = sc.cpu().numpy(), ix.cpu().numpy()
    bank.close()
    for q in range(0, Q, max(1, Q // 6)):
        wq = np.ones(D) if w is None else w[q].astype(np.float64)
        allv = O.item_scores(t[q].astype(np.float64), wq, z, metric, "min")
        ref_s, ref_i = O.topk(allv, k, metric)
        # bf16 operand rounding of the queries (and of w, w t, z z in K2w): 1e-3 at D >= 256, looser below
        # (fewer features to average the roundings over); MSE relative to the size of the contraction terms
        rel = 1e-3 if D >= 256 else 3e-3
        # (a typical score is about twice sum w t^2 / (D sum w); the median alone degenerates when the bank holds little
        # more than the query's own near neighbour, e.g. n = 1)
        fin = allv[np.isfinite(allv)]
        tq = t[q].astype(np.float64)
        term = 2.0 * float((wq * tq * tq).sum() / (D * wq.sum()))
        scale = max(2.0 * float(np.median(fin)), term) if (metric == "MSE" and fin.size) else None
        ok, msg = O.check_topk_parity(sc[q], ix[q], ref_s, ref_i, rel, all_scores=allv, scale=scale)
        assert ok, f"n={n} D={D} k={k} {metric} Q={Q} weighted={weighted} q{q}: {msg}"
        if ties and n > 10 and q == 0 and metric == "MSE" and not weighted and k >= 3:
            assert ix[q][:3].tolist() == [2, 7, 9]            # exact ties: lower index first
