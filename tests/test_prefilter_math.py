"""CPU: the conservative pre-filters of the batched tensor kernel (csrc/tc_batch.cu, batch_bounds_kernel + epilogue)
restated in numpy float32 -- a score that the exact fp32 test accepts must never be rejected by the cheap filter.
This probes the decision boundary with millions of adversarial (dot, |t|, |z|^2, bound) combinations."""
import numpy as np

F = np.float32


def fma(a, b, c):
    return (a.astype(np.float64) * b.astype(np.float64) + c.astype(np.float64)).astype(F)


def _cases(n, seed):
    rng = np.random.Generator(np.random.PCG64(seed))
    qq = rng.uniform(50.0, 3000.0, n).astype(F)          # |t|^2
    rn = rng.uniform(0.0, 3000.0, n).astype(F)           # |z|^2 (0 = null row)
    cos = rng.uniform(-1.0, 1.0, n)
    dot = (cos * np.sqrt(qq.astype(np.float64) * rn.astype(np.float64))).astype(F)
    return rng, qq, rn, dot


def test_mse_prefilter_never_rejects_an_exact_pass():
    D = 768
    dd = F(D) * F(D)
    inv_dd = F(1.0) / dd
    rng, qc, rn, dot = _cases(2_000_000, 1)
    sv = ((qc - F(2.0) * dot) + rn) * inv_dd                                  # the exact test's score (fp32, same order)
    # bounds right at, just above and just below the score of this very row: the hardest cases
    ulps = rng.integers(-3, 4, sv.shape[0])
    tau = sv.copy()
    for _ in range(3):
        tau = np.where(ulps > 0, np.nextafter(tau, F(np.inf)), np.where(ulps < 0, np.nextafter(tau, F(-np.inf)), tau)).astype(F)
        ulps = ulps - np.sign(ulps)
    exact_pass = sv <= tau                                                    # smaller is better; ties may pass
    td = tau * dd
    b1 = (F(0.5) * (qc - td) - F(2e-6) * (np.abs(qc) + np.abs(td))).astype(F)
    rterm = (F(0.5) * rn * (F(1.0) - F(4e-6))).astype(F)
    bound = (b1 + rterm).astype(F)
    pre_pass = ~(dot < bound)
    assert not np.any(exact_pass & ~pre_pass)
    # and the filter is tight: it lets through (almost) nothing that is clearly worse than the bound
    clearly_worse = sv > tau * F(1.001) + F(1e-6)
    assert (pre_pass & clearly_worse).mean() < 1e-3


def test_cosine_prefilter_never_rejects_an_exact_pass():
    rng, qq, rn, dot = _cases(2_000_000, 2)
    qc, mx = np.sqrt(qq).astype(F), np.sqrt(rn).astype(F)
    sv = (dot / fma(qc, mx, np.full_like(qc, F(1e-6)))).astype(F)
    ulps = rng.integers(-3, 4, sv.shape[0])
    tau = sv.copy()
    for _ in range(3):
        tau = np.where(ulps > 0, np.nextafter(tau, F(np.inf)), np.where(ulps < 0, np.nextafter(tau, F(-np.inf)), tau)).astype(F)
        ulps = ulps - np.sign(ulps)
    exact_pass = sv >= tau                                                    # larger is better
    f = np.where(tau >= 0, F(1.0) - F(4e-6), F(1.0) + F(4e-6)).astype(F)
    b1 = (tau * qc * f).astype(F)
    b2 = (tau * F(1e-6) * f).astype(F)
    bound = fma(b1, mx, b2)
    pre_pass = ~(dot < bound)
    assert not np.any(exact_pass & ~pre_pass)
    clearly_worse = sv < tau - np.abs(tau) * F(1e-3) - F(1e-6)
    assert (pre_pass & clearly_worse).mean() < 1e-3


def test_weighted_cosine_prefilter_never_rejects_an_exact_pass():
    """csrc/tc_weighted.cu: s = d1 / (q1 sqrt(d2) + 1e-6) >= th is implied by f(d1 + 1e-6) >= f(th q1) d2 with
    f(x) = x |x|, lowered by 8e-6 relative (no square root, no division in the filter)."""
    rng = np.random.Generator(np.random.PCG64(3))
    n = 2_000_000
    q1 = rng.uniform(0.05, 5.0, n).astype(F)                                  # sqrt(sum w t^2)
    d2 = rng.uniform(0.0, 25.0, n).astype(F)                                  # w.(z z)
    d2[rng.random(n) < 0.01] = F(-1e-7)                                       # rounding can leave it slightly negative
    cos = rng.uniform(-1.0, 1.0, n)
    d1 = (cos * q1.astype(np.float64) * np.sqrt(np.maximum(d2, 0).astype(np.float64))).astype(F)
    r = np.sqrt(np.maximum(d2, F(0))).astype(F)
    sv = (d1.astype(np.float64) / fma(q1, r, np.full_like(q1, F(1e-6))).astype(np.float64)).astype(F)
    sv = np.where(np.abs(sv) < 1.0001, sv, np.sign(sv) * F(1.0)).astype(F)    # __fdividef is within 2 ulp of this
    ulps = rng.integers(-3, 4, n)
    th = sv.copy()
    for _ in range(3):
        th = np.where(ulps > 0, np.nextafter(th, F(np.inf)), np.where(ulps < 0, np.nextafter(th, F(-np.inf)), th)).astype(F)
        ulps = ulps - np.sign(ulps)
    # the device divides with __fdividef (<= 2 ulp): count as an exact pass anything within 2 ulp of the threshold
    lo = th
    for _ in range(2):
        lo = np.nextafter(lo, F(-np.inf)).astype(F)
    exact_pass = sv >= lo
    x = (d1 + F(1e-6)).astype(F)
    t2 = (th * q1).astype(F)
    rhs = (t2 * np.abs(t2) * np.maximum(d2, F(0))).astype(F)
    rhs = fma(-np.abs(rhs), np.full_like(rhs, F(8e-6)), rhs)
    pre_pass = ~((x * np.abs(x)).astype(F) < rhs)
    assert not np.any(exact_pass & ~pre_pass)


def test_k2_cosine_prefilter_never_rejects_an_exact_pass():
    """csrc/tc_search.cu epilogue: bound = th (|t||z| + 1e-6), lowered by 1e-6 relative; pass iff !(dot < bound)."""
    rng, qq, rn, dot = _cases(2_000_000, 4)
    qc, mx = np.sqrt(qq).astype(F), np.sqrt(rn).astype(F)
    den = fma(qc, mx, np.full_like(qc, F(1e-6)))
    sv = (dot / den).astype(F)
    ulps = rng.integers(-3, 4, sv.shape[0])
    th = sv.copy()
    for _ in range(3):
        th = np.where(ulps > 0, np.nextafter(th, F(np.inf)), np.where(ulps < 0, np.nextafter(th, F(-np.inf)), th)).astype(F)
        ulps = ulps - np.sign(ulps)
    exact_pass = sv >= th
    bound = (th * den).astype(F)
    bound = fma(-np.abs(bound), np.full_like(bound, F(1e-6)), bound)
    assert not np.any(exact_pass & ~(~(dot < bound)))
