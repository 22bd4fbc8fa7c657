"""Statistical comparison helpers for Monte Carlo outputs."""
import numpy as np


def _t2_expectation(K, r):
    """E[t^2] of the two-sample statistic below for Gaussian cells whose per-repetition variances have the ratio r:
    E[(1 + r) / (x + r y)] with x, y ~ chi^2_{K-1}/(K-1).  Closed form nu/(nu-2), nu = 2(K-1), for r = 1; numerically
    (fixed-seed sampling, 4e6 draws: 3 digits) otherwise."""
    if abs(r - 1.0) < 1e-12:
        nu = 2.0 * (K - 1)
        return nu / (nu - 2.0)
    rng = np.random.default_rng(20261018)
    x = rng.chisquare(K - 1, 4000000) / (K - 1)
    y = rng.chisquare(K - 1, 4000000) / (K - 1)
    return float(np.mean((1.0 + r) / (x + r * y)))


def chi2_per_dof(a_runs, b_runs, min_rel=0.0, var_ratio=1.0):
    """a_runs, b_runs: [K, cells] independent repetitions of the same estimator from two implementations.
    Returns (chi2/dof, dof, relative difference of the totals, its sigma).  Per cell t^2 = (mean_a-mean_b)^2 /
    (var_a/K + var_b/K) with the sample variances; the expectation of t^2 for Gaussian cells (nu/(nu-2) with
    nu = 2(K-1) for equal variances) is divided out so that chi2/dof has expectation 1.  var_ratio = ratio of the
    per-repetition variances the comparison was designed with (e.g. 1/4 when every repetition of one side used four
    times the packets and was scaled down)."""
    a = np.asarray(a_runs, np.float64)
    b = np.asarray(b_runs, np.float64)
    K = a.shape[0]
    ma, mb = a.mean(0), b.mean(0)
    va, vb = a.var(0, ddof=1) / K, b.var(0, ddof=1) / K
    ok = (a > 0).all(0) & (b > 0).all(0) & (va + vb > 0)
    if min_rel > 0:
        ok &= ma > min_rel * ma.max()
    t2 = (ma[ok] - mb[ok]) ** 2 / (va[ok] + vb[ok])
    chi2 = t2.mean() / _t2_expectation(K, float(var_ratio))
    tot = abs(ma.sum() - mb.sum()) / mb.sum()
    tot_sigma = np.sqrt(a.sum(1).var(ddof=1) / K + b.sum(1).var(ddof=1) / K) / mb.sum()
    return chi2, int(ok.sum()), tot, tot_sigma
