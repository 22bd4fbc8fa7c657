"""Statistical comparison helpers for Monte Carlo outputs."""
import numpy as np


def chi2_per_dof(a_runs, b_runs, min_rel=0.0):
    """a_runs, b_runs: [K, cells] independent repetitions of the same estimator from two implementations.
    Returns (chi2/dof, dof, relative difference of the totals).  Per cell t^2 = (mean_a-mean_b)^2 /
    (var_a/K + var_b/K) with the sample variances; for Gaussian cells E[t^2] = nu/(nu-2) with nu = 2(K-1),
    which is divided out so that the expectation is 1."""
    a = np.asarray(a_runs, np.float64)
    b = np.asarray(b_runs, np.float64)
    K = a.shape[0]
    ma, mb = a.mean(0), b.mean(0)
    va, vb = a.var(0, ddof=1) / K, b.var(0, ddof=1) / K
    ok = (a > 0).all(0) & (b > 0).all(0) & (va + vb > 0)
    if min_rel > 0:
        ok &= ma > min_rel * ma.max()
    t2 = (ma[ok] - mb[ok]) ** 2 / (va[ok] + vb[ok])
    nu = 2.0 * (K - 1)
    chi2 = t2.mean() * (nu - 2.0) / nu
    tot = abs(ma.sum() - mb.sum()) / mb.sum()
    tot_sigma = np.sqrt(a.sum(1).var(ddof=1) / K + b.sum(1).var(ddof=1) / K) / mb.sum()
    return chi2, int(ok.sum()), tot, tot_sigma
