"""TEST INFRASTRUCTURE -- CPU restatement of what the reference's optimisers do to a whole population between two
evaluations of the objective (not shipped, not measured).

Follows /root/reference/src/pySODM/optimization/pso.py:207-216 (velocity / position update and the bound clamp) and
:229-236 (personal bests, `np.argmin`), written with the same NumPy expressions, and -- for the KDE move of the
reference's default cocktail (/root/reference/src/pySODM/optimization/mcmc.py:26; emcee's KDEMove is a third-party
dependency absent from this image) -- `scipy.stats.gaussian_kde.logpdf`, the function that move calls.

Pinning: `tests/test_optimization.py` replays the unmodified reference's seeded `pso.optimize` runs (golden fixtures)
through the host optimiser built on these same expressions; the KDE density is scipy's own.
"""
import numpy as np


def pso_move(x, v, p, g, lb, ub, omega, phip, phig, rp, rg):
    """pso.py:207-216 for x, v, p, rp, rg [S, D]; g, lb, ub [D].  Returns (x_new, v_new)."""
    v = omega * v + phip * rp * (p - x) + phig * rg * (g - x)
    x = x + v
    maskl = x < lb
    masku = x > ub
    x = x * (~np.logical_or(maskl, masku)) + lb * maskl + ub * masku
    return x, v


def pso_best(x, fx, p, fp, feasible=None):
    """pso.py:229-236: personal bests where fx < fp (and feasible), then the swarm's best.  Returns (p, fp, i_min)."""
    fs = np.ones(len(fx), dtype=bool) if feasible is None else np.asarray(feasible, dtype=bool)
    i_update = np.logical_and(fx < fp, fs)
    p, fp = p.copy(), fp.copy()
    p[i_update, :] = x[i_update, :].copy()
    fp[i_update] = fx[i_update]
    return p, fp, int(np.argmin(fp))


def kde_logpdf(data, queries, bw_method="scott"):
    """log-density of a Gaussian KDE of data[K, D] at queries[M, D] (scipy.stats.gaussian_kde)."""
    from scipy.stats import gaussian_kde
    return gaussian_kde(np.asarray(data).T, bw_method=bw_method).logpdf(np.asarray(queries).T)
