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


# ---- ensemble-MCMC moves (the reference's cocktail, /root/reference/src/pySODM/optimization/mcmc.py:22-28) --------------
# emcee is a third-party dependency absent from this image: these restate the published moves (Goodman & Weare 2010;
# ter Braak 2006; ter Braak & Vrugt 2008) one walker at a time, drawing from an explicit stream of 32-bit words so
# that the device kernels can be checked value for value.  "Parity unpinned" against emcee itself: the statistical
# tests (tests/test_mcmc_cpu.py, tests/test_gpu_facade.py) check that each move samples a known target.

class Words:
    """Uniforms, bounded integers and normals from a stream of uint32 words (two words per draw, high word first)."""

    def __init__(self, words):
        self.w, self.i = np.asarray(words, dtype=np.uint64).reshape(-1), 0

    def _u64(self):
        v = (int(self.w[self.i]) << 32) | int(self.w[self.i + 1])
        self.i += 2
        return v

    def uniform(self):                                  # [0, 1), 53 bits
        return (self._u64() >> 11) * 2.0 ** -53

    def below(self, n):                                 # 0 .. n-1: high half of the 64 x 64 bit product
        return (self._u64() * int(n)) >> 64

    def normal(self):                                   # Box-Muller, the cosine branch
        u1, u2 = 1.0 - self.uniform(), self.uniform()
        return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def walker_propose(kind, x, me, others, param, sigma, words):
    """One walker's proposal.  x [n, D]; `others` the index list of the complementary half; kind 'stretch' | 'de' |
    'snooker' | 'kde'.  Returns (q [D], log Hastings factor)."""
    r = Words(words)
    s, D = x[me], x.shape[1]
    if kind == "stretch":
        z = ((param - 1.0) * r.uniform() + 1.0) ** 2 / param
        c = x[others[r.below(len(others))]]
        return c - (c - s) * z, (D - 1.0) * np.log(z)
    if kind == "de":
        gamma = param if param > 0 else 2.38 / np.sqrt(2.0 * D)
        ia, ib = r.below(len(others)), r.below(len(others) - 1)
        ib += ib >= ia
        return s + gamma * (x[others[ia]] - x[others[ib]]) + sigma * r.normal(), 0.0
    if kind == "kde":                                   # param = lower Cholesky factor of the kernel covariance
        c = x[others[r.below(len(others))]]
        z = np.array([r.normal() for _ in range(D)])
        return c + np.asarray(param) @ z, None          # the factor is a density ratio: kde_logpdf(s) - kde_logpdf(q)
    i0, i1, i2 = r.below(len(others)), r.below(len(others) - 1), r.below(len(others) - 2)
    i1 += i1 >= i0
    lo, hi = min(i0, i1), max(i0, i1)
    i2 += i2 >= lo
    i2 += i2 >= hi
    z, z1, z2 = x[others[i0]], x[others[i1]], x[others[i2]]
    delta = s - z
    norm = np.linalg.norm(delta)
    u = delta / norm
    q = s + u * param * (np.dot(u, z1) - np.dot(u, z2))
    return q, (D - 1.0) * (np.log(np.linalg.norm(q - z)) - np.log(norm))


def walker_accepts(lp_old, lp_new, log_factor, words):
    """Metropolis-Hastings with u on (0, 1]; an infinite score never moves the walker."""
    log_u = np.log(1.0 - Words(words).uniform())
    return bool(np.isfinite(lp_new) and log_factor + lp_new - lp_old > log_u)
