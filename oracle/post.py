"""TEST INFRASTRUCTURE -- CPU restatement of the reference's ensemble post-processing (not shipped, not measured).

Follows /root/reference/src/pySODM/optimization/utils.py:11-95 (the three `add_*_noise` functions: one NumPy legacy
sampler call per state variable on the whole array) and the xarray reductions the tutorials apply to a finished
ensemble (`mean(dim='draws')`, `quantile(dim='draws', q=...)`: tutorials/influenza_1718/calibration.py:245-246), which
for a dense array are `np.mean` and `np.quantile(..., method='linear')` along the draws axis.

Pinning: the arithmetic is NumPy's own (`np.random.poisson`, `.normal`, `.negative_binomial`, `np.quantile`), the very
functions the reference calls -- nothing is re-derived here, so there is nothing to pin beyond argument order, which
`tests/test_oracle_golden.py::test_noise_oracle_matches_reference_utils` checks against the unmodified reference
functions (xarray stub) in this container.
"""
import numpy as np


def add_poisson_noise(values, rs=None):
    """utils.py:30 -- every value replaced by a Poisson draw with that mean."""
    rs = np.random if rs is None else rs
    return rs.poisson(values)


def add_gaussian_noise(values, sigma, relative=True, rs=None):
    """utils.py:59-64 -- absolute: N(x, sigma); relative: N(x, |sigma x|)."""
    rs = np.random if rs is None else rs
    values = np.asarray(values)
    if relative:
        return rs.normal(values, np.abs(sigma * values))
    return rs.normal(values, sigma)


def add_negative_binomial_noise(values, alpha, rs=None):
    """utils.py:91-92 -- NB with n = 1/alpha, p = n / (x + n): mean x, variance x + alpha x^2."""
    rs = np.random if rs is None else rs
    values = np.asarray(values)
    return rs.negative_binomial(1 / alpha, (1 / alpha) / (values + (1 / alpha)))


def mean_over_draws(values, axis=0):
    return np.mean(np.asarray(values, dtype=np.float64), axis=axis)


def quantile_over_draws(values, q, axis=0):
    """xarray's DataArray.quantile default: linear interpolation between order statistics (np.quantile / nanquantile)."""
    return np.quantile(np.asarray(values, dtype=np.float64), q, axis=axis, method="linear")


def bracketing_ranks(n, q):
    """0-based ranks of the two order statistics np.quantile(method='linear') interpolates between, and the weight of
    the upper one (numpy/lib/_function_base_impl.py: 'linear' -> virtual index (n - 1) q; `_get_indexes`: floor and
    floor + 1, both the LAST element once the virtual index reaches n - 1; `_get_gamma`: the fractional part)."""
    virtual = (n - 1) * float(q)
    lo = int(np.floor(virtual))
    hi = lo + 1
    if virtual >= n - 1:
        lo = hi = n - 1
    if virtual < 0:
        lo = hi = 0
    return lo, hi, virtual - np.floor(virtual)


def lerp_order_stats(lo, hi, gamma):
    """NumPy's `_lerp(a, b, t)`: a + (b - a) t, and b - (b - a)(1 - t) where t >= 0.5 -- restated so that the host
    side of the device quantile can be checked bit for bit."""
    lo, hi = np.asarray(lo, dtype=np.float64), np.asarray(hi, dtype=np.float64)
    diff = hi - lo
    out = lo + diff * gamma
    return np.where(gamma >= 0.5, hi - diff * (1 - gamma), out)
