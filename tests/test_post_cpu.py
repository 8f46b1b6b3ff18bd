"""Ensemble post-processing, CPU side: the oracle (oracle/post.py) pinned against outputs of the unmodified reference
(`tests/golden/post.npz`, made by tests/golden/make_golden.py::gen_post), and the LAW of the device samplers
(pysodm_b200/csrc/noise.cuh compiled for the host by tests/hostcheck) against NumPy's."""
import numpy as np
import pytest
from scipy import stats

import hostcheck_util as hc
from oracle import post


def test_noise_oracle_reproduces_reference_bit_exact(golden):
    g = golden["post"]
    for name in ("counts", "conc"):
        x = g[name]
        np.random.seed(101)
        # the reference walks the Dataset's variables in order, one sampler call each (utils.py:27-31)
        a, b = post.add_poisson_noise(g["counts"]), post.add_poisson_noise(g["conc"])
        assert np.array_equal({"counts": a, "conc": b}[name], g["poisson_" + name])
        np.random.seed(102)
        a, b = post.add_negative_binomial_noise(g["counts"], 0.03), post.add_negative_binomial_noise(g["conc"], 0.03)
        assert np.array_equal({"counts": a, "conc": b}[name], g["negbin_" + name])
        np.random.seed(103)
        a, b = post.add_gaussian_noise(g["counts"], 0.04, True), post.add_gaussian_noise(g["conc"], 0.04, True)
        assert np.array_equal({"counts": a, "conc": b}[name], g["gauss_rel_" + name])
        np.random.seed(104)
        a, b = post.add_gaussian_noise(g["counts"], 2.5, False), post.add_gaussian_noise(g["conc"], 2.5, False)
        assert np.array_equal({"counts": a, "conc": b}[name], g["gauss_abs_" + name])
        assert x.shape == g["poisson_" + name].shape


def test_lerp_restatement_is_numpys():
    rng = np.random.default_rng(2)
    for n in (2, 3, 17, 1000, 99991):
        x = rng.normal(size=n) * 10 ** rng.uniform(-3, 6)
        xs = np.sort(x)
        for q in (0.0, 0.025, 0.25, 0.5, 0.75, 0.975, 1.0, 1 / 3, 0.999999):
            lo, hi, gamma = post.bracketing_ranks(n, q)
            assert post.lerp_order_stats(xs[lo], xs[hi], gamma) == np.quantile(x, q), (n, q)


@pytest.mark.parametrize("kind,param,x", [
    (0, 0.0, 0.4), (0, 0.0, 3.7), (0, 0.0, 9.99), (0, 0.0, 10.0), (0, 0.0, 42.0), (0, 0.0, 5213.5), (0, 0.0, 7.3e6),
    (1, 2.5, 10.0), (1, 0.0, 3.0), (2, 0.04, 250.0), (2, 0.04, -17.0),
    (3, 0.03, 55.0), (3, 0.03, 2.0e5), (3, 0.5, 4.0), (3, 2.0, 900.0), (3, 1.0, 0.3)])
def test_device_noise_samplers_follow_numpy_law(kind, param, x):
    n = 40000
    mine, status = hc.noise(kind, param, x, n, seed=5, cell=3)
    assert status == 0
    rs = np.random.RandomState(77)
    if kind == 0:
        ref = post.add_poisson_noise(np.full(n, x), rs)
    elif kind == 1:
        ref = post.add_gaussian_noise(np.full(n, x), param, False, rs)
    elif kind == 2:
        ref = post.add_gaussian_noise(np.full(n, x), param, True, rs)
    else:
        ref = post.add_negative_binomial_noise(np.full(n, x), param, rs)
    if param == 0.0 and kind == 1:
        assert np.array_equal(mine, ref)
        return
    assert stats.ks_2samp(mine, ref).pvalue > 1e-3
    if kind in (0, 3):
        assert np.all(mine == np.round(mine)) and mine.min() >= 0
        # exact moments: Poisson mean = var = x; NB mean x, var x + alpha x^2
        var = x if kind == 0 else x + param * x * x
        assert abs(mine.mean() - x) < 5 * np.sqrt(var / n)


def test_device_noise_edge_cases():
    assert np.all(hc.noise(0, 0, 0.0, 100)[0] == 0)                     # Poisson(0)
    assert np.all(hc.noise(3, 0.03, 0.0, 100)[0] == 0)                  # NB with mean 0 (p = 1)
    for kind in (0, 3):
        out, status = hc.noise(kind, 0.03, -1.0, 10)
        assert status == 1                                               # NumPy raises ValueError there
        out, status = hc.noise(kind, 0.03, float("nan"), 10)
        assert status == 1
    a, _ = hc.noise(0, 0, 42.0, 1000, seed=9, cell=1)
    b, _ = hc.noise(0, 0, 42.0, 1000, seed=9, cell=1)
    c, _ = hc.noise(0, 0, 42.0, 1000, seed=9, cell=2)
    assert np.array_equal(a, b) and not np.array_equal(a, c)            # streams keyed by (seed, trajectory, cell)
