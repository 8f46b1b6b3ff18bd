"""Batched ensemble sampler (pysodm_b200/optimization/mcmc.py): every move leaves a known target invariant, the
objective is reached through its `batch` seam only, and the reference's front end (`run_EnsembleSampler`,
`perturbate_theta`, `/root/reference/src/pySODM/optimization/mcmc.py`) keeps its contract: samples file layout, restart
from a backend, settings validation."""
import datetime
import os

import numpy as np
import pytest

from pysodm_b200.optimization import mcmc
from pysodm_b200.samples import load_samples

MU = np.array([1.0, -2.0, 0.5])
COV = np.array([[1.0, 0.6, 0.0], [0.6, 2.0, -0.5], [0.0, -0.5, 0.5]])
PREC = np.linalg.inv(COV)


class Gaussian:
    """Correlated normal target with the attributes of log_posterior_probability the sampler front end reads."""
    parameter_shapes = {"a": (1,), "b": (2,)}
    expanded_labels = ["a", "b_0", "b_1"]

    def __init__(self):
        self.batch_calls, self.single_calls = 0, 0

    def __call__(self, x):
        self.single_calls += 1
        d = np.asarray(x) - MU
        return -0.5 * d @ PREC @ d

    def batch(self, X):
        self.batch_calls += 1
        d = np.asarray(X) - MU
        return -0.5 * np.einsum("ni,ij,nj->n", d, PREC, d)


def scipy_logpdf(data, queries, bw_method):
    """The oracle of `sodm_kde_logpdf` (scipy.stats.gaussian_kde), standing in for the device on the CPU."""
    from scipy.stats import gaussian_kde
    return gaussian_kde(data.T, bw_method=bw_method).logpdf(queries.T)


def cocktail():
    return [(m if not isinstance(m, mcmc.KDEMove) else mcmc.KDEMove(bw_method="scott", logpdf=scipy_logpdf), w)
            for m, w in mcmc.default_moves()]


def start(n, seed=0):
    return MU + np.random.default_rng(seed).normal(size=(n, 3))


@pytest.mark.parametrize("move", [mcmc.StretchMove(), mcmc.DEMove(), mcmc.DEMove(gamma0=1.0), mcmc.DESnookerMove(),
                                  mcmc.KDEMove(bw_method="scott", logpdf=scipy_logpdf), cocktail()],
                         ids=["stretch", "de", "de_gamma1", "snooker", "kde", "cocktail"])
def test_move_samples_the_target(move):
    f = Gaussian()
    n, iters, burn = 400, 260, 60
    s = mcmc.EnsembleSampler(n, 3, f, moves=move, seed=11)
    s.run_mcmc(start(n), iters)
    assert f.single_calls == 0 and f.batch_calls == 1 + 2 * iters == s.n_batches
    x = s.get_chain(discard=burn, flat=True)
    # 80 000 correlated samples; the effective sample size is several thousand for every move
    assert np.allclose(x.mean(axis=0), MU, atol=0.08), x.mean(axis=0)
    assert np.allclose(np.cov(x, rowvar=False), COV, atol=0.15), np.cov(x, rowvar=False)
    assert 0.05 < s.acceptance_fraction.mean() < 0.95
    assert np.array_equal(s.get_log_prob()[-1], f.batch(s.get_chain()[-1]))


def test_distinct_indices():
    rng = np.random.default_rng(3)
    idx = mcmc._pick_distinct(rng, 5, 20000, 3)
    assert idx.min() == 0 and idx.max() == 4
    assert np.all(idx[:, 0] != idx[:, 1]) and np.all(idx[:, 0] != idx[:, 2]) and np.all(idx[:, 1] != idx[:, 2])
    # every ordered triple equally likely: 60 of them
    code = idx[:, 0] * 25 + idx[:, 1] * 5 + idx[:, 2]
    counts = np.unique(code, return_counts=True)[1]
    assert len(counts) == 60 and np.all(np.abs(counts - 20000 / 60) < 5 * np.sqrt(20000 / 60))
    with pytest.raises(ValueError):
        mcmc._pick_distinct(rng, 2, 4, 3)


def test_integrated_time_of_ar1():
    rng = np.random.default_rng(5)
    rho, n_t, n_w = 0.8, 4000, 16
    x = np.zeros((n_t, n_w, 1))
    for t in range(1, n_t):
        x[t] = rho * x[t - 1] + rng.normal(size=(n_w, 1))
    tau = mcmc.integrated_time(x)
    assert abs(tau[0] - (1 + rho) / (1 - rho)) < 1.0                      # 9 for rho = 0.8


def test_infinite_log_prob_is_rejected():
    class Box(Gaussian):
        def batch(self, X):
            lp = super().batch(X)
            lp[np.asarray(X)[:, 0] < MU[0]] = -np.inf
            return lp
    x0 = start(64)
    x0[:, 0] = MU[0] + np.abs(x0[:, 0] - MU[0]) + 0.01
    s = mcmc.EnsembleSampler(64, 3, Box(), moves=cocktail(), seed=2)
    s.run_mcmc(x0, 50)
    assert np.all(s.get_chain()[:, :, 0] >= MU[0]) and np.all(np.isfinite(s.get_log_prob()))


def test_start_validation():
    f = Gaussian()
    with pytest.raises(ValueError, match="twice"):
        mcmc.EnsembleSampler(4, 3, f)
    s = mcmc.EnsembleSampler(8, 3, f)
    with pytest.raises(ValueError, match="dimensions"):
        s.run_mcmc(start(6), 1)
    with pytest.raises(ValueError, match="condition number"):
        s.run_mcmc(np.ones((8, 3)) * np.arange(8)[:, None], 1)


def test_run_ensemble_sampler_files_and_restart(tmp_path, monkeypatch, capsys):
    monkeypatch.chdir(tmp_path)
    f = Gaussian()
    np.random.seed(4)
    ndim, nwalkers, pos = mcmc.perturbate_theta([1.0, -2.0, 0.5], pert=[0.3, 0.3, 0.3], multiplier=4,
                                                bounds=[(-10, 10), (-10, 10), (-10, 10)], verbose=True)
    assert (ndim, nwalkers, pos.shape) == (3, 12, (12, 3))
    assert np.all(np.abs(pos / np.array([1.0, -2.0, 0.5]) - 1) <= 0.3)
    settings = {"start_calibration": 0, "starting_estimate": [1.0, -2.0, 0.5], "tag": "x"}
    sampler, samples = mcmc.run_EnsembleSampler(pos, 30, "user", f, (), samples_path="out/", fig_path="out/", print_n=10,
                                                discard=5, thin=2, settings_dict=settings, progress=False, seed=1,
                                                moves=cocktail())
    today = str(datetime.date.today())
    assert sampler.iteration == 30 and f.single_calls == 0
    assert samples["a"].shape == (12, 12) and samples["b"].shape == (12, 12, 2)
    assert np.array_equal(samples["b"][..., 1], sampler.get_chain(discard=5, thin=2)[..., 2])
    on_disk = load_samples(f"out/user_SAMPLES_{today}.nc")
    assert np.array_equal(on_disk["a"], samples["a"]) and np.array_equal(on_disk["b"], samples["b"])
    assert on_disk.attrs["discard"] == 5 and on_disk.attrs["thin"] == 2 and on_disk.attrs["tag"] == "x"
    assert "Non-convergence" in capsys.readouterr().out
    # restart: the reference passes the backend's file name (an .hdf5 name resolves to the .npz written beside it)
    s2, samples2 = mcmc.run_EnsembleSampler(pos, 20, "user", f, (), samples_path="out/", print_n=10, progress=False,
                                            backend=f"out/user_BACKEND_{today}.hdf5", settings_dict=settings, seed=2,
                                            moves=cocktail())
    assert s2.iteration == 50 and np.array_equal(s2.get_chain()[:30], sampler.get_chain())
    assert samples2["a"].shape == (50, 12)
    with pytest.raises(FileNotFoundError):
        mcmc.run_EnsembleSampler(pos, 5, "user", f, backend="out/none.hdf5", progress=False, moves=cocktail())
    with pytest.raises(TypeError, match="settings_dict"):
        mcmc.run_EnsembleSampler(pos, 5, "user", f, settings_dict={"when": datetime.date.today()}, progress=False)
    with pytest.raises(TypeError, match="homogeneous"):
        mcmc.run_EnsembleSampler(pos, 5, "user", f, settings_dict={"mixed": [1, "a"]}, progress=False)


def test_perturbate_theta_validation():
    with pytest.raises(Exception, match="same length"):
        mcmc.perturbate_theta([1.0, 2.0], pert=[0.1], bounds=[(0, 1), (0, 1)])
    with pytest.raises(Exception, match="tuple for every parameter"):
        mcmc.perturbate_theta([1.0, 2.0], pert=[0.1, 0.1], bounds=[(0, 1)])
    np.random.seed(0)
    _, _, pos = mcmc.perturbate_theta([5.0, 0.2], pert=[0.5, 0.5], multiplier=3, bounds=[(0.1, 6.0), (0.1, 1.0)])
    assert pos[:, 0].max() <= 6.0 and pos[:, 1].min() >= 0.1           # clipped before perturbing


@pytest.mark.parametrize("kind,param,sigma", [("stretch", 2.0, 0.0), ("de", 0.0, 1e-3), ("snooker", 1.7, 0.0)])
def test_word_stream_oracle_moves_sample_the_target(kind, param, sigma):
    """The per-walker restatement the device kernels are checked against (oracle/population.py: draws taken from an
    explicit stream of 32-bit words) is itself a valid sampler of the same Gaussian."""
    from oracle import population as opop
    rng = np.random.default_rng(4)
    n, iters, burn = 64, 500, 100
    x = start(n)
    f = Gaussian()
    lp = f.batch(x)
    keep = []
    for it in range(iters):
        order = rng.permutation(n)
        for mine, others in ((order[:n // 2], order[n // 2:]), (order[n // 2:], order[:n // 2])):
            for me in mine:
                q, lf = opop.walker_propose(kind, x, me, others, param, sigma, rng.integers(0, 2 ** 32, 16, dtype=np.uint64))
                new = f(q)
                if opop.walker_accepts(lp[me], new, lf, rng.integers(0, 2 ** 32, 4, dtype=np.uint64)):
                    x[me], lp[me] = q, new
        if it >= burn:
            keep.append(x.copy())
    chain = np.concatenate(keep)
    assert np.allclose(chain.mean(axis=0), MU, atol=0.12), chain.mean(axis=0)
    assert np.allclose(np.cov(chain, rowvar=False), COV, atol=0.2), np.cov(chain, rowvar=False)
