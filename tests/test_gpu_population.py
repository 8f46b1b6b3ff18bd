"""GPU parity tests of the optimiser-population kernels (sodm_pso_move, sodm_pso_best, sodm_kde_logpdf) against
oracle/population.py, and of the device-resident particle swarm against the host one."""
import numpy as np
import pytest

import hostcheck_util as hc
from oracle import population as opop

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    from pysodm_b200 import engine
    engine.device(0)
    return engine, torch


def _dev(eng, a):
    engine, torch = eng
    return torch.from_numpy(np.ascontiguousarray(a)).to(engine.device(0))


@pytest.mark.parametrize("D,K,M", [(1, 7, 1), (2, 256, 129), (3, 1000, 5000), (5, 4096, 3000), (9, 300, 257), (16, 513, 64)])
def test_kde_logpdf_matches_scipy(eng, D, K, M):
    engine, torch = eng
    rng = np.random.default_rng(D * 1000 + K)
    A = rng.normal(size=(D, D)) + 2 * np.eye(D)
    data = rng.normal(size=(K, D)) @ A.T + rng.normal(size=D)
    queries = rng.normal(size=(M, D)) @ A.T * 1.5
    queries[0] = data[0]                                      # on a kernel centre
    queries[-1] = data.mean(axis=0) + 60 * data.std(axis=0)   # far in the tail: densities ~ e^-1000 must not underflow
    for bw in ("scott", "silverman", 0.37):
        got = engine.gaussian_kde_logpdf(data, queries, bw)
        want = opop.kde_logpdf(data, queries, bw)
        assert np.all(np.isfinite(got))
        np.testing.assert_allclose(got, want, rtol=1e-10, atol=1e-10)


def test_kde_logpdf_validation(eng):
    engine, torch = eng
    from pysodm_b200._ffi import SodmError
    pts = _dev(eng, np.zeros((3, 8)))
    with pytest.raises(ValueError):
        engine.kde_logpdf(pts, _dev(eng, np.zeros((2, 8))))
    with pytest.raises(ValueError):
        engine.kde_logpdf(_dev(eng, np.zeros((17, 8))), _dev(eng, np.zeros((17, 8))))
    with pytest.raises(SodmError):
        engine.kde_logpdf(_dev(eng, np.zeros((3, 0))), pts)
    assert engine.kde_logpdf(pts, _dev(eng, np.zeros((3, 0)))).shape == (0,)


@pytest.mark.parametrize("S,D", [(1, 1), (257, 3), (100_000, 5), (4099, 16)])
def test_pso_move_matches_oracle(eng, S, D):
    """Same uniforms (Philox keyed by (seed, particle, iteration), one block per dimension: r_p from words 0-1, r_g from
    words 2-3, 53 bits each) -> the reference's update expression; FMA contraction on the device allows 1-2 ulp."""
    engine, torch = eng
    rng = np.random.default_rng(S + D)
    lb, ub = -1.0 - rng.random(D), 1.0 + rng.random(D)
    x, p = rng.uniform(lb, ub, size=(2, S, D))
    v = rng.uniform(-2, 2, size=(S, D))
    g = rng.uniform(lb, ub)
    seed, it, id0 = 0x1234_5678_9ABC_DEF0, 7, 10 ** 10
    xd, vd, pd = _dev(eng, x.T), _dev(eng, v.T), _dev(eng, p.T)
    engine.pso_move(xd, vd, pd, _dev(eng, g), lb, ub, 0.8, 0.7, 0.9, seed, it, member_id0=id0)
    check = np.unique(np.concatenate([[0, S - 1], rng.integers(0, S, 200)]))
    rp, rg = np.empty((len(check), D)), np.empty((len(check), D))
    for r, i in enumerate(check):
        w = hc.philox(seed, id0 + int(i), it, D).astype(np.uint64)
        rp[r] = ((w[:, 0] << np.uint64(32) | w[:, 1]) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
        rg[r] = ((w[:, 2] << np.uint64(32) | w[:, 3]) >> np.uint64(11)).astype(np.float64) * 2.0 ** -53
    want_x, want_v = opop.pso_move(x[check], v[check], p[check], g, lb, ub, 0.8, 0.7, 0.9, rp, rg)
    got_x, got_v = xd.cpu().numpy().T, vd.cpu().numpy().T
    np.testing.assert_allclose(got_v[check], want_v, rtol=1e-13, atol=1e-14)
    np.testing.assert_allclose(got_x[check], want_x, rtol=1e-13, atol=1e-14)
    assert np.all(got_x >= lb) and np.all(got_x <= ub)
    assert np.any(got_x == lb) or np.any(got_x == ub) or S < 100        # the clamp is exercised
    # a shard of the swarm reproduces its slice (streams keyed by the global particle index)
    if S > 100:
        lo = S // 3
        xs, vs = _dev(eng, x[lo:].T), _dev(eng, v[lo:].T)
        engine.pso_move(xs, vs, _dev(eng, p[lo:].T), _dev(eng, g), lb, ub, 0.8, 0.7, 0.9, seed, it, member_id0=id0 + lo)
        assert np.array_equal(xs.cpu().numpy(), got_x[lo:].T) and np.array_equal(vs.cpu().numpy(), got_v[lo:].T)


@pytest.mark.parametrize("S,D", [(1, 2), (1000, 3), (70_001, 5)])
def test_pso_best_matches_oracle(eng, S, D):
    engine, torch = eng
    rng = np.random.default_rng(S)
    x, p = rng.normal(size=(2, S, D))
    fx = np.round(rng.normal(size=S), 1)                      # many ties: the FIRST minimum wins
    fp = np.round(rng.normal(size=S), 1)
    fp[rng.integers(0, S, S // 10 + 1)] = np.inf
    feasible = rng.random(S) < 0.8
    for feas in (None, feasible):
        pd, fpd = _dev(eng, p.T), _dev(eng, fp)
        best, idx = engine.pso_best(_dev(eng, x.T), _dev(eng, fx), pd, fpd,
                                    feasible=None if feas is None else _dev(eng, feas))
        want_p, want_fp, i_min = opop.pso_best(x, fx, p, fp, feas)
        assert np.array_equal(pd.cpu().numpy().T, want_p) and np.array_equal(fpd.cpu().numpy(), want_fp)
        assert int(idx) == i_min
        assert np.array_equal(best.cpu().numpy(), np.concatenate([[want_fp[i_min]], want_p[i_min]]))
    # NaN scores: never better than a personal best (fx < fp is False), but np.argmin puts a NaN personal best first
    fp2 = fp.copy()
    if S > 10:
        fp2[[S // 2, S // 2 + 3]] = np.nan
        fx2 = fx.copy()
        fx2[:5] = np.nan
        pd, fpd = _dev(eng, p.T), _dev(eng, fp2)
        best, idx = engine.pso_best(_dev(eng, x.T), _dev(eng, fx2), pd, fpd)
        want_p, want_fp, i_min = opop.pso_best(x, fx2, p, fp2)
        assert int(idx) == i_min == S // 2
        assert np.array_equal(fpd.cpu().numpy(), want_fp, equal_nan=True)


def test_device_swarm_calibrates_like_the_host_swarm(eng, capsys):
    """The calibration of tests/test_gpu_facade.py::test_pso_calibration_recovers_parameter with the swarm kept on the
    device: same optimum; per iteration nothing but 1 + D numbers comes back."""
    import pandas as pd
    from pysodm_b200 import optimization, tutorial_models as tm
    from pysodm_b200.objective_functions import ll_poisson, log_posterior_probability
    truth = tm.ODE_SIR_rate({"S": 1_000_000 - 10, "I": 10}, {"beta": 0.9, "gamma": 0.2})
    out = truth.sim(40)
    t = np.arange(10, 41, 2)
    data = pd.Series(out["I"].values[t], index=pd.Index(t.astype(float), name="time"))
    model = tm.ODE_SIR_rate({"S": 1_000_000 - 10, "I": 10}, {"beta": 0.5, "gamma": 0.2})
    lpp = log_posterior_probability(model, ["beta", "gamma"], [(0.1, 2.0), (0.05, 1.0)], [data], ["I"], [ll_poisson], [[]],
                                    start_sim=0)
    np.random.seed(0)
    theta_h, score_h = optimization.pso(lpp, swarmsize=256, max_iter=40)
    np.random.seed(0)
    theta_d, score_d = optimization.pso(lpp, swarmsize=256, max_iter=40, device_update=True)
    assert "Swarm update on the device" in capsys.readouterr().out
    for theta in (theta_h, theta_d):
        assert abs(theta[0] - 0.9) / 0.9 < 0.02 and abs(theta[1] - 0.2) / 0.2 < 0.05
    assert abs(score_d - score_h) < 1e-3 * abs(score_h) + 1.0
    # the reported score is the objective at the reported position
    assert abs(-lpp(np.array(theta_d)) - score_d) <= 1e-9 * abs(score_d)
    # the device evaluation of a resident population equals the host-fed one
    engine, torch = eng
    X = np.column_stack([np.linspace(0.5, 1.5, 33), np.linspace(0.1, 0.4, 33)])
    assert np.array_equal(lpp.batch_device(_dev(eng, X.T)).cpu().numpy(), lpp.batch(X.copy()))
    # priors other than "uniform over the bounds" are evaluated on the host for their coordinate
    from pysodm_b200.objective_functions import log_prior_normal, log_prior_uniform
    lpp2 = log_posterior_probability(model, ["beta", "gamma"], [(0.1, 2.0), (0.05, 1.0)], [data], ["I"], [ll_poisson], [[]],
                                     start_sim=0, log_prior_prob_fnc=[log_prior_normal, log_prior_uniform],
                                     log_prior_prob_fnc_args=[{"avg": 0.9, "stdev": 0.1}, {"bounds": (0.15, 0.35)}])
    got, want = lpp2.batch_device(_dev(eng, X.T)).cpu().numpy(), lpp2.batch(X.copy())
    assert np.array_equal(got, want) and np.isinf(want).any() and np.isfinite(want).any()
    assert not np.array_equal(want, lpp.batch(X.copy()))
    with pytest.raises(ValueError, match="inside the objective's bounds"):
        optimization.pso(lpp, bounds=[(0.0, 3.0), (0.05, 1.0)], swarmsize=8, max_iter=1, device_update=True)
    with pytest.raises(NotImplementedError):
        optimization.pso(lpp, swarmsize=8, max_iter=1, device_update=True, transform_pars=lambda x: x)


# ---- ensemble-MCMC moves on the device --------------------------------------------------------------------------------

def _stream(seed, walker, purpose, n_blocks=4):
    return hc.philox(seed, int(walker), int(purpose), n_blocks).reshape(-1)


@pytest.mark.parametrize("kind,param,sigma", [("stretch", 2.0, 0.0), ("de", 0.0, 1e-5), ("de", 1.0, 0.3), ("snooker", 1.7, 0.0)])
@pytest.mark.parametrize("n,D", [(7, 1), (1001, 3), (100_000, 5)])
def test_walkers_propose_matches_oracle(eng, kind, param, sigma, n, D):
    """Same Philox words (seed, walker, 2 * iteration) -> the published move, walker by walker; FMA contraction and the
    device's log / cospi allow a few ulp."""
    engine, torch = eng
    rng = np.random.default_rng(n + D)
    x = rng.normal(size=(n, D)) * (1.0 + np.arange(D))
    order = rng.permutation(n)
    mine, others = order[: n // 2], order[n // 2:]
    seed, it = 0xFEED_FACE_CAFE_BEEF, 12345
    q, lf = engine.walkers_propose(kind, _dev(eng, x.T), torch.from_numpy(mine).cuda(), torch.from_numpy(others).cuda(),
                                   param, sigma, seed, it)
    q, lf = q.cpu().numpy().T, lf.cpu().numpy()
    assert q.shape == (len(mine), D) and np.all(np.isfinite(q)) and np.all(np.isfinite(lf))
    for j in np.unique(np.concatenate([[0, len(mine) - 1], rng.integers(0, len(mine), 150)])):
        want_q, want_lf = opop.walker_propose(kind, x, mine[j], others, param, sigma, _stream(seed, mine[j], 2 * it))
        np.testing.assert_allclose(q[j], want_q, rtol=1e-12, atol=1e-13)
        np.testing.assert_allclose(lf[j], want_lf, rtol=1e-11, atol=1e-12)
    if n >= 1000:                                       # partners are spread over the whole complementary half
        if kind == "stretch":
            z = np.exp(lf / (D - 1.0))
            assert 0.5 <= z.min() < 0.52 and 1.95 < z.max() <= 2.0
            assert abs(np.mean(z) - 7.0 / 6.0) < 0.02   # E z under g(z) ~ 1/sqrt(z) on [1/2, 2]


@pytest.mark.parametrize("n,D,K", [(40, 1, 20), (3000, 3, 700), (60_000, 5, 4096), (5000, 16, 2500)])
def test_walkers_kde_move_matches_oracle(eng, n, D, K):
    """KDE move: the draw c_k + L z from the same Philox words, and the Hastings factor density(s) - density(q) against
    scipy.stats.gaussian_kde over the same centres and bandwidth."""
    engine, torch = eng
    from pysodm_b200.optimization import mcmc
    rng = np.random.default_rng(n + D)
    A = rng.normal(size=(D, D))
    x = rng.normal(size=(n, D)) @ A.T
    order = rng.permutation(n)
    mine, centres = order[: n // 2], order[n // 2:][:K]
    move = mcmc.KDEMove(bw_method="scott")
    L = move.factor_of(x[centres])
    seed, it = 31337, 40
    q, lf = engine.walkers_propose_kde(_dev(eng, x.T), torch.from_numpy(mine).cuda(), torch.from_numpy(centres).cuda(), L, seed, it)
    q, lf = q.cpu().numpy().T, lf.cpu().numpy()
    check = np.unique(np.concatenate([[0, len(mine) - 1], rng.integers(0, len(mine), 150)]))
    for j in check:
        want_q, _ = opop.walker_propose("kde", x, mine[j], centres, L, 0.0, _stream(seed, mine[j], 2 * it, 2 + D))
        np.testing.assert_allclose(q[j], want_q, rtol=1e-12, atol=1e-12)
    dens = opop.kde_logpdf(x[centres], np.concatenate([x[mine[check]], q[check]]), move.bandwidth(len(centres), D))
    np.testing.assert_allclose(lf[check], dens[:len(check)] - dens[len(check):], rtol=1e-9, atol=1e-9)
    # the proposals are draws from the KDE: their covariance is (1 + h^2) times the centres' covariance
    if n >= 3000 and D <= 5:
        h2 = move.bandwidth(len(centres), D) ** 2
        np.testing.assert_allclose(np.cov(q, rowvar=False), (1 + h2) * np.cov(x[centres], rowvar=False), rtol=0.15,
                                   atol=0.1 * np.abs(np.cov(x[centres], rowvar=False)).max())


def test_walkers_propose_validation(eng):
    engine, torch = eng
    x = _dev(eng, np.zeros((3, 8)))
    idx = torch.arange(4, device="cuda")
    with pytest.raises(Exception, match="needs 3 walkers"):
        engine.walkers_propose("snooker", x, idx, idx[:2], 1.7)
    with pytest.raises(Exception, match="exceed 1"):
        engine.walkers_propose("stretch", x, idx, idx, 1.0)
    with pytest.raises(TypeError):
        engine.walkers_propose("stretch", x, idx.to(torch.int32), idx, 2.0)
    q, lf = engine.walkers_propose("stretch", x, idx[:0], idx, 2.0)
    assert q.shape == (3, 0) and lf.shape == (0,)


def test_population_clip_and_accept_match_oracle(eng):
    engine, torch = eng
    rng = np.random.default_rng(5)
    n, D = 20_001, 4
    lb, ub = -np.ones(D), np.array([1.0, 2.0, 0.5, 1.0])
    q = rng.normal(size=(n, D))
    qd = _dev(eng, q.T)
    pen = engine.population_clip(qd, lb, ub).cpu().numpy()
    out = np.any((q < lb) | (q > ub), axis=1)
    assert np.array_equal(qd.cpu().numpy().T, np.clip(q, lb, ub))
    assert np.array_equal(pen, np.where(out, -np.inf, 0.0)) and out.any() and not out.all()

    N = 2 * n + 1
    x, lp = rng.normal(size=(N, D)), rng.normal(size=N)
    mine = rng.permutation(N)[:n]
    new_lp, lf = rng.normal(size=n), 0.3 * rng.normal(size=n)
    new_lp[:50] = -np.inf
    new_lp[50:60] = np.inf
    seed, it = 99, 7
    xd, lpd, acc = _dev(eng, x.T), _dev(eng, lp), torch.zeros(N, dtype=torch.int64, device="cuda")
    status = torch.zeros(1, dtype=torch.int32, device="cuda")
    engine.walkers_accept(xd, lpd, torch.from_numpy(mine).cuda(), qd, _dev(eng, new_lp), _dev(eng, lf), acc,
                          penalty=_dev(eng, pen), seed=seed, iteration=it, status=status)
    want_x, want_lp, want_acc = x.copy(), lp.copy(), np.zeros(N, dtype=np.int64)
    for j, me in enumerate(mine):
        if opop.walker_accepts(lp[me], new_lp[j] if pen[j] == 0 else -np.inf, lf[j], _stream(seed, me, 2 * it + 1, 1)):
            want_x[me], want_lp[me], want_acc[me] = np.clip(q[j], lb, ub), new_lp[j], 1
    assert int(status) == 0
    assert np.array_equal(acc.cpu().numpy(), want_acc) and 0.05 < want_acc.sum() / n < 0.8
    assert np.array_equal(xd.cpu().numpy().T, want_x) and np.array_equal(lpd.cpu().numpy(), want_lp)
    # a NaN score is reported, not accepted
    new_lp[100] = np.nan
    engine.walkers_accept(xd, lpd, torch.from_numpy(mine).cuda(), qd, _dev(eng, new_lp), _dev(eng, lf), acc, status=status)
    assert int(status) == 1 and np.isfinite(lpd.cpu().numpy()).all()


class _GaussianTarget:
    """A correlated Gaussian with the interface the device sampler needs (`batch`, `batch_device`, `expanded_bounds`)."""

    def __init__(self, torch, D, bound=50.0):
        rng = np.random.default_rng(D)
        A = rng.normal(size=(D, D))
        self.cov = A @ A.T / D + np.eye(D) * 0.3
        self.prec = np.linalg.inv(self.cov)
        self.prec_d = torch.from_numpy(self.prec).cuda()
        self.expanded_bounds = [(-bound, bound)] * D
        self.torch = torch

    def batch(self, X):
        return -0.5 * np.einsum("ij,jk,ik->i", X, self.prec, X)

    def batch_device(self, X):                          # [D, P]
        return -0.5 * (X * (self.prec_d @ X)).sum(dim=0)


@pytest.mark.parametrize("move", ["stretch", "de", "snooker", "kde", "cocktail"])
def test_device_ensemble_samples_a_gaussian(eng, move):
    """Each device move on its own, and the reference's cocktail, leave a correlated
    Gaussian invariant: covariance of the pooled chain within Monte-Carlo error (the CPU sampler passes the same test,
    tests/test_mcmc_cpu.py)."""
    engine, torch = eng
    from pysodm_b200.optimization import mcmc
    D, n, iters, burn = 3, 512, 400, 100
    target = _GaussianTarget(torch, D)
    moves = {"stretch": mcmc.StretchMove(), "de": mcmc.DEMove(sigma=1e-3), "snooker": mcmc.DESnookerMove(),
             "kde": mcmc.KDEMove(bw_method="scott", max_points=200), "cocktail": None}[move]
    if move == "snooker":                               # the snooker alone mixes slowly (it is meant for cocktails)
        moves = [(mcmc.DESnookerMove(), 0.5), (mcmc.DEMove(sigma=1e-3), 0.5)]
    sampler = mcmc.EnsembleSampler(n, D, target, moves=moves, seed=11, device_update=True)
    x0 = np.random.default_rng(0).multivariate_normal(np.zeros(D), target.cov, size=n)
    sampler.run_mcmc(x0, iters)
    chain = sampler.get_chain(discard=burn, flat=True)
    assert chain.shape == ((iters - burn) * n, D)
    assert np.all(np.abs(chain.mean(axis=0)) < 0.1 * np.sqrt(np.diag(target.cov)))
    np.testing.assert_allclose(np.cov(chain, rowvar=False), target.cov, rtol=0.12, atol=0.04)
    assert 0.15 < sampler.acceptance_fraction.mean() < 0.95
    assert np.allclose(sampler.get_log_prob()[-1], target.batch(sampler.get_chain()[-1]))


def test_device_ensemble_respects_bounds(eng):
    """Proposals outside the objective's bounds are clipped and never accepted (objective_functions.py:301-307)."""
    engine, torch = eng
    from pysodm_b200.optimization import mcmc
    target = _GaussianTarget(torch, 2, bound=0.5)
    sampler = mcmc.EnsembleSampler(256, 2, target, moves=mcmc.StretchMove(), seed=2, device_update=True)
    x0 = np.random.default_rng(0).uniform(-0.4, 0.4, size=(256, 2))
    sampler.run_mcmc(x0, 100)
    chain = sampler.get_chain(flat=True)
    assert np.all(np.abs(chain) < 0.5) and np.abs(chain).max() > 0.45


def test_sharded_device_populations_match_one_process(tmp_path):
    """The device sampler (reference cocktail) and the device swarm under torch.distributed: replicas of the population on
    every rank, each rank scoring its range of the proposals, one all-gather of log-posteriors per batch -- chains, scores
    and the swarm's optimum equal the single-process run bit for bit, on every rank.  NCCL when the box has >= 2 GPUs,
    else two ranks on one GPU over gloo."""
    import socket
    import subprocess
    import sys
    import torch
    worker = __file__.replace("test_gpu_population.py", "dist_mcmc_worker.py")
    world = 2
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = str(s.getsockname()[1])
    single, sharded = str(tmp_path / "single.npz"), str(tmp_path / "sharded")
    procs = [subprocess.Popen([sys.executable, worker, "0", "1", port, backend, single], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True)]
    procs += [subprocess.Popen([sys.executable, worker, str(r), str(world), port, backend, sharded], stdout=subprocess.PIPE,
                               stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    want = np.load(single)
    assert want["chain"].shape == (24, 66, 2) and 0.05 < want["accepted"].mean() < 0.95
    for r in range(world):
        got = np.load(sharded + f".rank{r}.npz")
        for key in ("chain", "log_prob", "accepted", "theta", "score"):
            assert np.array_equal(got[key], want[key]), (r, key)
