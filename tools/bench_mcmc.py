"""C5 through the optimiser front ends: 10^5 walkers / particles on the influenza NB posterior (84-day window), one B200.

Wall-clock per iteration of `optimization.mcmc.EnsembleSampler` (two half-ensemble `batch` calls) and of
`optimization.pso` (one `batch` call), split into the batched objective (host -> device -> host, simulation + likelihood
kernels) and the host-side bookkeeping (proposals, acceptance, swarm update).  Usage: python tools/bench_mcmc.py [out.json]
"""
import io
import json
import os
import sys
import time
from contextlib import redirect_stdout

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch

from pysodm_b200 import optimization, tutorial_data as td, tutorial_models as tm
from pysodm_b200.objective_functions import ll_negative_binomial, log_posterior_probability
from pysodm_b200.optimization import mcmc


class Timed:
    """Wraps an objective's `batch`, accumulating the wall time spent inside it."""

    def __init__(self, lpp):
        self.lpp, self.seconds, self.calls, self.rows = lpp, 0.0, 0, 0
        self.expanded_bounds, self.parameter_shapes = lpp.expanded_bounds, lpp.parameter_shapes

    def batch(self, X):
        t = time.perf_counter()
        lp = self.lpp.batch(X)
        self.seconds += time.perf_counter() - t
        self.calls += 1
        self.rows += len(X)
        return lp

    def __call__(self, x):
        return self.lpp(x)


class TimedDevice(Timed):
    def batch_device(self, X):
        t = time.perf_counter()
        lp = self.lpp.batch_device(X)
        torch.cuda.synchronize()
        self.seconds += time.perf_counter() - t
        self.calls += 1
        self.rows += X.shape[1]
        return lp


def main():
    n, iters = 100_000, 10
    model = tm.JumpProcess_influenza_model(td.INFLUENZA_INITIAL_STATES, td.INFLUENZA_PARAMETERS,
                                           coordinates={"age_group": td.AGE_GROUPS},
                                           time_dependent_parameters={"N": td.influenza_contact_function})
    dates = pd.date_range("2017-12-01", periods=13, freq="7D")
    idx = pd.MultiIndex.from_product([dates, td.AGE_GROUPS], names=["date", "age_group"])
    rng = np.random.default_rng(0)
    data = pd.Series(rng.gamma(2.0, 150.0, size=len(idx)), index=idx, name="CASES")
    lpp = log_posterior_probability(model, ["beta", "f_ud"], [(1e-6, 0.08), (1e-3, 1 - 1e-3)], [data], ["Im_inc"],
                                    [ll_negative_binomial], [4 * [0.03]], simulation_kwargs={"tau": 0.5})
    pos = np.column_stack([rng.uniform(0.01, 0.03, n), rng.uniform(0.05, 0.95, size=(n, 4))])
    lpp.batch(pos.copy())                                     # warm-up (tables, allocator)
    torch.cuda.synchronize()
    res = {"n_walkers": n, "iterations": iters, "window_days": 84, "data_points": 52}

    for label, moves in (("cocktail", None), ("stretch", mcmc.StretchMove()), ("kde", mcmc.KDEMove(bw_method="scott"))):
        f = Timed(lpp)
        s = mcmc.EnsembleSampler(n, 5, f, moves=moves, seed=1)
        lp0 = f.batch(pos.copy())
        f.seconds, f.calls, f.rows = 0.0, 0, 0
        t = time.perf_counter()
        s.run_mcmc(pos, iters, log_prob0=lp0)
        wall = time.perf_counter() - t
        res[f"mcmc_{label}"] = {"ms_per_iteration": wall / iters * 1e3, "ms_in_batch": f.seconds / iters * 1e3,
                                "ms_host_bookkeeping": (wall - f.seconds) / iters * 1e3, "batch_calls": f.calls,
                                "walker_updates_per_s": n * iters / wall,
                                "acceptance": float(s.acceptance_fraction.mean())}

    # the same sampler with the ensemble resident on the device (sodm_walkers_propose / _accept, sodm_population_clip)
    for label, moves in (("cocktail", None), ("stretch", mcmc.StretchMove()), ("de", mcmc.DEMove()),
                         ("kde", mcmc.KDEMove(bw_method="scott"))):
        f = TimedDevice(lpp)
        s = mcmc.EnsembleSampler(n, 5, f, moves=moves, seed=1, device_update=True)
        lp0 = f.batch(pos.copy())
        f.seconds, f.calls, f.rows = 0.0, 0, 0
        torch.cuda.synchronize()
        t = time.perf_counter()
        s.run_mcmc(pos, iters, log_prob0=lp0)
        wall = time.perf_counter() - t
        res[f"mcmc_{label}_device_update"] = {"ms_per_iteration": wall / iters * 1e3, "ms_in_batch": f.seconds / iters * 1e3,
                                              "ms_move_kernels_and_host": (wall - f.seconds) / iters * 1e3,
                                              "batch_calls": f.calls, "walker_updates_per_s": n * iters / wall,
                                              "acceptance": float(s.acceptance_fraction.mean())}

    f = Timed(lpp)
    np.random.seed(0)
    t = time.perf_counter()
    with redirect_stdout(io.StringIO()):
        optimization.pso(f, swarmsize=n, max_iter=iters)
    wall = time.perf_counter() - t
    res["pso"] = {"ms_per_iteration": wall / f.calls * 1e3, "ms_in_batch": f.seconds / f.calls * 1e3,
                  "ms_host_bookkeeping": (wall - f.seconds) / f.calls * 1e3, "batch_calls": f.calls,
                  "particle_evaluations_per_s": f.rows / wall}

    f = TimedDevice(lpp)
    np.random.seed(0)
    t = time.perf_counter()
    with redirect_stdout(io.StringIO()):
        optimization.pso(f, swarmsize=n, max_iter=iters, device_update=True)
    wall = time.perf_counter() - t
    res["pso_device_update"] = {"ms_per_iteration": wall / f.calls * 1e3, "ms_in_batch": f.seconds / f.calls * 1e3,
                                "ms_swarm_kernels_and_host": (wall - f.seconds) / f.calls * 1e3, "batch_calls": f.calls,
                                "particle_evaluations_per_s": f.rows / wall}

    # the KDE proposal density alone: 4096 kernels x 10^5 points, 5 dimensions (device) against scipy on a slice
    from pysodm_b200 import engine
    from scipy.stats import gaussian_kde
    data, queries = pos[:4096], pos
    engine.gaussian_kde_logpdf(data, queries[:1000])
    t = time.perf_counter()
    got = engine.gaussian_kde_logpdf(data, queries)
    ms_dev = (time.perf_counter() - t) * 1e3
    L = np.linalg.cholesky(np.cov(data, rowvar=False) * 4096 ** (-2.0 / 9))
    pts = torch.from_numpy(np.ascontiguousarray(np.linalg.solve(L, data.T))).cuda()
    qs = torch.from_numpy(np.ascontiguousarray(np.linalg.solve(L, queries.T))).cuda()
    out = torch.empty(n, dtype=torch.float64, device="cuda")
    engine.kde_logpdf(pts, qs, out=out)
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    engine.kde_logpdf(pts, qs, out=out)
    b.record()
    torch.cuda.synchronize()
    t = time.perf_counter()
    want = gaussian_kde(data.T, bw_method="scott").logpdf(queries[:2000].T)
    ms_scipy = (time.perf_counter() - t) * 1e3 * (n / 2000)
    res["kde_logpdf_4096x1e5x5"] = {"ms_device_incl_whitening_and_copies": ms_dev, "ms_kernel": a.elapsed_time(b),
                                    "pairs_per_s": 4096 * n / a.elapsed_time(b) * 1e3,
                                    "ms_scipy_extrapolated_from_2000": ms_scipy,
                                    "max_abs_diff_vs_scipy": float(np.max(np.abs(got[:2000] - want)))}

    text = json.dumps(res, indent=1)
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")


if __name__ == "__main__":
    main()
