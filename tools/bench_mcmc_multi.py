"""C5 through the device-resident optimisers on N GPUs: torchrun --nproc-per-node N tools/bench_mcmc_multi.py [out.json]
(also runs as a single process).  10^5 walkers / particles on the influenza NB posterior; every rank holds a replica of
the population and scores its range of each batch; wall clock per iteration on rank 0 after a barrier.
SODM_BENCH_WALKERS sets the population size (default 10^5)."""
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
import torch.distributed as dist

from pysodm_b200 import engine, optimization, tutorial_data as td, tutorial_models as tm
from pysodm_b200.objective_functions import ll_negative_binomial, log_posterior_probability
from pysodm_b200.optimization import mcmc


def main():
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    engine.device(local)
    if ws > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    n, iters = int(os.environ.get("SODM_BENCH_WALKERS", "100000")), int(os.environ.get("SODM_BENCH_ITERS", "10"))
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
    lp0 = lpp.batch(pos.copy())
    res = {"n_gpus": ws, "n_walkers": n, "iterations": iters}

    def fence():
        torch.cuda.synchronize()
        if ws > 1:
            dist.barrier()

    # steady state: per-iteration wall times from the sampler's generator (median), so that the one-time set-up (broadcast of
    # the starting ensemble, first touch of the chain storage, NCCL / kernel warm-up) is reported apart from the iteration
    for label, moves in (("cocktail", None), ("stretch", mcmc.StretchMove())):
        for rep in range(2):                                  # the first pass warms kernels, allocator and NCCL up
            s = mcmc.EnsembleSampler(n, 5, lpp, moves=moves, seed=1, device_update=True)
            fence()
            t0 = t = time.perf_counter()
            laps = []
            for _ in s.sample(pos, iterations=iters, log_prob0=lp0):
                torch.cuda.synchronize()
                now = time.perf_counter()
                laps.append(now - t)
                t = now
            fence()
            wall = time.perf_counter() - t0
        res[f"mcmc_{label}_device_update"] = {"ms_per_iteration": float(np.median(laps[1:])) * 1e3,
                                              "ms_per_iteration_incl_setup": wall / iters * 1e3,
                                              "walker_updates_per_s": n / float(np.median(laps[1:])),
                                              "acceptance": float(s.acceptance_fraction.mean())}
    walls = {}
    for k in (iters, 2 * iters, iters, 2 * iters):            # steady state = (T(2 K) - T(K)) / K; the first pair warms up
        np.random.seed(0)
        fence()
        t = time.perf_counter()
        with redirect_stdout(io.StringIO()):
            optimization.pso(lpp, swarmsize=n, max_iter=k, device_update=True, minstep=0, minfunc=0)
        fence()
        walls[k] = time.perf_counter() - t
    steady = (walls[2 * iters] - walls[iters]) / iters
    res["pso_device_update"] = {"ms_per_iteration": steady * 1e3, "ms_per_iteration_incl_setup": walls[iters] / (iters + 1) * 1e3,
                                "particle_evaluations_per_s": n / steady}
    if rank == 0:
        text = json.dumps(res, indent=1)
        print(text)
        if len(sys.argv) > 1:
            open(sys.argv[1], "w").write(text + "\n")
    if ws > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
