"""Worker of the sharded optimiser test: `python dist_mcmc_worker.py <rank> <world> <port> <backend> <out.npz>`
(world = 1: the single-process run it is compared with; no process group).  Every rank holds a replica of the walkers /
particles on its device and scores its range of each proposal batch (`sharding.batch_device_sharded`)."""
import io
import os
import sys
from contextlib import redirect_stdout

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def run():
    import numpy as np
    import pandas as pd
    from pysodm_b200 import optimization, tutorial_models as tm
    from pysodm_b200.objective_functions import ll_poisson, log_posterior_probability
    from pysodm_b200.optimization import mcmc
    truth = tm.ODE_SIR_rate({"S": 1_000_000 - 10, "I": 10}, {"beta": 0.9, "gamma": 0.2})
    t = np.arange(10, 41, 2)
    data = pd.Series(truth.sim(40)["I"].values[t], index=pd.Index(t.astype(float), name="time"))
    model = tm.ODE_SIR_rate({"S": 1_000_000 - 10, "I": 10}, {"beta": 0.5, "gamma": 0.2})
    lpp = log_posterior_probability(model, ["beta", "gamma"], [(0.1, 2.0), (0.05, 1.0)], [data], ["I"], [ll_poisson], [[]],
                                    start_sim=0)
    n = 66                                                   # halves of 33: shards of different length
    x0 = np.column_stack([np.random.default_rng(1).uniform(0.85, 0.95, n), np.random.default_rng(2).uniform(0.18, 0.22, n)])
    sampler = mcmc.EnsembleSampler(n, 2, lpp, seed=5, device_update=True)        # the reference's cocktail
    sampler.run_mcmc(x0, 24)
    np.random.seed(3)
    with redirect_stdout(io.StringIO()):
        theta, score = optimization.pso(lpp, swarmsize=51, max_iter=12, device_update=True)
    return {"chain": sampler.get_chain(), "log_prob": sampler.get_log_prob(), "accepted": sampler.acceptance_fraction,
            "theta": np.asarray(theta), "score": np.asarray(score)}


def main():
    rank, world, port, backend, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    import numpy as np
    import torch
    if world == 1:
        np.savez(path, **run())
        return
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(backend, rank=rank, world_size=world)
    from pysodm_b200 import engine
    engine.device(rank % torch.cuda.device_count())
    out = run()
    np.savez(path + f".rank{rank}.npz", **out)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
