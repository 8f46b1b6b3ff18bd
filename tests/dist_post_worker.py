"""Worker of the sharded post-processing tests: `python dist_post_worker.py <rank> <world> <port> <backend> <out.npz>`.
Every rank simulates its shard of ONE ensemble on cuda:(rank % device_count), reduces it across ranks (mean, quantiles,
noise) and rank 0 writes the results for comparison with the single-process run."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def main():
    rank, world, port, backend, path = int(sys.argv[1]), int(sys.argv[2]), sys.argv[3], sys.argv[4], sys.argv[5]
    import numpy as np
    import torch
    import torch.distributed as dist
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=port, RANK=str(rank), WORLD_SIZE=str(world))
    torch.cuda.set_device(rank % torch.cuda.device_count())
    dist.init_process_group(backend, rank=rank, world_size=world)
    from pysodm_b200 import sharding, tutorial_data as td, tutorial_models as tm
    from pysodm_b200.optimization.utils import add_negative_binomial_noise
    model = tm.JumpProcess_influenza_model(td.INFLUENZA_INITIAL_STATES, td.INFLUENZA_PARAMETERS,
                                           coordinates={"age_group": td.AGE_GROUPS},
                                           time_dependent_parameters={"N": td.influenza_contact_function})

    def draw(parameters, N):
        rng = np.random.default_rng(0)
        parameters["beta"] = rng.uniform(0.015, 0.02, size=N)
        return parameters
    draw.vectorized = True
    n = 5001                                                    # odd: shards of different length
    ens = model.sim(["2017-12-01", "2018-01-10"], N=n, tau=0.5, seed=21, output="device", draw_function=draw,
                    device=rank % torch.cuda.device_count())
    lo, hi = sharding.shard_range(n)
    assert ens.buffer.shape[2] == hi - lo and ens.n_draws == n and ens["S"].shape == (n, 41, 4)
    qs = [0.0, 0.025, 0.5, 0.975, 1.0]
    mean, q = ens.mean(), ens.quantile(qs)
    noisy = add_negative_binomial_noise(ens, 0.03, seed=5)
    qn = noisy["Im_inc"].sel(age_group=td.AGE_GROUPS[2]).quantile(0.9)
    f = ens.like(ens.buffer.to(torch.float64) * 0.37 - 11.0)     # a float64 ensemble through the same reductions
    qf, mf = f["E"].quantile([0.1, 0.66]), f["E"].mean()
    local = ens.to_dataset()
    full = {k: sharding.all_gather_concat(torch.from_numpy(np.ascontiguousarray(local[k].values)).cuda(), n).cpu().numpy() for k in ("S", "Im_inc")}
    full_noisy = sharding.all_gather_concat(torch.from_numpy(np.ascontiguousarray(noisy["Im_inc"].values)).cuda(), n).cpu().numpy()
    if rank == 0:
        np.savez(path, mean_S=np.asarray(mean["S"].values), mean_Im=np.asarray(mean["Im_inc"].values),
                 q_S=np.asarray(q["S"].values), q_Im=np.asarray(q["Im_inc"].values), qn=np.asarray(qn.values),
                 qf=np.asarray(qf.values), mf=np.asarray(mf.values), full_S=full["S"], full_Im=full["Im_inc"],
                 full_noisy=full_noisy)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
