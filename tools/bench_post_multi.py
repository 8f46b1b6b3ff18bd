"""Sharded post-processing at scale, one rank per GPU (torchrun --nproc-per-node G tools/bench_post_multi.py [n_per_gpu]).
Weak scaling: every rank holds n_per_gpu influenza trajectories; times (CUDA events, max over ranks) of the cross-rank
mean and of the 2.5 % / 97.5 % order statistics (radix select in phases, one NCCL all-reduce of the histograms per pass)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

import bench
from pysodm_b200 import engine, sharding, tutorial_data as td


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    dev = engine.device(local)
    dist.init_process_group("nccl", device_id=dev)
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
    y0 = torch.from_numpy(td.influenza_flat_y0()).to(dev)
    tab = torch.from_numpy(td.CONTACT_TABLE.reshape(3, 16).copy()).to(dev)
    idx = torch.from_numpy(td.influenza_schedule()).to(dev)
    draws = torch.from_numpy(bench.synthetic_draws(n, 1718 + rank)).to(dev)
    ens = torch.empty((28, 162, n), dtype=torch.int32, device=dev)
    engine.jump_run("influenza_1718_jump", y0, td.influenza_flat_par(), n, 0, 162, 0.5, var=(bench.VAR_INDEX, draws),
                    tdpf=(tab, idx), seed=1, traj_id0=rank * n, out_dtype=torch.int32, out=ens, dev=dev)
    n_tot = n * world
    lo, hi = int(np.floor((n_tot - 1) * 0.025)), int(np.floor((n_tot - 1) * 0.975))
    ranks = [lo, lo + 1, hi, hi + 1]

    def timed(fn, reps=3):
        fn()
        ts = []
        for _ in range(reps):
            dist.barrier()
            torch.cuda.synchronize()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            out = fn()
            b.record()
            torch.cuda.synchronize()
            t = torch.tensor([a.elapsed_time(b)], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ts.append(float(t))
        return float(np.median(ts)), out
    ms_sum, s = timed(lambda: sharding.sum_sharded(engine.draws_sum(ens)))
    ms_q, q = timed(lambda: sharding.order_stats_sharded(ens, ranks))
    ms_local, _ = timed(lambda: engine.draws_order_stats(ens, [int((n - 1) * 0.025), int((n - 1) * 0.975)]))
    # sanity: the global order statistics lie between the extreme local ones, sums add up
    assert bool((q[0] <= q[3]).all())
    if rank == 0:
        print(json.dumps({"n_gpus": world, "n_traj_per_gpu": n, "ensemble_GB_per_gpu": ens.numel() * 4 / 1e9,
                          "mean_sharded_ms": ms_sum, "order_stats_sharded_4_ranks_ms": ms_q,
                          "order_stats_local_fused_2_ranks_ms": ms_local,
                          "histogram_allreduce_MB_per_pass": 28 * 162 * 4 * 2048 * 4 / 1e6}))
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
