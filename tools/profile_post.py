"""One launch each of the post-processing kernels on an influenza ensemble, for ncu (python tools/profile_post.py [n_traj])."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from pysodm_b200 import engine, tutorial_data as td

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 17
dev = engine.device(0)
y0 = torch.from_numpy(td.influenza_flat_y0()).to(dev)
tab = torch.from_numpy(td.CONTACT_TABLE.reshape(3, 16).copy()).to(dev)
idx = torch.from_numpy(td.influenza_schedule()).to(dev)
draws = torch.from_numpy(bench.synthetic_draws(n, 1718)).to(dev)
ens = torch.empty((28, 162, n), dtype=torch.int32, device=dev)
engine.jump_run("influenza_1718_jump", y0, td.influenza_flat_par(), n, 0, 162, 0.5, var=(bench.VAR_INDEX, draws),
                tdpf=(tab, idx), seed=1, out_dtype=torch.int32, out=ens, dev=dev)
noisy = torch.empty_like(ens)
for rep in range(2):
    engine.noise("negative_binomial", ens, 0.03, seed=rep, out=noisy)
    engine.noise("poisson", ens, 0.0, seed=rep, out=noisy)
    engine.draws_sum(noisy)
    lo, hi = int((n - 1) * 0.025), int((n - 1) * 0.975)
    engine.draws_order_stats(noisy, [lo, lo + 1, hi, hi + 1])
torch.cuda.synchronize()
print("ok")
g = torch.empty((28, 162, n), dtype=torch.float32, device=dev)
for rep in range(2):
    engine.noise("gaussian_relative", ens, 0.04, seed=rep, out=g)
torch.cuda.synchronize()
