"""One launch of the influenza tau-leap kernel for ncu (python tools/profile_kernel.py [n_traj] [compat])."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from pysodm_b200 import engine, tutorial_data as td
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 18
compat = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = engine.device(0)
y0 = torch.from_numpy(td.influenza_flat_y0()).to(dev)
tab = torch.from_numpy(td.CONTACT_TABLE.reshape(3, 16).copy()).to(dev)
idx = torch.from_numpy(td.influenza_schedule()).to(dev)
draws = torch.from_numpy(bench.synthetic_draws(n, 1718)).to(dev)
out = torch.empty((28, 162, n), dtype=torch.int32, device=dev)
for rep in range(3):
    engine.jump_run("influenza_1718_jump", y0, td.influenza_flat_par(), n, 0, 162, 0.5, var=(bench.VAR_INDEX, draws),
                    tdpf=(tab, idx), seed=rep, compat=compat, out_dtype=torch.int32, out=out, dev=dev)
torch.cuda.synchronize()
print("ok", int(out[24:28, -1].sum()))
