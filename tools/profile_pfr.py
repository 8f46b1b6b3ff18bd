"""One launch of the packed-bed reactor kernel for ncu (python tools/profile_pfr.py [n_traj])."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysodm_b200 import engine, tutorial_data as td
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
dev = engine.device(0)
rng = np.random.default_rng(1718)
post = rng.normal(1.0, 0.03, size=(5, n)) * np.array([7.88e-04, 0.055, 2.64, 0.079, 0.418])[:, None]
var = ([4, 5, 6, 8, 7], torch.from_numpy(post).to(dev))
out = torch.empty((200, 1, n), dtype=torch.float64, device=dev)
for rep in range(2):
    _, nfev, st = engine.ode_run("packed_pfr_ode", td.pfr_flat_y0(25), td.pfr_flat_par(25), n, 0, 8001, model_arg=25,
                                 out_last_only=True, var=var, out=out, dev=dev)
torch.cuda.synchronize()
print("ok", float(nfev.double().mean()), int(st.max()))
