import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pysodm_b200 import engine, tutorial_data as td
n = 100000
dev = engine.device(0)
rng = np.random.default_rng(1718)
post = rng.normal(1.0, 0.03, size=(5, n)) * np.array([7.88e-04, 0.055, 2.64, 0.079, 0.418])[:, None]
var = ([4, 5, 6, 8, 7], torch.from_numpy(post).to(dev))
out = torch.empty((200, 1, n), dtype=torch.float64, device=dev)
ts=[]
for rep in range(3):
    a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
    a.record()
    _, nfev, st = engine.ode_run("packed_pfr_ode", td.pfr_flat_y0(25), td.pfr_flat_par(25), n, 0, 8001, model_arg=25, out_last_only=True, var=var, out=out, dev=dev)
    b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
print("pfr 1e5 draws: %.1f ms  checksum %.6f nfev %.1f" % (min(ts), float(out.sum()), float(nfev.double().mean())))
