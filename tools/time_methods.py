"""C1 (tutorial SIR, 121 grid points) and the batch reactor (PPBB, 1501 grid points) with every adaptive method the device
offers, at 10^5 draws: python tools/time_methods.py  ->  one line per (model, method)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysodm_b200 import engine
dev = engine.device(0)
rng = np.random.default_rng(1718)
n = 100_000
betas = torch.from_numpy(rng.uniform(0.25, 0.45, size=n)).to(dev)[None, :]
vf = torch.from_numpy(rng.uniform(6e-4, 1e-3, size=n)).to(dev)[None, :]
cases = (("sir_ode", [1000, 1, 0], [0.35, 5], 121, ([0], betas)),
         ("ppbb_ode", [46.0, 61.0, 37.0, 0.0], [10, 7.95e-04, 0.047, 1.10, 0.131, 0.66], 1501, ([1], vf)))
for model, y0, par, n_out, var in cases:
    out = torch.empty((len(y0), n_out, n), dtype=torch.float64, device=dev)
    for method in ("RK45", "RK23", "DOP853", "Radau", "BDF"):
        ts = []
        for rep in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            _, nfev, st = engine.ode_run(model, y0, par, n, 0, n_out, var=var, out=out, dev=dev, method=method)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        ms = min(ts[1:])
        print(f"{model:9s} {method:7s} n = {n}: {ms:9.3f} ms  {n / ms * 1e3:.3e} traj/s  mean nfev {float(nfev.double().mean()):8.2f}  "
              f"status {int(st.abs().max())}", flush=True)
