"""C1 (tutorial SIR, RK45, 121 grid points) at several ensemble sizes: python tools/time_c1.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysodm_b200 import engine
dev = engine.device(0)
rng = np.random.default_rng(1718)
for n in (100, 100_000, 1_000_000, 4_000_000):
    betas = torch.from_numpy(rng.uniform(0.25, 0.45, size=n)).to(dev)[None, :]
    out = torch.empty((3, 121, n), dtype=torch.float64, device=dev)
    ts = []
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _, nfev, _ = engine.ode_run("sir_ode", [1000, 1, 0], [0.35, 5], n, 0, 121, var=([0], betas), out=out, dev=dev)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = min(ts[1:])
    print(f"n = {n:8d}: {ms:8.3f} ms  {n / ms * 1e3:.3e} traj/s  {n * 2904 / ms / 1e6:.1f} GB/s of output  mean nfev {float(nfev.double().mean()):.2f}")
