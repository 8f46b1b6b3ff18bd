"""SIR_SI SSA (populations / 1000, 120 days) at several ensemble sizes: python tools/time_ssa.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysodm_b200 import engine, tutorial_data as td
dev = engine.device(0)
small = np.array([607, 1329, 7352, 2204, 0, 0, 0, 0, 0, 0, 0, 0, 1000, 2], dtype=np.float64)
for n in (2048, 10_000, 100_000):
    out = torch.empty((14, 121, n), dtype=torch.float64, device=dev)
    ts = []
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        engine.jump_run("sir_si_jump", small, td.sirsi_flat_par(), n, 0, 121, method="SSA", seed=rep, out=out, dev=dev)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    print(f"n = {n:7d}: {min(ts[1:]):9.2f} ms  {n / min(ts[1:]) * 1e3:.3e} traj/s   mean R at the end {float(out[8:12, -1].sum(0).mean()):.1f}")
