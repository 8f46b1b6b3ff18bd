"""C4 timing: packed-bed reactor, RK45 to t = 8000 s, final profile only.  python tools/time_pfr.py [nx ...]"""
import os, sys
sys.path.insert(0, os.getcwd())
import numpy as np, torch
from pysodm_b200 import engine, tutorial_data as td
dev = engine.device(0)
rng = np.random.default_rng(1718)
for nx in [int(a) for a in sys.argv[1:]] or [25, 50, 100]:
    n = {25: 100000, 50: 32768, 100: 8192}.get(nx, 8192)
    post = rng.normal(1.0, 0.03, size=(5, n)) * np.array([7.88e-04, 0.055, 2.64, 0.079, 0.418])[:, None]
    var = ([4, 5, 6, 8, 7], torch.from_numpy(post).to(dev))
    out = torch.empty((8 * nx, 1, n), dtype=torch.float64, device=dev)
    ts = []
    for rep in range(3):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        _, nfev, st = engine.ode_run("packed_pfr_ode", td.pfr_flat_y0(nx), td.pfr_flat_par(nx), n, 0, 8001, model_arg=nx,
                                     out_last_only=True, var=var, out=out, dev=dev)
        b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    print("pfr nx=%d, %d draws: %.1f ms (%.3g draws/s)  checksum %.9f  nfev %.2f  status %d" %
          (nx, n, min(ts), n / min(ts) * 1e3, float(out.sum()), float(nfev.double().mean()), int(st.max())))
