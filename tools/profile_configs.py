"""One or two launches of a non-headline kernel for ncu:  python tools/profile_configs.py <pfr25|pfr100|sir_ode|sir_radau|sir_bdf|ssa|loglik|sirsi> [n]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysodm_b200 import engine, tutorial_data as td

which = sys.argv[1]
dev = engine.device(0)
rng = np.random.default_rng(1718)
if which in ("pfr25", "pfr100"):
    nx = 25 if which == "pfr25" else 100
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (16384 if nx == 25 else 2048)
    post = rng.normal(1.0, 0.03, size=(5, n)) * np.array([7.88e-04, 0.055, 2.64, 0.079, 0.418])[:, None]
    var = ([4, 5, 6, 8, 7], torch.from_numpy(post).to(dev))
    out = torch.empty((8 * nx, 1, n), dtype=torch.float64, device=dev)
    for rep in range(2):
        _, nfev, st = engine.ode_run("packed_pfr_ode", td.pfr_flat_y0(nx), td.pfr_flat_par(nx), n, 0, 8001, model_arg=nx,
                                     out_last_only=True, var=var, out=out, dev=dev)
    torch.cuda.synchronize()
    print("ok", float(nfev.double().mean()), int(st.max()))
elif which in ("sir_ode", "sir_radau", "sir_bdf"):
    n = int(sys.argv[2]) if len(sys.argv) > 2 else (1_000_000 if which == "sir_ode" else 100_000)
    method = {"sir_ode": "RK45", "sir_radau": "Radau", "sir_bdf": "BDF"}[which]
    betas = torch.from_numpy(rng.uniform(0.25, 0.45, size=n)).to(dev)[None, :]
    out = torch.empty((3, 121, n), dtype=torch.float64, device=dev)
    for rep in range(2):
        engine.ode_run("sir_ode", [1000, 1, 0], [0.35, 5], n, 0, 121, var=([0], betas), out=out, dev=dev, method=method)
    torch.cuda.synchronize()
    print("ok")
elif which == "ssa":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 10_000
    small = np.array([607, 1329, 7352, 2204, 0, 0, 0, 0, 0, 0, 0, 0, 1000, 2], dtype=np.float64)
    out = torch.empty((14, 121, n), dtype=torch.float64, device=dev)
    for rep in range(2):
        engine.jump_run("sir_si_jump", small, td.sirsi_flat_par(), n, 0, 121, method="SSA", seed=rep, out=out, dev=dev)
    torch.cuda.synchronize()
    print("ok")
elif which == "sirsi":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1 << 18
    out = torch.empty((14, 121, n), dtype=torch.int32, device=dev)
    for rep in range(2):
        engine.jump_run("sir_si_jump", td.sirsi_flat_y0(), td.sirsi_flat_par(), n, 0, 121, 0.5, seed=rep, out_dtype=torch.int32, out=out, dev=dev)
    torch.cuda.synchronize()
    print("ok")
elif which == "loglik":
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1_000_000
    traj = torch.randint(0, 3000, (4, 85, n), dtype=torch.int32, device=dev)
    ydata = rng.gamma(2.0, 150.0, size=(13, 4))
    lp = torch.zeros(n, dtype=torch.float64, device=dev)
    for rep in range(2):
        engine.loglik("negative_binomial", traj, 0.0, [[0], [1], [2], [3]], np.arange(13) * 7.0, ydata, ll_arg=4 * [0.03], lp=lp, dev=dev)
    torch.cuda.synchronize()
    print("ok")
