"""Stiff problems (generated Robertson / Van der Pol models) with Radau and BDF on the device against the reference fixtures:
nfev differences and value errors per draw.  python tools/check_stiff_gpu.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np
import gen_models as gm
gs = np.load(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "ode_stiff.npz"))
rob, vdp = gm.make_robertson(), gm.make_vanderpol()
def draw_k1(parameters, N):
    parameters["k1"] = gs["rob_k1"]; return parameters
draw_k1.vectorized = True
def draw_mu(parameters, N):
    parameters["mu"] = gs["vdp_mu"]; return parameters
draw_mu.vectorized = True
for method in ("Radau", "BDF"):
    out = rob.sim(40, N=6, draw_function=draw_k1, method=method)
    y = np.stack([out[s].values for s in ("y1", "y2", "y3")], axis=1)
    ref = gs[f"rob_{method}_y"]
    print(method, "rob nfev diff", rob.last_nfev.cpu().numpy() - gs[f"rob_{method}_nfev"], "maxrel", np.max(np.abs(y - ref) / (np.abs(ref) + 1e-12)))
    out = vdp.sim(200, N=6, draw_function=draw_mu, method=method)
    y = np.stack([out[s].values for s in ("x", "v")], axis=1)
    ref = gs[f"vdp_{method}_y"]
    print(method, "vdp nfev diff", vdp.last_nfev.cpu().numpy() - gs[f"vdp_{method}_nfev"], "maxabs per draw", np.abs(y - ref).max(axis=(1, 2)), "median", np.median(np.abs(y-ref)))
