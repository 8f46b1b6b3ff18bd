"""Do the 64-register and the 96-register build of the tau-leap kernel give the same counts?  One-wave launch (60 000
trajectories, full season) in two processes: default dispatch (96 registers) and SODM_LATENCY_BUILD=0 (64 registers).
python tools/ab_builds_identical.py   (PYSODM_B200_LIB selects a tuning build)"""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = f"""
import sys; sys.path.insert(0, {ROOT!r})
import numpy as np, torch
from pysodm_b200 import engine, tutorial_data as td
import bench
n = 60000
dev = engine.device(0)
draws = torch.from_numpy(bench.synthetic_draws(n, 1718)).to(dev)
out, st = engine.jump_run("influenza_1718_jump", td.influenza_flat_y0(), td.influenza_flat_par(), n, 0, 162, 0.5, var=(bench.VAR_INDEX, draws),
                          tdpf=(td.CONTACT_TABLE.reshape(3, 16), td.influenza_schedule()), seed=7, out_dtype=torch.int32, dev=dev)
np.save(sys.argv[1], out[:, -1, :].cpu().numpy())
"""
outs = []
for tag, val in (("latency", None), ("throughput", "0")):
    env = dict(os.environ)
    env.pop("SODM_LATENCY_BUILD", None)
    if val is not None:
        env["SODM_LATENCY_BUILD"] = val
    path = f"/tmp/ab_build_{tag}.npy"
    r = subprocess.run([sys.executable, "-c", code, path], env=env, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr[-2000:]
    outs.append(path)
import numpy as np
a, b = np.load(outs[0]), np.load(outs[1])
print("library:", os.environ.get("PYSODM_B200_LIB", "in-tree"), "-> trajectories whose final state differs between the builds:",
      int((a != b).any(axis=0).sum()), "of", a.shape[1])
