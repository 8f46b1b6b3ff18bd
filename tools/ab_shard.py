"""The C3 launch at the shard sizes of a strong-scaling run (10^6 / G trajectories, full season) for every tuning build under
.variants/ and the in-tree library.   python tools/ab_shard.py"""
import glob, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
code = f"""
import sys; sys.path.insert(0, {ROOT!r})
import torch
from pysodm_b200 import engine, tutorial_data as td
import bench
dev = engine.device(0)
y0 = torch.from_numpy(td.influenza_flat_y0()).to(dev)
tab = torch.from_numpy(td.CONTACT_TABLE.reshape(3, 16).copy()).to(dev)
idx = torch.from_numpy(td.influenza_schedule()).to(dev)
row = []
for n in (125000, 250000, 500000, 1000000):
    draws = torch.from_numpy(bench.synthetic_draws(n, 1718)).to(dev)
    out = torch.empty((28, 162, n), dtype=torch.int32, device=dev)
    ts = []
    for rep in range(5):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        engine.jump_run("influenza_1718_jump", y0, td.influenza_flat_par(), n, 0, 162, 0.5, var=(bench.VAR_INDEX, draws), tdpf=(tab, idx), seed=rep,
                        out_dtype=torch.int32, out=out, dev=dev)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    row.append(f"{{n}}: {{min(ts[1:]):.2f}} ms")
    del out
print("  " + "   ".join(row))
"""
for lib in [None] + sorted(glob.glob(os.path.join(ROOT, ".variants", "libsodm_*.so"))):
    env = dict(os.environ)
    if lib: env["PYSODM_B200_LIB"] = lib
    print(os.path.basename(lib) if lib else "in-tree libsodm.so", flush=True)
    r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True)
    print(r.stdout.rstrip() if r.returncode == 0 else "FAILED\n" + r.stderr[-1500:], flush=True)
