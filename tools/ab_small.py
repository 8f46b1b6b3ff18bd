"""Small / sharded launches: walk (default 64-register build / SODM_TUNE_LATENCY = 96-register build) against queue sampler (SODM_TUNE_QUEUE) for ensembles that do not fill the GPU.
python tools/ab_small.py   (PYSODM_B200_LIB selects a tuning build)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from pysodm_b200 import engine, tutorial_data as td
import bench

dev = engine.device(0)
y0 = torch.from_numpy(td.influenza_flat_y0()).to(dev)
tab = torch.from_numpy(td.CONTACT_TABLE.reshape(3, 16).copy()).to(dev)
for n_out, label in ((85, "C5 window (168 steps)"), (162, "full season (322 steps)")):
    idx = torch.from_numpy(td.influenza_schedule()[:2 * (n_out - 1)].copy()).to(dev)
    print(label)
    for n in (1024, 4096, 12500, 32768, 125000):
        draws = torch.from_numpy(bench.synthetic_draws(n, 1718)).to(dev)
        out = torch.empty((28, n_out, n), dtype=torch.int32, device=dev)
        row = []
        for compat in (3, 3 | 512, 3 | 256):
            ts = []
            for rep in range(6):
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record()
                engine.jump_run("influenza_1718_jump", y0, td.influenza_flat_par(), n, 0, n_out, 0.5, var=(bench.VAR_INDEX, draws),
                                tdpf=(tab, idx), seed=rep, compat=compat, out_dtype=torch.int32, out=out, dev=dev)
                b.record(); torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            row.append(min(ts[2:]))
        print(f"  n = {n:7d}: walk {row[0]:7.3f} ms   walk, SODM_TUNE_LATENCY {row[1]:7.3f} ms   queue {row[2]:7.3f} ms")
