"""A/B timing of kernel variants on the influenza tau-leap workload (CUDA events, device-resident)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from pysodm_b200 import engine, tutorial_data as td

def main(n=1 << 20):
    dev = engine.device(0)
    y0 = torch.from_numpy(td.influenza_flat_y0()).to(dev)
    par = td.influenza_flat_par()
    tab = torch.from_numpy(td.CONTACT_TABLE.reshape(3, 16).copy()).to(dev)
    idx = torch.from_numpy(td.influenza_schedule()).to(dev)
    out = torch.empty((28, 162, n), dtype=torch.int32, device=dev)
    for name, compat in (("per-lane walk (default)", 3), ("two-pass sampler", 3 | 256)):
        ts = []
        for rep in range(4):
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            engine.jump_run("influenza_1718_jump", y0, par, n, 0, 162, 0.5, tdpf=(tab, idx), seed=rep, compat=compat,
                            out_dtype=torch.int32, out=out, dev=dev)
            b.record(); torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        print(f"{name}: {min(ts[1:]):.1f} ms for {n} trajectories -> {n / min(ts[1:]) * 1e3 / 1e6:.2f} M traj/s; checksum {int(out[24:28, -1].sum())}")

if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20)
