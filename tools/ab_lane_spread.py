"""Small-ensemble launches of the tau-leap kernel: trajectories per warp (SODM_LANE_SHIFT=0..3, unset = automatic).
Usage: SODM_LANE_SHIFT=k python tools/ab_lane_spread.py  -> one JSON line {n: ms} + a checksum of the counts
(the C5 window: influenza, 168 steps, 85 outputs, per-trajectory beta / f_ud, Im_inc rows only)."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

from pysodm_b200 import engine, tutorial_data as td


def main():
    engine.device(0)
    rng = np.random.default_rng(0)
    sched = td.influenza_schedule()[:168]
    tdpf = (td.CONTACT_TABLE.reshape(3, 16), sched)
    res, check = {}, 0
    for n in (4096, 12_500, 25_000, 50_000, 100_000, 150_000, 1_000_000):
        var = {1: rng.uniform(0.01, 0.03, n)}
        run = lambda: engine.jump_run("influenza_1718_jump", td.influenza_flat_y0(), td.influenza_flat_par(), n, 0, 85, 0.5,
                                      tdpf=tdpf, var=var, seed=7, out_dtype=torch.int32, state_sel=[24, 25, 26, 27])
        out, st = run()
        torch.cuda.synchronize()
        check ^= int(out.to(torch.int64).sum().item()) * (n % 1000 + 1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        best = 1e9
        for _ in range(5):
            a.record()
            run()
            b.record()
            torch.cuda.synchronize()
            best = min(best, a.elapsed_time(b))
        res[n] = round(best, 3)
    print(json.dumps({"lane_shift": os.environ.get("SODM_LANE_SHIFT", "auto"), "ms": res, "checksum": check}))


if __name__ == "__main__":
    main()
