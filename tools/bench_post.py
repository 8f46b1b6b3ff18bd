"""Device timing of the post-processing kernels on the BASELINE C3 ensemble (python tools/bench_post.py [n_traj] [out.json]).
CUDA events on the launching stream, 1 warm-up + 3 timed repetitions each; GB/s = algorithmic bytes / time."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

import bench
from pysodm_b200 import engine, tutorial_data as td


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def main(n=1_000_000, path=None):
    dev = engine.device(0)
    y0 = torch.from_numpy(td.influenza_flat_y0()).to(dev)
    tab = torch.from_numpy(td.CONTACT_TABLE.reshape(3, 16).copy()).to(dev)
    idx = torch.from_numpy(td.influenza_schedule()).to(dev)
    draws = torch.from_numpy(bench.synthetic_draws(n, 1718)).to(dev)
    ens = torch.empty((28, 162, n), dtype=torch.int32, device=dev)
    engine.jump_run("influenza_1718_jump", y0, td.influenza_flat_par(), n, 0, 162, 0.5, var=(bench.VAR_INDEX, draws),
                    tdpf=(tab, idx), seed=1, out_dtype=torch.int32, out=ens, dev=dev)
    torch.cuda.synchronize()
    nbytes = ens.numel() * 4
    res = {"n_traj": n, "ensemble_GB": nbytes / 1e9}
    noisy = torch.empty_like(ens)
    for kind, param in (("poisson", 0.0), ("negative_binomial", 0.03), ("gaussian_relative", 0.04)):
        out = noisy if kind != "gaussian_relative" else torch.empty((28, 162, n), dtype=torch.float32, device=dev)
        ms = timed(lambda: engine.noise(kind, ens, param, seed=3, out=out))
        res["noise_" + kind] = {"ms": ms, "GBps": (nbytes + out.numel() * out.element_size()) / ms / 1e6,
                                "elements_per_s": ens.numel() / ms * 1e3}
        del out
    ms = timed(lambda: engine.draws_sum(noisy))
    res["draws_sum"] = {"ms": ms, "GBps": nbytes / ms / 1e6}
    n_lo, n_hi = int(np.floor((n - 1) * 0.025)), int(np.floor((n - 1) * 0.975))
    for label, ranks in (("1_rank", [n // 2]), ("4_ranks_q025_q975", [n_lo, n_lo + 1, n_hi, n_hi + 1])):
        ms = timed(lambda: engine.draws_order_stats(noisy, ranks))
        res["order_stats_" + label] = {"ms": ms, "passes": 3, "GBps_one_read": nbytes / ms / 1e6,
                                       "GBps_three_reads": 3 * nbytes / ms / 1e6}
    f64 = torch.rand((28 * 8, n), dtype=torch.float64, device=dev)
    ms = timed(lambda: engine.draws_order_stats(f64, [n_lo, n_lo + 1, n_hi, n_hi + 1]))
    res["order_stats_f64_224_cells"] = {"ms": ms, "passes": 6, "GBps_six_reads": 6 * f64.numel() * 8 / ms / 1e6}
    # the reference's host path beside it, on a bounded sample: the NumPy calls of optimization/utils.py:30,64,91-92 and
    # of xarray's mean / quantile, on `cells` columns of the same ensemble
    import time
    cells = 16
    sample = ens[24, :cells].cpu().numpy().astype(np.float64)                 # Im_inc of one age group, 16 dates x n draws
    cpu = {"sample": f"{cells} cells x {n} draws (float64, as the reference holds them), one host thread"}
    for name, fn in (("noise_negative_binomial", lambda: np.random.negative_binomial(1 / 0.03, (1 / 0.03) / (sample + 1 / 0.03))),
                     ("noise_poisson", lambda: np.random.poisson(sample)),
                     ("noise_gaussian_relative", lambda: np.random.normal(sample, np.abs(0.04 * sample))),
                     ("mean", lambda: np.mean(sample, axis=1)),
                     ("quantile_q025_q975", lambda: np.quantile(sample, [0.025, 0.975], axis=1))):
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        cpu[name] = {"elements_per_s": sample.size / dt, "s_for_full_ensemble_extrapolated": dt * ens.numel() / sample.size}
    res["cpu_numpy_baseline"] = cpu
    print(json.dumps(res, indent=1))
    if path:
        json.dump(res, open(path, "w"), indent=1)


if __name__ == "__main__":
    main(int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000, sys.argv[2] if len(sys.argv) > 2 else None)
