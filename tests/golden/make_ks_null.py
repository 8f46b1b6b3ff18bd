"""Null distribution of the ensemble KS parity test: how many (state, time) cells fall below
p = 0.01, and how small the smallest p-value gets, when BOTH samples come from the same law.

The parity tests compare an ensemble from the device with an ensemble from the reference over hundreds of (state, time)
cells that are strongly correlated (the same trajectories at neighbouring times, coupled compartments), so the number of
cells below 0.01 is NOT Binomial(cells, 0.01).  This script measures it: R independent pairs of ensembles from the pinned
oracle (oracle/jump.py, bit-exact with the unmodified reference on seeded runs), same sizes / states / output times as the
fixtures the tests use (tests/golden/jump_ensemble_20k.npz).  tests/cases.assert_ks_ok takes its allowance from the result
(tests/golden/ks_null.npz) instead of an argued constant.

    python tests/golden/make_ks_null.py [replicates] [processes]          (~5 min on 8 cores for 40 replicates)
"""
import os
import sys
from multiprocessing import Pool

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(HERE))


def _ensemble(args):
    which, n, seed, times = args
    from oracle import jump, models
    from pysodm_b200 import tutorial_data as td
    if which == "flu":
        par = {k: v for k, v in td.INFLUENZA_PARAMETERS.items() if k != "N"}
        y = jump.tau_leap_ensemble(models.InfluenzaJump, td.influenza_flat_y0(), par, 0, 161, 0.5, n,
                                   tdpf={"N": (td.CONTACT_TABLE, td.influenza_schedule())}, seed=seed)
    else:
        y = jump.tau_leap_ensemble(models.SIRSIJump, td.sirsi_flat_y0(), dict(td.SIRSI_PARAMETERS), 0, 120, 0.5, n, seed=seed)
    return np.sort(y[:, :, times], axis=0).astype(np.int32)         # [n, states, times]


def main():
    import cases
    R = int(sys.argv[1]) if len(sys.argv) > 1 else 40
    procs = int(sys.argv[2]) if len(sys.argv) > 2 else os.cpu_count()
    big = np.load(os.path.join(HERE, "jump_ensemble_20k.npz"))
    out = {}
    for which, times, n in (("flu", big["flu_times"], 20000), ("sirsi", big["sirsi_times"], 20000)):
        jobs = [(which, n, 900_000 + 2 * r + s, times) for r in range(R) for s in (0, 1)]
        with Pool(procs) as pool:
            ens = pool.map(_ensemble, jobs, chunksize=1)
        below, min_p, frac5 = [], [], []
        for r in range(R):
            a, b = ens[2 * r], ens[2 * r + 1]
            ps = cases.ks_pvalues(np.transpose(a, (1, 2, 0)).astype(np.float64), np.transpose(b, (1, 2, 0)).astype(np.float64)).ravel()
            below.append(int((ps < 0.01).sum()))
            frac5.append(int((ps < 0.05).sum()))
            min_p.append(float(ps.min()))
        out[which + "_cells"] = np.array(ps.size)
        out[which + "_members"] = np.array(n)
        out[which + "_below_001"] = np.array(below)
        out[which + "_below_005"] = np.array(frac5)
        out[which + "_min_p"] = np.array(min_p)
        print(which, "cells", ps.size, "below 0.01 per replicate:", below, "min p:", ["%.1e" % x for x in min_p])
    np.savez_compressed(os.path.join(HERE, "ks_null.npz"), **out)


if __name__ == "__main__":
    main()
