"""Stage the UNMODIFIED reference for timing on the GPU box's host cores -- TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE.

`/root/reference` exists in the build container only.  `stage()` (called by `__graft_entry__.build()`) copies the files the
influenza tau-leap run needs -- the pure-Python package `src/pySODM`, the tutorial's `models.py` and its data file -- byte
for byte into `oracle/_ref/reference/` (git-ignored: reference sources never enter the history; NOT gpurun-ignored, so the
copy travels to the GPU box like a built `.so`).  `time_unmodified()` then runs the reference's own public call,
`JumpProcess.sim([start, end], N, tau=0.5, processes=cores)` (base.py:442-571), set up exactly as the tutorial does
(tests/golden/make_golden.py:make_influenza restates calibration.py:82-117 for the fixtures), and reports trajectories/s.
xarray is not installed in this image; tests/golden/ref_loader.py replaces the final `_output_to_xarray_dataset` by the
identity (documented there) -- the simulation itself is untouched.
"""
import os
import shutil
import sys
import time

HERE = os.path.dirname(os.path.abspath(__file__))
STAGED = os.path.join(HERE, "_ref", "reference")
SOURCE = "/root/reference"
FILES = ["tutorials/influenza_1718/models.py", "tutorials/influenza_1718/data/interim/ILI_weekly_ABS.csv"]


def stage():
    """Copy the reference files into oracle/_ref/reference (no-op when /root/reference is absent). Returns the path or None."""
    if not os.path.isdir(os.path.join(SOURCE, "src", "pySODM")):
        return STAGED if staged() else None
    if os.path.isdir(STAGED):
        shutil.rmtree(STAGED)
    shutil.copytree(os.path.join(SOURCE, "src", "pySODM"), os.path.join(STAGED, "src", "pySODM"),
                    ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    for rel in FILES:
        dst = os.path.join(STAGED, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SOURCE, rel), dst)
    return STAGED


def staged():
    return os.path.isfile(os.path.join(STAGED, "src", "pySODM", "models", "base.py"))


def time_unmodified(n_traj, processes):
    """Wall clock of the unmodified reference's `sim` for `n_traj` influenza trajectories on `processes` worker processes.
    -> (trajectories/s, seconds, mean Im_inc at the last day -- a sanity figure)."""
    if not staged():
        raise RuntimeError("oracle/_ref/reference is not staged (run __graft_entry__.build() where /root/reference exists)")
    import numpy as np
    os.environ["PYSODM_REFERENCE_ROOT"] = STAGED
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "tests", "golden"))
    import make_golden                                   # imports the staged reference through ref_loader
    model, _, _, start = make_golden.make_influenza()
    np.random.seed(1718)
    model.sim([start, "2017-12-10"], N=processes, tau=0.5, processes=processes)      # warm-up: imports, pool start
    t = time.perf_counter()
    res = model.sim([start, "2018-05-11"], N=n_traj, tau=0.5, processes=processes)
    dt = time.perf_counter() - t
    y = np.stack([r["y"] for r in res])
    return n_traj / dt, dt, float(y[:, 24:28, -1].mean())


if __name__ == "__main__":
    print(stage())
    if len(sys.argv) > 1:
        print(time_unmodified(int(sys.argv[1]), os.cpu_count()))
