"""cProfile of the class-API call `JumpProcess.sim(N=10^6, draw_function=...)` (host side of bench.py's e2e_api).
python tools/profile_api_host.py [n]"""
import cProfile, os, pstats, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
from pysodm_b200 import tutorial_data as td, tutorial_models as tm

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
model = tm.JumpProcess_influenza_model(dict(td.INFLUENZA_INITIAL_STATES), dict(td.INFLUENZA_PARAMETERS),
                                       coordinates={"age_group": td.AGE_GROUPS},
                                       time_dependent_parameters={"N": td.influenza_contact_function})


def draw_vec(parameters, N, rng_seed):
    d = bench.synthetic_draws(N, rng_seed)
    parameters["beta"] = d[0]
    parameters["f_ud"] = np.ascontiguousarray(d[1:].T)
    return parameters
draw_vec.vectorized = True
span = [td.INFLUENZA_START, td.INFLUENZA_END]
model.sim(span, N=4096, tau=0.5, draw_function=draw_vec, draw_function_kwargs={"rng_seed": 1}, output_dtype="int32", seed=1)
for rep in range(3):
    t0 = time.perf_counter()
    pr = cProfile.Profile()
    pr.enable()
    ds = model.sim(span, N=n, tau=0.5, draw_function=draw_vec, draw_function_kwargs={"rng_seed": 7 + rep}, output_dtype="int32", seed=7)
    pr.disable()
    print(f"call {rep}: {time.perf_counter() - t0:.3f} s")
    del ds
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
