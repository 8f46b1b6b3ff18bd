"""Host-side profile (cProfile) of the device-resident sampler on the C5 posterior: where an iteration's wall time goes
outside the kernels.  Usage: python tools/profile_mcmc_host.py [n_walkers]"""
import cProfile
import os
import pstats
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import pandas as pd
import torch

from pysodm_b200 import tutorial_data as td, tutorial_models as tm
from pysodm_b200.objective_functions import ll_negative_binomial, log_posterior_probability
from pysodm_b200.optimization import mcmc


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 100_000
    model = tm.JumpProcess_influenza_model(td.INFLUENZA_INITIAL_STATES, td.INFLUENZA_PARAMETERS,
                                           coordinates={"age_group": td.AGE_GROUPS},
                                           time_dependent_parameters={"N": td.influenza_contact_function})
    dates = pd.date_range("2017-12-01", periods=13, freq="7D")
    idx = pd.MultiIndex.from_product([dates, td.AGE_GROUPS], names=["date", "age_group"])
    rng = np.random.default_rng(0)
    data = pd.Series(rng.gamma(2.0, 150.0, size=len(idx)), index=idx, name="CASES")
    lpp = log_posterior_probability(model, ["beta", "f_ud"], [(1e-6, 0.08), (1e-3, 1 - 1e-3)], [data], ["Im_inc"],
                                    [ll_negative_binomial], [4 * [0.03]], simulation_kwargs={"tau": 0.5})
    pos = np.column_stack([rng.uniform(0.01, 0.03, n), rng.uniform(0.05, 0.95, size=(n, 4))])
    lp0 = lpp.batch(pos.copy())
    s = mcmc.EnsembleSampler(n, 5, lpp, moves=mcmc.StretchMove(), seed=1, device_update=True)
    s.run_mcmc(pos, 5, log_prob0=lp0)
    torch.cuda.synchronize()
    s = mcmc.EnsembleSampler(n, 5, lpp, moves=mcmc.StretchMove(), seed=1, device_update=True)
    prof = cProfile.Profile()
    prof.enable()
    s.run_mcmc(pos, 20, log_prob0=lp0)
    torch.cuda.synchronize()
    prof.disable()
    st = pstats.Stats(prof)
    st.sort_stats("cumulative").print_stats(45)
    st.sort_stats("tottime").print_stats(25)
    import io
    from contextlib import redirect_stdout
    from pysodm_b200 import optimization
    prof = cProfile.Profile()
    np.random.seed(0)
    prof.enable()
    with redirect_stdout(io.StringIO()):
        optimization.pso(lpp, swarmsize=n, max_iter=10, device_update=True)
    torch.cuda.synchronize()
    prof.disable()
    st = pstats.Stats(prof)
    st.sort_stats("tottime").print_stats(25)


if __name__ == "__main__":
    main()
