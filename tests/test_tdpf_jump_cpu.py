"""Time-dependent parameter functions of JUMP processes that read the states / the time / a drawn parameter, and SSA with
time-dependent parameters (reference `func`: /root/reference/src/pySODM/models/base.py:376-380, :398-401) -- CPU side:
  * the oracle's restatement is pinned bit for bit to seeded runs of the unmodified reference (tests/golden/tdpf_jump.npz),
  * the kernel code (host build of the same __host__ __device__ sources) reproduces the reference's SSA ensemble of the
    influenza model with its calendar contact function, through the per-day table the front end builds,
  * the front end's decision: translated (runs on the device, may read states) vs tabulated (must not)."""
import os
import random
from datetime import datetime

import numpy as np
import pytest

import cases
import gen_models as gm
import hostcheck_util as hc
import tdpf_jump_models as tj
from oracle import jump, models
from pysodm_b200 import tutorial_data as td
from pysodm_b200 import tutorial_models as tm

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tdpf_jump.npz")
TDPFS = [("beta", tj.feedback, ["k_fb"]), ("gamma", tj.slower_recovery, ["t_switch"])]


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN)


def _reseed_like_reference(seed):
    np.random.seed(int(seed))
    np.random.seed(np.random.randint(0, 2 ** 32, size=1))


def test_oracle_tau_leap_with_tdpfs_bit_exact(g):
    for i, s in enumerate(g["tl_seeds"]):
        _reseed_like_reference(s)
        _, y = jump.solve_discrete(models.SIRJump, g["tl_y0"], tj.TAU_LEAP["parameters"], 0, 60, tau=0.5,
                                   tdpf_functions=TDPFS, model_params=("beta", "gamma"))
        assert np.array_equal(y, g["tl_y"][i])


def test_oracle_ssa_with_tdpfs_bit_exact(g):
    for i, s in enumerate(g["ssa_seeds"]):
        _reseed_like_reference(s)
        random.seed(1000 + int(s))
        _, y = jump.solve_discrete(models.SIRJump, g["ssa_y0"], tj.SSA["parameters"], 0, 40, method="SSA",
                                   tdpf_functions=TDPFS, model_params=("beta", "gamma"))
        assert np.array_equal(y, g["ssa_y"][i])


def _small_flu():
    m = tm.JumpProcess_influenza_model(td.INFLUENZA_INITIAL_STATES, td.INFLUENZA_PARAMETERS,
                                       coordinates={"age_group": td.AGE_GROUPS},
                                       time_dependent_parameters={"N": td.influenza_contact_function})
    return m


def test_kernel_code_ssa_influenza_with_day_table_matches_reference_ensemble(g):
    """hand-written influenza functor, SSA, contact matrix from the per-interval table (entry 0 = value at t0, entry 1 + k =
    value inside day k) vs 600 SSA runs of the unmodified reference on the same /1000 population."""
    m = _small_flu()
    n_out = int(g["flu_small_n_out"])
    table, index = m._tabulate_tdpf_intervals(0.0, float(n_out - 1), 1.0, datetime(2017, 12, 1), "D")
    out, st = hc.jump_run("influenza_1718_jump", g["flu_small_y0"], td.influenza_flat_par(), 3000, 0, n_out, tau=1.0,
                          method="SSA", tdpf=(table, index), seed=77)
    assert not st.any()
    ref = g["flu_small_ens"]                                   # [600, 28, n_times]
    mine = out[:, g["flu_small_times"], :]                     # [28, n_times, 3000]
    ps = cases.ks_pvalues(mine[:24], np.moveaxis(ref, 0, -1)[:24])
    cases.assert_ks_ok(ps, "influenza/1000 SSA, kernel code on the host vs reference")
    # Im_inc is the Ip->Id indicator of the LAST event before the output time, interpolated: compare its mean
    assert np.allclose(mine[24:].mean(axis=-1), np.moveaxis(ref, 0, -1)[24:].mean(axis=-1), atol=0.08)
    # population is conserved per age group at every output time
    tot = out[:24].reshape(6, 4, n_out, -1).sum(axis=0)
    assert np.allclose(tot, tot[:, :1])


def test_front_end_translates_what_it_can_and_tabulates_the_rest():
    m = gm.make_tdpf_jump()
    assert m._generated and m.device_model is not None
    assert m._traced_tdpfs == ["beta", "gamma"] and list(m._spec.tdpf_params) == []
    assert list(m._spec_params) == ["beta", "gamma", "k_fb", "t_switch"]          # TDPF-only parameters travel too
    assert m.device_model.layout["uses_time"]
    # calendar logic cannot be translated: tabulated, and the states guard applies
    flu = gm.GenInfluenza(td.INFLUENZA_INITIAL_STATES, td.INFLUENZA_PARAMETERS, coordinates={"age_group": td.AGE_GROUPS},
                          time_dependent_parameters={"N": td.influenza_contact_function})
    flu._bind_generated_functor(86400e6, [datetime(2017, 12, 1), datetime(2018, 5, 11)])
    assert flu._traced_tdpfs == [] and list(flu._spec.tdpf_params) == ["N"] and "N" in flu._tdpf_reason
    # the generated functor of the translated model reproduces the Python functions at a later time and another state
    from pysodm_b200 import plugin
    par = np.array([0.7, 0.2, 0.002, 20.0])
    for t, y in ((3.5, [19000, 700, 300]), (27.0, [5000, 2500, 12500])):
        n, p = plugin.host_propose(m.device_model, y, par, 0.5, quirk=True, t=t)
        states = {"S": np.array([y[0]], float), "I": np.array([y[1]], float), "R": np.array([y[2]], float)}
        beta = tj.feedback(t, states, 0.7, 0.002)
        gamma = tj.slower_recovery(t, states, 0.2, 20)
        rates = tj.compute_rates(t, **states, beta=beta, gamma=gamma)
        want = [1 - np.exp(-0.5 * float(rates["S"][0][0])), 1 - np.exp(-0.5 * float(rates["I"][0][0]))]
        want = [w / (1 + w) for w in want]
        assert np.allclose(p, want, rtol=2e-5) and list(n) == y[:2]
