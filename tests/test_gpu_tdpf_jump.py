"""GPU parity of jump processes whose time-dependent parameter functions read the STATES, the TIME or a parameter DRAWN per
trajectory, and of SSA with time-dependent parameters -- the reference's `func` evaluates them at every step / event with
the live states (/root/reference/src/pySODM/models/base.py:376-380, :398-401).  Everything is compared IN LAW (two-sample
KS per state and time) with ensembles made by the unmodified reference (tests/golden/make_golden.py::gen_tdpf_jump)."""
import os

import numpy as np
import pytest

import cases
import gen_models as gm
import tdpf_jump_models as tj

pytestmark = pytest.mark.gpu
GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "tdpf_jump.npz")


@pytest.fixture(scope="module")
def g():
    return np.load(GOLDEN)


def _flat(out, names, n, n_t):
    """Dataset -> [state row, time, draw]"""
    return np.concatenate([np.moveaxis(np.asarray(out[s].values, dtype=np.float64).reshape(n, n_t, -1), 0, -1).transpose(1, 0, 2)
                           for s in names], axis=0)


def test_tau_leap_state_and_time_dependent_tdpfs(g):
    m = gm.make_tdpf_jump()
    n = 20000
    out = m.sim(60, N=n, tau=0.5, seed=11, output_dtype="int32")
    mine = _flat(out, m.states_names, n, 61)[:, g["tl_times"], :]
    ref = np.transpose(g["tl_sorted"], (1, 2, 0)).astype(np.float64)             # [3, T, 2000]
    cases.assert_ks_ok(cases.ks_pvalues(mine, ref), "tau-leap with state- and time-dependent TDPFs vs reference")
    assert np.allclose(mine.mean(axis=-1), g["tl_mean"][:, g["tl_times"]], rtol=0.05, atol=3.0)
    # the feedback matters: without it the epidemic is far larger (the test has power)
    plain = gm.TdpfSIR(dict(tj.TAU_LEAP["initial_states"]), {"beta": 0.7, "gamma": 0.2})
    big = plain.sim(60, N=2000, tau=0.5, seed=12, output_dtype="int32")
    assert big["R"].values[:, -1].mean() > 1.5 * out["R"].values[:, -1].mean()


def test_drawn_parameter_feeds_a_tdpf(g):
    m = gm.make_tdpf_jump()
    n = 20000
    np.random.seed(5)
    out = m.sim(60, N=n, tau=0.5, seed=13, output_dtype="int32", draw_function=tj.draw_k_fb)
    mine = _flat(out, m.states_names, n, 61)[:, g["tl_times"], :]
    ref = np.transpose(g["tld_sorted"], (1, 2, 0)).astype(np.float64)
    cases.assert_ks_ok(cases.ks_pvalues(mine, ref), "per-trajectory k_fb feeding the TDPF of beta vs reference")
    # and it differs from the fixed-k_fb law (power): the spread of the final size is much wider
    fixed = np.transpose(g["tl_sorted"], (1, 2, 0)).astype(np.float64)
    assert mine[2, -1].std() > 1.5 * fixed[2, -1].std()
    assert m.parameters["k_fb"] == tj.TAU_LEAP["parameters"]["k_fb"]              # parameters restored after the run


def test_ssa_with_tdpfs_at_event_times(g):
    m = gm.make_tdpf_jump(tj.SSA)
    n = 6000
    out = m.sim(40, N=n, method="SSA", seed=17)
    mine = _flat(out, m.states_names, n, 41)[:, g["ssa_times"], :]
    ref = np.transpose(g["ssa_ens"], (1, 2, 0)).astype(np.float64)                # [3, T, 600]
    cases.assert_ks_ok(cases.ks_pvalues(mine, ref), "SSA with state- and time-dependent TDPFs vs reference")


def test_hand_written_influenza_ssa_with_calendar_contact_function(g):
    """SSA of the tutorial model (hand-written functor) on a population / 1000: the calendar contact function cannot be
    translated, it is tabulated per day (entry 0: its value at t0 itself) and read at event times."""
    from pysodm_b200 import tutorial_data as td
    from pysodm_b200 import tutorial_models as tm
    y0 = g["flu_small_y0"].reshape(7, 4)
    init = {s: y0[k] for k, s in enumerate(td.INFLUENZA_STATES)}
    m = tm.JumpProcess_influenza_model(init, td.INFLUENZA_PARAMETERS, coordinates={"age_group": td.AGE_GROUPS},
                                       time_dependent_parameters={"N": td.influenza_contact_function})
    n, n_out = 4000, int(g["flu_small_n_out"])
    out = m.sim([td.INFLUENZA_START, "2018-02-01"], N=n, method="SSA", seed=19)
    mine = _flat(out, td.INFLUENZA_STATES, n, n_out)[:, g["flu_small_times"], :]
    ref = np.transpose(g["flu_small_ens"], (1, 2, 0)).astype(np.float64)
    cases.assert_ks_ok(cases.ks_pvalues(mine[:24], ref[:24]), "influenza/1000 SSA (hand-written functor, day table) vs reference")
    assert np.allclose(mine[24:].mean(axis=-1), ref[24:].mean(axis=-1), atol=0.08)
    # the translated model of the same class gives the same law (its SSA functor comes from the traced rate graph)
    gen = gm.GenInfluenza(init, td.INFLUENZA_PARAMETERS, coordinates={"age_group": td.AGE_GROUPS},
                          time_dependent_parameters={"N": td.influenza_contact_function})
    out_g = gen.sim([td.INFLUENZA_START, "2018-02-01"], N=n, method="SSA", seed=23)
    mine_g = _flat(out_g, td.INFLUENZA_STATES, n, n_out)[:, g["flu_small_times"], :]
    cases.assert_ks_ok(cases.ks_pvalues(mine_g[:24], ref[:24]), "influenza/1000 SSA (translated model, day table) vs reference")


def test_hand_written_functor_hands_over_to_the_translated_model(g):
    """`sir_jump` has no table slot at all: a TDPF on beta / gamma moves the class to the translated route, same law."""
    from pysodm_b200 import tutorial_models as tm

    class HandSir(tm.JumpProcess_SIR):
        @staticmethod
        def compute_rates(t, S, I, R, beta, gamma):
            return tj.compute_rates(t, S, I, R, beta, gamma)
    m = HandSir(dict(tj.TAU_LEAP["initial_states"]), dict(tj.TAU_LEAP["parameters"]),
                time_dependent_parameters={"beta": tj.feedback, "gamma": tj.slower_recovery})
    assert m._generated and m._hand_written == "sir_jump"
    n = 20000
    out = m.sim(60, N=n, tau=0.5, seed=29, output_dtype="int32")
    mine = _flat(out, m.states_names, n, 61)[:, g["tl_times"], :]
    ref = np.transpose(g["tl_sorted"], (1, 2, 0)).astype(np.float64)
    cases.assert_ks_ok(cases.ks_pvalues(mine, ref), "hand-written class with TDPFs -> translated model vs reference")


def test_untranslatable_state_dependent_tdpf_is_refused():
    """No silent tabulation: a function that reads the states AND cannot be translated raises."""
    def calendar_and_states(t, states, param):
        return param * (0.5 if int(states["I"][0]) % 7 == 0 else 1.0)          # int() of a model value: not translatable
    m = gm.TdpfSIR(dict(tj.TAU_LEAP["initial_states"]), {"beta": 0.7, "gamma": 0.2},
                   time_dependent_parameters={"beta": calendar_and_states})
    with pytest.raises(NotImplementedError, match="cannot be translated"):
        m.sim(10, N=8, tau=0.5)
