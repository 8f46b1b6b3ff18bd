"""GPU parity tests of GENERATED models (codegen.py + plugin.py): model classes that declare no hand-written functor
run from their own Python functions.  Bars as for the built-in functors: tau-leap / SSA in law (KS against ensembles
produced by the unmodified reference, or by the pinned oracle for models the reference fixtures do not hold), ODE within
rtol 1e-6 of the reference's scipy RK45 and with the same step sequence as the hand-written functors."""
import numpy as np
import pytest

import cases
import pandas as pd
import gen_models as gm
from oracle import jump as ojump
from oracle import models as omodels

pytestmark = pytest.mark.gpu


def test_generated_influenza_matches_reference_ensemble(golden):
    from pysodm_b200 import tutorial_data as td
    e = golden["jump_ensemble"]
    model = gm.make_influenza()
    n = 40_000
    out = model.sim([td.INFLUENZA_START, td.INFLUENZA_END], N=n, tau=0.5, seed=3, output_dtype="int32")
    mine = np.concatenate([np.moveaxis(out[s].values, 0, -1)[e["flu_times"]] for s in td.INFLUENZA_STATES], axis=1)  # [T, 28, n]
    mine = np.moveaxis(mine, 0, 1)                                      # [28, T, n]
    ref = np.transpose(e["flu_ens"], (1, 2, 0))
    cases.assert_ks_ok(cases.ks_pvalues(mine, ref), "generated influenza vs reference ensemble")
    tot = sum(out[s].values.astype(np.int64) for s in ("S", "E", "Ip", "Iud", "Id", "R"))
    assert (tot == tot[:, :1, :]).all() and min(out[s].values.min() for s in td.INFLUENZA_STATES) >= 0
    # and it is the same law as the hand-written functor: first two moments at 4e4 draws
    from pysodm_b200 import tutorial_models as tm
    hand = tm.JumpProcess_influenza_model(td.INFLUENZA_INITIAL_STATES, td.INFLUENZA_PARAMETERS,
                                          coordinates={"age_group": td.AGE_GROUPS},
                                          time_dependent_parameters={"N": td.influenza_contact_function})
    ref2 = hand.sim([td.INFLUENZA_START, td.INFLUENZA_END], N=n, tau=0.5, seed=4, output_dtype="int32")
    a, b = out["Im_inc"].values.astype(np.float64), ref2["Im_inc"].values.astype(np.float64)
    se = np.sqrt((a.var(axis=0) + b.var(axis=0)) / n) + 1e-9
    assert np.all(np.abs(a.mean(axis=0) - b.mean(axis=0)) < 5.5 * se)


def test_generated_sir_si_tau_leap_and_ssa(golden):
    e, g = golden["jump_ensemble"], golden["ssa"]
    model = gm.make_sirsi()
    out = model.sim(120, N=20000, tau=0.5, seed=5, output_dtype="int32")
    flat = np.concatenate([np.moveaxis(out[s].values.reshape(20000, 121, -1), 0, -1) for s in model.states_names], axis=1)
    mine = np.moveaxis(flat[e["sirsi_times"]], 0, 1)                    # [14, T, n]
    cases.assert_ks_ok(cases.ks_pvalues(mine, np.transpose(e["sirsi_ens"], (1, 2, 0))), "generated SIR_SI vs reference")
    # SSA on the down-scaled populations the reference fixture was made with
    small = gm.GenSirSi({"S": g["sirsi_small_y0"][0:4], "I": g["sirsi_small_y0"][4:8], "R": g["sirsi_small_y0"][8:12],
                         "S_v": g["sirsi_small_y0"][12], "I_v": g["sirsi_small_y0"][13]}, dict(gm.SIRSI_PARAMS),
                        coordinates={"age_group": ["a", "b", "c", "d"]})
    assert small.device_model is model.device_model                     # same functions, same shapes: one plugin
    ssa = small.sim(30, N=10000, method="SSA", seed=6)
    flat = np.concatenate([np.moveaxis(ssa[s].values.reshape(10000, 31, -1), 0, -1) for s in small.states_names], axis=1)
    times = [1, 3, 7, 15, 30]
    ref = np.transpose(g["sirsi_small_ens"], (1, 2, 0)).astype(np.float64)
    cases.assert_ks_ok(cases.ks_pvalues(np.moveaxis(flat[times], 0, 1), ref[:, times, :]), "generated SIR_SI SSA vs reference")


class _SeirdOracle(omodels.JumpModelSpec):
    """The SEIRD test model in the oracle's batched form (same functions, leading batch axis)."""
    name = "seird"
    states = tuple((s, (3,)) for s in gm.SEIRD.states)
    params = tuple((p, np.shape(gm.SEIRD_PARAMS[p])) for p in ['beta', 'sigma', 'gamma', 'mu', 'eta', 'rho', 'M', 'h'])
    transitions = (("S", 1), ("E", 1), ("I", 3), ("H", 1))

    @staticmethod
    def rates(t, y, p):
        N = y["S"] + y["E"] + y["I"] + y["H"] + y["R"]
        foi = p["beta"] * np.einsum("ab,...b->...a", p["M"], y["I"] / N)
        one = np.ones(y["S"].shape)
        return {"S": [foi], "E": [p["sigma"] * one], "I": [p["gamma"] * (1 - p["h"]) * one, p["mu"] * one, p["eta"] * p["h"] * one],
                "H": [p["rho"] * one]}

    @staticmethod
    def apply(t, tau, T, y, p):
        return {"S": y["S"] - T["S"][0], "E": y["E"] + T["S"][0] - T["E"][0],
                "I": y["I"] + T["E"][0] - T["I"][0] - T["I"][1] - T["I"][2], "H": y["H"] + T["I"][2] - T["H"][0],
                "R": y["R"] + T["I"][0] + T["H"][0], "D": y["D"] + T["I"][1], "cases": y["cases"] + T["E"][0]}


def test_new_model_with_three_exits_matches_oracle():
    model = gm.make_seird()
    n = 16000
    out = model.sim(60, N=n, tau=0.5, seed=7, output_dtype="int32")
    y0 = np.concatenate([np.ravel(np.asarray(model.initial_states[k], dtype=np.float64)) for k in model.states_names])
    ref = ojump.tau_leap_ensemble(_SeirdOracle, y0, gm.SEIRD_PARAMS, 0, 60, 0.5, n, seed=13)      # [n, 21, 61]
    times = [1, 5, 15, 30, 60]
    mine = np.concatenate([np.moveaxis(out[s].values, 0, -1) for s in model.states_names], axis=1)      # [61, 21, n]
    cases.assert_ks_ok(cases.ks_pvalues(np.moveaxis(mine[times], 0, 1), np.transpose(ref[:, :, times], (1, 2, 0))),
                       "generated SEIRD vs oracle")
    pop = sum(out[s].values.astype(np.int64) for s in ("S", "E", "I", "H", "R", "D"))
    assert (pop == pop[:, :1, :]).all()
    assert (np.diff(out["cases"].values, axis=1) >= 0).all() and (np.diff(out["D"].values, axis=1) >= 0).all()
    # without the reference's stay-probability quirk the three-exit compartment follows a visibly different law
    plain = model.sim(60, N=n, tau=0.5, seed=7, output_dtype="int32", compat=0)
    ps = cases.ks_pvalues(np.moveaxis(plain["D"].values, 0, -1)[[30, 60]], np.transpose(ref[:, 15:18, :][:, :, [30, 60]], (2, 1, 0)))
    assert (ps < 0.01).mean() > 0.5


def test_generated_odes_follow_the_hand_written_functors(golden):
    from pysodm_b200 import engine, tutorial_data as td, tutorial_models as tm
    # SIR_SI ODE: generated vs hand-written, same step sequence
    gen = gm.make_sirsi(gm.GenOdeSirSi)
    hand = gm.make_sirsi(tm.ODE_SIR_SI)
    a, b = gen.sim(120), hand.sim(120)
    for s in gen.states_names:
        np.testing.assert_allclose(a[s].values, b[s].values, rtol=1e-9, atol=1e-9)
    assert np.array_equal(gen.last_nfev.cpu().numpy(), hand.last_nfev.cpu().numpy())
    # tutorial SIR against the reference's scipy RK45 fixture (rtol 1e-6, same number of function evaluations)
    g = golden["ode"]
    model = gm.GenOdeSir({"S": 1000, "I": 1}, {"beta": 0.35, "gamma": 5})
    betas = g["sir_betas"]

    def draw(parameters, N):
        parameters["beta"] = betas
        return parameters
    draw.vectorized = True
    out = model.sim(120, N=100, draw_function=draw)
    y = np.stack([out[s].values for s in ("S", "I", "R")], axis=1)
    np.testing.assert_allclose(y, g["sir_y"], rtol=1e-6, atol=1e-9)
    assert np.array_equal(model.last_nfev.cpu().numpy(), g["sir_nfev"])
    # the implicit methods run on the translated functor too (its Jacobian is scipy's finite-difference one, implicit.cuh)
    for method in ("Radau", "BDF"):
        gi = golden["ode_" + method.lower()]
        betas = gi["sir_betas"]
        out = model.sim(120, N=30, draw_function=draw, method=method)
        y = np.stack([out[s].values for s in ("S", "I", "R")], axis=1)
        np.testing.assert_allclose(y, gi["sir_y"], rtol=1e-6, atol=1e-7)
        assert np.array_equal(model.last_nfev.cpu().numpy(), gi["sir_nfev"])
        # stiff problems (Robertson, Van der Pol up to mu = 200) translated from Python, against the unmodified reference
        # running the same functions (tests/golden/ode_stiff.npz); see tests/test_hostcheck.py for the two draws left out
        gs = golden["ode_stiff"]
        rob, vdp = gm.make_robertson(), gm.make_vanderpol()

        def draw_k1(parameters, N):
            parameters["k1"] = gs["rob_k1"]
            return parameters
        draw_k1.vectorized = True
        out = rob.sim(40, N=6, draw_function=draw_k1, method=method)
        y = np.stack([out[s].values for s in ("y1", "y2", "y3")], axis=1)
        np.testing.assert_allclose(y, gs[f"rob_{method}_y"], rtol=1e-6, atol=1e-12)
        assert np.array_equal(rob.last_nfev.cpu().numpy(), gs[f"rob_{method}_nfev"])

        def draw_mu(parameters, N):
            parameters["mu"] = gs["vdp_mu"][2:]
            return parameters
        draw_mu.vectorized = True
        out = vdp.sim(200, N=4, draw_function=draw_mu, method=method)
        y = np.stack([out[s].values for s in ("x", "v")], axis=1)
        np.testing.assert_allclose(y, gs[f"vdp_{method}_y"][2:], rtol=1e-5, atol=1e-3)
        assert np.array_equal(vdp.last_nfev.cpu().numpy(), gs[f"vdp_{method}_nfev"][2:])
    # packed-bed reactor written with NumPy slicing: generated thread-per-draw kernel vs the warp-per-draw functor
    nx = 6
    gp, hp = gm.make_pfr(nx), gm.make_pfr(nx, tm.packed_PFR)
    ga, ha = gp.sim(600, output_stride=100), hp.sim(600, output_stride=100)
    for s in ("C_F", "C_S"):
        np.testing.assert_allclose(ga[s].values, ha[s].values, rtol=1e-8, atol=1e-12)
    assert np.array_equal(gp.last_nfev.cpu().numpy(), hp.last_nfev.cpu().numpy())
    # at nx = 25 (200 equations) the generated model takes the WARP-per-draw layout: equations grouped by expression shape,
    # one template per position for the whole warp -- against the hand-written warp functor, draws of the kinetic parameters
    nx = 25
    gp, hp = gm.make_pfr(nx), gm.make_pfr(nx, tm.packed_PFR)
    assert gp.device_model.layout.get("warp")
    with pytest.raises(NotImplementedError, match="dense Jacobian"):
        gp.sim(10, method="BDF")

    def draw(parameters, N):
        r = np.random.default_rng(3)
        parameters["Vf_Ks"] = 7.88e-04 * r.uniform(0.8, 1.2, size=N)
        parameters["K_eq"] = 0.418 * r.uniform(0.8, 1.2, size=N)
        return parameters
    draw.vectorized = True
    ga = gp.sim(2000, N=37, draw_function=draw, output_stride=500)
    ha = hp.sim(2000, N=37, draw_function=draw, output_stride=500)
    for s in ("C_F", "C_S"):
        np.testing.assert_allclose(ga[s].values, ha[s].values, rtol=1e-8, atol=1e-12)
    assert np.array_equal(gp.last_nfev.cpu().numpy(), hp.last_nfev.cpu().numpy())
    eu_g, eu_h = gp.sim(50, tau=0.5, output_last_only=True), hp.sim(50, tau=0.5, output_last_only=True)
    np.testing.assert_allclose(eu_g["C_F"].values, eu_h["C_F"].values, rtol=1e-10, atol=1e-14)
    # a model that exists nowhere else: against scipy on the host (the reference's own solver call, base.py:855)
    lv = gm.make_lotka()
    res = lv.sim(40, rtol=1e-8, atol=1e-10)
    from scipy.integrate import solve_ivp

    def rhs(t, y):
        d = gm.LotkaVolterra.integrate(t, y[:5], y[5:], **lv.parameters)
        return np.concatenate(d)
    y0 = np.concatenate([lv.initial_states["prey"], lv.initial_states["pred"]])
    ref = solve_ivp(rhs, [0, 40], y0, t_eval=np.arange(41), rtol=1e-8, atol=1e-10)
    np.testing.assert_allclose(np.concatenate([res["prey"].values, res["pred"].values], axis=1).T, ref.y, rtol=1e-6)
    # Euler with a fixed step too
    eu = lv.sim(10, tau=0.125)
    assert np.isfinite(eu["prey"].values).all()


def test_ode_time_dependent_parameter_functions(golden):
    """TDPFs of an ODE are evaluated at every RK stage time (base.py:769-779): translated with `integrate`, the device
    run follows the reference's scipy run step for step -- numeric time, dates with a date-valued parameter (also drawn
    per trajectory, tutorials/SIR/workflow_tutorial.py:201-223), and a ramp that reads a state."""
    import pandas as pd
    g = golden["tdpf"]
    num, date, ramp = gm.make_tdpf_models()
    out = num.sim(120)
    y = np.stack([out[s].values for s in ("S", "I", "R")])
    np.testing.assert_allclose(y, g["numeric_y"], rtol=1e-6, atol=1e-9)
    assert int(num.last_nfev[0]) == int(g["numeric_nfev"])
    out = num.sim(120, rtol=1e-8)
    np.testing.assert_allclose(np.stack([out[s].values for s in ("S", "I", "R")]), g["numeric_rtol1e-8_y"], rtol=1e-7, atol=1e-9)
    assert int(num.last_nfev[0]) == int(g["numeric_rtol1e-8_nfev"])
    assert num.parameters == {"beta": 0.35, "gamma": 5}

    # one launch for the four start dates: the date parameter varies per trajectory
    shifts = list(g["date_shift_days"])
    it = iter([shifts[0]] + shifts)                                  # the draw function is called once for validation first

    def draw(parameters):
        parameters["start_measures"] = pd.Timestamp("2020-02-05") + pd.Timedelta(days=float(next(it)))
        return parameters
    span = ["2020-01-01", "2020-04-30"]
    out = date.sim(span, N=len(shifts), draw_function=draw)
    assert tuple(out["I"].dims) == ("draws", "date")
    y = np.stack([out[s].values for s in ("S", "I", "R")], axis=1)          # [draws, state, time]
    np.testing.assert_allclose(y, g["date_y"], rtol=1e-6, atol=1e-9)
    assert np.array_equal(date.last_nfev.cpu().numpy(), g["date_nfev"])
    assert date.parameters["start_measures"] == pd.Timestamp("2020-02-05")     # restored after sim
    single = date.sim(span)
    np.testing.assert_allclose(single["I"].values, g["date_y"][0][1], rtol=1e-6, atol=1e-9)

    out = ramp.sim(span)
    np.testing.assert_allclose(np.stack([out[s].values for s in ("S", "I", "R")]), g["ramp_y"], rtol=1e-6, atol=1e-9)
    assert int(ramp.last_nfev[0]) == int(g["ramp_nfev"])
    # the same model on numeric time is a different translation (dates cannot be compared with numbers)
    with pytest.raises(NotImplementedError):
        ramp.sim(120)


def test_log_posterior_of_a_generated_ode_with_a_time_dependent_parameter():
    """`log_posterior_probability.batch` on a translated model: one launch for all thetas equals sim + the oracle's
    likelihood theta by theta (dates as time index, a date-valued TDPF parameter, calibrated beta and gamma)."""
    import pandas as pd
    from oracle import loglik as oll
    from pysodm_b200.objective_functions import ll_poisson, log_posterior_probability
    _, date, _ = gm.make_tdpf_models()
    dates = pd.date_range("2020-01-10", "2020-04-01", freq="7D")
    data = pd.Series(np.linspace(5, 400, len(dates)).round(), index=pd.Index(dates, name="date"))
    lpp = log_posterior_probability(date, ["beta", "gamma"], [(0.1, 1.0), (1.0, 10.0)], [data], ["I"], [ll_poisson], [[]],
                                    start_sim=pd.Timestamp("2020-01-01"))
    thetas = np.array([[0.3, 5.0], [0.35, 5.0], [0.4, 4.0], [0.5, 6.5]])
    lp = lpp.batch(thetas.copy())
    want = []
    for b, g_ in thetas:
        date.parameters.update(beta=b, gamma=g_)
        out = date.sim(["2020-01-01", "2020-03-27"])               # the objective simulates up to the last data point
        y = np.stack([out[s].values for s in ("S", "I", "R")])
        t_data = np.array([(d - pd.Timestamp("2020-01-01")).days for d in dates], dtype=np.float64)
        want.append(oll.log_likelihood("poisson", y, 0.0, [[1]], t_data, data.to_numpy(), None))
    np.testing.assert_allclose(lp, want, rtol=1e-9)


def test_rates_that_use_the_time_itself():
    """`compute_rates(t, ...)` with seasonal forcing in `t` (numeric time) and a date window (date time): the step's start
    time reaches the generated functor; parity against the oracle's tau-leap with the same Python function."""
    from datetime import datetime, timedelta
    model = gm.make_seasonal()
    n = 12000
    out = model.sim(100, N=n, tau=0.5, seed=3, output_dtype="int32")

    class Oracle(omodels.JumpModelSpec):
        name, states, params = "seasonal", (("S", ()), ("I", ()), ("R", ())), (("beta", ()), ("gamma", ()), ("amp", ()))
        transitions = (("S", 1), ("I", 1))
        when = staticmethod(lambda t: t)

        @classmethod
        def rates(cls, t, y, p):
            r = gm.SeasonalSIR.compute_rates(cls.when(t), y["S"], y["I"], y["R"], p["beta"], p["gamma"], p["amp"])
            return {"S": [np.asarray(r["S"][0], dtype=np.float64)], "I": [r["I"][0] * np.ones(np.shape(y["I"]))]}

        @staticmethod
        def apply(t, tau, T, y, p):
            return {"S": y["S"] - T["S"][0], "I": y["I"] + T["S"][0] - T["I"][0], "R": y["R"] + T["I"][0]}
    par = {"beta": 0.5, "gamma": 0.2, "amp": 0.6}
    ref = ojump.tau_leap_ensemble(Oracle, np.array([99000.0, 1000.0, 0.0]), par, 0, 100, 0.5, n, seed=4)
    times = [5, 20, 40, 70, 100]
    mine = np.stack([np.moveaxis(out[s].values, 0, -1) for s in ("S", "I", "R")])          # [3, 101, n]
    cases.assert_ks_ok(cases.ks_pvalues(mine[:, times], np.transpose(ref[:, :, times], (1, 2, 0))), "seasonal SIR (numeric time)")
    # the forcing matters: without it the epidemic is visibly different
    flat = gm.SeasonalSIR({'S': 99000, 'I': 1000}, {'beta': 0.5, 'gamma': 0.2, 'amp': 0.0}).sim(100, N=n, tau=0.5, seed=3, output_dtype="int32")
    a, b = np.median(flat["I"].values[:, 6]), np.median(out["I"].values[:, 6])
    assert b > 1.5 * a, (a, b)                       # forcing starts at 1 + amp = 1.6: a visibly faster take-off
    # SSA evaluates the rates at every EVENT time (base.py:392, :399): the time reaches `propensities` too.  Small population,
    # against the oracle's SSA loop over the same Python function.
    small = gm.SeasonalSIR({'S': 780, 'I': 20}, {'beta': 0.5, 'gamma': 0.2, 'amp': 0.6})
    ssa = small.sim(30, N=4000, method="SSA", seed=9)
    import random
    np.random.seed(12)
    random.seed(12)
    ref_ssa = np.stack([ojump.solve_discrete(Oracle, np.array([780.0, 20.0, 0.0]), par, 0, 30, method="SSA")[1] for _ in range(200)])
    mine_ssa = np.stack([np.moveaxis(ssa[s].values, 0, -1) for s in ("S", "I", "R")])      # [3, 31, n]
    t_ssa = [3, 8, 15, 22, 30]
    cases.assert_ks_ok(cases.ks_pvalues(mine_ssa[:, t_ssa], np.transpose(ref_ssa[:, :, t_ssa], (1, 2, 0))), "seasonal SIR, SSA")
    # over dates the same class takes the other branch of its compute_rates (a lockdown window between two dates)
    start = datetime(2021, 1, 1)
    dated = model.sim(["2021-01-01", "2021-04-11"], N=n, tau=0.5, seed=5, output_dtype="int32")
    Oracle.when = staticmethod(lambda t: pd.Timestamp(start + timedelta(days=float(t))))
    ref = ojump.tau_leap_ensemble(Oracle, np.array([99000.0, 1000.0, 0.0]), par, 0, 100, 0.5, n, seed=6, chunk=n)
    mine = np.stack([np.moveaxis(dated[s].values, 0, -1) for s in ("S", "I", "R")])
    cases.assert_ks_ok(cases.ks_pvalues(mine[:, times], np.transpose(ref[:, :, times], (1, 2, 0))), "seasonal SIR (dates)")
    back = model.sim(100, N=64, tau=0.5, seed=3, output_dtype="int32")                           # and back to numeric time
    assert np.array_equal(back["I"].values, out["I"].values[:64])
