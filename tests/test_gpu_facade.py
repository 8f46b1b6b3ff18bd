"""GPU tests of the pySODM-style front end: `JumpProcess.sim`, `ODE.sim`, `log_posterior_probability` (API contract
of the reference's own tests -- shapes, dims, initial-value echo, loose dynamics -- plus parity against the oracle)."""
from datetime import datetime

import numpy as np
import pandas as pd
import pytest

import cases

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def td():
    from pysodm_b200 import tutorial_data
    return tutorial_data


def _influenza(td):
    from pysodm_b200 import tutorial_models as tm
    return tm.JumpProcess_influenza_model(td.INFLUENZA_INITIAL_STATES, td.INFLUENZA_PARAMETERS,
                                          coordinates={"age_group": td.AGE_GROUPS},
                                          time_dependent_parameters={"N": td.influenza_contact_function})


def test_jump_sim_single_and_draws(td, golden):
    """Reference contract (src/tests/test_JumpProcess.py:144-151, :374-376, :508-512): dims, coords, S[0] echo."""
    model = _influenza(td)
    np.random.seed(3)
    out = model.sim([td.INFLUENZA_START, td.INFLUENZA_END], tau=0.5)
    S = out["S"]
    assert tuple(S.dims) == ("date", "age_group") and S.values.shape == (162, 4)
    assert list(out["date"][:2]) == [datetime(2017, 12, 1), datetime(2017, 12, 2)]
    assert np.array_equal(S.values[0], td.INFLUENZA_INITIAL_STATES["S"])
    assert S.values.dtype == np.float64
    before = dict(model.parameters)

    def draw_fcn(parameters, spread):
        parameters["beta"] = 0.0174 * np.random.uniform(1 - spread, 1 + spread)
        return parameters

    n = 512
    out = model.sim([td.INFLUENZA_START, td.INFLUENZA_END], N=n, tau=0.5, draw_function=draw_fcn,
                    draw_function_kwargs={"spread": 0.05}, output_dtype="int32", processes=4)
    Im = out["Im_inc"]
    assert tuple(Im.dims) == ("draws", "date", "age_group") and Im.values.shape == (n, 162, 4)
    assert Im.values.dtype == np.int32
    assert model.parameters["beta"] == before["beta"]                       # parameters restored (base.py:548,574)
    e = golden["jump_ensemble"]
    # beta jitter of +-5 % around the tutorial value: ensemble mean stays within a few % of the reference's
    ref_mean = e["flu_mean"][24:28].T                                       # [time, age]
    rel = np.abs(Im.values.mean(axis=0)[30:120] - ref_mean[30:120]) / (ref_mean[30:120] + 5)
    assert np.median(rel) < 0.15
    mean_via_api = out["Im_inc"].mean(dim="draws")
    assert mean_via_api.values.shape == (162, 4)
    sel = out["Im_inc"].sel(age_group=td.AGE_GROUPS[2]).isel(draws=0)
    assert sel.values.shape == (162,)


def test_jump_sim_vectorized_draws_and_seed_reproducibility(td):
    model = _influenza(td)

    def draw(parameters, N):
        rng = np.random.default_rng(0)
        parameters["beta"] = rng.uniform(0.015, 0.02, size=N)
        parameters["f_ud"] = np.clip(parameters["f_ud"][None, :] * rng.uniform(0.9, 1.1, size=(N, 4)), 1e-3, 0.999)
        return parameters
    draw.vectorized = True
    span = ["2017-12-01", "2017-12-31"]                      # the tutorial's contact function needs dates
    a = model.sim(span, N=300, tau=0.5, draw_function=draw, seed=11, output_dtype="int32", chunk=128)
    b = model.sim(span, N=300, tau=0.5, draw_function=draw, seed=11, output_dtype="int32", chunk=300)
    assert tuple(a["S"].dims) == ("draws", "date", "age_group") and a["S"].values.shape == (300, 31, 4)
    for s in ("S", "E", "Im_inc"):
        assert np.array_equal(a[s].values, b[s].values)                     # chunking does not change the result
    c = model.sim(span, N=300, tau=0.5, draw_function=draw, seed=12, output_dtype="int32")
    assert not np.array_equal(a["E"].values, c["E"].values)
    sub = model.sim(span, N=8, tau=0.5, states=["Im_inc", "S"], output_dtype="float32")
    assert list(sub.keys()) == ["Im_inc", "S"]
    from pysodm_b200 import tutorial_models as tm
    sir = tm.JumpProcess_SIR({"S": 1_000_000 - 10, "I": 10}, {"beta": 0.9, "gamma": 0.2})
    out = sir.sim(50, N=64, tau=1)                            # integer 'time' axis (src/tests/test_JumpProcess.py:144-151)
    assert tuple(out["S"].dims) == ("draws", "time") and np.array_equal(np.asarray(out["time"]), np.arange(51))
    assert (out["S"].values[:, 0] == 1_000_000 - 10).all() and np.median(out["S"].values[:, -1]) < 12000


def test_sir_si_mixed_state_shapes_and_ssa(td):
    from pysodm_b200 import tutorial_models as tm
    model = tm.JumpProcess_SIR_SI(td.SIRSI_INITIAL_STATES, td.SIRSI_PARAMETERS, coordinates={"age_group": td.AGE_GROUPS})
    out = model.sim(50, N=4, tau=0.5)
    assert out["S"].values.shape == (4, 51, 4) and out["S_v"].values.shape == (4, 51)
    assert tuple(out["I_v"].dims) == ("draws", "time")
    small = tm.JumpProcess_SIR_SI({"S": [607, 1329, 7352, 2204], "S_v": 1000, "I_v": 2}, td.SIRSI_PARAMETERS,
                                  coordinates={"age_group": td.AGE_GROUPS})
    out = small.sim(20, N=64, method="SSA")
    assert out["S"].values.shape == (64, 21, 4)
    assert np.any(out["I"].values != np.round(out["I"].values))               # F3: SSA output is interpolated
    with pytest.raises(ValueError, match="invalid solution method"):
        model.sim(10, method="euler")
    with pytest.raises(TypeError, match="'tau' must be of type int or float"):
        model.sim(10, tau="0.5")
    with pytest.raises(ValueError, match="start of simulation is chronologically after"):
        model.sim([10, 5])


def test_ode_sim_contract_and_parity(td, golden):
    from pysodm_b200 import tutorial_models as tm
    g = golden["ode"]
    model = tm.ODE_SIR({"S": 1000, "I": 1}, {"beta": 0.35, "gamma": 5})
    out = model.sim(120)
    assert tuple(out["S"].dims) == ("time",) and out["S"].values.shape == (121,)
    ref = g["sir_y"][np.argmin(np.abs(g["sir_betas"] - 0.35))]          # not the same beta: only a sanity bound
    assert abs(out["S"].values[-1] - ref[0, -1]) / ref[0, -1] < 0.5
    betas = g["sir_betas"]
    it = iter(betas)

    def draw_fcn(parameters):
        parameters["beta"] = next(it)
        return parameters
    # the draw function is called once for validation, then N times (Q12): start the iterator one early
    it = iter(np.concatenate([[betas[0]], betas]))
    out = model.sim(120, N=100, draw_function=draw_fcn)
    y = np.stack([out[s].values for s in ("S", "I", "R")], axis=1)          # [draws, state, time]
    np.testing.assert_allclose(y, g["sir_y"], rtol=1e-6, atol=1e-9)
    assert np.array_equal(model.last_nfev.cpu().numpy(), g["sir_nfev"])
    with pytest.raises(ValueError, match="without using a draw function"):
        model.sim(10, N=5)
    with pytest.raises(NotImplementedError, match="no device implementation"):
        model.sim(10, method="LSODA")
    eu = model.sim(60, tau=0.25)
    np.testing.assert_allclose(eu["I"].values, g["sir_euler_y"][1], rtol=1e-12)
    # method='RK23' / 'DOP853' through the class API, against the unmodified reference's runs with those methods
    for method in ("RK23", "DOP853", "Radau", "BDF"):
        gm = golden["ode_" + method.lower()]
        model.parameters["beta"] = float(gm["sir_betas"][3])
        r = model.sim(120, method=method)
        np.testing.assert_allclose(np.stack([r[s].values for s in ("S", "I", "R")]), gm["sir_y"][3], rtol=1e-6, atol=1e-7)
        assert int(model.last_nfev.cpu().numpy()[0]) == int(gm["sir_nfev"][3])


def test_packed_pfr_last_profile(td, golden):
    from pysodm_b200 import tutorial_models as tm
    g = golden["ode"]
    nx = 25
    model = tm.packed_PFR(td.pfr_initial_states(nx), td.pfr_parameters(nx),
                          {"species": ["S", "A", "Es", "W"], "x": np.linspace(0, 1, nx)})
    post = g["pfr_posterior"]

    def draw(parameters, N):
        pick = g["pfr25_pick"]
        for j, k in enumerate(("Vf_Ks", "R_AS", "R_AW", "R_Es", "K_eq")):
            parameters[k] = post[pick, j]
        return parameters
    draw.vectorized = True
    out = model.sim(8000, N=4, draw_function=draw, output_last_only=True)
    cf = out["C_F"]
    assert tuple(cf.dims) == ("draws", "time", "species", "x") and cf.values.shape == (4, 1, 4, nx)
    ref = g["pfr25_y"][:, :, -1]                                             # [draw, 200]
    np.testing.assert_allclose(cf.values[:, 0].reshape(4, -1), ref[:, :100], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(out["C_S"].values[:, 0].reshape(4, -1), ref[:, 100:], rtol=1e-6, atol=1e-7)


def test_log_posterior_batch_matches_oracle(td):
    """Config C5 in miniature: NB log-posterior of P particles; same trajectories scored by the oracle."""
    import torch
    from oracle import loglik as oll
    from pysodm_b200 import engine
    from pysodm_b200.objective_functions import ll_negative_binomial, ll_poisson, log_posterior_probability
    model = _influenza(td)
    dates = pd.date_range("2017-12-01", periods=13, freq="7D")
    idx = pd.MultiIndex.from_product([dates, td.AGE_GROUPS], names=["date", "age_group"])
    rng = np.random.default_rng(0)
    data = pd.Series(rng.gamma(2.0, 150.0, size=len(idx)), index=idx, name="CASES")
    lpp = log_posterior_probability(model, ["beta", "f_ud"], [(1e-6, 0.08), (1e-3, 1 - 1e-3)], [data], ["Im_inc"],
                                    [ll_negative_binomial], [4 * [0.03]], simulation_kwargs={"tau": 0.5})
    assert lpp.expanded_labels == ["beta", "f_ud_0", "f_ud_1", "f_ud_2", "f_ud_3"]
    assert lpp.start_sim == datetime(2017, 12, 1) and lpp.end_sim == dates[-1].to_pydatetime()
    P = 257
    thetas = np.column_stack([rng.uniform(0.01, 0.03, P), rng.uniform(0.05, 0.95, size=(P, 4))])
    thetas[0, 0] = 0.5                                                       # out of bounds: clipped, -inf (Q15)
    lp = lpp.batch(thetas, seed=77)
    assert lp.shape == (P,) and lp[0] == -np.inf and thetas[0, 0] == 0.08
    assert np.isfinite(lp[1:]).all()
    # the same trajectories, regenerated with the same seed, scored on the host by the oracle
    var = {1: thetas[:, 0], 4: thetas[:, 1], 5: thetas[:, 2], 6: thetas[:, 3], 7: thetas[:, 4]}
    n_out = 85
    traj, _ = engine.jump_run("influenza_1718_jump", td.influenza_flat_y0(), td.influenza_flat_par(), P, 0, n_out, 0.5,
                              var=var, tdpf=(td.CONTACT_TABLE.reshape(3, 16), td.influenza_schedule(n_steps=168)),
                              seed=77, out_dtype=torch.float64)
    traj = traj.cpu().numpy()
    ydata = data.to_numpy().reshape(13, 4)
    want = np.array([oll.log_likelihood("negative_binomial", traj[24:28, :, i], 0.0, [[0], [1], [2], [3]],
                                        np.arange(13) * 7.0, ydata, 4 * [0.03]) for i in range(P)])
    np.testing.assert_allclose(lp[1:], want[1:], rtol=1e-9)
    single = lpp(thetas[5].copy())
    assert np.isfinite(single)
    # aggregated data (no age axis): the model's age dimension is summed (objective_functions.py:250-252)
    total = data.groupby(level="date").sum()
    lpp2 = log_posterior_probability(model, ["beta"], [(1e-6, 0.08)], [total], ["Im_inc"], [ll_poisson], [[]],
                                     simulation_kwargs={"tau": 0.5})
    lp2 = lpp2.batch(thetas[:9, :1].copy(), seed=5)
    traj2, _ = engine.jump_run("influenza_1718_jump", td.influenza_flat_y0(), td.influenza_flat_par(), 9, 0, n_out, 0.5,
                               var={1: thetas[:9, 0]}, tdpf=(td.CONTACT_TABLE.reshape(3, 16), td.influenza_schedule(n_steps=168)),
                               seed=5, out_dtype=torch.float64)
    traj2 = traj2.cpu().numpy()
    want2 = [oll.log_likelihood("poisson", traj2[:, :, i], 0.0, [[24, 25, 26, 27]], np.arange(13) * 7.0,
                                total.to_numpy()[:, None]) for i in range(9)]
    np.testing.assert_allclose(lp2, want2, rtol=1e-9)


def test_log_posterior_ode_time_index(td):
    """`time`-indexed data at non-integer times (src/tests/test_calibration.py:20-36) on an ODE model."""
    from oracle import loglik as oll, models, ode
    from pysodm_b200 import tutorial_models as tm
    from pysodm_b200.objective_functions import ll_normal, log_posterior_probability
    model = tm.ODE_SIR_rate({"S": 1_000_000 - 10, "I": 10}, {"beta": 0.9, "gamma": 0.2})
    t = np.linspace(50, 60, 20)
    data = pd.Series(np.linspace(4000, 900, 20), index=pd.Index(t, name="time"))
    lpp = log_posterior_probability(model, ["beta"], [(0.1, 2.0)], [data], ["I"], [ll_normal], [250.0], start_sim=0)
    betas = np.array([[0.7], [0.9], [1.1]])
    lp = lpp.batch(betas.copy())
    want = []
    for b in betas[:, 0]:
        _, y, _ = ode.simulate(models.SIRRateODE, [1_000_000 - 10, 10, 0], {"beta": b, "gamma": 0.2}, 0, 60)
        want.append(oll.log_likelihood("normal", y, 0.0, [[1]], t, data.to_numpy(), 250.0))
    np.testing.assert_allclose(lp, want, rtol=1e-6)       # ODE trajectories agree to ~1e-12; ll amplifies by < 1e3


def test_pso_calibration_recovers_parameter(td, capsys):
    """The reference's calibration test in miniature (src/tests/test_calibration.py:440-479): PSO on an ODE model
    recovers the beta that generated the data; every swarm iteration is ONE batched device evaluation."""
    from pysodm_b200 import optimization, tutorial_models as tm
    from pysodm_b200.objective_functions import ll_poisson, log_posterior_probability
    truth = tm.ODE_SIR_rate({"S": 1_000_000 - 10, "I": 10}, {"beta": 0.9, "gamma": 0.2})
    out = truth.sim(40)
    t = np.arange(10, 41, 2)
    data = pd.Series(out["I"].values[t], index=pd.Index(t.astype(float), name="time"))
    model = tm.ODE_SIR_rate({"S": 1_000_000 - 10, "I": 10}, {"beta": 0.5, "gamma": 0.2})
    lpp = log_posterior_probability(model, ["beta"], [(0.1, 2.0)], [data], ["I"], [ll_poisson], [[]], start_sim=0)
    np.random.seed(0)
    theta, score = optimization.pso(lpp, swarmsize=128, max_iter=20)
    assert abs(theta[0] - 0.9) / 0.9 < 0.02
    theta_nm, _ = optimization.nelder_mead(lpp, np.array([0.6]), [0.2], max_iter=30)
    assert abs(theta_nm[0] - 0.9) / 0.9 < 0.02


@pytest.mark.parametrize("device_update", [False, True])
def test_ensemble_mcmc_samples_the_posterior(td, tmp_path, monkeypatch, device_update):
    """The reference's MCMC leg in miniature (src/tests/test_calibration.py:195-209): walkers started around an
    estimate sample the posterior of beta; every half-ensemble update is ONE batched device evaluation.  The chain's
    mean and spread are compared with the posterior tabulated on a grid through the same objective."""
    from pysodm_b200 import tutorial_models as tm
    from pysodm_b200.objective_functions import ll_negative_binomial, log_posterior_probability
    from pysodm_b200.optimization import mcmc
    monkeypatch.chdir(tmp_path)
    truth = tm.ODE_SIR_rate({"S": 1_000_000 - 10, "I": 10}, {"beta": 0.9, "gamma": 0.2})
    t = np.arange(10, 31, 4)
    data = pd.Series(truth.sim(30)["I"].values[t], index=pd.Index(t.astype(float), name="time"))
    model = tm.ODE_SIR_rate({"S": 1_000_000 - 10, "I": 10}, {"beta": 0.5, "gamma": 0.2})
    lpp = log_posterior_probability(model, ["beta"], [(0.1, 2.0)], [data], ["I"], [ll_negative_binomial], [0.05],
                                    start_sim=0)
    grid = np.linspace(0.8, 1.0, 4001)
    w = np.exp(lpp.batch(grid[:, None].copy()) - lpp(np.array([0.9])))
    w /= w.sum()
    mean = float(np.sum(w * grid))
    std = float(np.sqrt(np.sum(w * (grid - mean) ** 2)))
    assert 1e-3 < std < 0.03
    np.random.seed(1)
    ndim, nwalkers, pos = mcmc.perturbate_theta([0.9], pert=[0.02], multiplier=64, bounds=lpp.expanded_bounds)
    sampler, samples = mcmc.run_EnsembleSampler(pos, 150, "sir", lpp, samples_path="out/", print_n=50, discard=50,
                                                progress=False, seed=3, device_update=device_update)
    assert sampler.n_batches == 1 + 2 * sampler.iteration
    x = samples["beta"].ravel()
    assert x.shape == ((sampler.iteration - 50) * 64,)
    assert abs(x.mean() - mean) < 0.15 * std and abs(x.std() / std - 1) < 0.15
    assert 0.1 < sampler.acceptance_fraction.mean() < 0.9


@pytest.mark.parametrize("n", [1, 33, 100_003, 3_000_000])
def test_status_flags_are_ored_on_the_device(n):
    """`sodm_status_or`: the launch's flags are the OR over all trajectories (a maximum would let SODM_ST_TRUNCATED_Y0 on one
    trajectory hide SODM_ST_BAD_PROB on another)."""
    import torch
    from pysodm_b200 import _ffi, engine
    engine.device(0)
    st = torch.zeros(n, dtype=torch.int32, device="cuda")
    assert engine.status_bits(st) == 0
    st[n // 2] |= _ffi.ST_BAD_PROB
    st[0] |= _ffi.ST_TRUNCATED_Y0
    st[n - 1] |= _ffi.ST_MAX_STEPS
    want = _ffi.ST_BAD_PROB | _ffi.ST_TRUNCATED_Y0 | _ffi.ST_MAX_STEPS
    assert engine.status_bits(st) == want
    if n > 2:
        assert int(st.max()) & _ffi.ST_BAD_PROB == 0
    assert engine.status_bits(st[:0]) == 0


def test_log_posterior_with_aggregation_function(td):
    """`aggregation_function` (reference objective_functions.py:245-248; its test: src/tests/test_calibration.py:624-668): the
    doubly stratified SIR of that test, data by age group only.  The device value must equal the reference's recipe done by
    hand on the host: simulate, apply the function to out[state], interpolate to the data times, score."""
    import gen_models as gm
    from pysodm_b200.objective_functions import ll_negative_binomial, ll_normal, log_posterior_probability
    model = gm.make_double_sir()
    t = np.arange(0, 50, 7, dtype=float)
    idx = pd.MultiIndex.from_product([t, ["0-20", "20-120"]], names=["time", "age_groups"])
    rng = np.random.default_rng(3)
    data = pd.Series(rng.gamma(2.0, 40.0, size=len(idx)), index=idx, name="cases")

    def total(output):
        return output.sum(dim="spatial_units")

    def weighted(output):
        return output.sel(spatial_units=0) + 2.0 * output.sel(spatial_units=2)
    lpp = log_posterior_probability(model, ["beta"], [(1e-6, 1)], [data, data], ["I", "I"], [ll_negative_binomial, ll_normal],
                                    [0.03 * np.ones(2), 25.0], start_sim=0, aggregation_function=[total, weighted],
                                    simulation_kwargs={"rtol": 1e-9})
    thetas = np.array([[0.35, 0.30], [0.25, 0.4], [0.5, 0.2]])
    lp = lpp.batch(thetas.copy())
    want = []
    for th in thetas:
        model.parameters["beta"] = th
        out = model.sim(49, rtol=1e-9)
        ll = 0.0
        for agg, fn, arg in ((total, ll_negative_binomial, 0.03 * np.ones(2)), (weighted, ll_normal, 25.0)):
            a = agg(out["I"])                                         # dims (time, age_groups)
            ym = np.asarray(a.values, dtype=np.float64)[::7]          # data times are grid points
            ll += fn(ym, data.to_numpy().reshape(len(t), 2), arg)
        want.append(ll)
    np.testing.assert_allclose(lp, want, rtol=1e-8)
