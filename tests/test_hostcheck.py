"""Statistics of the KERNEL CODE itself, run on the CPU by the test-only harness (tests/hostcheck): the same
__host__ __device__ sampler / stepper / integrator sources the GPU executes, compiled for the host.  These tests
let the build container (no GPU) catch a wrong law before GPU time is spent; the `-m gpu` tests repeat the
comparisons on the device."""
import ctypes as C
import numpy as np
import pytest
from scipy import stats

import cases
import hostcheck_util as hc


def test_philox_known_answer():
    """Philox4x32-10 known-answer vectors (Random123 kat_vectors: counter/key all-zero and all-ones)."""
    out = hc.philox(0, 0, 0, 1)[0]
    assert [hex(int(x)) for x in out] == ["0x6627e8d5", "0xe169c58d", "0xbc57ac4c", "0x9b00dbd8"]
    # generic: distinct trajectories / purposes give distinct, well-mixed streams
    a, b, c = hc.philox(7, 1, 0, 4096), hc.philox(7, 2, 0, 4096), hc.philox(7, 1, 1, 4096)
    assert not np.array_equal(a, b) and not np.array_equal(a, c)
    bits = np.unpackbits(a.view(np.uint8))
    assert abs(bits.mean() - 0.5) < 5 * 0.5 / np.sqrt(bits.size)


def test_philox2x32_known_answer():
    """Philox2x32-10 known-answer vectors (Random123 kat_vectors)."""
    assert [hex(int(x)) for x in hc.philox2x32(0, 0, 0)] == ["0xff1dae59", "0x6cd10df2"]
    assert [hex(int(x)) for x in hc.philox2x32(0xffffffff, 0xffffffff, 0xffffffff)] == ["0x2c3f628b", "0xab4fd7ad"]
    assert [hex(int(x)) for x in hc.philox2x32(0x13198a2e, 0x243f6a88, 0x85a308d3)] == ["0xdd7ce038", "0xf62a4c12"]


@pytest.mark.parametrize("n,p", [(7317774, 1.2e-3), (4292, 0.28), (100, 0.3), (50, 0.5), (30, 0.4), (1000, 0.005),
                                 (7317774, 1e-7), (81, 0.13), (5, 0.9), (2130556, 0.02), (40, 0.26), (12, 0.97),
                                 (200000, 0.5), (1, 0.3), (20, 0.5), (1000000, 0.9999)])
def test_binomial_sampler_is_exact(n, p):
    """Chi-square goodness of fit against the exact binomial pmf, across the inversion / BTRS / flipped regimes."""
    x, attempts = hc.binomial(n, p, 300_000, seed=n % 97 + 1)
    p32 = float(np.float32(p))
    lo, hi = int(x.min()), int(x.max())
    assert 0 <= lo and hi <= n
    ks = np.arange(lo, hi + 1)
    obs = np.bincount(x - lo, minlength=len(ks)).astype(float)
    exp = stats.binom.pmf(ks, n, p32) * len(x)
    big = exp >= 8
    o = np.append(obs[big], obs[~big].sum())
    e = np.append(exp[big], len(x) - exp[big].sum())
    if e[-1] < 4:
        o, e = o[:-1], e[:-1]
    if len(o) > 1:
        chi2 = ((o - e) ** 2 / e).sum()
        assert stats.chi2.sf(chi2, len(o) - 1) > 1e-4, (n, p)
    assert attempts / len(x) < 1.6


@pytest.mark.parametrize("n,p", [(7317774, 1.2e-3), (4292, 0.28), (100, 0.3), (50, 0.5), (30, 0.4), (1000, 0.005),
                                 (7317774, 1e-7), (81, 0.13), (5, 0.9), (2130556, 0.02), (40, 0.26), (12, 0.97),
                                 (200000, 0.5), (1, 0.3), (20, 0.5), (1000000, 0.9999),
                                 # modes in the FP32 / FP64 hand-over band of the exact test (0.9 * 8192 .. 8192), n >= 2^24,
                                 # and deep inside the FP64 branch
                                 (7317774, 1.02e-3), (7317774, 1.1e-3), (2130556, 3.8e-3), (20000000, 4.0e-4),
                                 (16777216, 0.25), (50000000, 0.5), (7317774, 0.3)])
def test_two_pass_attempts_are_exact(n, p):
    """The default sampler's unit of work (counter-based words, box test, squeeze, postponed exact test): chi-square against
    the exact pmf."""
    x, attempts = hc.binomial_words(n, p, 300_000, seed=n % 89 + 3)
    p32 = float(np.float32(p))
    lo, hi = int(x.min()), int(x.max())
    assert 0 <= lo and hi <= n
    ks = np.arange(lo, hi + 1)
    obs = np.bincount(x - lo, minlength=len(ks)).astype(float)
    exp = stats.binom.pmf(ks, n, p32) * len(x)
    big = exp >= 8
    o = np.append(obs[big], obs[~big].sum())
    e = np.append(exp[big], len(x) - exp[big].sum())
    if e[-1] < 4:
        o, e = o[:-1], e[:-1]
    if len(o) > 1:
        chi2 = ((o - e) ** 2 / e).sum()
        assert stats.chi2.sf(chi2, len(o) - 1) > 1e-4, (n, p)
    assert attempts / len(x) < 1.6


def test_uniform_stays_inside_the_unit_interval():
    """a word of all ones must not map to u = 1.0f, on which sequential inversion runs to its cap."""
    for n, p in [(2000, 0.004), (50, 0.02), (7317774, 1e-6)]:
        x, _ = hc.binomial_words(n, p, 1, force_word=0xFFFFFFFF)
        assert x[0] < stats.binom.isf(1e-9, n, p) + 2, (n, p, x[0])
        x, _ = hc.binomial_words(n, p, 1, force_word=0xFFFFFF00)
        assert x[0] < stats.binom.isf(1e-9, n, p) + 2, (n, p, x[0])
    # Poisson noise below the inversion threshold has the same search loop
    out, _ = hc.noise(0, 0.0, 7.5, 200000, seed=3)
    assert out.max() < 40


def test_binomial_edge_cases():
    assert (hc.binomial(0, 0.3, 100)[0] == 0).all()
    assert (hc.binomial(50, 0.0, 100)[0] == 0).all()
    assert (hc.binomial(50, 1.0, 100)[0] == 50).all()
    assert (hc.binomial(50, float("nan"), 10)[0] == -1).all()
    assert (hc.binomial(50, 1.5, 10)[0] == -1).all()
    assert (hc.binomial(-3, 0.5, 10)[0] == -1).all()


def test_exact_log_pmf_ratio():
    from scipy.special import gammaln
    for n, p, k in [(100, 0.3, 41), (7317774, 1.2e-3, 9001), (50, 0.5, 10), (4292, 0.28, 1100)]:
        m = np.floor((n + 1) * p)
        want = (gammaln(m + 1) + gammaln(n - m + 1) - gammaln(k + 1) - gammaln(n - k + 1)
                + (k - m) * np.log(p / (1 - p)))
        assert hc.lib().hc_binomial_log_ratio(n, p, k) == pytest.approx(want, abs=2e-7 * max(1, abs(want)) + 1e-6)


def test_large_mode_log_pmf_ratio():
    """The FP64 multiply-add form used for modes >= 8192 against the exact product of pmf ratios (summed term by term:
    gammaln loses ~1e-7 at these magnitudes), over the range BTRS candidates take (|k - mode| up to 5 sigma), incl. n >= 2^24
    and p = 1/2; outside its range of validity it reports failure."""
    import math
    rng = np.random.default_rng(3)
    worst = 0.0
    for n, p in [(7317774, 1.2e-3), (7317774, 0.02), (2130556, 0.3), (50_000_000, 0.5), (20_000_000, 4.2e-4), (30000, 0.4)]:
        p = float(np.float32(p))
        m = int(np.floor((n + 1) * p))
        sd = np.sqrt(n * p * (1 - p))
        for k in np.unique(np.clip(np.rint(m + sd * rng.uniform(-5, 5, 60)), 0, n)).astype(np.int64):
            # f(j+1)/f(j) = (n - j) p / ((j + 1) q)
            j = np.arange(min(m, k), max(m, k), dtype=np.float64)
            total = math.fsum(np.log((n - j) / (j + 1.0)) + math.log(p / (1 - p)))
            want = total if k >= m else -total
            got = hc.lib().hc_binomial_log_ratio_large(n, p, float(k))
            worst = max(worst, abs(got - want))
    assert worst < 2e-9, worst
    assert np.isnan(hc.lib().hc_binomial_log_ratio_large(30000, 0.4, 100.0))       # |x| too large for the series: caller falls back


def test_tau_leap_influenza_law(golden):
    g, e = golden["jump_single"], golden["jump_ensemble"]
    out, st = hc.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 12000, 0, cases.FLU_N_OUT, 0.5,
                          tdpf=(g["flu_table"], g["flu_step_index"]), out_dtype=np.int32, seed=5)
    assert (st == 0).all()
    mine = out[:, e["flu_times"], :]
    ref = np.transpose(e["flu_ens"], (1, 2, 0))
    cases.assert_ks_ok(cases.ks_pvalues(mine, ref), "influenza kernel code (host) vs reference ensemble")
    tot = out[:24].reshape(6, 4, cases.FLU_N_OUT, -1).sum(axis=0)
    assert (tot == tot[:, :1, :]).all()


def test_queue_sampler_law_on_the_host(golden):
    """The queue sampler (SODM_TUNE_QUEUE: regime-separated batches, counter-based words per attempt) run as a warp of ONE lane:
    same law as the reference, population conserved, shard invariant."""
    g, e = golden["jump_single"], golden["jump_ensemble"]
    kw = dict(tdpf=(g["flu_table"], g["flu_step_index"]), out_dtype=np.int32, seed=6, compat=3 | 256)
    out, st = hc.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 12000, 0, cases.FLU_N_OUT, 0.5, **kw)
    assert (st == 0).all()
    ref = np.transpose(e["flu_ens"], (1, 2, 0))
    cases.assert_ks_ok(cases.ks_pvalues(out[:, e["flu_times"], :], ref), "queue sampler (host) vs reference ensemble")
    tot = out[:24].reshape(6, 4, cases.FLU_N_OUT, -1).sum(axis=0)
    assert (tot == tot[:, :1, :]).all() and out.min() >= 0
    part, _ = hc.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 50, 0, cases.FLU_N_OUT, 0.5, traj_id0=300, **kw)
    assert np.array_equal(part, out[:, :, 300:350])
    out, _ = hc.jump_run("sir_si_jump", g["sirsi_y0"], cases.SIRSI_PAR, 12000, 0, 121, 0.5, out_dtype=np.int32, compat=3 | 256)
    cases.assert_ks_ok(cases.ks_pvalues(out[:, e["sirsi_times"], :], np.transpose(e["sirsi_ens"], (1, 2, 0))),
                       "queue sampler, SIR_SI (host) vs reference ensemble")


def test_tau_leap_sir_si_law(golden):
    g, e = golden["jump_single"], golden["jump_ensemble"]
    out, st = hc.jump_run("sir_si_jump", g["sirsi_y0"], cases.SIRSI_PAR, 12000, 0, 121, 0.5, out_dtype=np.int32)
    mine = out[:, e["sirsi_times"], :]
    cases.assert_ks_ok(cases.ks_pvalues(mine, np.transpose(e["sirsi_ens"], (1, 2, 0))),
                       "SIR_SI kernel code (host) vs reference ensemble")


def test_interpolated_output_and_shard_invariance(golden):
    g = golden["jump_single"]
    tdpf = (g["flu_tau03_table"], g["flu_tau03_step_index"])
    full, _ = hc.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 64, 0, 46, 0.3, tdpf=tdpf, seed=3)
    part, _ = hc.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 16, 0, 46, 0.3, tdpf=tdpf, seed=3,
                          traj_id0=40)
    assert np.array_equal(part, full[:, :, 40:56])
    assert np.any(full != np.round(full))
    # every third step (0.9 -> ... 3.0) lands within 1 ulp of an integer time: interpolation weights stay in [0,1]
    assert full.min() >= 0


def test_ssa_law(golden):
    g = golden["ssa"]
    out, st = hc.jump_run("sir_si_jump", g["sirsi_small_y0"], cases.SIRSI_PAR, 4000, 0, 31, method="SSA", seed=2)
    assert (st == 0).all()
    times = [1, 3, 7, 15, 30]
    ref = np.transpose(g["sirsi_small_ens"], (1, 2, 0)).astype(np.float64)
    cases.assert_ks_ok(cases.ks_pvalues(out[:, times, :], ref[:, times, :]), "SSA kernel code (host) vs reference")


def test_rk45_follows_scipy_step_for_step(golden):
    g = golden["ode"]
    out, nfev, st = hc.ode_run("sir_ode", [1000, 1, 0], [0.35, 5], 100, 0, 121, var={0: g["sir_betas"]})
    assert np.array_equal(nfev, g["sir_nfev"]) and (st == 0).all()
    np.testing.assert_allclose(out, np.transpose(g["sir_y"], (1, 2, 0)), rtol=1e-9, atol=1e-9)
    out, _, _ = hc.ode_run("sir_ode", [1000, 1, 0], [0.35, 5], 1, 0, 61, method="euler", tau=0.25)
    assert np.array_equal(out[:, :, 0], g["sir_euler_y"])
    par = np.concatenate([g["pfr25_par_scalar"], g["pfr25_kL_a"], g["pfr25_D_ax"]])
    post = g["pfr_posterior"][g["pfr25_pick"]]
    out, nfev, st = hc.ode_run("packed_pfr_ode", g["pfr25_y0"], par, 4, 0, 8001, model_arg=25, out_stride=100,
                               var={4: post[:, 0], 5: post[:, 1], 6: post[:, 2], 8: post[:, 3], 7: post[:, 4]})
    assert np.array_equal(nfev, g["pfr25_nfev"])
    np.testing.assert_allclose(out, np.transpose(g["pfr25_y"], (1, 2, 0)), rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("method", ["RK23", "DOP853"])
def test_rk23_and_dop853_follow_scipy_step_for_step(golden, method):
    """`method='RK23'` (Bogacki-Shampine 3(2)) and `'DOP853'` (Hairer's 8(5,3) pair incl. the three extra evaluations of its
    dense output on steps that contain a grid point) with scipy's tableaux and controller, against runs of the UNMODIFIED
    reference with that method (tests/golden/ode_rk23.npz, ode_dop853.npz): identical `nfev` on every draw, values to 1e-7."""
    g = golden["ode_" + method.lower()]
    tol = dict(rtol=1e-7, atol=1e-7)
    out, nfev, st = hc.ode_run("sir_ode", [1000, 1, 0], [0.35, 5], 30, 0, 121, var={0: g["sir_betas"]}, method=method)
    assert np.array_equal(nfev, g["sir_nfev"]) and (st == 0).all()
    np.testing.assert_allclose(out, np.transpose(g["sir_y"], (1, 2, 0)), **tol)
    out, nfev, _ = hc.ode_run("sir_ode", [1000, 1, 0], [0.35, 5], 1, 0, 121, method=method, rtol=1e-7)
    assert int(nfev[0]) == int(g["sir_rtol1e-07_nfev"])
    np.testing.assert_allclose(out[:, :, 0], g["sir_rtol1e-07_y"], rtol=1e-9, atol=1e-9)
    out, nfev, _ = hc.ode_run("sir_si_ode", g["sirsi_y0"], cases.SIRSI_PAR, 1, 0, 121, method=method)
    assert int(nfev[0]) == int(g["sirsi_nfev"])
    np.testing.assert_allclose(out[:, :, 0], g["sirsi_y"], rtol=1e-7, atol=1e-5)
    out, nfev, _ = hc.ode_run("ppbb_ode", g["ppbb_y0"], g["ppbb_par"], 1, 0, 1501, method=method)
    assert int(nfev[0]) == int(g["ppbb_nfev"])
    np.testing.assert_allclose(out[:, :, 0], g["ppbb_y"], **tol)


def test_packed_pfr_nx100_kernel_code_vs_reference(golden):
    """The integrator + reactor functor at nx = 100 on the host against the unmodified reference over the full horizon; see
    tests/test_gpu_ode.py::test_packed_pfr_nx100_full_horizon for why the late-time tolerance is the reference's own band."""
    from pysodm_b200 import tutorial_data as td
    g = golden["ode_pfr100"]
    post = g["pfr100_posterior"][:1]
    out, nfev, st = hc.ode_run("packed_pfr_ode", g["pfr100_y0"], td.pfr_flat_par(100), 1, 0, 8001, model_arg=100, out_stride=100,
                               var={4: post[:, 0], 5: post[:, 1], 6: post[:, 2], 8: post[:, 3], 7: post[:, 4]})
    assert int(nfev[0]) == int(g["pfr100_nfev"][0]) and int(st[0]) == 0
    ref = g["pfr100_y"][0]
    np.testing.assert_allclose(out[:, :11, 0], ref[:, :11], rtol=1e-6, atol=1e-7)
    rel = np.abs(out[:, :, 0] - ref) / (np.abs(ref) + 1e-6)
    assert rel.max() <= 5 * float(g["pfr100_ref_self_max_rel"])


@pytest.mark.parametrize("method", ["Radau", "BDF"])
def test_radau_and_bdf_follow_scipy_step_for_step(golden, method):
    """`method='Radau'` / `'BDF'` (csrc/implicit.cuh: scipy's Radau IIA and variable-order NDF with its finite-difference
    Jacobian, simplified Newton and LU) against runs of the UNMODIFIED reference with that method (tests/golden/ode_radau.npz,
    ode_bdf.npz): identical `nfev` on every draw -- same Newton iteration counts, rejections and Jacobian refreshes --
    values to 1e-6 (observed 1e-8: the Newton iterates stop at newton_tol, so LAPACK-vs-own LU rounding is visible there)."""
    g = golden["ode_" + method.lower()]
    out, nfev, st = hc.ode_run("sir_ode", [1000, 1, 0], [0.35, 5], 30, 0, 121, var={0: g["sir_betas"]}, method=method)
    assert np.array_equal(nfev, g["sir_nfev"]) and (st == 0).all()
    np.testing.assert_allclose(out, np.transpose(g["sir_y"], (1, 2, 0)), rtol=1e-6, atol=1e-7)
    out, nfev, _ = hc.ode_run("sir_ode", [1000, 1, 0], [0.35, 5], 1, 0, 121, method=method, rtol=1e-7)
    assert int(nfev[0]) == int(g["sir_rtol1e-07_nfev"])
    np.testing.assert_allclose(out[:, :, 0], g["sir_rtol1e-07_y"], rtol=1e-9, atol=1e-9)
    out, nfev, _ = hc.ode_run("sir_si_ode", g["sirsi_y0"], cases.SIRSI_PAR, 1, 0, 121, method=method)      # 14 equations
    assert int(nfev[0]) == int(g["sirsi_nfev"])
    np.testing.assert_allclose(out[:, :, 0], g["sirsi_y"], rtol=1e-7, atol=1e-5)
    out, nfev, _ = hc.ode_run("ppbb_ode", g["ppbb_y0"], g["ppbb_par"], 1, 0, 1501, method=method)
    assert int(nfev[0]) == int(g["ppbb_nfev"])
    np.testing.assert_allclose(out[:, :, 0], g["ppbb_y"], rtol=1e-7, atol=1e-7)
    with pytest.raises(RuntimeError):                            # the reactor's node-per-lane layout has no dense Jacobian
        from pysodm_b200 import tutorial_data as td
        hc.ode_run("packed_pfr_ode", golden["ode"]["pfr25_y0"], td.pfr_flat_par(25), 1, 0, 11, model_arg=25, method=method)


def test_finite_difference_jacobian_matches_scipy_num_jac():
    """`num_jac` of the implicit steppers against scipy's own on Robertson's kinetics (entries from 0.04 to 3e7, zero state
    components: the increment search and the too-small-difference retry) and a random 12-equation chain, over three successive
    calls so that the persisted, adapted increment factors are compared too: J and factors bit-identical."""
    from scipy.integrate._ivp.common import num_jac
    L = hc.lib()
    L.hc_num_jac.restype = C.c_int
    L.hc_num_jac.argtypes = [C.c_int, C.c_void_p, C.c_double, C.c_void_p, C.c_double, C.c_int, C.c_void_p, C.c_void_p]

    def rob(t, y, par):
        r1 = par[0] * y[0]
        r2 = par[1] * y[1] * y[2]
        r3 = par[2] * y[1] * y[1]
        return np.array([-r1 + r2, r1 - r2 - r3, r3])

    def chain(t, y, par, n=12):
        dy = np.empty(n)
        for i in range(n):
            s = par[n + i] * y[i]
            if i > 0:
                s = par[i] * y[i - 1] + s
            if i < n - 1:
                s = s + par[2 * n + i] * y[i + 1]
            dy[i] = s - par[3 * n] * y[i] * y[i]
        return dy
    rng = np.random.default_rng(9)
    stiff = 10.0 ** rng.uniform(-1, 3, size=12)
    chain_par = np.concatenate([rng.uniform(0, 0.5, 12) * stiff, -stiff, rng.uniform(0, 0.5, 12) * stiff, [0.05]])
    for mid, fun, par, ys in ((0, rob, np.array([0.04, 1e4, 3e7]), [np.array([1.0, 0.0, 0.0]), np.array([0.98, 3.4e-5, 0.02]), np.array([0.5, 1e-6, 0.5])]),
                              (3, chain, chain_par, [rng.uniform(0.5, 2.0, 12), rng.uniform(0, 1e-3, 12), rng.uniform(0.5, 2.0, 12)])):
        n = ys[0].size
        factor_ref, factor = None, np.zeros(n)
        for call, y in enumerate(ys):
            f = fun(0.0, y, par)
            J_ref, factor_ref = num_jac(lambda t, Y: np.stack([fun(t, Y[:, k], par) for k in range(Y.shape[1])], axis=1), 0.0, y, f, 1e-6, factor_ref)
            J = np.zeros((n, n))
            y = np.ascontiguousarray(y)
            assert L.hc_num_jac(mid, par.ctypes.data, 0.0, y.ctypes.data, 1e-6, int(call > 0), factor.ctypes.data, J.ctypes.data) == 0
            assert np.array_equal(J, J_ref), (mid, call, np.abs(J - J_ref).max())
            assert np.array_equal(factor, factor_ref), (mid, call)


def test_bdf_difference_array_rescaling_matches_scipy():
    """`change_D` of the BDF stepper (R(factor) U formed for the maximal order with U = the signed Pascal matrix, applied under
    predicates on the order) against scipy's own `change_D` on random difference arrays, every order and a range of factors;
    rows above the order stay untouched."""
    from scipy.integrate._ivp.bdf import change_D, compute_R
    assert np.array_equal(compute_R(5, 1), [[(-1) ** k * __import__("math").comb(j, k) for j in range(6)] for k in range(6)])
    L = hc.lib()
    L.hc_bdf_change_D.restype = None
    L.hc_bdf_change_D.argtypes = [C.c_int, C.c_double, C.c_void_p]
    rng = np.random.default_rng(5)
    for order in range(1, 6):
        for factor in (0.2, 0.5, 0.731, 1.0, 1.9, 10.0):
            D = rng.normal(size=(8, 3)) * 10.0 ** rng.uniform(-6, 2, size=(8, 1))
            want = D.copy()
            change_D(want, order, factor)
            got = np.ascontiguousarray(D.copy())
            L.hc_bdf_change_D(order, factor, got.ctypes.data)
            assert np.array_equal(got[order + 1:], D[order + 1:])
            np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-12 * max(np.abs(D).max(), np.abs(want).max()))   # (sums cancel)


@pytest.mark.parametrize("n", [3, 8, 12])
def test_dense_lu_of_the_implicit_steppers(n):
    """LuReal / LuComplex (csrc/implicit.cuh) against LAPACK on matrices that need row interchanges at every column (largest
    entries below the diagonal), both index paths (n <= 8: compile-time indices with predicated swaps; 12: run-time indices)."""
    from scipy.linalg import lu_factor, lu_solve
    rng = np.random.default_rng(n)
    for trial in range(20):
        A = rng.normal(size=(n, n))
        A[np.arange(n), np.arange(n)] *= 1e-3                           # weak diagonal: partial pivoting must swap
        b = rng.normal(size=n)
        piv = lu_factor(A)[1]
        assert (piv != np.arange(n)).any()
        np.testing.assert_allclose(hc.lu_solve(A, b), lu_solve(lu_factor(A), b), rtol=1e-9, atol=1e-12)
        Ac = A + 1j * rng.normal(size=(n, n))
        bc = b + 1j * rng.normal(size=n)
        np.testing.assert_allclose(hc.lu_solve(Ac, bc), lu_solve(lu_factor(Ac), bc), rtol=1e-9, atol=1e-12)
    I = np.eye(n)
    assert np.array_equal(hc.lu_solve(I, np.arange(n, dtype=float)), np.arange(n, dtype=float))


@pytest.mark.parametrize("method", ["Radau", "BDF", "RK45", "RK23", "DOP853"])
@pytest.mark.parametrize("n", [3, 12])
def test_adaptive_methods_against_scipy_on_random_stiff_chains(method, n):
    """Random chains  y_i' = a_i y_(i-1) + b_i y_i + c_i y_(i+1) - q y_i^2  with decay rates spread over four decades, through
    the call the reference makes (`solve_ivp(fun, [t0, t1], y0, args=..., t_eval=grid, method=method, rtol=rtol)`, ref
    base.py:855) and through the kernel code: identical `nfev` -- i.e. the same accepted / rejected steps, Newton iterations and
    Jacobian refreshes -- for Jacobians of varied structure; 12 equations take the run-time-indexed path of implicit.cuh."""
    from scipy.integrate import solve_ivp

    def fun(t, y, par):
        dy = np.empty(n)
        for i in range(n):
            s = par[n + i] * y[i]
            if i > 0:
                s = par[i] * y[i - 1] + s
            if i < n - 1:
                s = s + par[2 * n + i] * y[i + 1]
            dy[i] = s - par[3 * n] * y[i] * y[i]
        return dy
    rng = np.random.default_rng(100 + n)
    implicit = method in ("Radau", "BDF")
    for c in range(8 if implicit else 3):
        stiff = 10.0 ** rng.uniform(-1, 3 if implicit else 1, size=n)
        par = np.concatenate([rng.uniform(0, 0.5, n) * stiff, -stiff, rng.uniform(0, 0.5, n) * stiff, [rng.uniform(0, 0.1)]])
        y0 = rng.uniform(0.5, 2.0, size=n)
        rtol = [1e-4, 1e-6][c % 2]
        ref = solve_ivp(fun, [0, 30], y0, args=(par,), t_eval=np.arange(31.0), method=method, rtol=rtol)
        out, nfev, st = hc.ode_run(f"test:chain{n}", y0, par, 1, 0, 31, method=method, rtol=rtol)
        assert int(st[0]) == 0 and ref.status == 0
        # scipy's own reproducibility: an explicit method at its stability limit (solution decayed below atol) accepts and rejects
        # chaotically -- one ulp on y0 moves scipy's nfev (observed 497 / 506 / 515); only a reproducible nfev can be matched
        band = {ref.nfev}
        for toward in (np.inf, -np.inf):
            yp = y0.copy()
            yp[0] = np.nextafter(yp[0], toward)
            band.add(solve_ivp(fun, [0, 30], yp, args=(par,), t_eval=np.arange(31.0), method=method, rtol=rtol).nfev)
        if len(band) == 1:
            assert int(nfev[0]) == ref.nfev, (method, n, c, int(nfev[0]), ref.nfev)
            np.testing.assert_allclose(out[:, :, 0], ref.y, rtol=1e-6, atol=1e-9)
        else:
            assert not implicit
            spread = max(band) - min(band)
            assert min(band) - spread <= int(nfev[0]) <= max(band) + spread, (method, n, c, int(nfev[0]), sorted(band))
            np.testing.assert_allclose(out[:, :, 0], ref.y, rtol=10 * rtol, atol=1e-4)


@pytest.mark.parametrize("method", ["Radau", "BDF"])
def test_radau_and_bdf_on_stiff_problems(golden, method):
    """Robertson's kinetics and the Van der Pol oscillator (mu = 0.5 .. 200, through its fast transitions) -- rejected steps,
    failed Newton iterations with Jacobian refreshes, BDF orders up to 5 -- against the unmodified reference running the same
    right-hand sides (tests/golden/ode_stiff.npz): identical `nfev` up to 17 121 evaluations.  The two mildly stiff oscillator
    draws under BDF are compared on the first 20 time units only: over thousands of steps scipy's own `nfev` there moves with
    a one-ulp change of y0 (6401 / 6418 / 6443), i.e. the step sequence is not reproducible across LU roundings."""
    g = golden["ode_stiff"]
    out, nfev, st = hc.ode_run("test:robertson", [1.0, 0, 0], [0.04, 1e4, 3e7], 6, 0, 41, var={0: g["rob_k1"]}, method=method)
    assert np.array_equal(nfev, g[f"rob_{method}_nfev"]) and (st == 0).all()
    np.testing.assert_allclose(out, np.transpose(g[f"rob_{method}_y"], (1, 2, 0)), rtol=1e-7, atol=1e-12)
    out, nfev, _ = hc.ode_run("test:robertson", [1.0, 0, 0], [0.04, 1e4, 3e7], 1, 0, 401, method=method, rtol=1e-6)
    assert int(nfev[0]) == int(g[f"rob_{method}_long_nfev"])
    np.testing.assert_allclose(out[:, :, 0], g[f"rob_{method}_long_y"], rtol=1e-8, atol=1e-13)
    out, nfev, st = hc.ode_run("test:vanderpol", [2.0, 0], [100.0], 6, 0, 201, var={0: g["vdp_mu"]}, method=method)
    ref = np.transpose(g[f"vdp_{method}_y"], (1, 2, 0))
    same = slice(0, 6) if method == "Radau" else slice(2, 6)
    assert np.array_equal(nfev[same], g[f"vdp_{method}_nfev"][same]) and (st == 0).all()
    # values: to the solver's tolerance where the solution jumps (a 1e-9 shift of a transition moves x by 1e-4 there), 1e-6 elsewhere
    np.testing.assert_allclose(out[:, :21], ref[:, :21], rtol=1e-8, atol=1e-9)
    np.testing.assert_allclose(out[..., same], ref[..., same], rtol=1e-5, atol=1e-3)
    assert np.median(np.abs(out[..., same] - ref[..., same])) < 1e-9


@pytest.mark.parametrize("method", ["RK45", "RK23", "DOP853", "Radau", "BDF"])
def test_windowed_output_driver_equals_the_plain_driver(golden, method):
    """The device's thread kernel delivers kept grid points in windows (rk45_run_windowed, ode_run.cuh); run here as a warp of
    one lane it must reproduce the plain driver bit for bit: window lengths that do and do not divide the kept count, strided
    grids, last point only, and runs abandoned at max_steps (status set, remaining points as the plain driver leaves them)."""
    g = golden["ode"]
    base = dict(var={0: g["sir_betas"][:12]}, method=method)
    for kw in (dict(n_out=121), dict(n_out=121, out_stride=7), dict(n_out=122, out_stride=11), dict(n_out=121, out_last_only=True),
               dict(n_out=2), dict(n_out=1), dict(n_out=121, max_steps=6)):
        n_out = kw.pop("n_out")
        want = hc.ode_run("sir_ode", [1000, 1, 0], [0.35, 5], 12, 0, n_out, **base, **kw)
        for window in (1, 5, 16, 200):
            got = hc.ode_run("sir_ode", [1000, 1, 0], [0.35, 5], 12, 0, n_out, window=window, **base, **kw)
            for a, b in zip(got, want):
                assert np.array_equal(a, b, equal_nan=True), (method, n_out, kw, window)
