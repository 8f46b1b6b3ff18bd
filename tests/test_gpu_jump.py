"""GPU parity tests of the jump-process kernels (through the C ABI, `pysodm_b200.engine` -> libsodm.so).

Stochastic parity bar (BASELINE.json): ensemble distributions per (state, time) agree with the reference's
tau-leap / SSA within Monte-Carlo error (two-sample KS, p > 0.01) -- checked against ensembles produced by the
unmodified reference (tests/golden/*.npz) and against the pinned oracle on draws with per-trajectory
parameters.  Integer bookkeeping (conservation, determinism, shard invariance) is checked exactly.
"""
import os

import numpy as np
import pytest

import cases

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    import torch
    from pysodm_b200 import engine
    engine.device(0)
    return engine


def _cpu(t):
    return t.cpu().numpy()


def test_influenza_matches_reference_ensemble(eng, golden):
    import torch
    g, e = golden["jump_single"], golden["jump_ensemble"]
    n = 100_000
    out, st = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, cases.FLU_N_OUT, 0.5,
                           tdpf=(g["flu_table"], g["flu_step_index"]), seed=1, out_dtype=torch.int32)
    out, st = _cpu(out), _cpu(st)
    assert (st == 0).all()
    assert np.array_equal(out[:, 0, 0], g["flu_y0"].astype(np.int32))
    mine = out[:, e["flu_times"], :]                                   # [28, T, n]
    ref = np.transpose(e["flu_ens"], (1, 2, 0))
    cases.assert_ks_ok(cases.ks_pvalues(mine, ref), "influenza GPU vs reference ensemble")
    # means and 2.5/50/97.5 % quantiles within Monte-Carlo error of the reference's 2000-member ensemble
    sd = ref.std(axis=-1) + 1.0
    assert np.all(np.abs(mine.mean(-1) - ref.mean(-1)) < 5 * sd / np.sqrt(ref.shape[-1]) + 1e-9)
    q_mine = np.quantile(out, [0.025, 0.5, 0.975], axis=-1)
    sd_all = np.transpose(e["flu_q"][2] - e["flu_q"][0], (0, 1)) / 3.92 + 1.0
    # the reference's tail quantiles come from 2000 members: s.e. ~ 0.06 sd (tails) / 0.03 sd (median); 13 608 cells
    assert np.all(np.abs(q_mine - e["flu_q"]) < 0.4 * sd_all + 2.0)
    assert np.mean(np.abs(q_mine - e["flu_q"]) < 0.15 * sd_all + 1.0) > 0.97
    # population of every age group is conserved exactly (S+E+Ip+Iud+Id+R)
    tot = out[:24].reshape(6, 4, cases.FLU_N_OUT, n).sum(axis=0)
    assert (tot == tot[:, :1, :]).all()
    assert out.min() >= 0


def test_influenza_without_quirk_differs_from_reference(eng, golden):
    """Power check: the mathematically intended stay probability (compat=0) is NOT the reference's law (F1)."""
    import torch
    g, e = golden["jump_single"], golden["jump_ensemble"]
    out, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 20000, 0, cases.FLU_N_OUT, 0.5,
                          tdpf=(g["flu_table"], g["flu_step_index"]), seed=2, out_dtype=torch.int32, compat=0)
    mine = _cpu(out)[4:24][:, e["flu_times"], :]
    ref = np.transpose(e["flu_ens"], (1, 2, 0))[4:24]
    assert np.mean(cases.ks_pvalues(mine, ref) < 0.01) > 0.5


def test_sir_si_matches_reference_ensemble(eng, golden):
    import torch
    g, e = golden["jump_single"], golden["jump_ensemble"]
    n = 100_000
    out, st = eng.jump_run("sir_si_jump", g["sirsi_y0"], cases.SIRSI_PAR, n, 0, 121, 0.5, seed=3,
                           out_dtype=torch.int32)
    out = _cpu(out)
    assert (_cpu(st) == 0).all()
    mine = out[:, e["sirsi_times"], :]
    ref = np.transpose(e["sirsi_ens"], (1, 2, 0))
    cases.assert_ks_ok(cases.ks_pvalues(mine, ref), "SIR_SI GPU vs reference ensemble")
    humans = out[:12].reshape(3, 4, 121, n).sum(axis=0)
    assert (humans == humans[:, :1, :]).all()
    assert ((out[12] + out[13]) == out[12, 0] + out[13, 0]).all()


def test_per_trajectory_draws_match_oracle(eng, golden):
    """Config C3: beta and f_ud differ per trajectory; reference law from the pinned batched oracle."""
    import torch
    from oracle import jump, models
    g = golden["jump_single"]
    n = 20000
    beta, f_ud = cases.flu_draws(n)
    var = {cases.FLU_IDX_BETA: beta}
    var.update({k: f_ud[:, i] for i, k in enumerate(cases.FLU_IDX_FUD)})
    n_out = 64
    idx = g["flu_step_index"][: 2 * (n_out - 1)]
    out, st = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, n_out, 0.5, var=var,
                           tdpf=(g["flu_table"], idx), seed=4, out_dtype=torch.int32)
    par = dict(cases.FLU_PAR_DICT, beta=beta, f_ud=f_ud)
    ref = jump.tau_leap_ensemble(models.InfluenzaJump, g["flu_y0"], par, 0, n_out - 1, 0.5, n,
                                 tdpf={"N": (g["flu_table"], idx)}, seed=11)
    times = [1, 2, 5, 10, 20, 40, 63]
    mine = _cpu(out)[:, times, :]
    cases.assert_ks_ok(cases.ks_pvalues(mine, np.transpose(ref[:, :, times], (1, 2, 0))), "C3 draws GPU vs oracle")


@pytest.mark.slow
def test_full_size_ensembles_match_the_oracle_per_cell(eng, golden):
    """BASELINE configs at their FULL size against large oracle ensembles (the 10^6 runs were only mean-checked).
    C3: 10^6 influenza trajectories, full season, per-trajectory beta / f_ud draws, against 10^5 trajectories of the pinned
    oracle with parameters from the same distribution -- two-sample KS per (state, time) cell, resolution ~0.4 % in the
    CDF.  C2: 10^6 SIR_SI tau-leap trajectories against 10^5.  Same bar as every ensemble test (cases.assert_ks_ok)."""
    import torch
    g = golden["jump_single"]
    n, m = 1_000_000, 100_000
    times = [1, 7, 28, 56, 84, 120, 161]
    beta, f_ud = cases.flu_draws(n)
    var = {cases.FLU_IDX_BETA: beta}
    var.update({k: f_ud[:, i] for i, k in enumerate(cases.FLU_IDX_FUD)})
    out, st = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, cases.FLU_N_OUT, 0.5, var=var,
                           tdpf=(g["flu_table"], g["flu_step_index"]), seed=31, out_dtype=torch.int32)
    assert int(st.abs().sum()) == 0
    mine = out[:, times, :].reshape(-1, n).contiguous()
    del out
    ref = cases.oracle_ensemble_parallel("flu", m, times, seed=9000, with_draws=True)             # [m, 28, T]
    ref_d = torch.from_numpy(np.ascontiguousarray(np.transpose(ref, (1, 2, 0)).reshape(-1, m))).cuda()
    cases.assert_ks_ok(cases.ks_pvalues_counts_device(mine, ref_d), "C3 at N = 10^6 vs 10^5 oracle trajectories")
    del mine, ref_d
    times2 = [1, 10, 30, 60, 90, 120]
    out, st = eng.jump_run("sir_si_jump", g["sirsi_y0"], cases.SIRSI_PAR, n, 0, 121, 0.5, seed=32, out_dtype=torch.int32)
    assert int(st.abs().sum()) == 0
    mine = out[:, times2, :].reshape(-1, n).contiguous()
    del out
    ref = cases.oracle_ensemble_parallel("sirsi", m, times2, seed=9500)
    ref_d = torch.from_numpy(np.ascontiguousarray(np.transpose(ref, (1, 2, 0)).reshape(-1, m))).cuda()
    cases.assert_ks_ok(cases.ks_pvalues_counts_device(mine, ref_d), "C2 at N = 10^6 vs 10^5 oracle trajectories")


def test_test_sir_final_size(eng, golden):
    """The reference's own assertion (src/tests/test_JumpProcess.py:144-151) + KS against the oracle."""
    import torch
    from oracle import jump, models
    g = golden["jump_single"]
    n = 20000
    out, _ = eng.jump_run("sir_jump", g["sir_y0"], cases.SIR_PAR, n, 0, 51, 1.0, seed=5, out_dtype=torch.float64)
    out = _cpu(out)
    assert out.shape == (3, 51, n)
    assert (out[0, 0] == 1_000_000 - 10).all()
    assert (out[0, -1] < 12000).mean() > 0.99          # a few trajectories die out early, as in the reference
    ref = jump.tau_leap_ensemble(models.SIRJump, g["sir_y0"], {"beta": 0.9, "gamma": 0.2}, 0, 50, 1.0, n, seed=3)
    times = [1, 5, 10, 20, 35, 50]
    cases.assert_ks_ok(cases.ks_pvalues(out[:, times, :], np.transpose(ref[:, :, times], (1, 2, 0))),
                       "test SIR GPU vs oracle")


def test_determinism_and_shard_invariance(eng, golden):
    """Philox streams are keyed by GLOBAL trajectory id: a shard reproduces its slice of the full run bit for bit."""
    import torch
    g = golden["jump_single"]
    kw = dict(tdpf=(g["flu_table"], g["flu_step_index"]), seed=99, out_dtype=torch.int32)
    full, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 4096, 0, cases.FLU_N_OUT, 0.5, **kw)
    again, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 4096, 0, cases.FLU_N_OUT, 0.5, **kw)
    assert torch.equal(full, again)
    part, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 1000, 0, cases.FLU_N_OUT, 0.5,
                           traj_id0=3000, **kw)
    assert torch.equal(part, full[:, :, 3000:4000])
    other, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 4096, 0, cases.FLU_N_OUT, 0.5,
                            **dict(kw, seed=100))
    assert not torch.equal(full, other)


def test_state_selection_and_dtypes(eng, golden):
    import torch
    g = golden["jump_single"]
    kw = dict(tdpf=(g["flu_table"], g["flu_step_index"]), seed=7)
    full, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 512, 0, cases.FLU_N_OUT, 0.5,
                           out_dtype=torch.int32, **kw)
    sel = [24, 25, 26, 27, 0]
    part, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 512, 0, cases.FLU_N_OUT, 0.5,
                           out_dtype=torch.float64, state_sel=sel, **kw)
    assert part.shape == (5, cases.FLU_N_OUT, 512)
    assert torch.equal(part.to(torch.int32), full[sel])
    f32, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 512, 0, cases.FLU_N_OUT, 0.5,
                          out_dtype=torch.float32, **kw)
    assert torch.equal(f32.to(torch.int32), full)


def test_non_dyadic_tau_interpolates_like_reference(eng, golden):
    """tau = 0.3: outputs are np.interp'ed between steps (F3/Q4/Q5).  Same Philox streams + same sampler code as
    the CPU harness would give identical counts; here the check is distributional against the pinned oracle plus
    the exact interpolation identity on the GPU's own step states."""
    import torch
    from oracle import jump, models
    g = golden["jump_single"]
    n = 16000
    tdpf = (g["flu_tau03_table"], g["flu_tau03_step_index"])
    out, st = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, 46, 0.3, tdpf=tdpf, seed=8,
                           out_dtype=torch.float64)
    out = _cpu(out)
    assert (_cpu(st) == 0).all()
    assert np.any(out != np.round(out))
    ref = jump.tau_leap_ensemble(models.InfluenzaJump, g["flu_y0"], cases.FLU_PAR_DICT, 0, 45, 0.3, n,
                                 tdpf={"N": tdpf}, seed=5)
    times = [1, 2, 3, 10, 30, 45]
    cases.assert_ks_ok(cases.ks_pvalues(out[:, times, :], np.transpose(ref[:, :, times], (1, 2, 0))),
                       "tau=0.3 GPU vs oracle")
    with pytest.raises(Exception, match="int32 output needs"):
        eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 8, 0, 46, 0.3, tdpf=tdpf,
                     out_dtype=torch.int32)


def test_ssa_matches_reference_ensemble(eng, golden):
    import torch
    g = golden["ssa"]
    n = 20000
    out, st = eng.jump_run("sir_si_jump", g["sirsi_small_y0"], cases.SIRSI_PAR, n, 0, 31, method="SSA", seed=9,
                           out_dtype=torch.float64)
    out = _cpu(out)
    assert (_cpu(st) == 0).all()
    times = [1, 3, 7, 15, 30]
    ref = np.transpose(g["sirsi_small_ens"], (1, 2, 0)).astype(np.float64)
    cases.assert_ks_ok(cases.ks_pvalues(out[:, times, :], ref[:, times, :]), "SSA GPU vs reference ensemble")
    # event budget -> status flag
    _, st = eng.jump_run("sir_si_jump", g["sirsi_small_y0"], cases.SIRSI_PAR, 64, 0, 31, method="SSA", seed=9,
                         max_events=10)
    assert (_cpu(st) & 8).all()


def test_ssa_tie_rule_flag(eng):
    """F2: with the reference rule the first of two bit-identical propensities always fires."""
    # I[0] == I[1] = 3 -> the two I->R channels have bit-identical propensity 3/gamma
    y0 = np.array([5, 5, 5, 5, 3, 3, 0, 0, 0, 0, 0, 0, 10, 0], dtype=np.float64)
    par = np.array([0.05, 0.1, 0.2, 0.15, 7.0, 5.0])
    n = 8192
    # propensities: I->R 0.6, 0.6 ; S_v->I_v 0.5625 ; S_v death 1.4286  => P(some I->R fires first) = 0.376
    ref_rule, _ = eng.jump_run("sir_si_jump", y0, par, n, 0, 2, method="SSA", seed=1, max_events=1)
    fair, _ = eng.jump_run("sir_si_jump", y0, par, n, 0, 2, method="SSA", seed=1, max_events=1, compat=1)
    r0, r1 = _cpu(ref_rule)[8, 1], _cpu(ref_rule)[9, 1]      # R[0], R[1] after exactly one event
    assert (r1 == 0).all() and 0.33 < (r0 > 0).mean() < 0.42
    f0, f1 = _cpu(fair)[8, 1], _cpu(fair)[9, 1]
    assert 0.15 < (f1 > 0).mean() < 0.23 and 0.15 < (f0 > 0).mean() < 0.23


def test_bad_probability_is_flagged(eng, golden):
    g = golden["jump_single"]
    par = cases.FLU_PAR.copy()
    par[1] = -1.0                                # negative beta -> negative rate -> p < 0: NumPy raises (Q3)
    _, st = eng.jump_run("influenza_1718_jump", g["flu_y0"], par, 256, 0, 8, 0.5,
                         tdpf=(g["flu_table"], g["flu_step_index"]))
    assert (_cpu(st) & 1).all()


def test_empty_and_ragged_launch(eng, golden):
    import torch
    g = golden["jump_single"]
    out, st = eng.jump_run("sir_si_jump", g["sirsi_y0"], cases.SIRSI_PAR, 0, 0, 5, 0.5)
    assert out.shape == (14, 5, 0)
    out, st = eng.jump_run("sir_si_jump", g["sirsi_y0"], cases.SIRSI_PAR, 131, 0, 1, 0.5, out_dtype=torch.int32)
    assert torch.equal(out[:, 0, :].cpu(), torch.from_numpy(g["sirsi_y0"].astype(np.int32))[:, None].expand(14, 131))


def test_full_size_properties(eng, golden):
    """N = 10^6 influenza trajectories over the full season (BASELINE config 3) -- size-independent properties:
    exact conservation per age group, non-negativity, initial-state echo, reproducibility checksum, and the
    ensemble mean against the reference's."""
    import torch
    g, e = golden["jump_single"], golden["jump_ensemble"]
    n = 1_000_000
    beta, f_ud = cases.flu_draws(n)
    kw = dict(tdpf=(g["flu_table"], g["flu_step_index"]), seed=1718, out_dtype=torch.int32)
    out, st = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, cases.FLU_N_OUT, 0.5, **kw)
    assert int(st.abs().sum()) == 0
    assert int(out.min()) >= 0
    tot = out[:24].view(6, 4, cases.FLU_N_OUT, n).sum(dim=0)
    assert bool((tot == tot[:, :1, :]).all())
    assert bool((out[:, 0, :] == torch.from_numpy(g["flu_y0"].astype(np.int32)).to(out.device)[:, None]).all())
    mean = out.to(torch.float64).mean(dim=-1).cpu().numpy()
    sd = (e["flu_q"][2] - e["flu_q"][0]) / 3.92 + 1.0
    assert np.all(np.abs(mean - e["flu_mean"]) < 6 * sd / np.sqrt(2000))
    checksum = int(out.to(torch.int64).sum())
    del out
    out2, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, cases.FLU_N_OUT, 0.5, **kw)
    assert int(out2.to(torch.int64).sum()) == checksum


@pytest.mark.parametrize("coop", [False, True])
@pytest.mark.parametrize("n,p", [(7_317_774, 1.2e-3), (4292, 0.28), (100, 0.3), (30, 0.4), (81, 0.13), (2_130_556, 0.02),
                                 (200_000, 0.45), (1000, 0.005), (7_000_000, 1e-7), (12, 0.7), (3885, 0.133),
                                 # modes in the FP32 / FP64 hand-over band of the exact test (7300 .. 8300), n >= 2^24 and
                                 # deep inside the FP64 branch
                                 (7_317_774, 1.02e-3), (7_317_774, 1.1e-3), (2_130_556, 3.8e-3), (20_000_000, 4.0e-4),
                                 (16_777_216, 0.25), (50_000_000, 0.5), (7_317_774, 0.3)])
def test_device_binomial_is_exact(eng, coop, n, p):
    """Chi-square goodness of fit of the DEVICE sampler (fast-math proposal, FP32/FP64 exact tests; `coop` = the queue
    sampler with counter-based words instead of the default per-lane walk) against the exact binomial pmf, driven through the public kernel: one tau-leap step of the dimensionless SIR with S = I = n, R = 0
    and beta = -2 ln(1 - p) makes the S->I count Bin(n, p) (compat without the stay-probability quirk)."""
    import torch
    from scipy import stats
    m = 400_000
    beta = -2.0 * np.log1p(-p)
    compat = 256 if coop else 0
    out, st = eng.jump_run("sir_jump", [n, n, 0], [beta, 0.0], m, 0, 2, 1.0, seed=n % 1000 + 7, compat=compat,
                           out_dtype=torch.int32)
    assert int(st.abs().sum()) == 0
    k = n - out[0, 1].cpu().numpy().astype(np.int64)
    p_eff = float(-np.expm1(np.float32(-1.0) * np.float32(beta) * (np.float32(n) / np.float32(2 * n))))
    lo, hi = int(k.min()), int(k.max())
    assert 0 <= lo and hi <= n
    ks = np.arange(lo, hi + 1)
    obs = np.bincount(k - lo, minlength=len(ks)).astype(float)
    exp = stats.binom.pmf(ks, n, p_eff) * m
    big = exp >= 8
    o = np.append(obs[big], obs[~big].sum())
    e = np.append(exp[big], m - exp[big].sum())
    if e[-1] < 4:
        o, e = o[:-1], e[:-1]
    if len(o) > 1:
        chi2 = ((o - e) ** 2 / e).sum()
        assert stats.chi2.sf(chi2, len(o) - 1) > 1e-4, (n, p, coop)


def test_queue_sampler_has_the_same_law(eng, golden):
    """SODM_TUNE_QUEUE changes the schedule and the RNG streams, not the law."""
    import torch
    g, e = golden["jump_single"], golden["jump_ensemble"]
    out, st = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 50_000, 0, cases.FLU_N_OUT, 0.5,
                           tdpf=(g["flu_table"], g["flu_step_index"]), seed=21, out_dtype=torch.int32, compat=3 | 256)
    assert int(st.abs().sum()) == 0
    mine = _cpu(out)[:, e["flu_times"], :]
    cases.assert_ks_ok(cases.ks_pvalues(mine, np.transpose(e["flu_ens"], (1, 2, 0))), "queue sampler vs reference")
    part, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 100, 0, cases.FLU_N_OUT, 0.5,
                           tdpf=(g["flu_table"], g["flu_step_index"]), seed=21, out_dtype=torch.int32, compat=3 | 256,
                           traj_id0=777)
    assert torch.equal(part, out[:, :, 777:877])          # shard invariance: an item's words depend on (trajectory, step, slot, attempt) only


def test_lane_spread_does_not_change_results(tmp_path):
    """Small ensembles are spread over more warps (32 >> lane_shift trajectories per warp, csrc/jump_kernels.cuh); Philox
    streams are keyed by trajectory, so every spread gives the same counts.  SODM_LANE_SHIFT is read once per process:
    one subprocess per setting."""
    import subprocess
    import sys
    code = (
        "import sys, hashlib, numpy as np, torch\n"
        f"sys.path.insert(0, {ROOT!r})\n"
        "from pysodm_b200 import engine, tutorial_data as td\n"
        "engine.device(0)\n"
        "n = 5003\n"
        "var = {1: np.random.default_rng(0).uniform(0.01, 0.03, n)}\n"
        "out, st = engine.jump_run('influenza_1718_jump', td.influenza_flat_y0(), td.influenza_flat_par(), n, 0, 41, 0.5,\n"
        "                          tdpf=(td.CONTACT_TABLE.reshape(3, 16), td.influenza_schedule()[:80]), var=var, seed=3,\n"
        "                          out_dtype=torch.int32)\n"
        "assert int(st.abs().sum()) == 0\n"
        "print(hashlib.sha256(out.cpu().numpy().tobytes()).hexdigest())\n")
    digests = []
    for shift in ("0", "1", "2", "3", None):
        env = dict(os.environ)
        env.pop("SODM_LANE_SHIFT", None)
        if shift is not None:
            env["SODM_LANE_SHIFT"] = shift
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=300)
        assert r.returncode == 0, r.stderr[-2000:]
        digests.append(r.stdout.strip().splitlines()[-1])
    assert len(set(digests)) == 1, digests


def test_peer_destinations_receive_the_same_rows(eng, golden):
    """`sodm_jump_args.out_peers`: the kernel stores every output element to `out` and to each further destination (peer
    GPUs' buffers in a sharded run; plain second buffers of this GPU here) -- aligned tau-leap, interpolated tau-leap, SSA."""
    import torch
    g = golden["jump_single"]
    full = torch.full((28, cases.FLU_N_OUT, 2000), -1, dtype=torch.int32, device="cuda")
    twin = [torch.full_like(full, -1) for _ in range(2)]
    lo, n = 500, 777
    kw = dict(tdpf=(g["flu_table"], g["flu_step_index"]), seed=99, out_dtype=torch.int32, traj_id0=lo)
    want, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, cases.FLU_N_OUT, 0.5, **kw)
    eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, cases.FLU_N_OUT, 0.5, out=full[:, :, lo:lo + n],
                 out_peers=[t[:, :, lo:].data_ptr() for t in twin], **kw)
    assert torch.equal(full[:, :, lo:lo + n], want)
    for t in twin:
        assert torch.equal(t, full) and int((t[:, :, :lo] != -1).sum()) == 0 and int((t[:, :, lo + n:] != -1).sum()) == 0
    tdpf = (g["flu_tau03_table"], g["flu_tau03_step_index"])
    a = torch.zeros((28, 46, 300), dtype=torch.float64, device="cuda")
    b = torch.zeros_like(a)
    eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 300, 0, 46, 0.3, tdpf=tdpf, seed=8, out=a, out_peers=[b.data_ptr()])
    assert torch.equal(a, b) and float(a.abs().sum()) > 0
    s = golden["ssa"]
    a = torch.zeros((14, 31, 256), dtype=torch.float64, device="cuda")
    b = torch.zeros_like(a)
    eng.jump_run("sir_si_jump", s["sirsi_small_y0"], cases.SIRSI_PAR, 256, 0, 31, method="SSA", seed=9, out=a, out_peers=[b.data_ptr()])
    assert torch.equal(a, b) and float(a.abs().sum()) > 0
    with pytest.raises(Exception, match="peer"):
        eng.jump_run("sir_si_jump", s["sirsi_small_y0"], cases.SIRSI_PAR, 8, 0, 31, out_peers=[0])


def test_fused_gather_over_peer_mappings(eng, golden, tmp_path):
    """Two or three ranks, each simulating its shard: the kernels store straight into every rank's copy of the ensemble (CUDA
    IPC mappings, `sharding.PeerEnsemble`) and rank 0's copy equals the single-process run bit for bit.  NCCL over distinct GPUs
    when the box has them, else two processes sharing one GPU over gloo (the IPC mapping is the same mechanism)."""
    import socket
    import subprocess
    import sys
    import torch
    world = 2 if torch.cuda.device_count() < 3 else 3
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    path = str(tmp_path / "peer.npz")
    with socket.socket() as sck:
        sck.bind(("127.0.0.1", 0))
        port = str(sck.getsockname()[1])
    worker = __file__.replace("test_gpu_jump.py", "dist_peer_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), port, backend, path], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]
    g = golden["jump_single"]
    want, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 3001, 0, cases.FLU_N_OUT, 0.5,
                           tdpf=(g["flu_table"], g["flu_step_index"]), seed=99, out_dtype=torch.int32)
    assert np.array_equal(np.load(path)["full"], _cpu(want))


def test_latency_build_is_the_same_law(eng, golden):
    """SODM_TUNE_LATENCY selects the 96-register build of the same kernel source (csrc/jump_kernels.cuh): reproducible and
    shard-invariant for the flag, the reference's law, and -- because ptxas contracts different multiply-add pairs under the two
    register budgets -- NOT bit-identical to the default build: a few trajectories in 10^5 round differently somewhere."""
    import torch
    from pysodm_b200 import _ffi
    g, e = golden["jump_single"], golden["jump_ensemble"]
    flag = _ffi.COMPAT_REFERENCE | _ffi.TUNE_LATENCY
    kw = dict(tdpf=(g["flu_table"], g["flu_step_index"]), seed=21, out_dtype=torch.int32)
    n = 50_000
    out, st = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, cases.FLU_N_OUT, 0.5, compat=flag, **kw)
    assert int(st.abs().sum()) == 0
    cases.assert_ks_ok(cases.ks_pvalues(_cpu(out)[:, e["flu_times"], :], np.transpose(e["flu_ens"], (1, 2, 0))), "latency build vs reference")
    part, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, 100, 0, cases.FLU_N_OUT, 0.5, compat=flag, traj_id0=777, **kw)
    assert torch.equal(part, out[:, :, 777:877])
    default, _ = eng.jump_run("influenza_1718_jump", g["flu_y0"], cases.FLU_PAR, n, 0, cases.FLU_N_OUT, 0.5, **kw)
    differ = int((out != default).any(dim=0).any(dim=0).sum())
    assert differ < n // 1000, differ                            # the same trajectories but for a handful
