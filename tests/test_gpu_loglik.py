"""GPU parity tests of the batched log-likelihood (sodm_loglik) and the layout kernel (sodm_to_draws_major).

Parity bar (BASELINE.json): log-posterior values within 1e-9 relative on identical trajectories."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu
REL = 1e-9


@pytest.fixture(scope="module")
def eng():
    from pysodm_b200 import engine
    engine.device(0)
    return engine


def _as_traj(ymodel, n_copies=3):
    """[n_t, n_col] model prediction -> trajectory tensor [n_col, n_t, n_copies] on an integer grid."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(ymodel.T)).cuda()
    return t[:, :, None].expand(-1, -1, n_copies).contiguous()


def test_golden_values(eng, golden):
    """Values produced by the unmodified reference `ll_*` functions (objective_functions.py:550-666)."""
    g = golden["loglik"]
    ym, yd = g["ymodel"], g["ydata"]
    traj = _as_traj(ym)
    cols = [[c] for c in range(4)]
    t = np.arange(13.0)
    cases = [("poisson", None, g["poisson"]),
             ("negative_binomial", 4 * [0.03], g["negbin_vec"]),
             ("negative_binomial", 0.07, g["negbin_scalar"]),
             ("normal", 12.0, g["normal_scalar"]),
             ("normal", g["sigma_col"], g["normal_col"]),
             ("normal", g["sigma_full"], g["normal_full"]),
             ("lognormal", 0.4, g["lognormal_scalar"]),
             ("lognormal", [0.2, 0.3, 0.4, 0.5], g["lognormal_col"])]
    for kind, arg, want in cases:
        lp, st = eng.loglik(kind, traj, 0.0, cols, t, yd, ll_arg=arg)
        got = lp.cpu().numpy()
        assert (st.cpu().numpy() == 0).all()
        np.testing.assert_allclose(got, float(want), rtol=REL, err_msg=f"{kind} {arg}")
    traj1 = _as_traj(g["ymodel1"])
    for kind, arg, want in [("poisson", None, g["poisson1"]), ("negative_binomial", 0.05, g["negbin1"]),
                            ("normal", 4.0, g["normal1"])]:
        lp, _ = eng.loglik(kind, traj1, 0.0, [[0]], np.arange(20.0), g["ydata1"], ll_arg=arg)
        np.testing.assert_allclose(lp.cpu().numpy(), float(want), rtol=REL)


def test_alignment_against_oracle(eng):
    """Sum over model dimensions absent from the data + linear time interpolation to non-grid data times
    (the reference's `linspace(50, 60, 20)` test data, src/tests/test_calibration.py:20-36), int32 and float64
    trajectories, weights and accumulation."""
    import torch
    from oracle import loglik as oll
    rng = np.random.default_rng(5)
    n, n_row, n_out = 257, 12, 40
    traj = rng.integers(0, 5000, size=(n_row, n_out, n)).astype(np.int32)
    cols = [[0, 4, 8], [1, 5, 9], [11]]
    t_data = np.concatenate([np.linspace(50, 60, 20), [50.0, 89.0]]) - 0.0
    t0 = 50.0
    ydata = rng.integers(0, 9000, size=(len(t_data), 3)).astype(np.float64)
    for dtype in (torch.int32, torch.float64):
        d = torch.from_numpy(traj).cuda().to(dtype)
        for kind, arg in [("poisson", None), ("negative_binomial", [0.03, 0.1, 0.5]), ("normal", 35.0),
                          ("lognormal", 0.3)]:
            lp = torch.full((n,), 2.5, dtype=torch.float64, device="cuda")
            eng.loglik(kind, d, t0, cols, t_data, ydata, ll_arg=arg, weight=0.7, lp=lp)
            want = np.array([2.5 + oll.log_likelihood(kind, traj[:, :, i], t0, cols, t_data, ydata, arg, 0.7)
                             for i in range(n)])
            np.testing.assert_allclose(lp.cpu().numpy(), want, rtol=REL, err_msg=kind)


def test_data_outside_simulation_is_flagged(eng):
    import torch
    traj = torch.ones((1, 5, 8), dtype=torch.float64, device="cuda")
    lp, st = eng.loglik("poisson", traj, 0.0, [[0]], [1.0, 4.5], [[3.0], [2.0]])
    assert (st.cpu().numpy() & 2).all() and np.isnan(lp.cpu().numpy()).all()


@pytest.mark.parametrize("n", [1, 33, 1000])
def test_to_draws_major(eng, n):
    import torch
    rng = np.random.default_rng(n)
    src = torch.from_numpy(rng.integers(-5, 10**6, size=(9, 37, n)).astype(np.int32)).cuda()
    for r0, n_dim in [(0, 4), (4, 4), (8, 1)]:
        want = src[r0:r0 + n_dim].permute(2, 1, 0).contiguous()
        got = eng.to_draws_major(src, r0, n_dim, dst_dtype=torch.float64)
        assert torch.equal(got, want.to(torch.float64))
        got = eng.to_draws_major(src, r0, n_dim, dst_dtype=torch.int32)
        assert torch.equal(got, want)
    host = torch.empty((n, 37, 4), dtype=torch.float64).pin_memory()
    eng.to_draws_major(src, 4, 4, dst=host)
    torch.cuda.synchronize()
    assert torch.equal(host, src[4:8].permute(2, 1, 0).contiguous().cpu().to(torch.float64))
