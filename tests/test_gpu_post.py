"""GPU parity tests of the ensemble post-processing kernels (sodm_noise, sodm_draws_sum, sodm_draws_order_stats) and
of the device-resident front end (`sim(output='device')`, `DeviceEnsemble`, `optimization.utils.add_*_noise`).

Oracle: oracle/post.py (NumPy's own samplers / np.mean / np.quantile, as the reference calls them).  Bars: noise =
law (two-sample KS against NumPy draws per cell); sums of counts and order statistics / quantiles = bit-exact."""
import numpy as np
import pytest
from scipy import stats

import cases
from oracle import post

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def eng():
    from pysodm_b200 import engine
    engine.device(0)
    return engine


def _cpu(t):
    return t.detach().cpu().numpy()


# --------------------------------------------------------------------------------------------------------- noise
@pytest.mark.parametrize("kind,param", [("poisson", 0.0), ("negative_binomial", 0.03), ("negative_binomial", 0.8),
                                        ("gaussian", 2.5), ("gaussian_relative", 0.04)])
def test_noise_law_matches_numpy(eng, kind, param):
    import torch
    n = 30000
    means = np.array([0.0, 0.7, 9.5, 10.0, 55.0, 980.0, 31234.0, 7.3e6])
    x = torch.from_numpy(np.repeat(means[:, None], n, axis=1).copy()).cuda()
    out, status = eng.noise(kind, x, param, seed=17)
    assert int(status.item()) == 0
    mine = _cpu(out)
    rs = np.random.RandomState(4)
    full = np.repeat(means[:, None], n, axis=1)
    ref = {"poisson": lambda: post.add_poisson_noise(full, rs),
           "negative_binomial": lambda: post.add_negative_binomial_noise(full, param, rs),
           "gaussian": lambda: post.add_gaussian_noise(full, param, False, rs),
           "gaussian_relative": lambda: post.add_gaussian_noise(full, param, True, rs)}[kind]()
    ps = cases.ks_pvalues(mine, ref.astype(np.float64))
    assert ps.min() > 1e-3, (kind, ps)
    if kind in ("poisson", "negative_binomial"):
        assert np.all(mine == np.round(mine)) and mine.min() >= 0
        var = means if kind == "poisson" else means + param * means ** 2
        assert np.all(np.abs(mine.mean(axis=1) - means) <= 5 * np.sqrt(var / n) + 1e-12)


def test_noise_integer_ensemble_shards_and_status(eng):
    import torch
    rng = np.random.default_rng(0)
    x = torch.from_numpy(rng.integers(0, 4000, size=(5, 7, 1000)).astype(np.int32)).cuda()
    full, _ = eng.noise("negative_binomial", x, 0.03, seed=9)
    assert full.dtype == torch.int32 and full.shape == x.shape
    lo, _ = eng.noise("negative_binomial", x[..., :384].contiguous(), 0.03, seed=9, traj_id0=0)
    hi, _ = eng.noise("negative_binomial", x[..., 384:].contiguous(), 0.03, seed=9, traj_id0=384)
    assert torch.equal(torch.cat([lo, hi], dim=-1), full)              # keyed by global draw id: shard-invariant
    view, _ = eng.noise("negative_binomial", x[..., :384], 0.03, seed=9)                       # strided view (ld = 1000)
    assert torch.equal(view, lo)
    other, _ = eng.noise("negative_binomial", x, 0.03, seed=10)
    assert not torch.equal(other, full)
    g, _ = eng.noise("gaussian_relative", x, 0.04, seed=1)
    assert g.dtype == torch.float64                                    # Gaussian noise on counts is real-valued
    bad = x.clone()
    bad[2, 3, 17] = -5
    for kind in ("poisson", "negative_binomial"):
        _, status = eng.noise(kind, bad, 0.03, seed=1)
        assert int(status.item()) == 1
    from pysodm_b200._ffi import SodmError
    with pytest.raises(SodmError):
        eng.noise("gaussian", x, -1.0)
    with pytest.raises(SodmError):
        eng.noise("negative_binomial", x, 0.0)
    empty, _ = eng.noise("poisson", x[..., :0], 0.0)
    assert empty.shape[-1] == 0


# ------------------------------------------------------------------------------------------------ sums / order stats
@pytest.mark.parametrize("dtype", [np.int32, np.float32, np.float64])
@pytest.mark.parametrize("n", [1, 2, 31, 512, 4099, 100003])
def test_sum_and_order_statistics_are_exact(eng, dtype, n):
    import torch
    rng = np.random.default_rng(n)
    cells = 13
    if dtype == np.int32:
        a = rng.integers(-2 ** 31, 2 ** 31 - 1, size=(cells, n), dtype=np.int64).astype(np.int32)
        a[1] = rng.integers(0, 7, size=n)                       # heavy ties
        a[2] = 12345
        a[3] = rng.integers(6_000_000, 7_400_000, size=n)       # clustered, as epidemic counts are
    else:
        a = (rng.normal(size=(cells, n)) * 10.0 ** rng.integers(-30, 30, size=(cells, 1))).astype(dtype)
        a[1] = rng.integers(-3, 4, size=n)
        a[2] = -0.0
        a[3, ::2] = 0.0
        if n > 2:
            a[4, 0], a[4, 1] = np.inf, -np.inf
    t = torch.from_numpy(a).cuda()
    s = _cpu(eng.draws_sum(t))
    if dtype == np.int32:
        assert np.array_equal(s, a.astype(np.int64).sum(axis=1).astype(np.float64))
    else:
        finite = np.isfinite(a).all(axis=1)
        want = a.astype(np.float64).sum(axis=1)
        scale = np.abs(a.astype(np.float64)).sum(axis=1)
        assert np.all(np.abs(s - want)[finite] <= 1e-13 * scale[finite] + 1e-300)
    ranks = sorted({0, n - 1, n // 2, (n - 1) // 3, min(n - 1, 7), max(0, n - 2)})
    got = _cpu(eng.draws_order_stats(t, ranks))
    want = np.sort(a, axis=1)[:, ranks].T.astype(np.float64)
    assert np.array_equal(got, want)


def test_order_statistics_nan_strides_and_errors(eng):
    import torch
    rng = np.random.default_rng(3)
    a = rng.normal(size=(4, 6, 1000))
    a[1, 2, 77] = np.nan
    t = torch.from_numpy(a).cuda()
    got = _cpu(eng.draws_order_stats(t, [0, 499, 999]))
    assert got.shape == (3, 4, 6)
    assert np.isnan(got[:, 1, 2]).all()                                 # np.quantile yields NaN for a slice with NaN
    ok = np.ones((4, 6), bool)
    ok[1, 2] = False
    assert np.array_equal(got[:, ok], np.sort(a, axis=-1)[..., [0, 499, 999]][ok].T)
    view = t[..., 100:900]                                              # ld = 1000, n_traj = 800, unaligned start
    got = _cpu(eng.draws_order_stats(view, [0, 400, 799]))
    want = np.sort(a[..., 100:900], axis=-1)[..., [0, 400, 799]]
    assert np.array_equal(got[:, ok], np.moveaxis(want, -1, 0)[:, ok])
    assert np.allclose(_cpu(eng.draws_sum(view))[ok], a[..., 100:900].sum(-1)[ok], rtol=1e-13)
    many = list(range(0, 1000, 97))                                     # more ranks than one launch takes
    assert np.array_equal(_cpu(eng.draws_order_stats(t[0], many)), np.sort(a[0], axis=-1)[:, many].T)
    from pysodm_b200._ffi import SodmError
    with pytest.raises(SodmError):
        eng.draws_order_stats(t, [1000])
    with pytest.raises(ValueError):
        eng.draws_order_stats(t.transpose(0, 2), [0])


# ------------------------------------------------------------------------------------------------------ front end
def _influenza(td):
    from pysodm_b200 import tutorial_models as tm
    return tm.JumpProcess_influenza_model(td.INFLUENZA_INITIAL_STATES, td.INFLUENZA_PARAMETERS,
                                          coordinates={"age_group": td.AGE_GROUPS},
                                          time_dependent_parameters={"N": td.influenza_contact_function})


def test_device_ensemble_matches_host_path_and_numpy():
    from pysodm_b200 import tutorial_data as td
    from pysodm_b200.ensemble import DeviceEnsemble
    from pysodm_b200.optimization.utils import add_negative_binomial_noise, add_poisson_noise, add_gaussian_noise
    model = _influenza(td)
    span = ["2017-12-01", "2018-01-15"]
    n = 3000
    host = model.sim(span, N=n, tau=0.5, seed=21)
    ens = model.sim(span, N=n, tau=0.5, seed=21, output="device")
    assert isinstance(ens, DeviceEnsemble) and list(ens.keys()) == list(host.keys())
    assert ens["S"].dims == ("draws", "date", "age_group") and ens["S"].shape == (n, 46, 4)
    full = ens.to_dataset()
    for name in host.keys():
        assert np.array_equal(np.asarray(full[name].values), np.asarray(host[name].values)), name   # same seed, same draws
    # mean / quantile over draws: bit-identical to NumPy on the host copy
    qs = [0.025, 0.5, 0.975]
    m, q = ens.mean(dim="draws"), ens.quantile(qs, dim="draws")
    for name in host.keys():
        h = np.asarray(host[name].values)
        assert np.array_equal(np.asarray(m[name].values), post.mean_over_draws(h)), name
        assert np.array_equal(np.asarray(q[name].values), post.quantile_over_draws(h, qs)), name
        assert tuple(q[name].dims) == ("quantile", "date", "age_group") and tuple(m[name].dims) == ("date", "age_group")
    # the tutorial's expression (calibration.py:245-246): one age class, scalar q
    one = ens["Im_inc"].sel(age_group=td.AGE_GROUPS[2])
    h = np.asarray(host["Im_inc"].values)[:, :, 2]
    assert np.array_equal(np.asarray(one.quantile(dim="draws", q=0.975).values), np.quantile(h, 0.975, axis=0))
    assert np.array_equal(np.asarray(one.mean(dim="draws").values), h.mean(axis=0))
    last = ens["S"].isel(date=-1)
    assert last.dims == ("draws", "age_group")
    assert np.array_equal(np.asarray(last.quantile(0.5).values), np.quantile(np.asarray(host["S"].values)[:, -1], 0.5, axis=0))
    assert np.array_equal(ens["R"].sel(date=slice("2017-12-05", "2017-12-09")).values, np.asarray(host["R"].values)[:, 4:9])
    # noise on the device ensemble: law against NumPy applied to the host copy
    np.random.seed(5)
    noisy = add_negative_binomial_noise(ens, 0.03)
    assert isinstance(noisy, DeviceEnsemble) and noisy.buffer.data_ptr() != ens.buffer.data_ptr()
    mine = noisy["Im_inc"].values
    ref = post.add_negative_binomial_noise(np.asarray(host["Im_inc"].values), 0.03, np.random.RandomState(1))
    times = [1, 10, 25, 45]
    ps = cases.ks_pvalues(np.moveaxis(mine[:, times], 0, -1), np.moveaxis(ref[:, times].astype(np.float64), 0, -1))
    cases.assert_ks_ok(ps, "NB noise on the influenza ensemble")
    np.random.seed(5)
    again = add_negative_binomial_noise(ens, 0.03)
    assert np.array_equal(again["Im_inc"].values, mine)                 # reproducible through NumPy's global seed
    band = noisy["Im_inc"].quantile([0.025, 0.975])
    assert np.all(np.asarray(band.values)[0] <= np.asarray(band.values)[1])
    assert np.array_equal(np.asarray(band.values), np.quantile(mine, [0.025, 0.975], axis=0))
    # noise on one variable only (what the tutorial plots): same law, a seventh of the work
    one_noisy = add_negative_binomial_noise(ens["Im_inc"].sel(age_group=td.AGE_GROUPS[1]), 0.03, seed=8)
    assert one_noisy.dims == ("draws", "date") and one_noisy.shape == (n, 46)
    ref1 = post.add_negative_binomial_noise(np.asarray(host["Im_inc"].values)[:, :, 1], 0.03, np.random.RandomState(2))
    ps = cases.ks_pvalues(one_noisy.values[:, times].T, ref1[:, times].T.astype(np.float64))
    cases.assert_ks_ok(ps, "NB noise on one state variable")
    assert np.array_equal(np.asarray(one_noisy.quantile(0.5).values), np.quantile(one_noisy.values, 0.5, axis=0))
    # host Datasets go through the device as well (drop-in form of the reference call)
    small = model.sim(span, N=50, tau=0.5, seed=2)
    for fn, args in ((add_poisson_noise, ()), (add_negative_binomial_noise, (0.03,)), (add_gaussian_noise, (0.04,))):
        res = fn(small, *args)
        assert np.asarray(res["S"].values).shape == np.asarray(small["S"].values).shape
        assert not np.array_equal(np.asarray(res["Im_inc"].values), np.asarray(small["Im_inc"].values))
    assert np.asarray(add_poisson_noise(small)["S"].values).dtype == np.int64


def test_device_ensemble_ode_and_full_size_properties():
    from pysodm_b200 import tutorial_data as td, tutorial_models as tm
    from pysodm_b200.optimization.utils import add_gaussian_noise
    # ODE ensemble (float64 buffer): SIR with drawn beta (tutorials/SIR/workflow_tutorial.py:49-70)
    model = tm.ODE_SIR({"S": 1000 - 1, "I": 1}, {"beta": 0.35, "gamma": 5})

    def draw(parameters, N):
        parameters["beta"] = np.random.default_rng(1).uniform(0.25, 0.45, size=N)
        return parameters
    draw.vectorized = True
    host = model.sim(120, N=500, draw_function=draw)
    ens = model.sim(120, N=500, draw_function=draw, output="device")
    for name in ("S", "I", "R"):
        h = np.asarray(host[name].values)
        assert np.array_equal(ens[name].values, h)
        assert np.array_equal(np.asarray(ens[name].quantile(0.3).values), np.quantile(h, 0.3, axis=0))
        assert np.allclose(np.asarray(ens[name].mean().values), h.mean(axis=0), rtol=1e-13)
    noisy = add_gaussian_noise(ens, 0.04, relative=True, seed=3)
    z = (noisy["I"].values - np.asarray(host["I"].values)) / (0.04 * np.asarray(host["I"].values))
    assert abs(z.mean()) < 0.01 and abs(z.std() - 1) < 0.01
    # full-size influenza ensemble: properties that need no oracle
    flu = _influenza(td)
    n = 200_000
    ens = flu.sim([td.INFLUENZA_START, td.INFLUENZA_END], N=n, tau=0.5, seed=1, output="device")
    m = ens.mean()
    total = sum(np.asarray(m[s].values) for s in ("S", "E", "Ip", "Iud", "Id", "R"))
    pop = sum(np.asarray(td.INFLUENZA_INITIAL_STATES.get(s, np.zeros(4)), dtype=np.float64)
              for s in ("S", "E", "Ip", "Iud", "Id", "R"))
    assert np.allclose(total, pop, rtol=1e-14, atol=0)                   # population conserved in every draw => in the mean
    exact = ens.buffer[:24].view(6, 4, 162, n).sum(dim=0)                # S+E+Ip+Iud+Id+R per (age, time, draw), int64 on device
    assert bool((exact == exact[:, :1, :1]).all())
    q = ens["Im_inc"].quantile([0.0, 0.025, 0.5, 0.975, 1.0])
    qv = np.asarray(q.values)
    assert np.all(np.diff(qv, axis=0) >= 0)
    mv = np.asarray(m["Im_inc"].values)
    assert np.all(qv[0] <= mv) and np.all(mv <= qv[-1])


def _run_ranks(world, backend, path):
    import socket
    import subprocess
    import sys
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = str(s.getsockname()[1])
    worker = __file__.replace("test_gpu_post.py", "dist_post_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), str(world), port, backend, path], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT, text=True) for r in range(world)]
    outs = [p.communicate(timeout=600)[0] for p in procs]
    for p, o in zip(procs, outs):
        assert p.returncode == 0, o[-3000:]


def test_sharded_ensemble_reduces_like_one_gpu(tmp_path):
    """An ensemble sharded by trajectory over the ranks of a torch.distributed job: mean, quantiles (radix select in
    phases with one all-reduce of the histograms per pass) and noise equal the single-process results bit for bit.
    Runs with as many GPUs as the box has (NCCL when there are >= 2, else two ranks on one GPU over gloo)."""
    import torch
    from pysodm_b200 import tutorial_data as td, tutorial_models as tm
    from pysodm_b200.optimization.utils import add_negative_binomial_noise
    world = 2 if torch.cuda.device_count() < 3 else 3
    backend = "nccl" if torch.cuda.device_count() >= world else "gloo"
    path = str(tmp_path / "sharded.npz")
    _run_ranks(world, backend, path)
    got = np.load(path)
    model = tm.JumpProcess_influenza_model(td.INFLUENZA_INITIAL_STATES, td.INFLUENZA_PARAMETERS,
                                           coordinates={"age_group": td.AGE_GROUPS},
                                           time_dependent_parameters={"N": td.influenza_contact_function})

    def draw(parameters, N):
        rng = np.random.default_rng(0)
        parameters["beta"] = rng.uniform(0.015, 0.02, size=N)
        return parameters
    draw.vectorized = True
    n = 5001
    ens = model.sim(["2017-12-01", "2018-01-10"], N=n, tau=0.5, seed=21, output="device", draw_function=draw)
    full = ens.to_dataset()
    assert np.array_equal(got["full_S"], np.asarray(full["S"].values))                 # union of the shards == one GPU
    assert np.array_equal(got["full_Im"], np.asarray(full["Im_inc"].values))
    qs = [0.0, 0.025, 0.5, 0.975, 1.0]
    for key, name in (("S", "S"), ("Im", "Im_inc")):
        h = np.asarray(full[name].values)
        assert np.array_equal(got["mean_" + key], h.mean(axis=0))
        assert np.array_equal(got["q_" + key], np.quantile(h, qs, axis=0))
    noisy = add_negative_binomial_noise(ens, 0.03, seed=5)
    assert np.array_equal(got["full_noisy"], noisy["Im_inc"].values)                   # noise keyed by global draw id
    assert np.array_equal(got["qn"], np.quantile(noisy["Im_inc"].values[:, :, 2], 0.9, axis=0))
    e = np.asarray(full["E"].values) * 0.37 - 11.0
    assert np.array_equal(got["qf"], np.quantile(e, [0.1, 0.66], axis=0))
    assert np.allclose(got["mf"], e.mean(axis=0), rtol=1e-13)
