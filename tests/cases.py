"""Shared inputs of the parity tests (the BASELINE.json configurations)."""
import numpy as np

# influenza_1718 (tutorials/influenza_1718/calibration.py:105-117): flat parameter vector
#   alpha, beta, gamma, delta, f_ud[4], N[16] (N comes from the time-dependent table)
FLU_PAR = np.concatenate([[1.0, 0.0174, 1.0, 3.0], [0.01, 0.64, 0.905, 0.60], np.zeros(16)])
FLU_PAR_DICT = {"alpha": 1, "beta": 0.0174, "gamma": 1, "delta": 3, "f_ud": np.array([0.01, 0.64, 0.905, 0.60])}
FLU_N_OUT = 162          # 2017-12-01 .. 2018-05-11
FLU_IDX_BETA, FLU_IDX_FUD = 1, (4, 5, 6, 7)

# SIR_SI (tutorials/SIR_SI/calibration.py:31-32): alpha[4], beta, gamma
SIRSI_PAR = np.array([0.05, 0.1, 0.2, 0.15, 7.0, 5.0])
SIRSI_PAR_DICT = {"alpha": np.array([0.05, 0.1, 0.2, 0.15]), "gamma": 5, "beta": 7}

# test SIR (src/tests/test_JumpProcess.py:144-151): beta, gamma (rate)
SIR_PAR = np.array([0.9, 0.2])


def flu_draws(n, seed=1718):
    """Synthetic per-trajectory draws of config C3: beta_i, f_ud_i."""
    rng = np.random.default_rng(seed)
    z = rng.standard_normal(n)
    u = rng.uniform(0.9, 1.1, size=(n, 4))
    beta = np.clip(0.0174 * (1 + 0.05 * z), 1e-6, 0.08)
    f_ud = np.clip(np.array([0.01, 0.64, 0.905, 0.60]) * u, 1e-3, 1 - 1e-3)
    return beta, f_ud


def _oracle_chunk(job):
    which, n, seed, times, with_draws = job
    from oracle import jump, models
    from pysodm_b200 import tutorial_data as td
    if which == "flu":
        par = dict(FLU_PAR_DICT)
        if with_draws:
            beta, f_ud = flu_draws(n, seed)
            par.update(beta=beta, f_ud=f_ud)
        y = jump.tau_leap_ensemble(models.InfluenzaJump, td.influenza_flat_y0(), par, 0, FLU_N_OUT - 1, 0.5, n,
                                   tdpf={"N": (td.CONTACT_TABLE, td.influenza_schedule())}, seed=seed, chunk=1024)
    else:
        y = jump.tau_leap_ensemble(models.SIRSIJump, td.sirsi_flat_y0(), dict(td.SIRSI_PARAMETERS), 0, 120, 0.5, n, seed=seed, chunk=2048)
    return y[:, :, times].astype(np.int32)


def oracle_ensemble_parallel(which, n, times, seed=5000, with_draws=False, per_job=2048):
    """n trajectories of the pinned oracle (oracle/jump.py, reference law) on all host cores -> int32 [n, states, len(times)].
    Config C3's per-trajectory draws (`flu_draws`) are re-drawn per job from the job's seed: the same DISTRIBUTION of
    parameters as the device run, not the same values."""
    import multiprocessing as mp
    import os
    jobs = [(which, min(per_job, n - lo), seed + 7 * k, np.asarray(times), with_draws) for k, lo in enumerate(range(0, n, per_job))]
    with mp.get_context("fork").Pool(min(len(jobs), os.cpu_count())) as pool:
        parts = pool.map(_oracle_chunk, jobs, chunksize=1)
    return np.concatenate(parts, axis=0)


def ks_pvalues_counts_device(a, b):
    """Two-sample KS p-values of integer samples, computed on the device: a[cells, na], b[cells, nb] (torch int32 CUDA tensors)
    -> NumPy [cells].  Statistic from the two cumulated histograms over the cell's value range; p from the asymptotic
    Kolmogorov law with en = na nb / (na + nb) (scipy.stats.ks_2samp's `asymp` mode, which it uses at these sizes)."""
    import torch
    from scipy.stats import distributions
    na, nb = a.shape[1], b.shape[1]
    ps = np.ones(a.shape[0])
    for r in range(a.shape[0]):
        lo = int(min(a[r].min(), b[r].min()))
        hi = int(max(a[r].max(), b[r].max()))
        if lo == hi:
            continue
        ca = torch.cumsum(torch.bincount((a[r] - lo).long(), minlength=hi - lo + 1), 0).double() / na
        cb = torch.cumsum(torch.bincount((b[r] - lo).long(), minlength=hi - lo + 1), 0).double() / nb
        d = float((ca - cb).abs().max())
        ps[r] = float(np.clip(distributions.kstwo.sf(d, np.round(na * nb / (na + nb))), 0, 1))
    return ps


def ks_pvalues(a, b):
    """Two-sample KS p-values along the last axis of a[..., na] vs b[..., nb]; constant-equal cells give 1."""
    from scipy import stats
    a2 = a.reshape(-1, a.shape[-1])
    b2 = b.reshape(-1, b.shape[-1])
    ps = np.ones(a2.shape[0])
    for r in range(a2.shape[0]):
        if a2[r].min() == a2[r].max() and b2[r].min() == b2[r].max() and a2[r, 0] == b2[r, 0]:
            continue
        ps[r] = stats.ks_2samp(a2[r], b2[r]).pvalue
    return ps.reshape(a.shape[:-1])


_NULL_FRACTION = None


def null_fraction(level=0.985):
    """Fraction of (state, time) cells below p = 0.01 that a SAME-LAW comparison stays under with probability `level`, from
    the committed null fixture tests/golden/ks_null.npz (tests/golden/make_ks_null.py: 200 independent pairs of 20 000-member
    ensembles of the pinned oracle per model, same states / output times as the parity fixtures; pooled over the influenza
    and SIR_SI models).  The null is NOT Binomial(cells, 0.01): the cells are strongly correlated (the same trajectories at
    neighbouring times, coupled compartments), so most replicates have 0-2 cells below 0.01 and a few have a whole cluster
    (influenza: up to 22 of 224; SIR_SI: 4 of 200 replicates with 36-46 of 84)."""
    global _NULL_FRACTION
    if _NULL_FRACTION is None:
        import os
        d = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "ks_null.npz"))
        frac = np.concatenate([d[w + "_below_001"] / float(d[w + "_cells"]) for w in ("flu", "sirsi")])
        _NULL_FRACTION = (float(np.quantile(frac, level)), float(min(d["flu_min_p"].min(), d["sirsi_min_p"].min())))
    return _NULL_FRACTION


def assert_ks_ok(ps, what, p_single=0.01):
    """The parity bar (BASELINE.json): KS p > 0.01 per (state, time).  Over hundreds of correlated cells some p-values fall
    below 0.01 under the null itself, in clusters; how many is MEASURED, not argued: the allowance is the 98.5 % point of the
    committed oracle-vs-oracle null (`null_fraction`, ~4 % of the cells), or the 99.9 % point of Binomial(cells, 0.01) when
    there are few cells -- i.e. a correct sampler fails a test with probability ~1.5 %, while a wrong law (dropping the
    stay-probability quirk, say) fails > 90 % of the cells.  No cell may fall a decade below the smallest p-value the null
    fixture has seen over 200 x 308 cells (1.5e-4)."""
    from scipy import stats
    ps = np.asarray(ps).ravel()
    n_fail = int(np.sum(ps < p_single))
    frac, null_min_p = null_fraction()
    allowed = max(int(np.ceil(frac * ps.size)), int(stats.binom.isf(1e-3, ps.size, p_single)))
    assert n_fail <= allowed, f"{what}: {n_fail} of {ps.size} KS tests have p < {p_single} (allowed {allowed})"
    assert ps.min() > 0.1 * null_min_p, f"{what}: min KS p-value {ps.min():.3e} over {ps.size} cells"
