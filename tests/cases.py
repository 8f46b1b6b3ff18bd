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


def assert_ks_ok(ps, what, p_single=0.01, min_global=1e-4):
    """The parity bar (BASELINE.json): KS p > 0.01 per (state, time).  With hundreds of cells a few p-values
    below 0.01 occur by chance under the null (1 % of them), and the cells are strongly CORRELATED (the same
    trajectories seen at neighbouring times / in coupled compartments), so chance failures come in clusters:
    pinned-oracle-vs-pinned-oracle runs of the same size show 0.5-4 % of cells below 0.01.  The test therefore
    requires (i) the NUMBER below 0.01 to stay within that null band (4 % of the cells, or the 99.9 % point of
    Binomial(cells, 0.01) when there are few cells; a wrong law -- e.g. dropping the stay-probability quirk -- fails
    > 90 % of the cells) and (ii) no cell to be astronomically small."""
    from scipy import stats
    ps = np.asarray(ps).ravel()
    n_fail = int(np.sum(ps < p_single))
    # allowed = the larger of 4 % of the cells (clustered chance failures) and the 99.9 % point of Binomial(cells, 1 %)
    allowed = max(int(np.ceil(0.04 * ps.size)), int(stats.binom.isf(1e-3, ps.size, p_single)))
    assert n_fail <= allowed, f"{what}: {n_fail} of {ps.size} KS tests have p < {p_single} (allowed {allowed})"
    assert ps.min() > min_global / max(1, ps.size) * 10, f"{what}: min KS p-value {ps.min():.3e} over {ps.size} cells"
