"""Generate the golden fixtures under tests/golden/ by RUNNING THE UNMODIFIED REFERENCE.

Run in the build container only (needs /root/reference):   python tests/golden/make_golden.py [what ...]
`what` in {jump_single, jump_ensemble, ssa, ode, loglik, schedule}; default = all.

The fixtures pin the oracle (oracle/*.py) -- see oracle/__init__.py -- and are what the GPU parity
tests compare against on a box where the reference is absent.
"""
import os
import random
import sys
import time
from datetime import datetime

import numpy as np
import pandas as pd

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_loader  # noqa: E402

base, objf = ref_loader.load()

# ------------------------------------------------------------------------------------------------
# model set-ups, as the tutorials / tests make them
# ------------------------------------------------------------------------------------------------

# tutorials/influenza_1718/calibration.py:82-95
N_NOHOLIDAY_WEEK = np.transpose(np.array([[10.47, 3.16, 10.65, 0.55], [1.65, 19.41, 11.79, 0.70],
                                          [0.93, 1.98, 14.09, 0.96], [0.18, 0.45, 3.67, 3.09]]))
N_NOHOLIDAY_WEEKEND = np.transpose(np.array([[1.81, 2.46, 8.92, 1.52], [1.28, 7.07, 14.76, 1.64],
                                             [0.78, 2.49, 15.26, 1.17], [0.51, 1.05, 4.45, 2.00]]))
N_HOLIDAY_WEEK = np.transpose(np.array([[4.54, 2.47, 10.86, 0.35], [1.29, 5.76, 9.64, 1.47],
                                        [0.96, 1.63, 13.08, 1.38], [0.12, 0.94, 5.23, 5.10]]))
CONTACTS = {'holiday': {'week': N_HOLIDAY_WEEK, 'weekend': N_NOHOLIDAY_WEEKEND},
            'no_holiday': {'week': N_NOHOLIDAY_WEEK, 'weekend': N_NOHOLIDAY_WEEKEND}}
AGE_GROUPS = pd.IntervalIndex.from_tuples([(0, 5), (5, 15), (15, 65), (65, 120)], closed='left')


def influenza_data():
    path = os.path.join(ref_loader.REFERENCE_ROOT, 'tutorials/influenza_1718/data/interim/ILI_weekly_ABS.csv')
    data = pd.read_csv(path, index_col=[0, 1], parse_dates=True, date_format='%Y-%m-%d').squeeze()
    dates = data.index.get_level_values('DATE').unique()
    index = pd.MultiIndex.from_product([dates, AGE_GROUPS], names=['date', 'age_group'])
    return pd.Series(index=index, name='CASES', data=data.values) / 7       # calibration.py:62


def make_influenza():
    flu = ref_loader.load_tutorial_module('tutorials/influenza_1718/models.py', 'ref_flu_models')
    df = influenza_data()
    start = df.index.get_level_values('date').unique()[8].strftime('%Y-%m-%d')
    params = {'alpha': 1, 'beta': 0.0174, 'gamma': 1, 'delta': 3,
              'f_ud': np.array([0.01, 0.64, 0.905, 0.60]), 'N': CONTACTS['holiday']['week']}
    row = df.loc[start, slice(None)]
    init = {'S': [620914, 1306826, 7317774, 2130556],
            'E': list(np.rint((1 / (1 - params['f_ud'])) * row)),
            'Ip': list(np.rint((1 / (1 - params['f_ud'])) * row)),
            'Iud': list(np.rint((params['f_ud'] / (1 - params['f_ud'])) * row)),
            'Id': list(np.rint(row)),
            'Im_inc': list(np.rint(row))}
    cf = flu.make_contact_matrix_function(CONTACTS).contact_function
    model = flu.JumpProcess_influenza_model(initial_states=init, parameters=params,
                                            coordinates={'age_group': AGE_GROUPS},
                                            time_dependent_parameters={'N': cf})
    return model, cf, df, start


def make_sir_si(jump=True, scale=1.0):
    m = ref_loader.load_tutorial_module('tutorials/SIR_SI/models.py', 'ref_sirsi_models')
    params = {'alpha': np.array([0.05, 0.1, 0.2, 0.15]), 'gamma': 5, 'beta': 7}   # SIR_SI/calibration.py:31
    S = [606938, 1328733, 7352492, 2204478]
    init = {'S': [round(s * scale) for s in S], 'S_v': round(1e6 * scale), 'I_v': 2}
    cls = m.JumpProcess_SIR_SI if jump else m.ODE_SIR_SI
    return cls(init, params, coordinates={'age_group': AGE_GROUPS})


def make_test_sir_jump():
    class SIR(base.JumpProcess):                 # src/tests/test_JumpProcess.py:11-27
        states = ['S', 'I', 'R']
        parameters = ['beta', 'gamma']

        @staticmethod
        def compute_rates(t, S, I, R, beta, gamma):
            return {'S': [beta * (I / (S + I + R)), ], 'I': [np.array([gamma]), ]}

        @staticmethod
        def apply_transitionings(t, tau, transitionings, S, I, R, beta, gamma):
            S_new = S - transitionings['S'][0]
            I_new = I + transitionings['S'][0] - transitionings['I'][0]
            R_new = R + transitionings['I'][0]
            return S_new, I_new, R_new
    return SIR(initial_states={'S': 1_000_000 - 10, 'I': 10}, parameters={'beta': 0.9, 'gamma': 0.2})


def contact_schedule(cf, start, n_steps, tau):
    """Evaluate the reference TDPF at every tau-leap step time, as base.py:365-380 does."""
    start_date = datetime.strptime(start, '%Y-%m-%d')
    mats = [CONTACTS['no_holiday']['week'], CONTACTS['no_holiday']['weekend'], CONTACTS['holiday']['week']]
    idx = np.empty(n_steps, dtype=np.int32)
    t = 0
    for k in range(n_steps):
        date = base.time_to_date(float(t), start_date, 'D')
        M = cf(date, {}, None)
        idx[k] = [i for i, m in enumerate(mats) if M is m][0]
        t = t + tau
    return np.stack(mats), idx


def flat_y0(model):
    return np.concatenate([np.ravel(v) for v in model.initial_states.values()]).astype(np.float64)


def save(name, **arrays):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **arrays)
    print(f'wrote {path}  ({os.path.getsize(path) / 1e6:.2f} MB)')


# ------------------------------------------------------------------------------------------------
def gen_jump_single():
    """Seeded single trajectories (global NumPy RNG seeded right before each sim call)."""
    out = {}
    flu, cf, df, start = make_influenza()
    table, idx = contact_schedule(cf, start, 322, 0.5)
    out['flu_table'], out['flu_step_index'], out['flu_y0'] = table, idx, flat_y0(flu)
    seeds = [11, 12, 13]
    ys = []
    for s in seeds:
        np.random.seed(s)
        r = flu.sim([start, '2018-05-11'], N=1, tau=0.5)
        ys.append(r['y'])
    out['flu_seeds'], out['flu_y'] = np.array(seeds), np.stack(ys)
    # non-dyadic tau -> real interpolation (F3/Q4/Q5)
    np.random.seed(21)
    out['flu_tau03_y'] = flu.sim([start, '2018-01-15'], N=1, tau=0.3)['y']
    t, n = 0, 0
    while t < 45:
        t = t + 0.3
        n += 1
    out['flu_tau03_table'], out['flu_tau03_step_index'] = contact_schedule(cf, start, n, 0.3)

    sirsi = make_sir_si()
    out['sirsi_y0'] = flat_y0(sirsi)
    ys = []
    for s in seeds:
        np.random.seed(s)
        ys.append(sirsi.sim(120, N=1, tau=0.5)['y'])
    out['sirsi_y'] = np.stack(ys)
    np.random.seed(5)
    out['sirsi_tau1_y'] = sirsi.sim(60, N=1, tau=1)['y']

    sir = make_test_sir_jump()
    out['sir_y0'] = flat_y0(sir)
    ys = []
    for s in seeds:
        np.random.seed(s)
        ys.append(sir.sim(50, N=1, tau=1)['y'])
    out['sir_y'] = np.stack(ys)
    save('jump_single.npz', **out)


def gen_ssa():
    """Seeded SSA runs; population / 1000 so that a run is ~6e3 events."""
    out = {}
    sirsi = make_sir_si(scale=1e-3)
    sirsi.initial_states['I_v'] = np.array(2.0)
    out['sirsi_small_y0'] = flat_y0(sirsi)
    ys = []
    seeds = [3, 4]
    for s in seeds:
        np.random.seed(s)
        random.seed(1000 + s)
        ys.append(sirsi.sim(30, N=1, method='SSA')['y'])
    out['seeds'], out['sirsi_small_y'] = np.array(seeds), np.stack(ys)
    # ensemble for KS (serial; 30 days of the /1000 population is ~1e3 events per run)
    t = time.time()
    np.random.seed(99)
    random.seed(99)
    n = 400
    res = sirsi.sim(30, N=n, method='SSA')
    out['sirsi_small_ens'] = np.stack([r['y'] for r in res]).astype(np.float32)
    print(f'SSA ensemble n={n}: {time.time() - t:.1f}s')
    save('ssa.npz', **out)


def gen_jump_ensemble(n=2000, processes=8):
    """Reference ensembles for KS tests of the batched oracle (and, transitively, of the GPU)."""
    out = {}
    flu, cf, df, start = make_influenza()
    t = time.time()
    np.random.seed(1718)
    res = flu.sim([start, '2018-05-11'], N=n, tau=0.5, processes=processes)
    Y = np.stack([r['y'] for r in res])                      # [n, 28, 162]
    print(f'influenza reference ensemble n={n}: {time.time() - t:.1f}s ({n / (time.time() - t):.1f} traj/s)')
    times = np.array([1, 3, 7, 14, 28, 42, 56, 70, 84, 100, 120, 140, 161])
    out['flu_times'], out['flu_ens'] = times, Y[:, :, times].astype(np.int32)
    out['flu_mean'] = Y.mean(axis=0)
    out['flu_q'] = np.quantile(Y, [0.025, 0.5, 0.975], axis=0)

    sirsi = make_sir_si()
    t = time.time()
    np.random.seed(4242)
    res = sirsi.sim(120, N=n, tau=0.5, processes=processes)
    Y = np.stack([r['y'] for r in res])                      # [n, 14, 121]
    print(f'SIR_SI reference ensemble n={n}: {time.time() - t:.1f}s')
    times = np.array([1, 5, 10, 20, 40, 60, 80, 100, 120])
    out['sirsi_times'], out['sirsi_ens'] = times, Y[:, :, times].astype(np.int32)
    out['sirsi_mean'] = Y.mean(axis=0)
    out['sirsi_q'] = np.quantile(Y, [0.025, 0.5, 0.975], axis=0)
    save('jump_ensemble.npz', **out)


def gen_jump_ensemble_big(n=20000, processes=8):
    """20 000-member reference ensembles (a 2 000-member ensemble resolves only ~0.036 in CDF).
    Per (state, time) cell only the MARGINAL law enters a KS test, so each cell's sample is stored sorted (int32; sorted
    columns compress ~4x better) at a few output times.  ~10 min on 8 cores."""
    out = {}
    flu, cf, df, start = make_influenza()
    t = time.time()
    np.random.seed(171820)
    res = flu.sim([start, '2018-05-11'], N=n, tau=0.5, processes=processes)
    Y = np.stack([r['y'] for r in res])                      # [n, 28, 162]
    print(f'influenza reference ensemble n={n}: {time.time() - t:.1f}s ({n / (time.time() - t):.1f} traj/s)')
    times = np.array([1, 7, 28, 56, 84, 112, 140, 161])
    out['flu_times'] = times
    out['flu_sorted'] = np.sort(Y[:, :, times], axis=0).astype(np.int32)          # [n, 28, 8] sorted along members
    out['flu_mean'] = Y.mean(axis=0)
    out['flu_var'] = Y.var(axis=0)
    del res, Y

    sirsi = make_sir_si()
    t = time.time()
    np.random.seed(424220)
    res = sirsi.sim(120, N=n, tau=0.5, processes=processes)
    Y = np.stack([r['y'] for r in res])                      # [n, 14, 121]
    print(f'SIR_SI reference ensemble n={n}: {time.time() - t:.1f}s')
    times = np.array([1, 10, 30, 60, 90, 120])
    out['sirsi_times'] = times
    out['sirsi_sorted'] = np.sort(Y[:, :, times], axis=0).astype(np.int32)
    out['sirsi_mean'] = Y.mean(axis=0)
    out['sirsi_var'] = Y.var(axis=0)
    save('jump_ensemble_20k.npz', **out)


def gen_schedule():
    """Contact-matrix schedule of the reference TDPF over a wide date range and several tau."""
    flu, cf, df, start = make_influenza()
    out = {}
    for label, (s, n, tau) in {'season': ('2017-12-01', 322, 0.5), 'autumn': ('2017-10-01', 400, 1.0),
                               'spring': ('2018-02-01', 600, 0.25)}.items():
        table, idx = contact_schedule(cf, s, n, tau)
        out[label + '_index'] = idx
        out[label + '_meta'] = np.array([datetime.strptime(s, '%Y-%m-%d').toordinal(), n, tau])
    out['table'] = table
    save('contact_schedule.npz', **out)


# ------------------------------------------------------------------------------------------------
def gen_ode():
    out = {}
    # C1: tutorial SIR, N=100 draws of beta (tutorials/SIR/workflow_tutorial.py:49-70)
    class ODE_SIR(base.ODE):
        states = ['S', 'I', 'R']
        parameters = ['beta', 'gamma']

        @staticmethod
        def integrate(t, S, I, R, beta, gamma):
            N = S + I + R
            dS = -beta * S * I / N
            dI = beta * S * I / N - 1 / gamma * I
            dR = 1 / gamma * I
            return dS, dI, dR
    sir = ODE_SIR(initial_states={'S': 1000, 'I': 1}, parameters={'beta': 0.35, 'gamma': 5})
    rng = np.random.default_rng(1718)
    betas = rng.uniform(0.25, 0.45, size=100)
    ys, nfev = [], []
    for b in betas:
        sir.parameters['beta'] = b
        r = sir.sim(120)
        ys.append(r['y'])
        nfev.append(r['nfev'])
    out['sir_betas'], out['sir_y'], out['sir_nfev'] = betas, np.stack(ys), np.array(nfev)
    sir.parameters['beta'] = 0.35
    for rtol in (1e-6, 1e-9):
        r = sir.sim(120, rtol=rtol)
        out[f'sir_rtol{rtol:g}_y'], out[f'sir_rtol{rtol:g}_nfev'] = r['y'], np.array(r['nfev'])
    out['sir_euler_y'] = sir.sim(60, tau=0.25)['y']
    out['sir_t20_121_y'] = sir.sim([20, 121])['y']

    # SIR_SI ODE
    sirsi = make_sir_si(jump=False)
    r = sirsi.sim(120)
    out['sirsi_y0'], out['sirsi_y'], out['sirsi_nfev'] = flat_y0(sirsi), r['y'], np.array(r['nfev'])

    # PPBB (tutorials/enzyme_kinetics/models.py:53-69); values in the range of calibrate_intrinsic_kinetics.py
    enz = ref_loader.load_tutorial_module('tutorials/enzyme_kinetics/models.py', 'ref_enzyme_models')
    ppbb_par = {'c_enzyme': 10, 'Vf_Ks': 7.95e-04, 'R_AS': 0.047, 'R_AW': 1.10, 'R_Es': 0.131, 'K_eq': 0.66}
    ppbb = enz.PPBB_model({'S': 46.0, 'A': 61.0, 'W': 37.0, 'Es': 0.0}, ppbb_par)
    r = ppbb.sim(1500)
    out['ppbb_y0'], out['ppbb_y'], out['ppbb_nfev'] = flat_y0(ppbb), r['y'], np.array(r['nfev'])
    out['ppbb_par'] = np.array([ppbb_par[k] for k in ('c_enzyme', 'Vf_Ks', 'R_AS', 'R_AW', 'R_Es', 'K_eq')])

    # C4: packed PFR nx=25, a few posterior draws (simulate_1D_PFR.py:64-120)
    import math
    from scipy.io import netcdf_file
    nc = netcdf_file(os.path.join(ref_loader.REFERENCE_ROOT,
                                  'tutorials/enzyme_kinetics/data/username_SAMPLES_2025-02-05.nc'), 'r', mmap=False)
    names = ['Vf_Ks', 'R_AS', 'R_AW', 'R_Es', 'K_eq']
    samples = np.stack([np.array(nc.variables[k][:], dtype=np.float64).ravel() for k in names], axis=1)
    out['pfr_posterior'] = samples                                 # [900, 5] (iteration-major, chain)
    for nx in (25,):
        l, Q, dt = 1, (0.20 / 60) * 10 ** -6, 0.0024
        dp, mu, rho_B, rho_F = 0.0004755, 3.35e-03, 545, 775
        D_AB = np.array([3.47e-07, 2.33e-07, 1.95e-07, 1.39e-06])
        epsilon = 0.39 + 1.74 / (dt / dp + 1.140) ** 2
        A = (dt / 2) ** 2 * math.pi
        U = Q / A
        u = U / epsilon
        Re = dp * U * rho_F / mu / (1 - epsilon)
        a = 6 * (1 - epsilon) / dp
        kL = (0.7 * D_AB + (dp * U) / (0.18 + 0.008 * Re ** 0.59))
        D_ax = ((1.09 / 100) * (D_AB / dp) ** (2 / 3) * (U) ** (1 / 3))
        params = {'epsilon': epsilon, 'kL_a': kL * a, 'D_ax': D_ax, 'delta_x': l / nx, 'u': u, 'rho_B': rho_B,
                  'Vf_Ks': 7.88e-04, 'R_AS': 0.055, 'R_AW': 2.64, 'R_Es': 0.079, 'K_eq': 0.418}
        coords = {'species': ['S', 'A', 'Es', 'W'], 'x': np.linspace(start=0, stop=l, num=nx)}
        C_F = np.zeros([4, nx])
        C_S = np.zeros([4, nx])
        C_F[:, 0] = [30, 60, 0, 28]
        C_S[:, 0] = [30, 60, 0, 28]
        C_F[3, :] = 28
        C_S[3, :] = 28
        pfr = enz.packed_PFR({'C_F': C_F, 'C_S': C_S}, params, coords)
        out[f'pfr{nx}_y0'] = flat_y0(pfr)
        order = ('delta_x', 'epsilon', 'u', 'rho_B', 'Vf_Ks', 'R_AS', 'R_AW', 'K_eq', 'R_Es')
        out[f'pfr{nx}_par_scalar'] = np.array([params[k] for k in order], dtype=np.float64)
        out[f'pfr{nx}_kL_a'], out[f'pfr{nx}_D_ax'] = params['kL_a'], params['D_ax']
        pick = [0, 123, 456, 899]
        ys, nf = [], []
        t = time.time()
        for i in pick:
            for j, k in enumerate(names):
                pfr.parameters[k] = samples[i, j]
            r = pfr.sim(8000)
            ys.append(r['y'][:, ::100])                              # stride-100 subset of the 8001-point grid
            nf.append(r['nfev'])
        print(f'PFR nx={nx}: {len(pick)} draws in {time.time() - t:.1f}s, nfev={nf}')
        out[f'pfr{nx}_pick'], out[f'pfr{nx}_y'], out[f'pfr{nx}_nfev'] = np.array(pick), np.stack(ys), np.array(nf)
    save('ode.npz', **out)


# ------------------------------------------------------------------------------------------------
def gen_ode_rk23(method='RK23', name='ode_rk23.npz'):
    """`method='RK23'` / `'DOP853'` of the reference's ODE.sim (validation.py:158 -> solve_ivp, base.py:855): the tutorial SIR for
    30 draws of beta, SIR_SI and the enzyme-kinetics batch reactor -> tests/golden/ode_rk23.npz / ode_dop853.npz."""
    out = {}

    class ODE_SIR(base.ODE):
        states = ['S', 'I', 'R']
        parameters = ['beta', 'gamma']

        @staticmethod
        def integrate(t, S, I, R, beta, gamma):
            N = S + I + R
            dS = -beta * S * I / N
            dI = beta * S * I / N - 1 / gamma * I
            dR = 1 / gamma * I
            return dS, dI, dR
    sir = ODE_SIR(initial_states={'S': 1000, 'I': 1}, parameters={'beta': 0.35, 'gamma': 5})
    betas = np.random.default_rng(23).uniform(0.05, 3.0, size=30)          # from no epidemic to a very fast one
    ys, nfev = [], []
    for b in betas:
        sir.parameters['beta'] = b
        r = sir.sim(120, method=method)
        ys.append(r['y'])
        nfev.append(r['nfev'])
    out['sir_betas'], out['sir_y'], out['sir_nfev'] = betas, np.stack(ys), np.array(nfev)
    sir.parameters['beta'] = 0.35
    r = sir.sim(120, method=method, rtol=1e-7)
    out['sir_rtol1e-07_y'], out['sir_rtol1e-07_nfev'] = r['y'], np.array(r['nfev'])
    sirsi = make_sir_si(jump=False)
    r = sirsi.sim(120, method=method)
    out['sirsi_y0'], out['sirsi_y'], out['sirsi_nfev'] = flat_y0(sirsi), r['y'], np.array(r['nfev'])
    enz = ref_loader.load_tutorial_module('tutorials/enzyme_kinetics/models.py', 'ref_enzyme_models')
    ppbb_par = {'c_enzyme': 10, 'Vf_Ks': 7.95e-04, 'R_AS': 0.047, 'R_AW': 1.10, 'R_Es': 0.131, 'K_eq': 0.66}
    ppbb = enz.PPBB_model({'S': 46.0, 'A': 61.0, 'W': 37.0, 'Es': 0.0}, ppbb_par)
    r = ppbb.sim(1500, method=method)
    out['ppbb_y0'], out['ppbb_par'] = flat_y0(ppbb), np.array([ppbb_par[k] for k in ('c_enzyme', 'Vf_Ks', 'R_AS', 'R_AW', 'R_Es', 'K_eq')])
    out['ppbb_y'], out['ppbb_nfev'] = r['y'], np.array(r['nfev'])
    save(name, **out)


def gen_ode_dop853():
    gen_ode_rk23('DOP853', 'ode_dop853.npz')


def gen_ode_radau():
    """`method='Radau'` (scipy's Radau IIA, finite-difference Jacobian) of the unmodified reference -> tests/golden/ode_radau.npz"""
    gen_ode_rk23('Radau', 'ode_radau.npz')


def gen_ode_bdf():
    """`method='BDF'` (scipy's variable-order NDF, finite-difference Jacobian) of the unmodified reference -> tests/golden/ode_bdf.npz"""
    gen_ode_rk23('BDF', 'ode_bdf.npz')


def gen_ode_stiff():
    """Stiff problems through the reference's ODE.sim with `method='Radau'` / `'BDF'` (base.py:855 -> solve_ivp, numerical
    Jacobian): Robertson's kinetics (draws of k1) and the Van der Pol oscillator (draws of mu from mildly to very stiff, through
    its fast transitions: rejected steps, failed Newton iterations, Jacobian refreshes, BDF orders up to 5)
    -> tests/golden/ode_stiff.npz.  The same functors are compiled into the CPU harness (tests/hostcheck/hostcheck.cu) and
    translated from Python for the device (tests/gen_models.py)."""
    class Robertson(base.ODE):
        states = ['y1', 'y2', 'y3']
        parameters = ['k1', 'k2', 'k3']

        @staticmethod
        def integrate(t, y1, y2, y3, k1, k2, k3):
            r1 = k1 * y1
            r2 = k2 * y2 * y3
            r3 = k3 * y2 * y2
            return -r1 + r2, r1 - r2 - r3, r3

    class VanDerPol(base.ODE):
        states = ['x', 'v']
        parameters = ['mu']

        @staticmethod
        def integrate(t, x, v, mu):
            return v, mu * (1 - x * x) * v - x
    out = {}
    rng = np.random.default_rng(77)
    out['rob_k1'] = rng.uniform(0.02, 0.08, size=6)
    out['vdp_mu'] = np.array([0.5, 5.0, 20.0, 60.0, 100.0, 200.0])
    rob = Robertson(initial_states={'y1': 1.0, 'y2': 0.0, 'y3': 0.0}, parameters={'k1': 0.04, 'k2': 1e4, 'k3': 3e7})
    vdp = VanDerPol(initial_states={'x': 2.0, 'v': 0.0}, parameters={'mu': 100.0})
    for method in ('Radau', 'BDF'):
        ys, nf = [], []
        for k1 in out['rob_k1']:
            rob.parameters['k1'] = k1
            r = rob.sim(40, method=method)
            ys.append(r['y']); nf.append(r['nfev'])
        out[f'rob_{method}_y'], out[f'rob_{method}_nfev'] = np.stack(ys), np.array(nf)
        ys, nf = [], []
        for mu in out['vdp_mu']:
            vdp.parameters['mu'] = mu
            r = vdp.sim(200, method=method)
            ys.append(r['y']); nf.append(r['nfev'])
        out[f'vdp_{method}_y'], out[f'vdp_{method}_nfev'] = np.stack(ys), np.array(nf)
        rob.parameters['k1'] = 0.04
        r = rob.sim(400, method=method, rtol=1e-6)
        out[f'rob_{method}_long_y'], out[f'rob_{method}_long_nfev'] = r['y'], np.array(r['nfev'])
        print(method, 'Robertson nfev', out[f'rob_{method}_nfev'], 'VdP nfev', out[f'vdp_{method}_nfev'], 'long', r['nfev'])
    save('ode_stiff.npz', **out)


def gen_pfr100():
    """C4 at nx = 100 (docs/enzyme_kinetics.md:408-412): the reference's packed_PFR through solve_ivp RK45 for three
    posterior draws, stride-100 subset of the 8001-point grid + nfev (nx = 100 had no reference parity)."""
    import math
    from scipy.io import netcdf_file
    enz = ref_loader.load_tutorial_module('tutorials/enzyme_kinetics/models.py', 'ref_enzyme_models')
    nc = netcdf_file(os.path.join(ref_loader.REFERENCE_ROOT,
                                  'tutorials/enzyme_kinetics/data/username_SAMPLES_2025-02-05.nc'), 'r', mmap=False)
    names = ['Vf_Ks', 'R_AS', 'R_AW', 'R_Es', 'K_eq']
    samples = np.stack([np.array(nc.variables[k][:], dtype=np.float64).ravel() for k in names], axis=1)
    nx = 100
    l, Q, dt = 1, (0.20 / 60) * 10 ** -6, 0.0024
    dp, mu, rho_B, rho_F = 0.0004755, 3.35e-03, 545, 775
    D_AB = np.array([3.47e-07, 2.33e-07, 1.95e-07, 1.39e-06])
    epsilon = 0.39 + 1.74 / (dt / dp + 1.140) ** 2
    A = (dt / 2) ** 2 * math.pi
    U = Q / A
    u = U / epsilon
    Re = dp * U * rho_F / mu / (1 - epsilon)
    a = 6 * (1 - epsilon) / dp
    kL = (0.7 * D_AB + (dp * U) / (0.18 + 0.008 * Re ** 0.59))
    D_ax = ((1.09 / 100) * (D_AB / dp) ** (2 / 3) * (U) ** (1 / 3))
    params = {'epsilon': epsilon, 'kL_a': kL * a, 'D_ax': D_ax, 'delta_x': l / nx, 'u': u, 'rho_B': rho_B,
              'Vf_Ks': 7.88e-04, 'R_AS': 0.055, 'R_AW': 2.64, 'R_Es': 0.079, 'K_eq': 0.418}
    coords = {'species': ['S', 'A', 'Es', 'W'], 'x': np.linspace(start=0, stop=l, num=nx)}
    C_F = np.zeros([4, nx])
    C_S = np.zeros([4, nx])
    C_F[:, 0] = [30, 60, 0, 28]
    C_S[:, 0] = [30, 60, 0, 28]
    C_F[3, :] = 28
    C_S[3, :] = 28
    pfr = enz.packed_PFR({'C_F': C_F, 'C_S': C_S}, params, coords)
    out = {'pfr100_y0': flat_y0(pfr)}
    pick = [0, 456, 899]
    ys, nf = [], []
    t = time.time()
    for i in pick:
        for j, k in enumerate(names):
            pfr.parameters[k] = samples[i, j]
        r = pfr.sim(8000)
        ys.append(r['y'][:, ::100].astype(np.float64))
        nf.append(r['nfev'])
        print(f'PFR nx=100 draw {i}: nfev={r["nfev"]} ({time.time() - t:.0f}s)', flush=True)
    out['pfr100_pick'], out['pfr100_y'], out['pfr100_nfev'] = np.array(pick), np.stack(ys), np.array(nf)
    # How reproducible is the reference itself?  The explicit RK45 runs this 800-equation system at its stability limit
    # (19 778 evaluations), where the step-size controller amplifies round-off: changing rtol by one part in 10^14 -- less than
    # the difference between two BLAS summation orders in scipy's error norm -- moves the reference's own output by ~1e-4
    # after t ~ 2000 s (nfev unchanged).  Stored so that the parity tests can state their tolerance against this band.
    r = pfr.sim(8000, rtol=1e-4 * (1 + 1e-14))
    y_pert = r['y'][:, ::100]
    rel = np.abs(y_pert - ys[-1]) / (np.abs(ys[-1]) + 1e-6)
    out['pfr100_ref_self_max_rel'] = np.array(rel.max())
    out['pfr100_ref_self_max_rel_t1000'] = np.array(rel[:, :11].max())
    out['pfr100_ref_self_nfev'] = np.array(r['nfev'])
    print('reference vs reference with rtol * (1 + 1e-14): max rel', rel.max(), 'up to t = 1000:', rel[:, :11].max(), 'nfev', r['nfev'])
    out['pfr100_posterior'] = samples[pick]
    save('ode_pfr100.npz', **out)


# ------------------------------------------------------------------------------------------------
def gen_loglik():
    rng = np.random.default_rng(7)
    out = {}
    ymodel = rng.gamma(2.0, 200.0, size=(13, 4))
    ydata = rng.poisson(ymodel).astype(np.float64) / 7
    out['ymodel'], out['ydata'] = ymodel, ydata
    out['poisson'] = np.array(objf.ll_poisson(ymodel, ydata))
    out['negbin_vec'] = np.array(objf.ll_negative_binomial(ymodel, ydata, 4 * [0.03]))
    out['negbin_scalar'] = np.array(objf.ll_negative_binomial(ymodel, ydata, 0.07))
    sig_col = [3.0, 10.0, 25.0, 7.5]
    sig_full = rng.uniform(1, 30, size=(13, 4))
    out['sigma_col'], out['sigma_full'] = np.array(sig_col), sig_full
    out['normal_scalar'] = np.array(objf.ll_normal(ymodel, ydata, 12.0))
    out['normal_col'] = np.array(objf.ll_normal(ymodel, ydata, sig_col))
    out['normal_full'] = np.array(objf.ll_normal(ymodel, ydata, sig_full))
    out['lognormal_scalar'] = np.array(objf.ll_lognormal(ymodel, ydata, 0.4))
    out['lognormal_col'] = np.array(objf.ll_lognormal(ymodel, ydata, [0.2, 0.3, 0.4, 0.5]))
    # one-column data (the `not additional_axes_data` branch, objective_functions.py:256-269)
    ym1 = rng.gamma(3.0, 50.0, size=(20, 1))
    yd1 = rng.poisson(ym1).astype(np.float64)
    out['ymodel1'], out['ydata1'] = ym1, yd1
    out['poisson1'] = np.array(objf.ll_poisson(ym1, yd1))
    out['negbin1'] = np.array(objf.ll_negative_binomial(ym1, yd1, 0.05))
    out['normal1'] = np.array(objf.ll_normal(ym1, yd1, 4.0))
    # uniform prior (objective_functions.py:343-367)
    out['prior_inside'] = np.array(objf.log_prior_uniform(0.5, bounds=(0, 1)))
    out['prior_outside'] = np.array(objf.log_prior_uniform(1.5, bounds=(0, 1)))
    save('loglik.npz', **out)


def optimizer_objective(x):
    """Deterministic stand-in for a log posterior (to be MAXIMISED); shared with tests/test_optimization.py."""
    x = np.asarray(x)
    return -(np.sum((x - np.array([0.3, -1.2, 2.0])) ** 2) + 0.5 * np.sin(3 * x[0]) * x[1])


OPTIMIZER_BOUNDS = [(-2, 2), (-3, 3), (0, 5)]


def gen_optimizers():
    """Seeded runs of the reference's pso.optimize / nelder_mead.optimize (pso.py:28, nelder_mead.py:19)."""
    import contextlib
    import io
    import pySODM.optimization.nelder_mead as rnm
    import pySODM.optimization.pso as rpso
    out = {}
    with contextlib.redirect_stdout(io.StringIO()):
        for seed, S, it in [(1, 20, 15), (2, 64, 40)]:
            np.random.seed(seed)
            th, sc = rpso.optimize(optimizer_objective, bounds=OPTIMIZER_BOUNDS, swarmsize=S, max_iter=it)
            out[f'pso_{seed}_theta'], out[f'pso_{seed}_score'] = np.asarray(th, dtype=float), np.array(sc)
        th, sc = rnm.optimize(optimizer_objective, np.array([1.0, 1.0, 1.0]), [0.1, 0.1, 0.1], bounds=OPTIMIZER_BOUNDS, max_iter=60)
        out['nm_theta'], out['nm_score'] = np.asarray(th, dtype=float), np.array(sc)
        th, sc = rnm.optimize(optimizer_objective, np.array([1.9, 2.5, 0.2]), [0.3, 0.3, 0.3], bounds=OPTIMIZER_BOUNDS, max_iter=25)
        out['nm2_theta'], out['nm2_score'] = np.asarray(th, dtype=float), np.array(sc)
    save('optimizers.npz', **out)


def tdpf_compliance(t, states, param):
    """Step change on NUMERIC time (the pattern of src/tests/test_ODE.py:541-545)."""
    if t < 30:
        return param
    else:
        return param / 1.5


def tdpf_lower_infectivity(t, states, param, start_measures):
    """Date-valued switch with a date PARAMETER (tutorials/SIR/workflow_tutorial.py:201-205)."""
    if pd.to_datetime(t) < start_measures:
        return param
    else:
        return param / 2


def tdpf_ramp(t, states, param, ramp_days):
    """Linear ramp between two dates + dependence on the state (reads I)."""
    t0 = pd.Timestamp('2020-02-10')
    if t < t0:
        return param
    elif t < t0 + pd.Timedelta(days=ramp_days):
        frac = (t - t0) / pd.Timedelta(days=1) / ramp_days
        return param * (1 - 0.6 * frac)
    return param * 0.4 * (1 + 1e-4 * states['I'])


def gen_tdpf():
    """ODE runs of the unmodified reference with time-dependent parameter functions (evaluated at every RK stage,
    base.py:769-779): numeric time, date time with a date-valued parameter, a ramp that also reads a state."""
    class ODE_SIR(base.ODE):
        states = ['S', 'I', 'R']
        parameters = ['beta', 'gamma']

        @staticmethod
        def integrate(t, S, I, R, beta, gamma):
            N = S + I + R
            dS = -beta * S * I / N
            dI = beta * S * I / N - 1 / gamma * I
            dR = 1 / gamma * I
            return dS, dI, dR
    out = {}
    m = ODE_SIR({'S': 1000, 'I': 1}, {'beta': 0.35, 'gamma': 5}, time_dependent_parameters={'beta': tdpf_compliance})
    r = m.sim(120)
    out['numeric_y'], out['numeric_nfev'] = r['y'], np.array(r['nfev'])
    r = m.sim(120, rtol=1e-8)
    out['numeric_rtol1e-8_y'], out['numeric_rtol1e-8_nfev'] = r['y'], np.array(r['nfev'])
    m = ODE_SIR({'S': 1000, 'I': 1}, {'beta': 0.35, 'gamma': 5, 'start_measures': pd.Timestamp('2020-02-05')},
                time_dependent_parameters={'beta': tdpf_lower_infectivity})
    ys, nf, starts = [], [], []
    for shift in (0, 3.5, 11.25, 40):
        m.parameters['start_measures'] = pd.Timestamp('2020-02-05') + pd.Timedelta(days=shift)
        r = m.sim(['2020-01-01', '2020-04-30'])
        ys.append(r['y']); nf.append(r['nfev']); starts.append(shift)
    out['date_y'], out['date_nfev'], out['date_shift_days'] = np.stack(ys), np.array(nf), np.array(starts)
    m = ODE_SIR({'S': 1000, 'I': 1}, {'beta': 0.35, 'gamma': 5, 'ramp_days': 14.0}, time_dependent_parameters={'beta': tdpf_ramp})
    r = m.sim(['2020-01-01', '2020-04-30'])
    out['ramp_y'], out['ramp_nfev'] = r['y'], np.array(r['nfev'])
    save('tdpf.npz', **out)


class _FakeVariable:
    def __init__(self, values):
        self.values = values


class _FakeDataset(dict):
    """The three members of xarray.Dataset that utils.py:26-32 touches: copy(), data_vars.items(), ds[name].values."""
    def copy(self):
        return _FakeDataset({k: _FakeVariable(v.values.copy()) for k, v in self.items()})

    @property
    def data_vars(self):
        return self


def gen_post():
    """Seeded outputs of the reference's add_*_noise (optimization/utils.py:11-95) on a small fake Dataset, plus the
    np.mean / np.quantile reductions the tutorials apply along 'draws'.  utils.py imports matplotlib (absent here) at
    module top for its plotting helpers; a stub module stands in -- harness glue, no arithmetic involved."""
    import types
    ref_loader.load()
    if 'matplotlib' not in sys.modules:
        try:
            import matplotlib.pyplot  # noqa: F401
        except ImportError:
            mpl, plt = types.ModuleType('matplotlib'), types.ModuleType('matplotlib.pyplot')
            mpl.pyplot = plt
            plt.Axes = plt.Figure = object                     # annotations only (utils.py:190)
            sys.modules['matplotlib'], sys.modules['matplotlib.pyplot'] = mpl, plt
    import pySODM.optimization.utils as ref_utils
    rng = np.random.default_rng(11)
    counts = rng.integers(0, 5000, size=(64, 7, 4)).astype(np.float64)        # (draws, time, age)
    conc = rng.gamma(2.0, 3.0, size=(64, 9))
    out = {'counts': counts, 'conc': conc}
    ds = _FakeDataset({'counts': _FakeVariable(counts), 'conc': _FakeVariable(conc)})
    np.random.seed(101)
    r = ref_utils.add_poisson_noise(ds)
    out['poisson_counts'], out['poisson_conc'] = r['counts'].values, r['conc'].values
    np.random.seed(102)
    r = ref_utils.add_negative_binomial_noise(ds, 0.03)
    out['negbin_counts'], out['negbin_conc'] = r['counts'].values, r['conc'].values
    np.random.seed(103)
    r = ref_utils.add_gaussian_noise(ds, 0.04, relative=True)
    out['gauss_rel_counts'], out['gauss_rel_conc'] = r['counts'].values, r['conc'].values
    np.random.seed(104)
    r = ref_utils.add_gaussian_noise(ds, 2.5, relative=False)
    out['gauss_abs_counts'], out['gauss_abs_conc'] = r['counts'].values, r['conc'].values
    save('post.npz', **out)


# ------------------------------------------------------------------------------------------------
# jump processes with state- / time-dependent TDPFs, SSA with TDPFs, a drawn parameter feeding a TDPF (base.py:376-380)
# ------------------------------------------------------------------------------------------------
sys.path.insert(0, os.path.dirname(HERE))
import tdpf_jump_models as tj  # noqa: E402


class TdpfSIR(base.JumpProcess):                 # module level: the reference's process pool pickles the model
    states = tj.STATES
    parameters = tj.PARAMETERS
    compute_rates = staticmethod(tj.compute_rates)
    apply_transitionings = staticmethod(tj.apply_transitionings)


def make_tdpf_jump(cfg):
    return TdpfSIR(initial_states=dict(cfg['initial_states']), parameters=dict(cfg['parameters']),
                   time_dependent_parameters={'beta': tj.feedback, 'gamma': tj.slower_recovery}), tj


def make_influenza_small():
    """The tutorial model on a population / 1000 (an SSA run is ~10^4 events) with its calendar contact function."""
    flu, cf, df, start = make_influenza()
    for k, v in flu.initial_states.items():
        flu.initial_states[k] = np.rint(np.asarray(v, dtype=np.float64) / 1000.0)
    flu.initial_states['Ip'] = np.array([2.0, 3.0, 6.0, 2.0])
    flu.initial_states['E'] = np.array([2.0, 4.0, 8.0, 2.0])
    return flu, cf, start


def gen_tdpf_jump(n=2000, n_ssa=600, processes=8):
    out = {}
    # ---- tau-leap, both TDPFs; seeded single runs (oracle: bit-exact) + ensemble (device: KS)
    m, _ = make_tdpf_jump(tj.TAU_LEAP)
    out['tl_y0'] = flat_y0(m)
    ys = []
    for sd in (31, 32):
        np.random.seed(sd)
        ys.append(m.sim(60, N=1, tau=0.5)['y'])
    out['tl_seeds'], out['tl_y'] = np.array([31, 32]), np.stack(ys)
    t = time.time()
    np.random.seed(3100)
    res = m.sim(60, N=n, tau=0.5, processes=processes)
    Y = np.stack([r['y'] for r in res])
    print(f'TDPF tau-leap ensemble n={n}: {time.time() - t:.1f}s')
    times = np.array([2, 5, 10, 15, 20, 25, 30, 40, 50, 60])
    out['tl_times'], out['tl_sorted'] = times, np.sort(Y[:, :, times], axis=0).astype(np.int32)
    out['tl_mean'] = Y.mean(axis=0)
    # ---- the same with k_fb drawn per trajectory (feeds the TDPF of beta)
    t = time.time()
    np.random.seed(3200)
    res = m.sim(60, N=n, tau=0.5, draw_function=tj.draw_k_fb, processes=processes)
    Y = np.stack([r['y'] for r in res])
    print(f'TDPF tau-leap ensemble with drawn k_fb n={n}: {time.time() - t:.1f}s')
    out['tld_sorted'] = np.sort(Y[:, :, times], axis=0).astype(np.int32)
    out['tld_mean'] = Y.mean(axis=0)
    # ---- SSA, both TDPFs evaluated at event times
    m, _ = make_tdpf_jump(tj.SSA)
    out['ssa_y0'] = flat_y0(m)
    ys = []
    for sd in (41, 42):
        np.random.seed(sd)
        random.seed(1000 + sd)
        ys.append(m.sim(40, N=1, method='SSA')['y'])
    out['ssa_seeds'], out['ssa_y'] = np.array([41, 42]), np.stack(ys)
    t = time.time()
    np.random.seed(4100)
    random.seed(4100)
    res = m.sim(40, N=n_ssa, method='SSA', processes=processes)
    Y = np.stack([r['y'] for r in res])
    print(f'TDPF SSA ensemble n={n_ssa}: {time.time() - t:.1f}s')
    times = np.array([2, 5, 8, 12, 16, 20, 25, 30, 40])
    out['ssa_times'], out['ssa_ens'] = times, Y[:, :, times].astype(np.float32)
    # ---- influenza / 1000, SSA, calendar contact function on dates (tabulated per day on the device side)
    flu, cf, start = make_influenza_small()
    out['flu_small_y0'] = flat_y0(flu)
    t = time.time()
    np.random.seed(5100)
    random.seed(5100)
    res = flu.sim([start, '2018-02-01'], N=n_ssa, method='SSA', processes=processes)
    Y = np.stack([r['y'] for r in res])
    print(f'influenza/1000 SSA ensemble n={n_ssa}: {time.time() - t:.1f}s')
    times = np.array([3, 7, 14, 21, 28, 35, 42, 49, 56, 62])
    out['flu_small_times'], out['flu_small_ens'] = times, Y[:, :, times].astype(np.float32)
    out['flu_small_n_out'] = np.array(Y.shape[-1])
    table, idx = contact_schedule(cf, start, Y.shape[-1] - 1, 1.0)
    out['flu_small_table'], out['flu_small_day_index'] = table, idx
    save('tdpf_jump.npz', **out)


if __name__ == '__main__':
    what = sys.argv[1:] or ['jump_single', 'ssa', 'jump_ensemble', 'schedule', 'ode', 'loglik', 'optimizers', 'post', 'tdpf']
    for w in what:
        globals()['gen_' + w]()
