"""Model classes WITHOUT a hand-written device functor: pysodm_b200 translates their Python functions (codegen.py).
Shared by the CPU and GPU tests of the generic-model path."""
import numpy as np

from pysodm_b200 import tutorial_data as td
from pysodm_b200 import tutorial_models as tm
from pysodm_b200.base import ODE, JumpProcess


class GenInfluenza(tm.JumpProcess_influenza_model):
    device_model = None


class GenSirSi(tm.JumpProcess_SIR_SI):
    device_model = None


class GenSir(tm.JumpProcess_SIR):
    device_model = None


class GenOdeSirSi(tm.ODE_SIR_SI):
    device_model = None


class GenOdeSir(tm.ODE_SIR):
    device_model = None


class GenPfr(tm.packed_PFR):
    device_model = None


class SEIRD(JumpProcess):
    """Not among the built-in functors: three competing exits from I (recovery, death, hospitalisation), an
    age-stratified force of infection through a contact matrix parameter, a cumulative counter, scalar + vector states."""
    states = ['S', 'E', 'I', 'H', 'R', 'D', 'cases']
    parameters = ['beta', 'sigma', 'gamma', 'mu', 'eta', 'rho', 'M']
    stratified_parameters = ['h']
    dimensions = ['age']

    @staticmethod
    def compute_rates(t, S, E, I, H, R, D, cases, beta, sigma, gamma, mu, eta, rho, M, h):
        N = S + E + I + H + R
        foi = beta * (M @ (I / N))
        size = np.ones(S.shape)
        return {'S': [foi], 'E': [sigma * size], 'I': [gamma * (1 - h), mu * size, eta * h], 'H': [rho * size]}

    @staticmethod
    def apply_transitionings(t, tau, transitionings, S, E, I, H, R, D, cases, beta, sigma, gamma, mu, eta, rho, M, h):
        T = transitionings
        S_new = S - T['S'][0]
        E_new = E + T['S'][0] - T['E'][0]
        I_new = I + T['E'][0] - T['I'][0] - T['I'][1] - T['I'][2]
        H_new = H + T['I'][2] - T['H'][0]
        R_new = R + T['I'][0] + T['H'][0]
        D_new = D + T['I'][1]
        return S_new, E_new, I_new, H_new, R_new, D_new, cases + T['E'][0]


SEIRD_AGES = ['young', 'adult', 'old']
SEIRD_STATES = {'S': np.array([40000., 90000., 30000.]), 'E': np.array([20., 60., 10.]), 'I': np.array([15., 40., 8.])}
SEIRD_PARAMS = {'beta': 0.045, 'sigma': 0.4, 'gamma': 0.25, 'mu': 0.01, 'eta': 0.2, 'rho': 0.1,
                'M': np.array([[9.0, 3.0, 1.0], [3.0, 6.0, 1.5], [1.0, 1.5, 2.0]]), 'h': np.array([0.02, 0.08, 0.3])}


def make_seird():
    return SEIRD(SEIRD_STATES, SEIRD_PARAMS, coordinates={'age': SEIRD_AGES})


class LotkaVolterra(ODE):
    """Not among the built-in functors: exp / sqrt / power, a loop over coupled patches filled element by element."""
    states = ['prey', 'pred']
    parameters = ['a', 'b', 'c', 'd', 'mig']
    dimensions = ['patch']

    @staticmethod
    def integrate(t, prey, pred, a, b, c, d, mig):
        n = len(prey)
        dprey = np.zeros(n)
        dpred = np.zeros(n)
        for j in range(n):
            left, right = prey[(j - 1) % n], prey[(j + 1) % n]
            season = 1 + 0.1 * np.sin(0.0172 * t)
            dprey[j] = a * season * prey[j] - b * prey[j] * pred[j] + mig * (left - 2 * prey[j] + right)
            dpred[j] = -c * pred[j] + d * np.sqrt(prey[j] ** 2 + 1e-12) * pred[j] * np.exp(-1e-3 * pred[j])
        return dprey, dpred


class Robertson(ODE):
    """Robertson's stiff kinetics; tests/golden/make_golden.py::gen_ode_stiff runs the same function through the reference."""
    states = ['y1', 'y2', 'y3']
    parameters = ['k1', 'k2', 'k3']

    @staticmethod
    def integrate(t, y1, y2, y3, k1, k2, k3):
        r1 = k1 * y1
        r2 = k2 * y2 * y3
        r3 = k3 * y2 * y2
        return -r1 + r2, r1 - r2 - r3, r3


class VanDerPol(ODE):
    """Van der Pol oscillator, stiff for large mu (gen_ode_stiff)."""
    states = ['x', 'v']
    parameters = ['mu']

    @staticmethod
    def integrate(t, x, v, mu):
        return v, mu * (1 - x * x) * v - x


def make_robertson():
    return Robertson({'y1': 1.0, 'y2': 0.0, 'y3': 0.0}, {'k1': 0.04, 'k2': 1e4, 'k3': 3e7})


def make_vanderpol():
    return VanDerPol({'x': 2.0, 'v': 0.0}, {'mu': 100.0})


def make_lotka(n=5):
    rng = np.random.default_rng(0)
    return LotkaVolterra({'prey': 10 + rng.random(n) * 5, 'pred': 3 + rng.random(n)},
                         {'a': 1.1, 'b': 0.4, 'c': 0.4, 'd': 0.1, 'mig': 0.05}, coordinates={'patch': list(range(n))})


def make_influenza():
    return GenInfluenza(td.INFLUENZA_INITIAL_STATES, td.INFLUENZA_PARAMETERS, coordinates={"age_group": td.AGE_GROUPS},
                        time_dependent_parameters={"N": td.influenza_contact_function})


SIRSI_STATES = {"S": np.array([606938., 1328733., 7352492., 2204478.]), "S_v": 1e6, "I_v": 2}
SIRSI_PARAMS = {"alpha": np.array([0.05, 0.1, 0.2, 0.15]), "gamma": 5, "beta": 7}


def make_sirsi(cls=GenSirSi, scale=1.0):
    st = {k: np.floor(np.asarray(v, dtype=np.float64) * scale) for k, v in SIRSI_STATES.items()}
    if scale != 1.0:
        st["I_v"] = 2.0
    return cls(st, dict(SIRSI_PARAMS), coordinates={"age_group": td.AGE_GROUPS})


def make_pfr(nx, cls=GenPfr):
    return cls(td.pfr_initial_states(nx), td.pfr_parameters(nx),
               coordinates={"species": ["S", "A", "Es", "W"], "x": list(range(nx))})


# ---- time-dependent parameter functions of an ODE (same bodies as tests/golden/make_golden.py::gen_tdpf) -------------
import pandas as pd  # noqa: E402


def tdpf_compliance(t, states, param):
    if t < 30:
        return param
    else:
        return param / 1.5


def tdpf_lower_infectivity(t, states, param, start_measures):
    if pd.to_datetime(t) < start_measures:
        return param
    else:
        return param / 2


def tdpf_ramp(t, states, param, ramp_days):
    t0 = pd.Timestamp('2020-02-10')
    if t < t0:
        return param
    elif t < t0 + pd.Timedelta(days=ramp_days):
        frac = (t - t0) / pd.Timedelta(days=1) / ramp_days
        return param * (1 - 0.6 * frac)
    return param * 0.4 * (1 + 1e-4 * states['I'])


def make_tdpf_models():
    """-> (numeric-time step model, date-switch model with a date parameter, ramp model reading a state).  The tutorial
    SIR class HAS a hand-written functor; with a time-dependent parameter function it is translated instead."""
    y0 = {"S": 1000, "I": 1}
    a = tm.ODE_SIR(y0, {"beta": 0.35, "gamma": 5}, time_dependent_parameters={"beta": tdpf_compliance})
    b = tm.ODE_SIR(y0, {"beta": 0.35, "gamma": 5, "start_measures": pd.Timestamp("2020-02-05")},
                   time_dependent_parameters={"beta": tdpf_lower_infectivity})
    c = tm.ODE_SIR(y0, {"beta": 0.35, "gamma": 5, "ramp_days": 14.0}, time_dependent_parameters={"beta": tdpf_ramp})
    return a, b, c


class SeasonalSIR(JumpProcess):
    """Seasonal forcing written INTO compute_rates (uses `t`), plus a lockdown between two dates when run over dates."""
    states = ['S', 'I', 'R']
    parameters = ['beta', 'gamma', 'amp']

    @staticmethod
    def compute_rates(t, S, I, R, beta, gamma, amp):
        if isinstance(t, (int, float)) or not hasattr(t, 'abs') and not hasattr(t, 'year'):
            season = 1 + amp * np.cos(2 * np.pi * t / 50.0)
        else:
            season = 1.0
            if t >= pd.Timestamp('2021-02-01') and t < pd.Timestamp('2021-02-20'):
                season = 1 - amp
        return {'S': [beta * season * I / (S + I + R)], 'I': [gamma]}

    @staticmethod
    def apply_transitionings(t, tau, transitionings, S, I, R, beta, gamma, amp):
        return S - transitionings['S'][0], I + transitionings['S'][0] - transitionings['I'][0], R + transitionings['I'][0]


def make_seasonal():
    return SeasonalSIR({'S': 99000, 'I': 1000}, {'beta': 0.5, 'gamma': 0.2, 'amp': 0.6})


# ---- jump process whose time-dependent parameter functions read the states, the time and a drawn parameter ------------
import tdpf_jump_models as tj  # noqa: E402


class TdpfSIR(JumpProcess):
    states = tj.STATES
    parameters = tj.PARAMETERS
    compute_rates = staticmethod(tj.compute_rates)
    apply_transitionings = staticmethod(tj.apply_transitionings)


def make_tdpf_jump(cfg=None):
    cfg = cfg or tj.TAU_LEAP
    return TdpfSIR(dict(cfg['initial_states']), dict(cfg['parameters']),
                   time_dependent_parameters={'beta': tj.feedback, 'gamma': tj.slower_recovery})


# ---- doubly stratified SIR (reference: src/tests/test_calibration.py:571-588), for aggregation functions -----------------
class SIRdoublestratified(ODE):
    states = ['S', 'I', 'R']
    parameters = ['gamma']
    stratified_parameters = [['beta'], []]
    dimensions = ['age_groups', 'spatial_units']

    @staticmethod
    def integrate(t, S, I, R, gamma, beta):
        beta = np.expand_dims(beta, axis=1)
        N = S + I + R
        dS = -beta * S * I / N
        dI = beta * S * I / N - gamma * I
        dR = gamma * I
        return dS, dI, dR


def make_double_sir():
    init = {"S": [[500_000 - 1, 400_000 - 5, 300_000 - 9], [250_000 - 2, 350_000 - 4, 450_000 - 8]],
            "I": [[1, 5, 9], [2, 4, 8]]}
    return SIRdoublestratified(init, {"gamma": 0.2, "beta": np.array([0.35, 0.3])},
                               coordinates={'age_groups': ['0-20', '20-120'], 'spatial_units': [0, 1, 2]})
