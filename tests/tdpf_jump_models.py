"""Jump-process test models whose time-dependent parameter functions (TDPFs) read the STATES, the TIME and a parameter that is
drawn per trajectory -- what the reference's `func` supports at every step / event (`/root/reference/src/pySODM/models/base.py:376-380`)
and round 1 refused on the device.  Plain functions only (no base class), so that the very same functions define

  * the reference-side models in `tests/golden/make_golden.py` (subclasses of the UNMODIFIED reference's `JumpProcess`) and
  * the device-side models in `tests/gen_models.py` (subclasses of `pysodm_b200.base.JumpProcess`, translated by codegen.py).
"""
import numpy as np

STATES = ['S', 'I', 'R']
PARAMETERS = ['beta', 'gamma']


def compute_rates(t, S, I, R, beta, gamma):
    return {'S': [beta * (I / (S + I + R)), ], 'I': [gamma * np.ones(1), ]}


def apply_transitionings(t, tau, transitionings, S, I, R, beta, gamma):
    S_new = S - transitionings['S'][0]
    I_new = I + transitionings['S'][0] - transitionings['I'][0]
    R_new = R + transitionings['I'][0]
    return S_new, I_new, R_new


def feedback(t, states, param, k_fb):
    """Behavioural feedback: transmission falls with the number of infectious people -- reads the states."""
    return param / (1.0 + k_fb * states['I'])


def slower_recovery(t, states, param, t_switch):
    """Recovery rate halves after t_switch -- reads the time."""
    if t < t_switch:
        return param
    return 0.5 * param


TAU_LEAP = dict(initial_states={'S': 20000 - 20, 'I': 20}, parameters={'beta': 0.7, 'gamma': 0.2, 'k_fb': 0.002, 't_switch': 20})
SSA = dict(initial_states={'S': 1500 - 10, 'I': 10}, parameters={'beta': 0.8, 'gamma': 0.25, 'k_fb': 0.01, 't_switch': 12})


def draw_k_fb(parameters):
    """k_fb feeds the TDPF of beta and is drawn per trajectory."""
    parameters['k_fb'] = np.random.uniform(0.0005, 0.004)
    return parameters
