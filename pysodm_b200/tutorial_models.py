"""The reference's tutorial / test models as `pysodm_b200` model classes, each bound to its device functor.

The Python `compute_rates` / `apply_transitionings` / `integrate` bodies are the models as a user writes them
(`/root/reference/tutorials/*/models.py`, `src/tests/test_*.py`); on the GPU they are executed by the hand-written
functor named in `device_model` (csrc/jump_models.cuh, csrc/ode_models.cuh).
"""
import numpy as np

from .base import ODE, JumpProcess


class JumpProcess_SIR(JumpProcess):
    """Stochastic SIR without dimensions; gamma is a rate (`src/tests/test_JumpProcess.py:11-27`)."""
    states = ['S', 'I', 'R']
    parameters = ['beta', 'gamma']
    device_model = "sir_jump"

    @staticmethod
    def compute_rates(t, S, I, R, beta, gamma):
        return {'S': [beta * (I / (S + I + R)), ], 'I': [np.array([gamma]), ]}

    @staticmethod
    def apply_transitionings(t, tau, transitionings, S, I, R, beta, gamma):
        S_new = S - transitionings['S'][0]
        I_new = I + transitionings['S'][0] - transitionings['I'][0]
        R_new = R + transitionings['I'][0]
        return S_new, I_new, R_new


class JumpProcess_SIR_SI(JumpProcess):
    """Age-stratified SIR for humans + SI for a disease vector (`tutorials/SIR_SI/models.py:4-40`)."""
    states = ['S', 'I', 'R', 'S_v', 'I_v']
    parameters = ['beta', 'gamma']
    stratified_parameters = ['alpha']
    dimensions = ['age_group']
    dimensions_per_state = [['age_group'], ['age_group'], ['age_group'], [], []]
    device_model = "sir_si_jump"

    @staticmethod
    def compute_rates(t, S, I, R, S_v, I_v, alpha, beta, gamma):
        N = S + I + R
        N_v = S_v + I_v
        return {'S': [alpha * (I_v / N_v), ],
                'I': [(1 / gamma) * np.ones(N.shape), ],
                'S_v': [np.sum(alpha * (I / N)) * np.ones(N_v.shape), (1 / beta) * np.ones(N_v.shape), ],
                'I_v': [(1 / beta) * np.ones(N_v.shape), ]}

    @staticmethod
    def apply_transitionings(t, tau, transitionings, S, I, R, S_v, I_v, alpha, beta, gamma):
        S_new = S - transitionings['S'][0]
        I_new = I + transitionings['S'][0] - transitionings['I'][0]
        R_new = R + transitionings['I'][0]
        S_v_new = S_v - transitionings['S_v'][0] + transitionings['I_v'][0]
        I_v_new = I_v + transitionings['S_v'][0] - transitionings['I_v'][0]
        return S_new, I_new, R_new, S_v_new, I_v_new


class JumpProcess_influenza_model(JumpProcess):
    """Stochastic SEIR with undetected carriers, 4 age groups (`tutorials/influenza_1718/models.py:32-69`)."""
    states = ['S', 'E', 'Ip', 'Iud', 'Id', 'R', 'Im_inc']
    parameters = ['alpha', 'beta', 'gamma', 'delta', 'N']
    stratified_parameters = ['f_ud']
    dimensions = ['age_group']
    device_model = "influenza_1718_jump"

    @staticmethod
    def compute_rates(t, S, E, Ip, Iud, Id, R, Im_inc, alpha, beta, gamma, delta, N, f_ud):
        T = S + E + Ip + Iud + Id + R
        return {'S': [beta * np.matmul(N, (Ip + Iud + 0.22 * Id) / T), ],
                'E': [1 / alpha * np.ones(T.shape), ],
                'Ip': [f_ud * (1 / gamma), (1 - f_ud) * (1 / gamma)],
                'Iud': [(1 / delta) * np.ones(T.shape), ],
                'Id': [(1 / delta) * np.ones(T.shape), ]}

    @staticmethod
    def apply_transitionings(t, tau, transitionings, S, E, Ip, Iud, Id, R, Im_inc, alpha, beta, gamma, delta, N, f_ud):
        S_new = S - transitionings['S'][0]
        E_new = E + transitionings['S'][0] - transitionings['E'][0]
        Ip_new = Ip + transitionings['E'][0] - transitionings['Ip'][0] - transitionings['Ip'][1]
        Iud_new = Iud + transitionings['Ip'][0] - transitionings['Iud'][0]
        Id_new = Id + transitionings['Ip'][1] - transitionings['Id'][0]
        R_new = R + transitionings['Iud'][0] + transitionings['Id'][0]
        Im_inc_new = transitionings['Ip'][1]
        return S_new, E_new, Ip_new, Iud_new, Id_new, R_new, Im_inc_new


class ODE_SIR(ODE):
    """Tutorial SIR, gamma is a duration (`tutorials/SIR/workflow_tutorial.py:49-67`)."""
    states = ['S', 'I', 'R']
    parameters = ['beta', 'gamma']
    device_model = "sir_ode"

    @staticmethod
    def integrate(t, S, I, R, beta, gamma):
        N = S + I + R
        dS = -beta * S * I / N
        dI = beta * S * I / N - 1 / gamma * I
        dR = 1 / gamma * I
        return dS, dI, dR


class ODE_SIR_rate(ODE):
    """Test-suite SIR, gamma is a rate (`src/tests/test_ODE.py:11-25`)."""
    states = ['S', 'I', 'R']
    parameters = ['beta', 'gamma']
    device_model = "sir_rate_ode"

    @staticmethod
    def integrate(t, S, I, R, beta, gamma):
        N = S + I + R
        dS = -beta * I * S / N
        dI = beta * I * S / N - gamma * I
        dR = gamma * I
        return dS, dI, dR


class ODE_SIR_SI(ODE):
    """`tutorials/SIR_SI/models.py:42-66`."""
    states = ['S', 'I', 'R', 'S_v', 'I_v']
    parameters = ['beta', 'gamma']
    stratified_parameters = ['alpha']
    dimensions = ['age_group']
    dimensions_per_state = [['age_group'], ['age_group'], ['age_group'], [], []]
    device_model = "sir_si_ode"

    @staticmethod
    def integrate(t, S, I, R, S_v, I_v, alpha, beta, gamma):
        N = S + I + R
        N_v = S_v + I_v
        dS = -alpha * (I_v / N_v) * S
        dI = alpha * (I_v / N_v) * S - 1 / gamma * I
        dR = 1 / gamma * I
        dS_v = -np.sum(alpha * (I / N) * S_v) + (1 / beta) * N_v - (1 / beta) * S_v
        dI_v = np.sum(alpha * (I / N) * S_v) - (1 / beta) * I_v
        return dS, dI, dR, dS_v, dI_v


class PPBB_model(ODE):
    """Enzymatic esterification S + A <-> Es + W in a batch reactor (`tutorials/enzyme_kinetics/models.py:53-69`)."""
    states = ['S', 'A', 'Es', 'W']
    parameters = ['c_enzyme', 'Vf_Ks', 'R_AS', 'R_AW', 'R_Es', 'K_eq']
    device_model = "ppbb_ode"

    @staticmethod
    def integrate(t, S, A, Es, W, c_enzyme, Vf_Ks, R_AS, R_AW, R_Es, K_eq):
        v = c_enzyme * (Vf_Ks * (S * A - (1 / K_eq) * Es * W) / (A + R_AS * S + R_AW * W + R_Es * Es))
        return -v, -v, v, v


class packed_PFR(ODE):
    """1-D packed-bed plug-flow reactor with axial dispersion (`tutorials/enzyme_kinetics/models.py:5-51`)."""
    states = ['C_F', 'C_S']
    parameters = ['delta_x', 'epsilon', 'u', 'rho_B', 'Vf_Ks', 'R_AS', 'R_AW', 'K_eq', 'R_Es']
    stratified_parameters = [['kL_a', 'D_ax'], []]
    dimensions = ['species', 'x']
    device_model = "packed_pfr_ode"

    @staticmethod
    def integrate(t, C_F, C_S, delta_x, epsilon, u, rho_B, Vf_Ks, R_AS, R_AW, K_eq, R_Es, kL_a, D_ax):
        dC_F = np.zeros(C_F.shape, dtype=np.float64)
        dC_S = np.zeros(C_S.shape, dtype=np.float64)
        sign = np.array([-1.0, -1.0, 1.0, 1.0])[:, None]
        v = (Vf_Ks * (C_S[0, 1:] * C_S[1, 1:] - (1 / K_eq) * C_S[2, 1:] * C_S[3, 1:])
             / (C_S[1, 1:] + R_AS * C_S[0, 1:] + R_AW * C_S[3, 1:] + R_Es * C_S[2, 1:])) / 60
        left, mid = C_F[:, :-1], C_F[:, 1:]
        right = np.concatenate([C_F[:, 2:], C_F[:, -2:-1]], axis=1)          # mirrored outlet node
        dC_F[:, 1:] = ((D_ax[:, None] / delta_x ** 2) * (left - 2 * mid + right) - (u / delta_x) * (mid - left)
                       + (kL_a[:, None] / epsilon) * (C_S[:, 1:] - mid))
        dC_S[:, 1:] = -(kL_a[:, None] / (1 - epsilon)) * (C_S[:, 1:] - mid) + rho_B * sign * v[None, :]
        return dC_F, dC_S
