"""Inputs of the reference's tutorial configurations (parameters, initial conditions, the influenza social-contact
schedule), restated from the tutorial scripts so that benchmarks and tests run where /root/reference is absent.

Sources (all under /root/reference/tutorials):
  influenza_1718/calibration.py:82-117   contact matrices, parameters, initial condition derived from
                                         data/interim/ILI_weekly_ABS.csv (week of 2017-12-01, divided by 7)
  influenza_1718/models.py:72-144        holiday / weekend rules of the contact-matrix function
  SIR_SI/calibration.py:31-32            SIR-SI parameters and initial condition
  enzyme_kinetics/simulate_1D_PFR.py:64-120   packed-bed reactor design variables
"""
import math
from datetime import datetime, timedelta

import numpy as np

AGE_GROUPS = ["[0, 5)", "[5, 15)", "[15, 65)", "[65, 120)"]

# ---- influenza 2017-2018 ---------------------------------------------------------------------------------
N_NOHOLIDAY_WEEK = np.transpose(np.array([[10.47, 3.16, 10.65, 0.55], [1.65, 19.41, 11.79, 0.70],
                                          [0.93, 1.98, 14.09, 0.96], [0.18, 0.45, 3.67, 3.09]]))
N_NOHOLIDAY_WEEKEND = np.transpose(np.array([[1.81, 2.46, 8.92, 1.52], [1.28, 7.07, 14.76, 1.64],
                                             [0.78, 2.49, 15.26, 1.17], [0.51, 1.05, 4.45, 2.00]]))
N_HOLIDAY_WEEK = np.transpose(np.array([[4.54, 2.47, 10.86, 0.35], [1.29, 5.76, 9.64, 1.47],
                                        [0.96, 1.63, 13.08, 1.38], [0.12, 0.94, 5.23, 5.10]]))
# table order used throughout: 0 = no-holiday week, 1 = weekend (same matrix in and out of holidays), 2 = holiday week
CONTACT_TABLE = np.stack([N_NOHOLIDAY_WEEK, N_NOHOLIDAY_WEEKEND, N_HOLIDAY_WEEK])

INFLUENZA_START = "2017-12-01"
INFLUENZA_END = "2018-05-11"
INFLUENZA_END_CALIBRATION = "2018-03-01"
INFLUENZA_PARAMETERS = {"alpha": 1, "beta": 0.0174, "gamma": 1, "delta": 3,
                        "f_ud": np.array([0.01, 0.64, 0.905, 0.60]), "N": N_HOLIDAY_WEEK}
INFLUENZA_INITIAL_STATES = {
    "S": [620914, 1306826, 7317774, 2130556],
    "E": [16, 37, 4292, 81],
    "Ip": [16, 37, 4292, 81],
    "Iud": [0, 24, 3885, 49],
    "Id": [16, 13, 408, 32],
    "Im_inc": [16, 13, 408, 32],
}
INFLUENZA_STATES = ["S", "E", "Ip", "Iud", "Id", "R", "Im_inc"]

# (start, end] intervals during which school holidays apply (models.py:120-144; strict < on the left, <= on the right)
_HOLIDAYS = [(datetime(2017, 10, 30), datetime(2017, 11, 5)), (datetime(2017, 12, 22), datetime(2018, 1, 7)),
             (datetime(2018, 2, 12), datetime(2018, 2, 18)), (datetime(2018, 4, 2), datetime(2018, 4, 15))]


def contact_matrix_index(date: datetime) -> int:
    """Row of CONTACT_TABLE in force at `date` (the tutorial's `contact_function`, restated without pandas)."""
    if date.weekday() >= 5:
        return 1
    for lo, hi in _HOLIDAYS:
        if lo < date <= hi:
            return 2
    return 0


def influenza_contact_function(t, states, param):
    """pySODM-compatible TDPF `fun(t, states, param)` returning the 4x4 contact matrix at datetime `t`."""
    return CONTACT_TABLE[contact_matrix_index(t)]


def influenza_flat_y0():
    y = np.zeros(28)
    for k, name in enumerate(INFLUENZA_STATES):
        if name in INFLUENZA_INITIAL_STATES:
            y[4 * k:4 * k + 4] = INFLUENZA_INITIAL_STATES[name]
    return y


def influenza_flat_par():
    p = INFLUENZA_PARAMETERS
    return np.concatenate([[p["alpha"], p["beta"], p["gamma"], p["delta"]], p["f_ud"], np.ravel(p["N"])]).astype(float)


def influenza_schedule(start=INFLUENZA_START, n_steps=322, tau=0.5):
    """step k -> CONTACT_TABLE row, with t_k accumulated in FP64 like the reference's loop (base.py:428-434)."""
    start_date = datetime.strptime(start, "%Y-%m-%d")
    idx = np.empty(n_steps, dtype=np.int32)
    t = 0
    for k in range(n_steps):
        idx[k] = contact_matrix_index(start_date + timedelta(microseconds=float(t) * 86400e6))
        t = t + tau
    return idx


# ---- SIR-SI ----------------------------------------------------------------------------------------------------
SIRSI_PARAMETERS = {"alpha": np.array([0.05, 0.1, 0.2, 0.15]), "gamma": 5, "beta": 7}
SIRSI_INITIAL_STATES = {"S": [606938, 1328733, 7352492, 2204478], "S_v": 1e6, "I_v": 2}


def sirsi_flat_y0():
    y = np.zeros(14)
    y[0:4] = SIRSI_INITIAL_STATES["S"]
    y[12], y[13] = SIRSI_INITIAL_STATES["S_v"], SIRSI_INITIAL_STATES["I_v"]
    return y


def sirsi_flat_par():
    p = SIRSI_PARAMETERS
    return np.concatenate([p["alpha"], [p["beta"], p["gamma"]]]).astype(float)


# ---- packed-bed plug-flow reactor ------------------------------------------------------------------------------
def pfr_parameters(nx=25):
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
    return {"epsilon": epsilon, "kL_a": kL * a, "D_ax": D_ax, "delta_x": l / nx, "u": u, "rho_B": rho_B,
            "Vf_Ks": 7.88e-04, "R_AS": 0.055, "R_AW": 2.64, "R_Es": 0.079, "K_eq": 0.418}


PFR_PARAM_ORDER = ("delta_x", "epsilon", "u", "rho_B", "Vf_Ks", "R_AS", "R_AW", "K_eq", "R_Es")


def pfr_flat_par(nx=25):
    p = pfr_parameters(nx)
    return np.concatenate([[p[k] for k in PFR_PARAM_ORDER], p["kL_a"], p["D_ax"]]).astype(float)


def pfr_initial_states(nx=25):
    C_F = np.zeros([4, nx])
    C_S = np.zeros([4, nx])
    C_F[:, 0] = [30, 60, 0, 28]
    C_S[:, 0] = [30, 60, 0, 28]
    C_F[3, :] = 28
    C_S[3, :] = 28
    return {"C_F": C_F, "C_S": C_S}


def pfr_flat_y0(nx=25):
    s = pfr_initial_states(nx)
    return np.concatenate([s["C_F"].ravel(), s["C_S"].ravel()])
