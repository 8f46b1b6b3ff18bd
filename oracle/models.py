"""NumPy restatements of the tutorial / test model definitions (oracle side; see oracle/__init__.py).

Each model works on dicts of float64 arrays with arbitrary leading batch axes `...` (one entry
per trajectory), the model's own state shape being the trailing axes.  With no batch axis the
arithmetic is the same sequence of NumPy calls the reference model makes, so a seeded
single-trajectory run is bit-identical to the reference (checked in tests/test_oracle_golden.py).

Flat state order = state declaration order, C-order ravel within a state
(`/root/reference/src/pySODM/models/utils.py:45-55`, `models/base.py:464-467`).
"""
import numpy as np


def _matvec(M, x):
    """M @ x over trailing axes; batch axes broadcast."""
    if M.ndim == 2 and x.ndim == 1:
        return np.matmul(M, x)
    return np.einsum("...ij,...j->...i", M, x)


def _col(v):
    """A per-trajectory scalar (shape = batch axes) gets a trailing axis so that it multiplies state
    vectors of shape batch + (k,); plain scalars (no batch) are returned unchanged."""
    v = np.asarray(v, dtype=np.float64)
    return v[..., None] if v.ndim > 0 else v


class JumpModelSpec:
    name = None
    states = ()            # ((name, shape), ...)
    params = ()            # ((name, shape), ...) in the flat parameter order of the C ABI
    transitions = ()       # ((state_name, n_transitions), ...) in compute_rates' dict order

    @classmethod
    def n_flat(cls):
        return int(sum(int(np.prod(s)) for _, s in cls.states))

    @classmethod
    def unflatten(cls, y):
        """y[..., n_flat] -> dict of arrays [..., *shape]"""
        out, off = {}, 0
        for nm, shp in cls.states:
            n = int(np.prod(shp))
            out[nm] = y[..., off:off + n].reshape(y.shape[:-1] + tuple(shp))
            off += n
        return out

    @classmethod
    def flatten(cls, d):
        parts = []
        for nm, shp in cls.states:
            v = np.asarray(d[nm], dtype=np.float64)
            nb = v.ndim - len(shp)
            parts.append(v.reshape(v.shape[:nb] + (-1,)))
        lead = np.broadcast_shapes(*[p.shape[:-1] for p in parts])
        return np.concatenate([np.broadcast_to(p, lead + p.shape[-1:]) for p in parts], axis=-1)


class SIRJump(JumpModelSpec):
    """Dimensionless stochastic SIR; gamma is a RATE. `/root/reference/src/tests/test_JumpProcess.py:11-27`."""
    name = "sir_jump"
    states = (("S", ()), ("I", ()), ("R", ()))
    params = (("beta", ()), ("gamma", ()))
    transitions = (("S", 1), ("I", 1))

    @staticmethod
    def rates(t, y, p):
        S, I, R = y["S"], y["I"], y["R"]
        return {"S": [p["beta"] * (I / (S + I + R))],
                "I": [np.asarray(p["gamma"], dtype=np.float64) * np.ones(np.shape(I))]}

    @staticmethod
    def apply(t, tau, T, y, p):
        S, I, R = y["S"], y["I"], y["R"]
        return {"S": S - T["S"][0], "I": I + T["S"][0] - T["I"][0], "R": R + T["I"][0]}


class SIRSIJump(JumpModelSpec):
    """Vector-borne SIR (4 age groups) + SI (vector). `/root/reference/tutorials/SIR_SI/models.py:4-40`.

    Note: S_v's second transition (natural death, rate 1/beta) is sampled but never
    applied; being S_v's LAST transition it still sets the normaliser of the tau-leap law.
    """
    name = "sir_si_jump"
    states = (("S", (4,)), ("I", (4,)), ("R", (4,)), ("S_v", ()), ("I_v", ()))
    params = (("alpha", (4,)), ("beta", ()), ("gamma", ()))
    transitions = (("S", 1), ("I", 1), ("S_v", 2), ("I_v", 1))

    @staticmethod
    def rates(t, y, p):
        S, I, R, S_v, I_v = y["S"], y["I"], y["R"], y["S_v"], y["I_v"]
        alpha = p["alpha"]
        N = S + I + R
        N_v = S_v + I_v
        frac_v = I_v / N_v
        return {
            "S": [alpha * _col(frac_v)],
            "I": [_col(1 / np.asarray(p["gamma"], dtype=np.float64)) * np.ones(N.shape)],
            "S_v": [np.sum(alpha * (I / N), axis=-1) * np.ones(np.shape(N_v)),
                    (1 / np.asarray(p["beta"], dtype=np.float64)) * np.ones(np.shape(N_v))],
            "I_v": [(1 / np.asarray(p["beta"], dtype=np.float64)) * np.ones(np.shape(N_v))],
        }

    @staticmethod
    def apply(t, tau, T, y, p):
        S, I, R, S_v, I_v = y["S"], y["I"], y["R"], y["S_v"], y["I_v"]
        return {"S": S - T["S"][0],
                "I": I + T["S"][0] - T["I"][0],
                "R": R + T["I"][0],
                "S_v": S_v - T["S_v"][0] + T["I_v"][0],
                "I_v": I_v + T["S_v"][0] - T["I_v"][0]}


class InfluenzaJump(JumpModelSpec):
    """Age-structured stochastic SEIR with undetected carriers.
    `/root/reference/tutorials/influenza_1718/models.py:32-69`.  `Im_inc` is the number of Ip->Id
    transitions of the LAST tau-step.  The contact matrix `N` is time dependent
    (`models.py:72-144`); the oracle receives it already evaluated for the current step.
    """
    name = "influenza_1718_jump"
    states = tuple((nm, (4,)) for nm in ("S", "E", "Ip", "Iud", "Id", "R", "Im_inc"))
    params = (("alpha", ()), ("beta", ()), ("gamma", ()), ("delta", ()), ("f_ud", (4,)), ("N", (4, 4)))
    transitions = (("S", 1), ("E", 1), ("Ip", 2), ("Iud", 1), ("Id", 1))

    @staticmethod
    def rates(t, y, p):
        S, E, Ip, Iud, Id, R = (y[k] for k in ("S", "E", "Ip", "Iud", "Id", "R"))
        f_ud = p["f_ud"]
        T = S + E + Ip + Iud + Id + R
        ones = np.ones(T.shape)
        inv_gamma = _col(1 / np.asarray(p["gamma"], dtype=np.float64))
        return {
            "S": [_col(p["beta"]) * _matvec(p["N"], (Ip + Iud + 0.22 * Id) / T)],
            "E": [_col(1 / np.asarray(p["alpha"], dtype=np.float64)) * ones],
            "Ip": [f_ud * inv_gamma, (1 - f_ud) * inv_gamma],
            "Iud": [_col(1 / np.asarray(p["delta"], dtype=np.float64)) * ones],
            "Id": [_col(1 / np.asarray(p["delta"], dtype=np.float64)) * ones],
        }

    @staticmethod
    def apply(t, tau, T, y, p):
        S, E, Ip, Iud, Id, R = (y[k] for k in ("S", "E", "Ip", "Iud", "Id", "R"))
        return {"S": S - T["S"][0],
                "E": E + T["S"][0] - T["E"][0],
                "Ip": Ip + T["E"][0] - T["Ip"][0] - T["Ip"][1],
                "Iud": Iud + T["Ip"][0] - T["Iud"][0],
                "Id": Id + T["Ip"][1] - T["Id"][0],
                "R": R + T["Iud"][0] + T["Id"][0],
                "Im_inc": T["Ip"][1] + 0.0 * S}


JUMP_MODELS = {m.name: m for m in (SIRJump, SIRSIJump, InfluenzaJump)}


# --------------------------------------------------------------------------------------------
# ODE right-hand sides.  Signature rhs(t, y_flat[n], p) -> dy_flat[n]  (single trajectory; the
# ODE oracle loops over draws because scipy's step controller is per trajectory).
# --------------------------------------------------------------------------------------------

class ODEModelSpec:
    name = None
    states = ()
    params = ()

    @classmethod
    def n_flat(cls):
        return int(sum(int(np.prod(s)) for _, s in cls.states))


class SIRODE(ODEModelSpec):
    """Tutorial SIR; gamma is a DURATION (1/gamma*I). `/root/reference/tutorials/SIR/workflow_tutorial.py:49-67`."""
    name = "sir_ode"
    states = (("S", ()), ("I", ()), ("R", ()))
    params = (("beta", ()), ("gamma", ()))

    @staticmethod
    def rhs(t, y, p):
        S, I, R = y
        beta, gamma = p["beta"], p["gamma"]
        N = S + I + R
        dS = -beta * S * I / N
        dI = beta * S * I / N - 1 / gamma * I
        dR = 1 / gamma * I
        return [dS, dI, dR]


class SIRRateODE(ODEModelSpec):
    """Test-suite SIR; gamma is a RATE. `/root/reference/src/tests/test_ODE.py:11-25`."""
    name = "sir_rate_ode"
    states = (("S", ()), ("I", ()), ("R", ()))
    params = (("beta", ()), ("gamma", ()))

    @staticmethod
    def rhs(t, y, p):
        S, I, R = y
        beta, gamma = p["beta"], p["gamma"]
        N = S + I + R
        dS = -beta * I * S / N
        dI = beta * I * S / N - gamma * I
        dR = gamma * I
        return [dS, dI, dR]


class SIRSIODE(ODEModelSpec):
    """`/root/reference/tutorials/SIR_SI/models.py:42-66`."""
    name = "sir_si_ode"
    states = (("S", (4,)), ("I", (4,)), ("R", (4,)), ("S_v", ()), ("I_v", ()))
    params = (("alpha", (4,)), ("beta", ()), ("gamma", ()))

    @staticmethod
    def rhs(t, y, p):
        y = np.asarray(y, dtype=np.float64)
        S, I, R, S_v, I_v = y[0:4], y[4:8], y[8:12], y[12], y[13]
        alpha, beta, gamma = p["alpha"], p["beta"], p["gamma"]
        N = S + I + R
        N_v = S_v + I_v
        dS = -alpha * (I_v / N_v) * S
        dI = alpha * (I_v / N_v) * S - 1 / gamma * I
        dR = 1 / gamma * I
        dS_v = -np.sum(alpha * (I / N) * S_v) + (1 / beta) * N_v - (1 / beta) * S_v
        dI_v = np.sum(alpha * (I / N) * S_v) - (1 / beta) * I_v
        return list(dS) + list(dI) + list(dR) + [dS_v, dI_v]


class PPBBODE(ODEModelSpec):
    """Batch enzyme kinetics S + A <-> Es + W. `/root/reference/tutorials/enzyme_kinetics/models.py:53-69`."""
    name = "ppbb_ode"
    states = (("S", ()), ("A", ()), ("Es", ()), ("W", ()))
    params = (("c_enzyme", ()), ("Vf_Ks", ()), ("R_AS", ()), ("R_AW", ()), ("R_Es", ()), ("K_eq", ()))

    @staticmethod
    def rhs(t, y, p):
        S, A, Es, W = y
        v = p["c_enzyme"] * (p["Vf_Ks"] * (S * A - (1 / p["K_eq"]) * Es * W)
                             / (A + p["R_AS"] * S + p["R_AW"] * W + p["R_Es"] * Es))
        return [-v, -v, v, v]


class PackedPFRODE(ODEModelSpec):
    """1-D packed-bed plug-flow reactor, method of lines, 4 species x nx nodes for the fluid (C_F)
    and solid (C_S) phases.  `/root/reference/tutorials/enzyme_kinetics/models.py:5-51`:
    node 0 is a Dirichlet inlet (derivative 0), upwind convection, central dispersion with a
    mirrored outlet node, first-order film transfer, reaction on the solid only.
    """
    name = "packed_pfr_ode"
    params = (("delta_x", ()), ("epsilon", ()), ("u", ()), ("rho_B", ()), ("Vf_Ks", ()), ("R_AS", ()),
              ("R_AW", ()), ("K_eq", ()), ("R_Es", ()), ("kL_a", (4,)), ("D_ax", (4,)))

    def __init__(self, nx):
        self.nx = nx
        self.states = (("C_F", (4, nx)), ("C_S", (4, nx)))

    def n_flat(self):
        return 8 * self.nx

    def rhs(self, t, y, p):
        nx = self.nx
        y = np.asarray(y, dtype=np.float64)
        C_F = y[:4 * nx].reshape(4, nx)
        C_S = y[4 * nx:].reshape(4, nx)
        dx, eps, u, rho_B = p["delta_x"], p["epsilon"], p["u"], p["rho_B"]
        kL_a = np.asarray(p["kL_a"], dtype=np.float64)[:, None]
        D_ax = np.asarray(p["D_ax"], dtype=np.float64)[:, None]
        sto = np.array([-1.0, -1.0, 1.0, 1.0])[:, None]
        dC_F = np.zeros_like(C_F)
        dC_S = np.zeros_like(C_S)
        v = (p["Vf_Ks"] * (C_S[0, 1:] * C_S[1, 1:] - (1 / p["K_eq"]) * C_S[2, 1:] * C_S[3, 1:])
             / (C_S[1, 1:] + p["R_AS"] * C_S[0, 1:] + p["R_AW"] * C_S[3, 1:] + p["R_Es"] * C_S[2, 1:])) / 60
        right = np.concatenate([C_F[:, 2:], C_F[:, -2:-1]], axis=1)   # outlet mirrors node X-2
        mid, left = C_F[:, 1:], C_F[:, :-1]
        dC_F[:, 1:] = ((D_ax / dx ** 2) * (left - 2 * mid + right)
                       - (u / dx) * (mid - left)
                       + (kL_a / eps) * (C_S[:, 1:] - mid))
        dC_S[:, 1:] = -(kL_a / (1 - eps)) * (C_S[:, 1:] - mid) + rho_B * sto * v[None, :]
        return np.concatenate([dC_F.ravel(), dC_S.ravel()])


ODE_MODELS = {m.name: m for m in (SIRODE, SIRRateODE, SIRSIODE, PPBBODE)}
