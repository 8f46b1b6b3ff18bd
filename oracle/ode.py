"""Oracle for the ODE runtime (see oracle/__init__.py -- test infrastructure only).

The reference integrates with a THIRD-PARTY routine that is not under /root/reference:
`scipy.integrate.solve_ivp(fun, [t0, t1], y0, args=[params], t_eval=arange(t0, t1 + 1), method, rtol)`
(`/root/reference/src/pySODM/models/base.py:855`; scipy is unpinned in `setup.py:18`, 1.18.1 in this
image; `atol` is not passed, so scipy's default 1e-6 applies).  The oracle makes that same call on the
NumPy restatements of the tutorial models in oracle/models.py; it is pinned by `tests/golden/ode.npz`,
generated from the unmodified reference classes (bit-exact for the scalar models).

`euler` restates `ODE._solve_discrete` (base.py:800-819).
"""
import numpy as np
from scipy.integrate import solve_ivp


def simulate(model, y0, params, t0, t1, method="RK45", rtol=1e-4, atol=1e-6):
    """Returns (t_eval, y[n_flat, n_t], nfev) for one parameter set."""
    t_eval = np.arange(start=t0, stop=t1 + 1, step=1)
    sol = solve_ivp(lambda t, y: model.rhs(t, y, params), [t0, t1], np.asarray(y0, dtype=np.float64),
                    t_eval=t_eval, method=method, rtol=rtol, atol=atol)
    return t_eval, sol.y, sol.nfev


def euler(model, y0, params, t0, t1, tau):
    t_eval = np.arange(start=t0, stop=t1 + 1, step=1)
    y = np.asarray(y0, dtype=np.float64).copy()
    cols, t_list = [y.copy()], [t_eval[0]]
    t = t_eval[0]
    while t < t_eval[-1]:
        step = np.asarray(model.rhs(t, y, params), dtype=np.float64) * tau
        y = y + step
        cols.append(y.copy())
        t = t + tau
        t_list.append(t)
    Y = np.stack(cols, axis=1)
    out = np.empty((Y.shape[0], len(t_eval)))
    for r in range(Y.shape[0]):
        out[r] = np.interp(t_eval, t_list, Y[r])
    return t_eval, out


def rk4(model, y0, params, t0, t1, tau):
    """Classical fixed-step Runge-Kutta 4 inside the reference's OWN discrete driver (`ODE._solve_discrete`, base.py:800-819:
    fixed step tau, states appended per step, `np.interp` onto the integer grid) -- the north star's "batched fixed-step
    RK4".  The reference ships Euler only, so there is nothing of its own to pin this against: the tableau is the textbook
    one (Kutta 1901: k1 = f(t, y), k2 = f(t + tau/2, y + tau/2 k1), k3 = f(t + tau/2, y + tau/2 k2), k4 = f(t + tau, y + tau k3),
    y += tau/6 (k1 + 2 k2 + 2 k3 + k4)), and the tests check it two ways: the device kernel against this restatement
    (same arithmetic, rtol 1e-13) and this restatement's 4th-order convergence to scipy's RK45 at rtol 1e-12."""
    t_eval = np.arange(start=t0, stop=t1 + 1, step=1)
    y = np.asarray(y0, dtype=np.float64).copy()
    f = lambda t, v: np.asarray(model.rhs(t, v, params), dtype=np.float64)      # noqa: E731
    cols, t_list = [y.copy()], [t_eval[0]]
    t = t_eval[0]
    while t < t_eval[-1]:
        k1 = f(t, y)
        k2 = f(t + 0.5 * tau, y + 0.5 * tau * k1)
        k3 = f(t + 0.5 * tau, y + 0.5 * tau * k2)
        k4 = f(t + tau, y + tau * k3)
        y = y + (tau / 6.0) * (k1 + 2.0 * k2 + 2.0 * k3 + k4)
        cols.append(y.copy())
        t = t + tau
        t_list.append(t)
    Y = np.stack(cols, axis=1)
    out = np.empty((Y.shape[0], len(t_eval)))
    for r in range(Y.shape[0]):
        out[r] = np.interp(t_eval, t_list, Y[r])
    return t_eval, out


def ensemble(model, y0, params, t0, t1, draws, **kw):
    """Loop over parameter draws: draws = {name: array[N, ...]}.  Returns y[N, n_flat, n_t], nfev[N]."""
    n = len(next(iter(draws.values())))
    ys, nf = [], []
    for i in range(n):
        p = dict(params)
        p.update({k: v[i] for k, v in draws.items()})
        _, y, nfev = simulate(model, y0, p, t0, t1, **kw)
        ys.append(y)
        nf.append(nfev)
    return np.stack(ys), np.asarray(nf)
