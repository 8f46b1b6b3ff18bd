"""Oracle for the jump-process steppers (see oracle/__init__.py -- test infrastructure only).

Follows `/root/reference/src/pySODM/models/base.py`:
  * `tau_leap_draw`   -- `_tau_leap` (:309-348) + `_multinomial_rvs` (:283-307), including the
                         reference's stay-probability quirk: the probability vector is
                         [p_0 .. p_{m-1}, 1 - sum_{i<m-1} p_i] and is normalised by its own sum.
  * `ssa_draw`        -- `_SSA` (:206-281) including the first-match-by-value tie rule (F2) and the
                         R == 0 -> tau = 1 rule (Q6).
  * `solve_discrete`  -- `_solve_discrete` (:421-440): every step stored, `t` accumulated in FP64
                         (Q5), rows linearly interpolated onto `t_eval = arange(t0, t1 + 1)` (:449).

All steppers accept arbitrary leading batch axes.  With no batch axis, `tau_leap_draw` issues
exactly the reference's sequence of `np.random.binomial` calls, so seeding NumPy's global
generator reproduces a reference trajectory bit for bit.
"""
import random as _pyrandom

import numpy as np


# --------------------------------------------------------------------------------------------
# tau-leap
# --------------------------------------------------------------------------------------------

def transition_probabilities(rate_list, tau, stay_quirk=True):
    """p[..., m+1] for one state: m leave-probabilities followed by the 'stay' weight (base.py:338-344).

    stay_quirk=True  : stay = 1 - sum_{i < m-1} p_i   (what the reference computes, F1)
    stay_quirk=False : stay = 1 - sum_{i < m}   p_i   (the mathematically intended complement)
    """
    p = 1 - np.exp(-tau * np.stack(rate_list, axis=-1, dtype=np.float64))
    m = p.shape[-1]
    acc = np.zeros(p.shape[:-1], dtype=np.float64)
    for i in range(m - 1 if stay_quirk else m):
        acc += p[..., i]
    stay = 1 - acc
    return np.concatenate((p, stay[..., None]), axis=-1)


def multinomial_by_conditional_binomials(count, p, binomial=None):
    """Multinomial(count; p / sum(p)) drawn as the reference does (base.py:296-307): categories are
    peeled off from the LAST to the second, each Binomial(remaining, p_i / sum_{j<=i} p_j), the first
    category takes what is left.  0/0 conditional probabilities are 0 (base.py:299-301)."""
    binomial = np.random.binomial if binomial is None else binomial
    count = np.array(count, dtype=np.int64)
    out = np.zeros(p.shape, dtype=np.int64)
    csum = np.cumsum(p, axis=-1)
    with np.errstate(divide="ignore", invalid="ignore"):
        cond = p / csum
    cond[np.isnan(cond)] = 0.0
    for i in range(p.shape[-1] - 1, 0, -1):
        draw = binomial(count, cond[..., i])
        out[..., i] = draw
        count = count - draw
    out[..., 0] = count
    return out


def tau_leap_draw(model, y, rates, tau, stay_quirk=True, binomial=None):
    """Number of transitionings per state: dict state -> [T_0, .., T_{m-1}, stayers] (base.py:336-346)."""
    T = {}
    for name, rate_list in rates.items():
        p = transition_probabilities(rate_list, tau, stay_quirk)
        n = np.array(y[name], np.int64)                       # truncation as base.py:345 (Q2)
        if n.shape != p.shape[:-1]:                           # batch axes present on only one of the two
            shape = np.broadcast_shapes(n.shape, p.shape[:-1])
            n = np.broadcast_to(n, shape).copy()
            p = np.broadcast_to(p, shape + p.shape[-1:])
        draws = multinomial_by_conditional_binomials(n, p, binomial)
        T[name] = [draws[..., i] for i in range(draws.shape[-1])]
    return T


# --------------------------------------------------------------------------------------------
# SSA (single trajectory: event times differ per trajectory, so the oracle loops)
# --------------------------------------------------------------------------------------------

def ssa_draw(model, y, rates, uniform=None, choose=None, first_match=True):
    """One Gillespie event (base.py:231-281).  Returns (T, tau).

    `uniform()` supplies u1 (default np.random.random, base.py:239); `choose(weights)` supplies the
    weighted index (default: Python's random.choices as base.py:254).  first_match=True maps the
    chosen propensity VALUE to the first channel holding that value (the reference's
    `list.index(...)`, F2)."""
    uniform = np.random.random if uniform is None else uniform
    R = 0
    for name, rate_list in rates.items():
        for r in rate_list:
            R += np.sum(y[name] * r, axis=None)
    shapes = {name: [np.shape(np.asarray(y[name] * r)) for r in rate_list] for name, rate_list in rates.items()}
    if R != 0:
        tau = 1 / R * np.log(1 / uniform())
        flat = []
        for name, rate_list in rates.items():
            for r in rate_list:
                flat += list(np.ravel(y[name] * r / R))
        if choose is None:
            picked = _pyrandom.choices(range(len(flat)), flat)[0]
        else:
            picked = choose(flat)
        idx = flat.index(flat[picked]) if first_match else picked
    else:
        tau, idx = 1, -1
    T, pos = {}, 0
    for name, rate_list in rates.items():
        T[name] = []
        for shp in shapes[name]:
            n = int(np.prod(shp)) if len(shp) else 1
            one_hot = np.zeros(n)
            if pos <= idx < pos + n:
                one_hot[idx - pos] = 1
            pos += n
            T[name].append(one_hot.reshape(shp))
    return T, tau


# --------------------------------------------------------------------------------------------
# time loop + output grid
# --------------------------------------------------------------------------------------------

def make_params_at(params, tdpf=None):
    """Returns f(step_index, t) -> params dict, substituting tabulated time-dependent parameters.

    tdpf = {param_name: (table[n_values, *shape], step_index[n_steps])}: the value used at step k is
    table[step_index[k]] -- the reference calls the TDPF with the ORIGINAL parameter value at every
    step (base.py:376-380, Q8), so tabulating per step is exact for state-independent TDPFs."""
    if not tdpf:
        return lambda k, t: params

    def at(k, t):
        p = dict(params)
        for name, (table, index) in tdpf.items():
            p[name] = np.asarray(table)[int(index[k])]
        return p
    return at


def apply_tdpf_functions(t, states, pars, tdpf_functions, model_params=None):
    """The reference's update of time-dependent parameters in front of `compute_rates` (base.py:376-386): every function
    gets the time, the live states and the ORIGINAL value of its own parameter; its extra arguments come from the dict
    being updated (so a later function sees an earlier one's result); parameters that only the functions use are dropped.

    tdpf_functions: [(parameter name, fn(t, states, param, **extras), [extra names])] in declaration order."""
    params = dict(pars)
    for name, fn, extras in tdpf_functions:
        params[name] = fn(t, states, pars[name], **{k: params[k] for k in extras})
    if model_params is not None:
        params = {k: v for k, v in params.items() if k in model_params}
    return params


def solve_discrete(model, y0, params, t0, t1, tau=0.5, method="tau_leap", tdpf=None,
                   stay_quirk=True, first_match=True, binomial=None, tdpf_functions=None, model_params=None,
                   tdpf_interval=None):
    """Run one trajectory (or a batch, tau_leap only) and return (t_eval, y_eval[..., n_flat, n_t]).

    y0: flat [..., n_flat] float64.  Mirrors base.py:421-440 + :442-472.
    tdpf_functions: time-dependent parameter functions called at every step / event as `func` does (:376-380).
    tdpf_interval: SSA with a TABULATED function (tdpf = {name: (table, index)}): index[0] is the row in force at t0 (the
    first event is evaluated there), index[1 + floor((t - t0) / tdpf_interval)] the row in force at a later time t -- exact
    for functions that are piecewise constant on the open intervals in between."""
    t_eval = np.arange(start=t0, stop=t1 + 1, step=1)
    params_at = make_params_at(params, tdpf)
    if tdpf_interval is not None:
        by_step = params_at
        n_rows = min(len(ix) for _, ix in tdpf.values())
        params_at = lambda k, t: by_step(0 if t == t_eval[0] else min(1 + int(np.floor((t - t_eval[0]) / tdpf_interval)), n_rows - 1), t)   # noqa: E731
    y = np.asarray(y0, dtype=np.float64)
    batch = y.shape[:-1]
    if method == "SSA" and batch:
        raise ValueError("the SSA oracle runs one trajectory at a time")
    cols, t_list = [y], [t_eval[0]]
    t, k = t_eval[0], 0
    while t < t_eval[-1]:
        p = params_at(k, t)
        states = model.unflatten(y)
        if tdpf_functions:
            p = apply_tdpf_functions(t, states, p, tdpf_functions, model_params)
        rates = model.rates(t, states, p)
        if method == "tau_leap":
            T, step = tau_leap_draw(model, states, rates, tau, stay_quirk, binomial), tau
        elif method == "SSA":
            T, step = ssa_draw(model, states, rates, first_match=first_match)
        else:
            raise ValueError(method)
        y = model.flatten(model.apply(t, step, T, states, p))
        cols.append(y)
        t = t + step
        t_list.append(t)
        k += 1
    Y = np.stack(cols, axis=-1)                                   # [..., n_flat, n_steps+1]
    t_arr = np.asarray(t_list, dtype=np.float64)
    flat = Y.reshape(-1, Y.shape[-1])
    out = np.empty((flat.shape[0], len(t_eval)))
    if _grid_is_subsample(t_arr, t_eval):
        pick = np.searchsorted(t_arr, t_eval)
        out[:] = flat[:, pick]
    else:
        for r in range(flat.shape[0]):
            out[r] = np.interp(t_eval, t_arr, flat[r])
    return t_eval, out.reshape(Y.shape[:-1] + (len(t_eval),))


def _grid_is_subsample(t_arr, t_eval):
    """True when every output time is hit exactly by a step (then np.interp is a pure gather)."""
    pos = np.searchsorted(t_arr, t_eval)
    return bool(np.all(pos < len(t_arr)) and np.all(t_arr[np.minimum(pos, len(t_arr) - 1)] == t_eval))


def tau_leap_ensemble(model, y0, params, t0, t1, tau, n_traj, tdpf=None, stay_quirk=True, seed=0,
                      chunk=4096):
    """Reference-law ensemble of n_traj trajectories, batched; returns y[n_traj, n_flat, n_t] float64.

    Per-trajectory parameters: arrays in `params` whose leading axis has length n_traj are sliced
    per chunk; everything else is broadcast.  Uses a private RandomState (legacy NumPy binomial =
    the reference's sampler, base.py:303)."""
    rs = np.random.RandomState(seed)
    model_shapes = dict(model.params)
    outs = []
    for lo in range(0, n_traj, chunk):
        hi = min(n_traj, lo + chunk)
        p = {}
        for name, v in params.items():
            v = np.asarray(v, dtype=np.float64)
            shp = tuple(model_shapes.get(name, v.shape))
            if v.ndim == len(shp) + 1 and v.shape[0] == n_traj:
                p[name] = v[lo:hi]
            else:
                p[name] = v
        y = np.asarray(y0, dtype=np.float64)
        if y.ndim == 2 and y.shape[0] == n_traj:
            yb = y[lo:hi]
        else:
            yb = np.broadcast_to(y, (hi - lo, y.shape[-1])).copy()
        _, yy = solve_discrete(model, yb, p, t0, t1, tau, "tau_leap", tdpf, stay_quirk, binomial=rs.binomial)
        outs.append(yy)
    return np.concatenate(outs, axis=0)
