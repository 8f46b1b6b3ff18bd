"""`log_posterior_probability` for whole populations of parameter vectors at once.

Mirrors `/root/reference/src/pySODM/optimization/objective_functions.py`: the constructor signature (:92-108),
`__call__(thetas) -> float` (:290-335: sum of log priors, clipping to the bounds with a -inf penalty, one simulation,
sum of weighted log-likelihoods) and the `ll_*` / `log_prior_*` names (:343-666).  What differs is the unit of
work: `batch(thetas[P, D]) -> lp[P]` simulates all P parameter vectors in ONE `sodm_jump_run` / `sodm_ode_run`
launch (one trajectory per vector, as the reference's `sim(N=1)` per call) and reduces every trajectory against the
data with `sodm_loglik` on the device -- no trajectory leaves the GPU; only `lp[P]` comes back.  PSO / Nelder-Mead
loops (`pso.py:177-183`, `nelder_mead.py:131-138`) and emcee (`vectorize=True`) can call `batch` directly;
`__call__` is `batch` on a population of one.

The model/data alignment the reference does through xarray (:244-288) is compiled ONCE at construction into an
"alignment plan": per dataset, the flat state rows that are summed into each data column (dimensions absent from
the data are summed, :250-252), the data's time points in simulation time (linear interpolation, :254) and the
`(time, *data_axes)` ordered data matrix (:256-259, :271-278).
"""
import inspect
import itertools
from datetime import datetime

import numpy as np
import pandas as pd

from . import _ffi
from .validation import TIME_UNIT_US

# ---- log-likelihood "functions": on the device the choice is an enum; these callables keep the reference's names
#      (and work on host arrays, for users who call them directly as in the reference's tutorials) ----------------


def _device_kind(kind):
    def deco(fn):
        fn.device_kind = kind
        return fn
    return deco


@_device_kind("poisson")
def ll_poisson(ymodel, ydata):
    """Poisson log-likelihood on +1-shifted model and data (reference :613-633), host evaluation."""
    from scipy.special import gammaln
    ym = np.asarray(ymodel, dtype=np.float64) + 1
    yd = np.asarray(ydata, dtype=np.float64) + 1
    return -np.sum(ym) + np.sum(yd * np.log(ym)) - np.sum(gammaln(yd))


@_device_kind("negative_binomial")
def ll_negative_binomial(ymodel, ydata, alpha):
    """Negative-binomial log-likelihood, dispersion `alpha` scalar or one per data column (reference :636-666)."""
    from scipy.special import gammaln
    a = np.array(alpha)[np.newaxis, ...]
    ym = np.asarray(ymodel, dtype=np.float64) + 1
    yd = np.asarray(ydata, dtype=np.float64) + 1
    return (np.sum(yd * np.log(ym)) - np.sum((yd + 1 / a) * np.log(1 + a * ym)) + np.sum(yd * np.log(a))
            + np.sum(gammaln(yd + 1 / a)) - np.sum(gammaln(yd + 1)) - np.sum(yd.shape[0] * gammaln(1 / a)))


def _sigma_like(sigma, ymodel, name):
    s = np.array(sigma)
    if s.shape != np.shape(ymodel):
        s = s[np.newaxis, ...]
    if np.any(s <= 0):
        raise ValueError(f"the standard deviation used in `{name}` contains values smaller than or equal to zero")
    return s


@_device_kind("normal")
def ll_normal(ymodel, ydata, sigma):
    """Gaussian log-likelihood (reference :582-610)."""
    s = _sigma_like(sigma, ymodel, "ll_normal")
    return -1 / 2 * np.sum((np.asarray(ydata) - np.asarray(ymodel)) ** 2 / s ** 2 + np.log(2 * np.pi * s ** 2))


@_device_kind("lognormal")
def ll_lognormal(ymodel, ydata, sigma):
    """Lognormal log-likelihood (reference :550-579)."""
    s = _sigma_like(sigma, ymodel, "ll_lognormal")
    yd, ym = np.asarray(ydata, dtype=np.float64), np.asarray(ymodel, dtype=np.float64)
    return -1 / 2 * np.sum((np.log(yd + 1) - np.log(ym + 1)) ** 2 / s ** 2 + np.log(2 * np.pi * s ** 2) + np.log(yd + 1))


# ---- log priors: D <= ~10 numbers per particle -> vectorised on the host -----------------------

def log_prior_uniform(x, bounds=None, weight=1):
    x = np.asarray(x, dtype=np.float64)
    return np.where((bounds[0] <= x) & (x <= bounds[1]), 0.0 * weight, -np.inf)


def log_prior_triangle(x, low=None, high=None, mode=None, weight=1):
    from scipy.stats import triang
    return weight * triang.logpdf(x, loc=low, scale=high - low, c=(mode - low) / (high - low))


def log_prior_normal(x, avg=None, stdev=None, weight=1):
    from scipy.stats import norm
    return weight * norm.logpdf(x, loc=avg, scale=stdev)


def log_prior_lognormal(x, s=None, scale=None, weight=1):
    from scipy.stats import lognorm
    return weight * lognorm.logpdf(x, s=s, scale=scale)


def log_prior_gamma(x, a=None, loc=None, scale=None, weight=1):
    from scipy.stats import gamma
    return weight * gamma.logpdf(x, a=a, loc=loc, scale=scale)


def log_prior_beta(x, a=None, b=None, loc=None, scale=None, weight=1):
    from scipy.stats import beta
    return weight * beta.logpdf(x, a, b, loc=loc, scale=scale)


_AGG_FORMATS = ("Valid formats of aggregation functions are: 1) a list containing one function, 2) a list containing a number "
                "of functions equal to the number of datasets, 3) a callable function.")


def _validate_aggregation_function(aggregation_function, n_datasets):
    """The reference's `validate_aggregation_function` (:1181-1225): same accepted forms, same messages."""
    if isinstance(aggregation_function, list):
        if len(aggregation_function) not in (n_datasets, 1):
            raise ValueError("number of aggregation functions must be equal to one or the number of datasets")
        if len(aggregation_function) == 1:
            if not inspect.isfunction(aggregation_function[0]):
                raise ValueError(_AGG_FORMATS)
            return [aggregation_function[0]] * n_datasets
        if not all(inspect.isfunction(f) or f is None for f in aggregation_function):
            raise ValueError(_AGG_FORMATS)
        return list(aggregation_function)
    if inspect.isfunction(aggregation_function):
        return [aggregation_function] * n_datasets
    raise ValueError(_AGG_FORMATS)


def linear_map_of(aggregation_function, state, shape, dims, coords, time_index):
    """The matrix of a user aggregation function (reference :245-248: `out_copy = aggregation_function(out[states])`).

    On the device the model output never becomes an xarray object, so the function cannot be CALLED per particle; what the
    tutorials and the reference's tests do with it -- sum or weight a dimension, add up coordinate slices, rescale -- is a
    LINEAR map of the state's cells that acts on every time point alike, and such a map is fully described by its matrix.
    The function is therefore applied ONCE, at construction, to a DataArray whose "time" axis enumerates the unit basis of
    the state's cells (plus an all-zero slice and two random slices); the result gives the matrix column by column, the
    zero slice proves there is no offset, the random slices prove linearity and that time points are not mixed.  A function
    that fails those checks is refused -- there is no host fallback.
    -> (A[n_new_cells, n_cells], new_dims, new_coords, new_shape)"""
    from .dataset import make_dataset
    n_cells = int(np.prod(shape)) if shape else 1
    rng = np.random.default_rng(20260101)
    probe = np.zeros((n_cells + 3,) + tuple(shape))
    flat = probe.reshape(n_cells + 3, n_cells)
    flat[np.arange(n_cells), np.arange(n_cells)] = 1.0
    flat[n_cells + 1:] = rng.uniform(0.5, 2.0, size=(2, n_cells))
    when = (list(pd.date_range("2000-01-01", periods=n_cells + 3)) if time_index == "date" else np.arange(n_cells + 3))
    arr = make_dataset({state: probe}, {state: list(dims)}, {state: list(coords)}, time_index, when, 1)[state]
    try:
        out = aggregation_function(arr)
    except Exception as exc:  # noqa: BLE001
        raise NotImplementedError(f"the aggregation function of state '{state}' failed on the probe DataArray "
                                  f"(dims {tuple(arr.dims)}): {type(exc).__name__}: {exc}") from exc
    if not hasattr(out, "dims") or time_index not in tuple(out.dims):
        raise NotImplementedError(f"the aggregation function of state '{state}' must return a DataArray that keeps the "
                                  f"'{time_index}' dimension")
    new_dims = [d for d in out.dims if d != time_index]
    out = out.transpose(time_index, *new_dims)
    vals = np.asarray(out.values, dtype=np.float64)
    if vals.shape[0] != n_cells + 3:
        raise NotImplementedError(f"the aggregation function of state '{state}' changes the '{time_index}' axis")
    new_shape = vals.shape[1:]
    new_coords = [list(np.asarray(out.coords[d].values if hasattr(out.coords[d], "values") else out.coords[d])) for d in new_dims]
    v = vals.reshape(n_cells + 3, -1)
    A = v[:n_cells].T.copy()
    ok = np.allclose(v[n_cells], 0.0, atol=1e-12) and np.allclose(v[n_cells + 1:], flat[n_cells + 1:] @ A.T, rtol=1e-10, atol=1e-12)
    if not ok:
        raise NotImplementedError(f"the aggregation function of state '{state}' is not a linear map of the state at each time "
                                  "point (sums, weighted sums and selections are); it cannot be folded into the device likelihood")
    return A, new_dims, new_coords, new_shape


# ================================================================================================================
class log_posterior_probability:
    def __init__(self, model, parameter_names, bounds, data, states, log_likelihood_fnc, log_likelihood_fnc_args,
                 start_sim=None, weights=None, log_prior_prob_fnc=None, log_prior_prob_fnc_args=None,
                 initial_states=None, aggregation_function=None, labels=None, simulation_kwargs={}):
        n = len(data)
        if weights is None:
            weights = n * [1, ]
        if not isinstance(weights, (list, np.ndarray)) or np.ndim(weights) > 1:
            raise TypeError("Expected a list/1D np.array for input argument `weights`")
        lists = [states, log_likelihood_fnc, log_likelihood_fnc_args, weights] + ([initial_states] if initial_states else [])
        if any(len(lst) != n for lst in lists):
            raise ValueError(f"The number of datasets ({n}), model states ({len(states)}), log likelihood functions "
                             f"({len(log_likelihood_fnc)}), the extra arguments of the log likelihood function "
                             f"({len(log_likelihood_fnc_args)}), and weights ({len(weights)}) must be of equal length")
        # aggregation functions (reference :181-188, validate_aggregation_function :1181-1225): one per dataset or None
        self.aggregation_function = _validate_aggregation_function(aggregation_function, n) if aggregation_function else n * [None]
        self.model, self.states, self.weights = model, list(states), [float(w) for w in weights]
        self.parameters_names = list(parameter_names)
        self.log_likelihood_fnc = list(log_likelihood_fnc)
        self.simulation_kwargs = dict(simulation_kwargs)
        for k in self.simulation_kwargs:
            if k not in ("tau", "method", "rtol", "atol", "time_unit", "compat"):
                raise ValueError(f"'{k}' is not a valid simulation keyword for a batched device simulation")
        self.initial_states = initial_states

        # ---- datasets: time axis + additional axes ---------------------------------------------------------------
        self.data = [self._as_series(d, i) for i, d in enumerate(data)]
        kinds = {("date" if "date" in d.index.names else "time") for d in self.data}
        if len(kinds) != 1:
            raise ValueError("all datasets must use the same time index name: 'time' or 'date'")
        self.time_index = kinds.pop()
        self.additional_axes_data = [[nm for nm in d.index.names if nm != self.time_index] for d in self.data]

        # ---- simulation window (reference :769-848) ---------------------------------------------------------------
        firsts = [d.index.get_level_values(self.time_index).min() for d in self.data]
        lasts = [d.index.get_level_values(self.time_index).max() for d in self.data]
        if self.time_index == "time":
            data_start, self.end_sim = float(min(firsts)), float(max(lasts))
            if start_sim is None:
                start_sim = data_start
            if isinstance(start_sim, bool) or not isinstance(start_sim, (int, float)):
                raise TypeError("'start_sim' must be of type int/float")
            start_sim = float(start_sim)
        else:
            data_start, self.end_sim = min(firsts).to_pydatetime(), max(lasts).to_pydatetime()
            if start_sim is None:
                start_sim = data_start
            if isinstance(start_sim, str):
                start_sim = datetime.strptime(start_sim, "%Y-%m-%d")
            if not isinstance(start_sim, datetime):
                raise TypeError("'start_sim' must be of type datetime.datetime or a string representation of a date ('yyyy-mm-dd')")
        assert start_sim <= data_start, (f"start of the earliest dataset ({data_start}) before 'start_sim' ({start_sim}). make sure "
                                         "'start_sim' is chronologically before the start of the earliest datapoint.")
        self.start_sim = start_sim

        # ---- calibrated parameters, bounds, labels (reference :851-1021) ------------------------------------------
        self.parameter_shapes, sizes = {}, []
        for name in self.parameters_names:
            if name not in model.parameters:
                raise Exception(f"To be calibrated parameter '{name}' is not a valid model parameter!")
            shape = np.shape(model.parameters[name])
            self.parameter_shapes[name] = shape
            sizes.append(int(np.prod(shape)) if shape else 1)
        if len(bounds) != len(self.parameters_names):
            raise Exception(f"The number of provided bounds ({len(bounds)}) must match the number of calibrated "
                            f"parameters ({len(self.parameters_names)})")
        for lo_hi in bounds:
            if lo_hi[0] >= lo_hi[1]:
                raise ValueError(f"bounds {tuple(lo_hi)}: lower bound must be smaller than upper bound")
        labels = list(labels) if labels else list(self.parameters_names)
        self.expanded_bounds, self.expanded_labels, self.parameters_names_postprocessing = [], [], []
        for name, size, bnd, lab in zip(self.parameters_names, sizes, bounds, labels):
            for idx in (itertools.product(*[range(s) for s in self.parameter_shapes[name]]) if size > 1 or self.parameter_shapes[name] else [()]):
                suffix = "_" + "_".join(str(i) for i in idx) if idx != () else ""
                self.expanded_bounds.append(tuple(bnd))
                self.expanded_labels.append(lab + suffix)
                self.parameters_names_postprocessing.append(name + suffix)
        self._sizes = sizes
        self.n_theta = sum(sizes)

        # ---- priors -------------------------------------------------------------------------------------------------
        if log_prior_prob_fnc is None:
            self.log_prior_prob_fnc = [log_prior_uniform] * self.n_theta
            self.log_prior_prob_fnc_args = [{"bounds": b} for b in self.expanded_bounds]
        else:
            if len(log_prior_prob_fnc) != len(self.parameters_names) or len(log_prior_prob_fnc_args) != len(self.parameters_names):
                raise Exception("The number of prior functions / prior arguments must match the number of calibrated parameters")
            self.log_prior_prob_fnc, self.log_prior_prob_fnc_args = [], []
            for fn, kw, size in zip(log_prior_prob_fnc, log_prior_prob_fnc_args, sizes):
                self.log_prior_prob_fnc += [fn] * size
                self.log_prior_prob_fnc_args += [dict(kw)] * size

        # ---- likelihood functions and their arguments --------------------------------------------------------------
        self._kinds = []
        for fn in self.log_likelihood_fnc:
            kind = getattr(fn, "device_kind", None) or {"ll_poisson": "poisson", "ll_negative_binomial": "negative_binomial",
                                                         "ll_normal": "normal", "ll_lognormal": "lognormal"}.get(getattr(fn, "__name__", ""))
            if kind is None:
                raise NotImplementedError(f"log-likelihood function {fn} has no device implementation "
                                          "(available: ll_poisson, ll_negative_binomial, ll_normal, ll_lognormal)")
            self._kinds.append(kind)
        self.log_likelihood_fnc_args = []
        for kind, arg in zip(self._kinds, log_likelihood_fnc_args):
            if kind == "poisson":
                if arg not in (None, [], ()):
                    raise ValueError("the likelihood function `ll_poisson` takes no additional arguments")
                self.log_likelihood_fnc_args.append(None)
            elif isinstance(arg, pd.Series):                  # aligned with the data's own index in _build_plan
                self.log_likelihood_fnc_args.append(arg)
            else:
                self.log_likelihood_fnc_args.append(np.asarray(arg, dtype=np.float64))

        # ---- alignment plan ------------------------------------------------------------------------------------------
        self._plan = [self._build_plan(i) for i in range(n)]

    # -------------------------------------------------------------------------------------------------------------
    def _as_series(self, d, i):
        if isinstance(d, pd.DataFrame):
            if d.shape[1] != 1:
                raise ValueError(f"{i}th dataset: a DataFrame must have exactly one column")
            d = d.iloc[:, 0]
        if not isinstance(d, pd.Series):
            raise TypeError(f"{i}th dataset must be a pd.Series or single-column pd.DataFrame, found {type(d)}")
        if "time" not in d.index.names and "date" not in d.index.names:
            raise ValueError(f"Index of {i}th dataset does not have 'time' or 'date' as index level.")
        if d.isna().any():
            raise ValueError(f"{i}th dataset contains nans")
        return d

    def _state_rows(self, state):
        names = self.model.states_names
        if state not in names:
            raise ValueError(f"'{state}' is not a model state {names}")
        r0 = 0
        for nm in names:
            n = int(np.prod(self.model.state_shapes[nm])) if self.model.state_shapes[nm] else 1
            if nm == state:
                return r0, self.model.state_shapes[nm], self.model.dimensions_per_state[nm], self.model.state_coordinates[nm]
            r0 += n
        raise AssertionError

    def _build_plan(self, i):
        d, state, axes = self.data[i], self.states[i], self.additional_axes_data[i]
        r0, shape, dims, coords = self._state_rows(state)
        A = None
        if self.aggregation_function[i]:
            # the data is matched against the AGGREGATED state (reference :245-252, compare_data_model_coordinates :1316-1321)
            A, dims, coords, shape = linear_map_of(self.aggregation_function[i], state, shape, dims, coords, self.time_index)
        for ax in axes:
            if ax not in dims:
                raise Exception(f"{i}th dataset: index level '{ax}' is not a dimension of model state '{state}' {dims}")
        times = d.index.get_level_values(self.time_index).unique()
        if self.time_index == "date":
            unit = TIME_UNIT_US[self.simulation_kwargs.get("time_unit", "D")]
            t_data = np.array([(t.to_pydatetime() - self.start_sim).total_seconds() * 1e6 / unit for t in times])
        else:
            t_data = np.asarray(times, dtype=np.float64)
        axis_coords = [list(d.index.get_level_values(ax).unique()) for ax in axes]
        for ax, cs in zip(axes, axis_coords):
            model_cs = coords[dims.index(ax)]
            missing = [c for c in cs if c not in model_cs]
            if missing:
                raise Exception(f"{i}th dataset: coordinates {missing} of '{ax}' are not coordinates of the model")
        # data matrix in (time, *axes) order, columns = C-order product of the data's coordinates
        series = d.reorder_levels([self.time_index] + axes) if axes else d
        full = pd.MultiIndex.from_product([times] + axis_coords) if axes else times
        ydata = series.reindex(full).to_numpy(dtype=np.float64).reshape(len(times), -1)
        if np.isnan(ydata).any():
            raise ValueError(f"{i}th dataset does not contain a value for every combination of its index levels")
        # model rows summed into each data column: fix the data's axes, sum over the state's other dimensions
        columns = []
        for combo in (itertools.product(*axis_coords) if axes else [()]):
            fixed = {ax: coords[dims.index(ax)].index(c) for ax, c in zip(axes, combo)}
            rows = []
            for idx in itertools.product(*[range(s) for s in shape]) if shape else [()]:
                if all(idx[dims.index(ax)] == k for ax, k in fixed.items()):
                    rows.append(r0 + (int(np.ravel_multi_index(idx, shape)) if shape else 0))
            columns.append(rows)
        weights = None
        if A is not None:
            # `rows` above index cells of the aggregated state: column = sum over them of (row of A) . (cells of the model state)
            cols_w = [A[[r - r0 for r in rows]].sum(axis=0) for rows in columns]
            columns = [[r0 + int(c) for c in np.flatnonzero(w)] for w in cols_w]
            weights = [[float(w[c]) for c in np.flatnonzero(w)] for w in cols_w]
            if all(x == 1.0 for w in weights for x in w):
                weights = None
        # the extra argument (sigma / alpha): a scalar, one value per data column, or one value per datapoint -- an array of the
        # data's shape or a pd.Series on the data's own index (the reference's weighted-least-squares use of ll_normal,
        # objective_functions.py:1396-1454), brought into the (time, *axes) order of `ydata` like the data itself
        arg = self.log_likelihood_fnc_args[i]
        if isinstance(arg, pd.Series):
            if not arg.index.equals(d.index):
                raise ValueError(f"index of pd.Series containing the extra argument of the {i}th log-likelihood function must "
                                 f"match the index of the {i}th dataset")
            a = arg.reorder_levels([self.time_index] + axes) if axes else arg
            arg = a.reindex(full).to_numpy(dtype=np.float64).reshape(ydata.shape)
        elif arg is not None and arg.ndim == 1 and ydata.shape[1] == 1 and arg.shape[0] == ydata.shape[0] and arg.size != 1:
            arg = arg.reshape(ydata.shape)                       # one value per datapoint of a dataset without extra axes
        if arg is not None:
            if arg.ndim > 0 and arg.shape not in ((ydata.shape[1],), ydata.shape) and arg.size != 1:
                raise ValueError(f"the extra argument of the {i}th log-likelihood function has shape {arg.shape}, "
                                 f"which does not match the data {ydata.shape}")
            if not np.all(arg > 0):                              # sigma = 0 would score +inf on the device
                raise ValueError(f"the extra argument of the {i}th log-likelihood function ({self._kinds[i]}: standard deviation / "
                                 "overdispersion) must be strictly positive")
            self.log_likelihood_fnc_args[i] = arg
        return {"columns": columns, "weights": weights, "t_data": t_data, "ydata": ydata}

    # -------------------------------------------------------------------------------------------------------------
    def compute_log_prior_probability(self, thetas):
        thetas = np.atleast_2d(np.asarray(thetas, dtype=np.float64))
        lp = np.zeros(thetas.shape[0])
        for j, (fn, kw) in enumerate(zip(self.log_prior_prob_fnc, self.log_prior_prob_fnc_args)):
            lp = lp + np.asarray(fn(thetas[:, j], **kw), dtype=np.float64)
        return lp

    def batch(self, thetas, seed=None, device=None, return_device=False):
        """lp[P] for thetas[P, D] (D = number of expanded parameters).  Like the reference's `__call__`, values
        outside the bounds are clipped IN PLACE and penalised with -inf (:301-307)."""
        import torch
        from . import engine
        thetas = np.asarray(thetas)
        if thetas.ndim != 2 or thetas.shape[1] != self.n_theta:
            raise ValueError(f"thetas must have shape [P, {self.n_theta}]")
        lp = self.compute_log_prior_probability(thetas)
        for j, (lo, hi) in enumerate(self.expanded_bounds):
            above, below = thetas[:, j] > hi, thetas[:, j] < lo
            thetas[above, j], thetas[below, j] = hi, lo
            lp = lp + np.where(above | below, -np.inf, 0.0)
        dev = engine.device(device)
        lp_d = self._device_log_likelihood([np.ascontiguousarray(thetas[:, j], dtype=np.float64) for j in range(self.n_theta)],
                                           thetas.shape[0], seed, dev)
        if return_device:
            return lp_d + torch.from_numpy(lp).to(dev)
        return lp + lp_d.cpu().numpy()

    def batch_device(self, thetas, seed=None, member_id0=0):
        """lp[P] (float64 CUDA tensor) for a population that LIVES on the device: `thetas` is a float64 CUDA tensor
        [D, P] (structure-of-arrays, one row per expanded parameter -- the layout the simulators read per-trajectory
        parameters in, so nothing is transposed or copied).  For optimisers that keep their population on the device
        (`optimization.pso(..., device_update=True)`, `mcmc.EnsembleSampler(..., device_update=True)`).  The caller keeps
        the population inside the bounds (the device swarm update clamps, the sampler clips and penalises); uniform priors
        over those bounds contribute 0 and are skipped, any other prior is evaluated on the host for its coordinate.
        `member_id0` is the global index of the first column when the population is sharded over ranks (it keys the Philox streams of a stochastic model: `sharding.batch_device_sharded`)."""
        import torch
        if not (isinstance(thetas, torch.Tensor) and thetas.is_cuda and thetas.dtype == torch.float64 and
                thetas.dim() == 2 and thetas.shape[0] == self.n_theta):
            raise ValueError(f"thetas must be a float64 CUDA tensor [{self.n_theta}, P]")
        lp = self._device_log_likelihood([thetas[j] for j in range(self.n_theta)], thetas.shape[1], seed, thetas.device,
                                         member_id0=member_id0)
        # priors that are not "uniform over the parameter's own bounds" (0 inside) are host functions of one coordinate
        # (scipy.stats log-densities, user callables): their rows make one round trip, 8 bytes per member each way
        rows = [j for j, (fn, kw) in enumerate(zip(self.log_prior_prob_fnc, self.log_prior_prob_fnc_args))
                if not (fn is log_prior_uniform and tuple(kw.get("bounds", ())) == tuple(self.expanded_bounds[j]))]
        if rows and thetas.shape[1]:
            host = thetas[rows].cpu().numpy()
            prior = np.zeros(thetas.shape[1])
            for r, j in enumerate(rows):
                prior = prior + np.asarray(self.log_prior_prob_fnc[j](host[r], **self.log_prior_prob_fnc_args[j]), dtype=np.float64)
            lp += torch.from_numpy(prior).to(thetas.device)
        return lp

    def _device_log_likelihood(self, columns, P, seed, dev, member_id0=0):
        """Sum of the weighted log-likelihoods of P parameter vectors given as one column (host array or CUDA tensor of P
        values) per expanded parameter -> float64 CUDA tensor [P]."""
        import torch
        from . import engine
        model = self.model
        kw = self.simulation_kwargs
        time_unit = kw.get("time_unit", "D")
        if getattr(model, "_generated", False) and model.time_dependent_parameters and (
                model._kind == "ode" or model.device_model is None or getattr(model, "_traced_tdpfs", None)):
            # translated time-dependent parameter functions exist per time mode (numbers / dates), base.py
            dated = self.time_index == "date"
            model._bind_generated_functor(TIME_UNIT_US[time_unit] if dated else None, [self.start_sim if dated else None])
        # parameter vector -> flat device parameter elements
        par, _, y0, _ = model._collect_draws(1, None, {})
        if "__t0__" in model._param_offsets:
            from .codegen import date_to_units
            par[model._param_offsets["__t0__"][0]] = date_to_units(self.start_sim, model._time_unit_us)
        var, col = {}, 0
        for name, size in zip(self.parameters_names, self._sizes):
            if name not in model._param_offsets:
                raise NotImplementedError(f"calibrated parameter '{name}' is not a parameter of the device functor")
            if name in model._spec.tdpf_params:              # tabulated per step (jump processes): no per-particle value
                raise NotImplementedError(f"calibrated parameter '{name}' is time dependent")
            for k_t, t_name in enumerate(model.time_dependent_parameters or {}):
                # a TABULATED function is evaluated once for all particles (with model.parameters): a calibrated argument
                # would silently keep its base value inside it.  Translated functions read the particle's own.
                if model._kind == "jump" and t_name in model._spec.tdpf_params and name in model._function_parameters[k_t]:
                    raise NotImplementedError(f"calibrated parameter '{name}' is an argument of the time-dependent parameter "
                                              f"function of '{t_name}', which could only be tabulated "
                                              f"({getattr(model, '_tdpf_reason', {}).get(t_name, 'hand-written functor')})")
            off = model._param_offsets[name][0]
            for e in range(size):
                var[off + e] = columns[col + e]
            col += size
        if self.time_index == "date":
            unit = TIME_UNIT_US[time_unit]
            t0, t1 = 0, int(np.ceil((self.end_sim - self.start_sim).total_seconds() * 1e6 / unit))
            start_date = self.start_sim
        else:
            t0, t1, start_date = round(self.start_sim), round(self.end_sim), None
        n_out = int(t1 - t0 + 1)
        lp_d = torch.zeros(P, dtype=torch.float64, device=dev)
        st_ll = torch.zeros(P, dtype=torch.int32, device=dev)
        y0_sets = [y0] if not self.initial_states else [
            model._flat_y0({**model.initial_states, **{k: np.asarray(v, dtype=np.float64) for k, v in s.items()}})
            for s in self.initial_states]
        if seed is None:
            seed = int(np.random.randint(0, 2 ** 32)) | (int(np.random.randint(0, 2 ** 31)) << 32)
        for k, y in enumerate(y0_sets):
            if model._kind == "jump":
                tau, method = kw.get("tau", 1), kw.get("method", "tau_leap")
                n_steps = int(_ffi.load().sodm_jump_n_steps(float(t0), n_out, float(tau))) if method == "tau_leap" else 0
                # the table depends on the window and on the model's own parameter values only: one entry is kept
                key = (method, t0, n_out, float(tau), start_date, time_unit,
                       tuple((nm, np.asarray(v).tobytes()) for nm, v in model.parameters.items()))
                if getattr(self, "_tdpf_cache", (None, None))[0] != key:
                    from .base import _StatesTouched
                    try:
                        table = (model._tabulate_tdpfs(t0, n_steps, tau, start_date, time_unit) if method == "tau_leap"
                                 else model._tabulate_tdpf_intervals(float(t0), float(t1), 1.0, start_date, time_unit))
                    except _StatesTouched:
                        raise NotImplementedError("a time-dependent parameter function that could only be tabulated reads the "
                                                  f"model states: {getattr(model, '_tdpf_reason', {})}") from None
                    self._tdpf_cache = (key, table)
                tdpf = self._tdpf_cache[1]
                if model._spec.tdpf_params and tdpf is None:
                    flat = np.concatenate([model._flat_param(nm, model.parameters[nm]) for nm in model._spec.tdpf_params])
                    tdpf = (flat[None, :], np.zeros(max(n_steps, 1), dtype=np.int32))
                if method != "tau_leap":
                    tau = 1.0                                 # SSA: width of a table interval (include/sodm.h)
                # only the rows the datasets look at leave the simulator; exact integer counts when tau = 2^-k
                which_sets = range(len(self.data)) if not self.initial_states else [k]
                needed = sorted({r for i in which_sets for c in self._plan[i]["columns"] for r in c})
                remap = {r: j for j, r in enumerate(needed)}
                inv = 1.0 / float(tau)
                aligned = method == "tau_leap" and 0 < tau <= 1 and inv == int(inv) and (int(inv) & (int(inv) - 1)) == 0
                traj, st = engine.jump_run(model.device_model, y, par, P, t0, n_out, tau, method=method, var=var,
                                           tdpf=tdpf, compat=kw.get("compat", _ffi.COMPAT_REFERENCE), seed=seed + k, traj_id0=member_id0,
                                           out_dtype=torch.int32 if aligned else torch.float64, state_sel=needed, dev=dev)
                model._raise_on_status(st)
            else:
                traj, _, st = engine.ode_run(model.device_model, y, par, P, t0, n_out, method="euler" if kw.get("tau") else "RK45",
                                             rtol=kw.get("rtol", 1e-4), atol=kw.get("atol", 1e-6), tau=kw.get("tau") or 0.0,
                                             var=var, model_arg=model._model_arg, dev=dev)
                remap = None
            which = range(len(self.data)) if not self.initial_states else [k]
            for i in which:
                plan = self._plan[i]
                columns = plan["columns"] if remap is None else [[remap[r] for r in c] for c in plan["columns"]]
                engine.loglik(self._kinds[i], traj, float(t0), columns, plan["t_data"], plan["ydata"],
                              ll_arg=self.log_likelihood_fnc_args[i], weight=self.weights[i], lp=lp_d, status=st_ll, dev=dev,
                              column_weights=plan["weights"])
        if engine.status_bits(st_ll) & _ffi.ST_NONFINITE:
            raise ValueError("simulation output contains nan, likely due to numerical unstability. try using more conservative bounds.")
        return lp_d

    def __call__(self, thetas):
        thetas = np.asarray(thetas)
        view = thetas.reshape(1, -1)                     # a view: clipping mutates the caller's array like the reference
        return float(self.batch(view)[0])
