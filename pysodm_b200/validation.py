"""Host-side input checks of the model front end.

The reference validates class declarations and `sim()` arguments in `src/pySODM/models/validation.py`
(`validate_simulation_time` :77-116, `validate_solution_methods_*` :133-190, `validate_draw_function` :193-272,
`validate_dimensions*` :281-377, `validate_initial_states` :445-490, ...).  This module re-implements the checks the
hot path depends on -- same accepted inputs, same exception TYPES, equivalent messages -- because they define the
flat state layout and the time grid the kernels run on.  It is init-time bookkeeping, not part of the timed path.
"""
import copy
import inspect
import math
from collections import OrderedDict
from datetime import datetime, timedelta

import numpy as np

TIME_UNIT_US = {"us": 1, "ms": 1000, "s": 10 ** 6, "min": 60 * 10 ** 6, "h": 3600 * 10 ** 6, "D": 86400 * 10 ** 6,
                "W": 7 * 86400 * 10 ** 6}
_NUMBER = (int, float, np.int32, np.float32, np.int64, np.float64)


def time_to_date(t, start_date, time_unit):
    return start_date + timedelta(microseconds=t * TIME_UNIT_US[time_unit])


def _chronological(time):
    if time[1] < time[0]:
        raise ValueError("start of simulation is chronologically after end of simulation")
    if time[0] == time[1]:
        raise ValueError("start of simulation is the same as the end of simulation")


def simulation_time(time, time_unit):
    """-> ([t0, t1] as numbers, date_range or [None]).  Accepts what the reference accepts (validation.py:77-116):
    a number (`[0, round(time)]`), `[start, stop]` numbers (rounded), or two 'YYYY-MM-DD' strings / datetimes
    (`[0, ceil(delta / unit)]` with the list of output datetimes)."""
    date_range = [None]
    if isinstance(time, float):
        time = [0, round(time)]
    elif isinstance(time, int):
        time = [0, time]
    elif isinstance(time, list):
        if len(time) != 2:
            raise ValueError(f"wrong length of list-like simulation start and stop (length: {len(time)}). "
                             "correct format: time=[start, stop] (length: 2).")
        if all(isinstance(x, _NUMBER) for x in time):
            time = [round(x) for x in time]
        elif all(isinstance(x, (str, datetime)) for x in time) and len({type(x) for x in time}) == 1:
            if isinstance(time[0], str):
                time = [datetime.strptime(x, "%Y-%m-%d") for x in time]
            _chronological(time)
            if time_unit not in TIME_UNIT_US:
                raise ValueError(f"Invalid time unit '{time_unit}'. Choose from {list(TIME_UNIT_US)}")
            delta_us = (time[1] - time[0]).total_seconds() * 1e6
            n = math.ceil(delta_us / TIME_UNIT_US[time_unit])
            date_range = [time[0] + timedelta(microseconds=i * TIME_UNIT_US[time_unit]) for i in range(n + 1)]
            time = [0, n]
        else:
            raise ValueError("simulation start and stop must have the format: time=[start, stop]. 'start' and 'stop' "
                             "must have the same datatype: int/float, str ('yyyy-mm-dd'), or datetime. mixing of types "
                             f"is not allowed. you supplied: {[type(t) for t in time]} ")
    else:
        raise TypeError("'time' must be 1) a single int/float representing the end of the simulation, "
                        "2) a list of format: time=[start, stop].")
    _chronological(time)
    return time, date_range


def jump_solution_method(method, tau):
    if not isinstance(method, str):
        raise TypeError("solver method 'method' must be of type string")
    if method not in ("tau_leap", "SSA"):
        raise ValueError(f"invalid solution method '{method}'. valid methods: 'SSA' and 'tau_leap'")
    if isinstance(tau, bool) or not isinstance(tau, (int, float)):
        raise TypeError("discrete timestep 'tau' must be of type int or float")


def ode_solution_method(rtol, method, tau):
    if not isinstance(rtol, float):
        raise TypeError("relative solver tolerance 'rtol' must be of type float")
    if not isinstance(method, str):
        raise TypeError("solver method 'method' must be of type string")
    if method not in ("RK23", "RK45", "DOP853", "Radau", "BDF", "LSODA", "RK4"):
        raise ValueError(f"invalid solution method '{method}'. valid methods: 'RK23', 'RK45', 'DOP853', 'Radau', "
                         "'BDF', 'LSODA'")
    if tau is not None and (isinstance(tau, bool) or not isinstance(tau, (int, float))):
        raise TypeError("discrete timestep 'tau' must be of type int or float")


def names(states, parameters, stratified, dimensions):
    for label, lst in (("states", states), ("parameters", parameters)):
        if not isinstance(lst, (list, tuple)) or not all(isinstance(x, str) for x in lst):
            raise TypeError(f"not all elements in '{label}' are of type str.")
    if dimensions is not None and not all(isinstance(x, str) for x in dimensions):
        raise TypeError("not all elements in 'dimensions' are of type str.")
    if stratified:
        if not dimensions:
            raise TypeError("a model without dimensions cannot have stratified parameters.")
    for label, lst in (("state_names", states), ("parameter_names", parameters), ("dimension_names", dimensions or [])):
        if len(set(lst)) != len(lst):
            raise ValueError(f"duplicates in '{label}': {sorted({x for x in lst if list(lst).count(x) > 1})}")


def flat_stratified(stratified, dimensions):
    """['a','b'] (1 dimension) or [['a'],[],['b']] (n dimensions) -> [(name, dimension), ...]"""
    if not stratified:
        return []
    if all(isinstance(x, str) for x in stratified):
        if len(dimensions) != 1:
            raise TypeError(f"'stratified_parameters' must be a list containing {len(dimensions)} sublists.")
        return [(x, dimensions[0]) for x in stratified]
    if len(stratified) != len(dimensions) or not all(isinstance(x, (list, tuple)) for x in stratified):
        raise TypeError(f"'stratified_parameters' must be a list containing {len(dimensions)} sublists.")
    out = []
    for dim, sub in zip(dimensions, stratified):
        for x in sub:
            if not isinstance(x, str):
                raise TypeError("each sublist of 'stratified_parameters' must either be empty or contain only str.")
            out.append((x, dim))
    return out


def state_layout(state_names, dimensions, dimensions_per_state, coordinates):
    """-> (state_shapes, dims_per_state, coords_per_state) as OrderedDicts (reference: build_state_sizes_dimensions)."""
    if dimensions:
        if not coordinates:
            raise ValueError("dimension name provided in integrate function but no coordinates were given")
        for d in dimensions:
            if d not in coordinates:
                raise ValueError(f"no coordinates provided for dimension '{d}'")
        for d in coordinates:
            if d not in dimensions:
                raise ValueError(f"coordinates provided for '{d}', which is not a model dimension {list(dimensions)}")
    elif coordinates:
        raise ValueError("you have provided a dictionary of coordinates but there are no model dimensions")
    if dimensions_per_state is not None:
        if len(dimensions_per_state) != len(state_names):
            raise ValueError(f"length of list/dict containing the dimensions of every model state "
                             f"({len(dimensions_per_state)}) must equal the number of model states ({len(state_names)})")
        per_state = [list(x) for x in dimensions_per_state]
        for lst in per_state:
            for d in lst:
                if d not in (coordinates or {}):
                    raise ValueError(f"dimension '{d}' in 'dimensions_per_state' has no coordinates")
    else:
        per_state = [list(dimensions or []) for _ in state_names]
    shapes, dims, coords = OrderedDict(), OrderedDict(), OrderedDict()
    for name, lst in zip(state_names, per_state):
        shapes[name] = tuple(len(coordinates[d]) for d in lst)
        dims[name] = lst
        coords[name] = [list(coordinates[d]) for d in lst]
    return shapes, dims, coords


def initial_states(given, state_shapes):
    """Fill missing states with zeros and check shapes -> OrderedDict name -> float64 ndarray (reference :445-490)."""
    if not isinstance(given, dict):
        raise TypeError("initial states must be: 1) a dictionary, 2) a callable function generating a dictionary")
    unknown = [k for k in given if k not in state_shapes]
    if unknown:
        raise ValueError(f"the initial states contain names that are not model states: {unknown}. "
                         f"valid states: {list(state_shapes)}")
    out = OrderedDict()
    for name, shape in state_shapes.items():
        if name in given:
            v = given[name]
            if not isinstance(v, (list, tuple, np.ndarray, int, float, np.integer, np.floating)):
                raise TypeError(f"initial state '{name}' must be of type int, float, list or np.ndarray; found {type(v)}")
            arr = np.asarray(v, dtype=np.float64)
            if arr.shape != tuple(shape) and not (arr.size == 1 and not shape):
                raise ValueError(f"the desired shape of model state '{name}' is {tuple(shape)}, but provided "
                                 f"initial state has shape {arr.shape}")
            out[name] = arr.reshape(shape)
        else:
            out[name] = np.zeros(shape, dtype=np.float64)
    return out


def draw_function(fn, kwargs, parameters, icf, state_shapes, icf_args):
    """Same contract as the reference (validation.py:193-272); calls the function once."""
    if not callable(fn):
        raise TypeError("a 'draw function' must be callable (a function)")
    args = list(inspect.signature(fn).parameters)
    if not args or args[0] != "parameters":
        raise ValueError(f"your draw function '{fn.__name__}' must have 'parameters' as its first input. "
                         f"Its current inputs are: '{args}'")
    if not isinstance(kwargs, dict):
        raise TypeError(f"your `draw_function_kwargs` must be of type 'dict' but are of type '{type(kwargs)}'")
    extra = [a for a in args[1:] if a not in ("N",) or a in kwargs]
    if extra and not kwargs:
        raise ValueError(f"the draw function '{fn.__name__}' has {len(extra)} arguments in addition to the mandatory "
                         "'parameters' argument\nhave you forgotten to pass `draw_function_kwargs` to the sim() function?")
    if set(extra) != set(kwargs):
        raise ValueError(f"incorrect arguments passed to draw function '{fn.__name__}'\nkeys missing in "
                         f"'draw_function_kwargs': {set(extra) - set(kwargs)}. redundant keys: {set(kwargs) - set(extra)}")
    if getattr(fn, "vectorized", False):
        return
    out = fn(copy.deepcopy(parameters), **kwargs)
    if not isinstance(out, dict):
        raise TypeError("a draw function must return one dictionary containing the model's parameters. "
                        f"found type '{type(out)}'")
    if set(out) != set(parameters):
        raise ValueError("a draw function must return the dictionary of model parameters; keys missing in draw function "
                         f"output: {set(parameters) - set(out)}. redundant keys: {set(out) - set(parameters)}")
    if icf:
        try:
            initial_states(icf(**{k: out[k] for k in icf_args}), state_shapes)
        except Exception as e:  # noqa: BLE001
            raise RuntimeError("draw function --> drawn parameter dictionary --> initial condition function --> "
                               f"invalid initial state.\nfound error: {e}") from e


def callable_signature(fn, label, state_names, parameter_names, leading, extra_allowed=()):
    """`compute_rates(t, *states, *parameters)` / `integrate(...)` / `apply_transitionings(t, tau, transitionings, ...)`"""
    args = list(inspect.signature(fn).parameters)
    if args[:len(leading)] != list(leading):
        raise ValueError(f"The first {len(leading)} argument(s) of '{label}' must be {list(leading)}; found {args[:len(leading)]}")
    rest = args[len(leading):]
    if rest[:len(state_names)] != list(state_names):
        raise ValueError(f"The states in '{label}' ({rest[:len(state_names)]}) don't match the model's states {list(state_names)}")
    pars = rest[len(state_names):]
    missing = [p for p in parameter_names if p not in pars]
    unknown = [p for p in pars if p not in parameter_names and p not in extra_allowed]
    if missing or unknown:
        raise ValueError(f"The parameters in '{label}' {pars} don't match the model's parameters {list(parameter_names)}")
    return pars
