"""pySODM's model front end (`ODE`, `JumpProcess`) with the simulation loop replaced by the B200 library.

The class API mirrors `/root/reference/src/pySODM/models/base.py`: class attributes `states`, `parameters`,
`stratified_parameters`, `dimensions`, `dimensions_per_state` (:92-96, :636-640); constructor
`(initial_states: dict | callable, parameters: dict, coordinates=None, time_dependent_parameters=None)` (:98-102);
`sim(time, N=1, draw_function=None, draw_function_kwargs={}, processes=None, method, tau | rtol, time_unit='D')`
(:485-494, :870-880) returning one `xarray.Dataset`-like object with dims `(['draws',] 'time'|'date', *state_dims)`
(:969-1042).  What differs is WHERE the N simulations run: the reference loops / forks a process pool over
`_mp_sim_single` (:558-566, :950-957); here all N trajectories go to the GPU in one batched call
(`sodm_jump_run` / `sodm_ode_run`), chunked through a double-buffered host pipeline (pipeline.py).

A model class is bound to a hand-written `__device__` functor by `device_model = "<registered name>"`.  A class
that declares none is TRANSLATED: its own Python functions (`integrate`, or `compute_rates` + `apply_transitionings`) are
traced into straight-line C++ (codegen.py), compiled by nvcc into a plugin library (plugin.py, include/sodm_plugin.h)
and driven by the same kernels.  A function that cannot be translated (value-dependent control flow, non-integral
bookkeeping) raises at construction: there is deliberately no CPU fallback.

Extensions over the reference signature (keyword-only, defaults = reference behaviour): `output_dtype`
('float64' like the reference; 'int32'/'float32' halve the 18-36 GB of config C3), `states` (subset of states to
return), `seed`, `device`, `chunk`, `compat` (tau-leap / SSA quirk flags), and for ODEs `atol`, `output_stride`,
`output_last_only` (the reference grid of the packed-bed config is 1.28 TB).
"""
import copy
import inspect
import random as _pyrandom
from collections import OrderedDict

import numpy as np

from . import _ffi, registry, validation
from .dataset import make_dataset


def _is_vectorized(fn):
    return bool(getattr(fn, "vectorized", False))


def _fresh_parameters(base):
    """What `copy.deepcopy(parameters)` gives a draw function, at a fraction of the cost: a new dict whose mutable values
    (arrays, lists, dicts) are copies and whose immutable ones (numbers, strings, timestamps, functions) are shared."""
    out = {}
    for k, v in base.items():
        if isinstance(v, np.ndarray):
            out[k] = v.copy()
        elif isinstance(v, (int, float, complex, str, bytes, bool, type(None), np.generic)) or callable(v):
            out[k] = v
        elif isinstance(v, list) and all(isinstance(x, (int, float, np.generic)) for x in v):
            out[k] = list(v)
        else:
            out[k] = copy.deepcopy(v)
    return out


class _StatesTouched(Exception):
    """A time-dependent parameter function that is being tabulated on the host read the model states."""


class _NeedsTracedTdpf(Exception):
    """The run needs a time-dependent parameter function evaluated per trajectory (it reads the states, or one of its
    arguments is drawn per trajectory): only a TRANSLATED function can do that."""


class _GuardedStates:
    """Stands in for `states` while a time-dependent parameter function is tabulated once per step for ALL trajectories:
    such a function must not depend on the states, and instead of probing it with a few made-up states (a function that
    happens to agree on the probes would be tabulated wrongly) every way of looking at the states raises."""

    def _touched(self, *a, **k):
        raise _StatesTouched()
    __getitem__ = __iter__ = __len__ = __contains__ = __bool__ = __eq__ = __ne__ = _touched
    __hash__ = None

    def __getattr__(self, name):
        raise _StatesTouched()


def _bounded_chunk(chunk, n_traj, bytes_per_traj, budget=4 << 30):
    """Trajectories per pipeline chunk: at most `chunk`, at most N, and device buffers of at most ~4 GiB each."""
    cap = max(32, int(budget // max(1, bytes_per_traj)) // 32 * 32)
    return max(1, min(int(chunk), max(int(n_traj), 1), cap))


class _DeviceBackedModel:
    states = None
    parameters = None
    stratified_parameters = None
    dimensions = None
    dimensions_per_state = None
    device_model = None            # name of the registered device functor (see registry.py / include/sodm.h)
    _kind = None
    _leading_args = ("t",)
    _rhs_name = None

    def __init__(self, initial_states, parameters, coordinates=None, time_dependent_parameters=None):
        self.states_names = list(self.states) if self.states is not None else None
        self.parameters_names = list(self.parameters) if self.parameters is not None else None
        if self.states_names is None or self.parameters_names is None:
            raise TypeError("a model class must declare 'states' and 'parameters'")
        self.parameters_stratified_names = self.stratified_parameters
        self.dimensions_names = list(self.dimensions) if self.dimensions else None
        self.coordinates = coordinates
        self.time_dependent_parameters = time_dependent_parameters
        validation.names(self.states_names, self.parameters_names, self.parameters_stratified_names, self.dimensions_names)
        self._stratified = validation.flat_stratified(self.parameters_stratified_names, self.dimensions_names or [])
        self.parameters_names_modeldeclaration = self.parameters_names + [n for n, _ in self._stratified]
        overlap = set(self.states_names) & set(self.parameters_names_modeldeclaration)
        if overlap:
            raise ValueError(f"overlap detected between 'state_names' and 'parameter_names + parameter_stratified_names': {sorted(overlap)}")
        self.state_shapes, self.dimensions_per_state, self.state_coordinates = validation.state_layout(
            self.states_names, self.dimensions_names, type(self).dimensions_per_state, coordinates)

        # time-dependent parameter functions: fun(t, states, param, *extra)
        self._function_parameters, self._extra_params_TDPF = [], []
        for name, fn in (time_dependent_parameters or {}).items():
            if name not in self.parameters_names_modeldeclaration:
                raise ValueError(f"The specified time-dependent parameter '{name}' is not a model parameter")
            args = list(inspect.signature(fn).parameters)
            if len(args) < 3 or args[0] != "t" or args[1] != "states" or args[2] != "param":
                raise ValueError(f"the signature of the time-dependent parameter function of '{name}' must start with "
                                 f"(t, states, param); found {args}")
            self._function_parameters.append(args[3:])
            self._extra_params_TDPF += [a for a in args[3:] if a not in self._extra_params_TDPF]

        # user functions' signatures
        rhs = getattr(self, self._rhs_name)
        validation.callable_signature(rhs, self._rhs_name, self.states_names, self.parameters_names_modeldeclaration,
                                      self._leading_args)
        self._check_extra_signatures()

        # initial condition function
        self.initial_states_function, self.initial_states_function_args = None, []
        if callable(initial_states):
            self.initial_states_function = initial_states
            self.initial_states_function_args = list(inspect.signature(initial_states).parameters)
        elif not isinstance(initial_states, dict):
            raise TypeError("initial states must be: 1) a dictionary, 2) a callable function generating a dictionary. "
                            f"found '{type(initial_states)}'")
        self._extra_params_ICF = list(self.initial_states_function_args)

        # parameters
        if not isinstance(parameters, dict):
            raise TypeError("'parameters' should be a dictionary.")
        required = list(self.parameters_names_modeldeclaration)
        required += [p for p in self._extra_params_TDPF + self._extra_params_ICF if p not in required]
        missing = [p for p in required if p not in parameters]
        if missing:
            raise ValueError(f"The specified parameters don't exactly match the predefined parameters. "
                             f"Missing parameters: {missing}")
        redundant = [p for p in parameters if p not in required]
        if redundant:
            raise ValueError(f"The specified parameters don't exactly match the predefined parameters. "
                             f"Redundant parameters: {redundant}")
        self.parameters = {k: parameters[k] for k in required}
        for name, dim in self._stratified:
            v = np.asarray(self.parameters[name])
            if v.ndim != 1 or v.shape[0] != len(coordinates[dim]):
                raise ValueError(f"The coordinates provided for dimension '{dim}' indicates a dimension size of "
                                 f"{len(coordinates[dim])}, but stratified parameter '{name}' has shape {v.shape}")
            self.parameters[name] = v

        # initial states
        if self.initial_states_function:
            given = self.initial_states_function(**{k: self.parameters[k] for k in self.initial_states_function_args})
        else:
            given = initial_states
        self.initial_states = validation.initial_states(given, self.state_shapes)

        self._bind_device_functor()

    def _check_extra_signatures(self):
        pass

    # ---------------------------------------------------------------------------------------------------------
    def _bind_device_functor(self):
        self._generated, self._generated_cache, self._date_params, self._time_unit_us = False, {}, [], None
        if self._kind == "ode" and self.time_dependent_parameters:
            # time-dependent parameter functions of an ODE run at every Runge-Kutta stage time (base.py:769-779): they are
            # translated together with `integrate`; whether they see numbers or dates is known at sim() -> built there
            self._generated, self.device_model = True, None
            return None
        if self.device_model and self._kind == "jump" and self.time_dependent_parameters:
            spec = registry.get(self.device_model)
            if any(name not in spec.tdpf_params for name in self.time_dependent_parameters):
                # the hand-written functor reads only `spec.tdpf_params` from the per-step table: a time-dependent parameter
                # function on any other parameter runs on the device as part of the TRANSLATED model
                self._hand_written, self.device_model = self.device_model, None
        if not self.device_model:
            self._generated = True
            try:
                return self._bind_generated_functor()
            except NotImplementedError as exc:
                if "date" not in str(exc):
                    raise
                return None                                  # the functions work on dates: translated at sim(), given dates
        spec = registry.get(self.device_model)
        if spec.kind != self._kind:
            raise TypeError(f"device functor '{spec.name}' is a {spec.kind} model, {type(self).__name__} is a {self._kind} model")
        self._model_arg = len(self.coordinates[spec.sized_by]) if spec.sized_by else 0
        states, params = spec.resolve(self._model_arg)
        if list(states) != self.states_names or any(tuple(states[k]) != tuple(self.state_shapes[k]) for k in states):
            raise ValueError(f"states of {type(self).__name__} {dict(self.state_shapes)} do not match device functor "
                             f"'{spec.name}' {dict(states)}")
        if set(params) != set(self.parameters_names_modeldeclaration):
            raise ValueError(f"parameters of {type(self).__name__} {self.parameters_names_modeldeclaration} do not match "
                             f"device functor '{spec.name}' {list(params)}")
        self._spec, self._spec_params = spec, params
        self._param_offsets, self._n_param_flat = spec.param_offsets(self._model_arg)

    def _bind_generated_functor(self, date_unit_us=None, date_range=None):
        """No hand-written functor (or an ODE with time-dependent parameter functions): translate the class's own Python
        functions (codegen.py) and compile them into a plugin library (plugin.py).  Flat parameter order: declaration
        order, then parameters only the TDPFs use; a jump process's time-dependent parameters go last (they are read from
        the per-step table); in date mode one hidden element `__t0__` (absolute time of t = 0) closes the vector.
        date_unit_us: None = model functions receive numbers for `t`; else they receive dates (sim was given dates)."""
        from . import plugin
        key = date_unit_us
        if key in self._generated_cache:
            plug, spec, shapes, self._date_params = self._generated_cache[key]
            self._traced_tdpfs = list(getattr(spec, "traced_tdpfs", []))
        else:
            tdpfs = self.time_dependent_parameters or {}
            decl = list(self.parameters_names_modeldeclaration)
            extras_of = {name: list(self._function_parameters[i]) for i, name in enumerate(tdpfs)}

            def shapes_of(order, lenient=()):
                shapes, date_params = OrderedDict(), []
                for name in order:
                    value = self.parameters[name]
                    if plugin.codegen._is_datelike(value):
                        if date_unit_us is None:
                            if name in lenient:
                                continue
                            raise NotImplementedError(f"parameter '{name}' is a date; simulate over dates (time=[start, stop])")
                        shapes[name] = ()
                        date_params.append(name)
                        continue
                    if callable(value) or isinstance(value, (str, dict)):
                        if name in lenient:
                            continue
                        raise NotImplementedError(f"parameter '{name}' is not numeric; generated device models take scalars, arrays and dates")
                    shapes[name] = tuple(np.shape(np.asarray(value, dtype=np.float64)))
                return shapes, date_params
            traced, self._tdpf_reason = [], {}
            if self._kind == "ode":
                tab = []
                order = decl + [n for n in self._extra_params_TDPF if n not in decl]
            else:
                # Which functions can be translated?  Those are evaluated on the device at every step / event with the live
                # states and the trajectory's own parameters, as the reference does (base.py:376-380); the others (calendar
                # logic, look-ups keyed by the date) are tabulated on the host and must be functions of the time alone.
                only_tdpf = [n for n in self._extra_params_TDPF if n not in decl]
                probe, probe_dates = shapes_of(decl + only_tdpf, lenient=only_tdpf)
                if date_unit_us is not None:
                    probe["__t0__"] = ()
                for name, fn in tdpfs.items():
                    ok, why = False, "an argument is not numeric"
                    if all(a in probe for a in extras_of[name]) and name not in probe_dates:
                        ok, why = plugin.codegen.tdpf_traceable(fn, extras_of[name], OrderedDict(
                            (k, tuple(self.state_shapes[k])) for k in self.states_names), probe, name, date_unit_us, probe_dates)
                    if ok:
                        traced.append(name)
                    else:
                        self._tdpf_reason[name] = why
                tab = [n for n in decl if n in tdpfs and n not in traced]
                for name in tab:                              # the host tabulation uses the ORIGINAL parameter values
                    if any(a in tdpfs for a in extras_of[name]):
                        raise NotImplementedError(f"the time-dependent parameter function of '{name}' cannot be translated "
                                                  f"({self._tdpf_reason[name]}) and takes another time-dependent parameter as an argument")
                need = [a for name in traced for a in extras_of[name] if a not in decl]
                order = [n for n in decl if n not in tab] + list(dict.fromkeys(need)) + tab
            shapes, date_params = shapes_of(order)
            if date_unit_us is not None:                      # hidden: absolute time of t = 0 (before the tabulated TDPF block)
                tail = [(k, shapes.pop(k)) for k in list(shapes) if k in tab]
                shapes["__t0__"] = ()
                shapes.update(tail)
            self._date_params = date_params
            self._traced_tdpfs = traced
            state_shapes = OrderedDict((k, tuple(self.state_shapes[k])) for k in self.states_names)
            n_par = sum(int(np.prod(v)) if v else 1 for v in shapes.values())
            n_st = sum(int(np.prod(v)) if v else 1 for v in state_shapes.values())
            if n_par > _ffi.MAX_PARAM or (n_st > _ffi.MAX_STATE and self._kind == "jump"):
                raise NotImplementedError(f"generated models are limited to {_ffi.MAX_PARAM} flat parameter elements (found "
                                          f"{n_par}) and, for jump processes, {_ffi.MAX_STATE} flat state elements (found {n_st})")
            origin = f" of {type(self).__module__}.{type(self).__qualname__}"
            if self._kind == "ode":
                spec_t = [(name, fn, self._function_parameters[i]) for i, (name, fn) in enumerate(tdpfs.items())]
                plug = plugin.build_ode(self.integrate, state_shapes, shapes, origin,
                                        reference=self._reference_rhs(date_range, date_unit_us), example_states=self.initial_states,
                                        example_params=self._example_flat(shapes, date_unit_us, date_range),
                                        tdpf=spec_t, model_params=decl, date_unit_us=date_unit_us, date_params=date_params)
            else:
                spec_t = [(name, tdpfs[name], extras_of[name]) for name in tdpfs if name in traced]
                plug = plugin.build_jump(self.compute_rates, self.apply_transitionings, state_shapes, shapes, tab,
                                         self.initial_states, self.parameters, origin, date_unit_us=date_unit_us,
                                         date_params=date_params, start_date=date_range[0] if date_range else None,
                                         tdpf=spec_t, model_params=decl)
            spec = registry.DeviceSpec(plug.name, self._kind, list(state_shapes.items()), list(shapes.items()), tdpf_params=tab)
            spec.traced_tdpfs = list(traced)
            self._generated_cache[key] = (plug, spec, shapes, date_params)
        self.device_model = plug
        self._model_arg = 0
        self._spec, self._spec_params = spec, shapes
        self._param_offsets, self._n_param_flat = spec.param_offsets(0)
        self._time_unit_us = date_unit_us

    def _example_flat(self, shapes, date_unit_us, date_range):
        """Flat parameter vector of the model's current parameters (for the build-time self-check of a generated ODE)."""
        from .codegen import date_to_units
        parts = []
        for name, shape in shapes.items():
            if name == "__t0__":
                parts.append(np.array([date_to_units(date_range[0], date_unit_us)]))
            elif name in self._date_params:
                parts.append(np.array([date_to_units(self.parameters[name], date_unit_us)]))
            else:
                parts.append(np.ravel(np.asarray(self.parameters[name], dtype=np.float64)))
        return np.concatenate(parts) if parts else np.zeros(0)

    def _reference_rhs(self, date_range, date_unit_us):
        """The right-hand side as the reference's `func` evaluates it (base.py:746-796), on NumPy values: TDPFs with the
        original parameter values, TDPF-only parameters dropped, then `integrate` -- used ONCE per build to check the
        translation, never to simulate."""
        tdpfs = self.time_dependent_parameters or {}
        decl = list(self.parameters_names_modeldeclaration)

        def rhs(t):
            when = t
            if date_unit_us is not None:
                when = validation.time_to_date(float(t), date_range[0], {v: k for k, v in validation.TIME_UNIT_US.items()}[int(date_unit_us)])
            states = {k: np.asarray(v, dtype=np.float64) for k, v in self.initial_states.items()}
            pars = dict(self.parameters)
            params = dict(pars)
            for i, (name, fn) in enumerate(tdpfs.items()):
                params[name] = fn(when, states, pars[name], **{k: pars[k] for k in self._function_parameters[i]})
            return self.integrate(when, **states, **{k: params[k] for k in decl})
        return rhs

    def _flat_y0(self, initial_states):
        return np.concatenate([np.ravel(np.asarray(v, dtype=np.float64)) for v in initial_states.values()])

    def _flat_param(self, name, value):
        off, n, shape = self._param_offsets[name]
        if name in self._date_params:                           # a date: absolute time in simulation units
            from .codegen import date_to_units
            return np.array([date_to_units(value, self._time_unit_us)])
        arr = np.asarray(value, dtype=np.float64)
        if arr.shape != tuple(shape) and not (arr.size == 1 and not shape):
            raise ValueError(f"parameter '{name}' must have shape {tuple(shape)}, found {arr.shape}")
        return arr.reshape(-1)

    def _collect_draws(self, N, draw_function, draw_function_kwargs):
        """Evaluate the draw function (reference: N calls in the parent, base.py:550-556) and split the result
        into the shared flat parameter vector, the per-trajectory rows, and the (possibly per-trajectory) y0."""
        base = self.parameters
        par = np.zeros(self._n_param_flat)
        tabulated = set(self._spec.tdpf_params)                # read from the per-step table (jump processes), not from `par`
        for name in self._spec_params:
            if name in tabulated or name == "__t0__":
                continue
            par[self._param_offsets[name][0]:][:self._param_offsets[name][1]] = self._flat_param(name, base[name])
        y0 = self._flat_y0(self.initial_states)
        if not draw_function:
            return par, {}, y0, [base]
        if _is_vectorized(draw_function):
            drawn = draw_function(copy.deepcopy(base), N=N, **draw_function_kwargs)
            stacked = {}
            for name, value in drawn.items():
                ref = np.asarray(base[name]) if not callable(base[name]) else None
                v = np.asarray(value) if ref is not None else None
                if ref is not None and v.ndim == ref.ndim + 1 and v.shape[0] == N:
                    stacked[name] = v
            dicts = None
        else:
            # The reference calls the draw function N times on a deep copy of the parameters (base.py:550-556).  At N = 10^6
            # a `deepcopy` per call plus a Python-level comparison of every parameter of every draw costs more than the
            # simulation, so: every call gets a fresh copy made by `_fresh_parameters` (arrays / lists copied, immutable
            # values shared), the returned values are only APPENDED per name, and which names vary is decided afterwards --
            # by identity where the function handed the same object back every time, else by ONE vectorised comparison of
            # the stacked values.
            dicts = [draw_function(_fresh_parameters(base), **draw_function_kwargs) for _ in range(N)]
            stacked = {}
            for name in base:
                first = dicts[0][name]
                if all(d[name] is first for d in dicts):
                    same = True
                else:
                    try:
                        if name in self._date_params:
                            col = np.stack([self._flat_param(name, d[name]) for d in dicts])[:, 0]
                        else:
                            col = np.stack([np.asarray(d[name], dtype=np.float64) for d in dicts])
                        same = bool(np.array_equal(col, np.broadcast_to(col[0], col.shape), equal_nan=True))
                    except Exception:  # noqa: BLE001  (non-numeric / ragged parameter values: compare one by one)
                        col = None
                        try:
                            same = all(np.array_equal(d[name], first) for d in dicts[1:])
                        except Exception:  # noqa: BLE001
                            same = False
                    if not same and col is not None and (name in self._date_params or name in self._spec_params or
                                                         name in self.initial_states_function_args):
                        stacked[name] = col
                if same and name in self._spec_params and name not in tabulated:
                    par[self._param_offsets[name][0]:][:self._param_offsets[name][1]] = self._flat_param(name, first)
        var = {}
        feeds_table = set(tabulated)                           # a tabulated function is evaluated once per step for ALL trajectories
        for i, tname in enumerate(self.time_dependent_parameters or {}):
            if tname in tabulated:
                feeds_table.update(self._function_parameters[i])
        for name, values in stacked.items():
            if name in feeds_table:
                raise _NeedsTracedTdpf(f"parameter '{name}' is drawn per trajectory and feeds the time-dependent parameter "
                                       f"function of {sorted(tabulated)}")
            if name in self._spec_params:
                off, n, shape = self._param_offsets[name]
                flat = np.asarray(values, dtype=np.float64).reshape(N, n)
                for e in range(n):
                    var[off + e] = np.ascontiguousarray(flat[:, e])
        if self.initial_states_function and any(a in stacked for a in self.initial_states_function_args):
            cols = []
            for i in range(N):
                kw = {a: (stacked[a][i] if a in stacked else base[a]) for a in self.initial_states_function_args}
                cols.append(self._flat_y0(validation.initial_states(self.initial_states_function(**kw), self.state_shapes)))
            y0 = np.ascontiguousarray(np.stack(cols, axis=1))
        return par, var, y0, dicts

    def _select_rows(self, states):
        """-> (pipeline variables, flat rows) for the requested subset of states (default: all, declaration order)."""
        from .pipeline import Variable
        names = self.states_names if states is None else list(states)
        offsets, r = {}, 0
        for name in self.states_names:
            n = int(np.prod(self.state_shapes[name])) if self.state_shapes[name] else 1
            offsets[name] = (r, n)
            r += n
        variables, rows, out_r = [], [], 0
        for name in names:
            if name not in offsets:
                raise ValueError(f"'{name}' is not a model state {self.states_names}")
            r0, n = offsets[name]
            variables.append(Variable(name, out_r, self.state_shapes[name]))
            rows += list(range(r0, r0 + n))
            out_r += n
        return variables, rows

    def _sim_on_device(self, run, var, y0, variables, n_rows, n_kept, N, sim_dtype, dev, date_range, t0, t1, raise_on_status,
                       time_coord=None):
        """One launch for this rank's trajectories into a resident [row][time][n_local] buffer -> DeviceEnsemble.
        Under torch.distributed the N draws are sharded as contiguous ranges (sharding.shard_range): every rank evaluated
        the same N draws on the host (same NumPy seed on every rank = same ensemble as one GPU would give), simulates its
        slice with Philox streams keyed by the GLOBAL draw index, and the ensemble's mean / quantile reduce across ranks."""
        import torch
        from . import sharding
        from .ensemble import DeviceEnsemble
        lo, hi = sharding.shard_range(N)
        n = hi - lo
        buf = torch.empty((n_rows, n_kept, n), dtype=sim_dtype, device=dev)
        keys = sorted(var)
        v = (keys, torch.from_numpy(np.stack([var[k][lo:hi] for k in keys])).to(dev)) if keys else None
        y0 = np.ascontiguousarray(y0 if y0.ndim == 1 else y0[:, lo:hi])
        y = torch.from_numpy(y0).to(dev)
        status = torch.zeros(n, dtype=torch.int32, device=dev)
        run(buf, v, y, status, lo, n)
        raise_on_status(status)
        time_name = "date" if date_range[0] is not None else "time"
        if time_coord is None:
            time_coord = date_range if date_range[0] is not None else np.arange(t0, t1 + 1)
        return DeviceEnsemble(buf, variables, self.dimensions_per_state, self.state_coordinates, time_name, time_coord,
                              traj_id0=lo, n_global=N)

    def _to_dataset(self, host, variables, N, time_name, time_coord):
        arrays = OrderedDict()
        for v in variables:
            a = host[v.name].numpy()
            arrays[v.name] = a[0] if N == 1 else a
        return make_dataset(arrays, self.dimensions_per_state, self.state_coordinates, time_name, time_coord, N)


# ================================================================================================================
class JumpProcess(_DeviceBackedModel):
    """Stochastic jump-process model solved with Gillespie's tau-leaping or SSA on the GPU.

    Subclasses declare `states`, `parameters`, optional `stratified_parameters` / `dimensions` /
    `dimensions_per_state`, the static methods `compute_rates(t, *states, *parameters)` and
    `apply_transitionings(t, tau, transitionings, *states, *parameters)`, and optionally `device_model`, a registered
    hand-written functor that restates them on the device; without it the two functions are translated (codegen.py)."""
    _kind = "jump"
    _rhs_name = "compute_rates"

    @staticmethod
    def compute_rates():
        raise NotImplementedError

    @staticmethod
    def apply_transitionings():
        raise NotImplementedError

    def _check_extra_signatures(self):
        validation.callable_signature(self.apply_transitionings, "apply_transitionings", self.states_names,
                                      self.parameters_names_modeldeclaration, ("t", "tau", "transitionings"))

    def _tdpf_values_at(self, t, start_date, time_unit):
        """Flat values of the TABULATED time-dependent parameters at time t: each function is called as the reference's
        `func` calls it -- with the ORIGINAL parameter value (base.py:376-380, Q8) -- except that `states` is a guard: a
        function that is evaluated once for all trajectories must not read them (`_StatesTouched`)."""
        when = validation.time_to_date(float(t), start_date, time_unit) if start_date is not None else t
        parts = []
        for name in self._spec.tdpf_params:
            if name in (self.time_dependent_parameters or {}):
                fn = self.time_dependent_parameters[name]
                i = list(self.time_dependent_parameters).index(name)
                kw = {a: self.parameters[a] for a in self._function_parameters[i]}
                val = fn(when, _GuardedStates(), self.parameters[name], **kw)
            else:
                val = self.parameters[name]
            parts.append(self._flat_param(name, val))
        return np.concatenate(parts)

    def _tabulate_tdpfs(self, t0, n_steps, tau, start_date, time_unit):
        """tau-leap: the tabulated functions at every step time t_k (accumulated in FP64 as base.py:433), de-duplicated into
        (table[n_values][n_elements], index[n_steps])."""
        if not self._spec.tdpf_params:
            return None
        rows, index, seen = [], np.empty(n_steps, dtype=np.int32), {}
        t = t0
        for k in range(n_steps):
            flat = self._tdpf_values_at(t, start_date, time_unit)
            key = flat.tobytes()
            if key not in seen:
                seen[key] = len(rows)
                rows.append(flat)
            index[k] = seen[key]
            t = t + tau
        return np.stack(rows), index

    def _tabulate_tdpf_intervals(self, t0, t1, width, start_date, time_unit):
        """SSA: the reference evaluates the functions at every EVENT time (base.py:376-380, :399).  A function that can only be
        tabulated must therefore be piecewise constant on the OPEN intervals (t0 + k w, t0 + (k+1) w) of the table (the
        tutorial's contact matrix changes at midnight, and its `a < t <= b` windows put the boundary itself on the earlier
        side: w = 1 day); that is verified at four interior points of every interval, and a function that varies inside one
        is refused (translate it, or pass a smaller `tdpf_resolution`).  Which side a boundary belongs to is immaterial --
        an event time hits one with probability zero -- except for the FIRST event, which is evaluated at t0 itself:
        index[0] = row at t0, index[1 + k] = row on interval k (include/sodm.h)."""
        if not self._spec.tdpf_params:
            return None
        n = max(1, int(np.ceil((t1 - t0) / width)))
        rows, index, seen = [], np.empty(n + 1, dtype=np.int32), {}

        def row_of(flat):
            key = flat.tobytes()
            if key not in seen:
                seen[key] = len(rows)
                rows.append(flat)
            return seen[key]
        index[0] = row_of(self._tdpf_values_at(t0, start_date, time_unit))
        for k in range(n):
            flat = self._tdpf_values_at(t0 + (k + 0.19) * width, start_date, time_unit)
            for frac in (2.0 ** -20, 0.37, 0.81, 1.0 - 2.0 ** -20):
                if not np.array_equal(flat, self._tdpf_values_at(t0 + (k + frac) * width, start_date, time_unit), equal_nan=True):
                    raise NotImplementedError(
                        f"SSA: a tabulated time-dependent parameter function ({list(self._spec.tdpf_params)}) changes inside the "
                        f"interval ({t0 + k * width}, {t0 + (k + 1) * width}); functions the device cannot translate must be "
                        "piecewise constant on intervals of `tdpf_resolution` time units")
            index[1 + k] = row_of(flat)
        return np.stack(rows), index

    def _switch_to_translated(self, why, want_unit, date_range):
        """Leave the hand-written functor (which can only read tabulated time-dependent parameters) for the translated
        model, whose time-dependent parameter functions run on the device per trajectory.  Raises when a function that
        would have to run there cannot be translated."""
        if not self._generated:
            self._hand_written = self.device_model
            self.device_model, self._generated, self._generated_cache = None, True, {}
        self._bind_generated_functor(want_unit, date_range)
        return why

    def sim(self, time, N=1, draw_function=None, draw_function_kwargs={}, processes=None, method="tau_leap", tau=1,
            time_unit="D", *, output_dtype="float64", states=None, seed=None, device=None, chunk=65536,
            compat=_ffi.COMPAT_REFERENCE, output="host", tdpf_resolution=1.0):
        """Simulate N trajectories over `time` (see module docstring).  `processes` is accepted and ignored.
        output='device' keeps the ensemble in HBM and returns a `DeviceEnsemble` (ensemble.py) for on-device noise /
        mean / quantiles instead of copying `(draws, time, *dims)` arrays to the host."""
        import torch
        from . import engine, pipeline
        validation.jump_solution_method(method, tau)
        time, date_range = validation.simulation_time(time, time_unit)
        if draw_function:
            validation.draw_function(draw_function, draw_function_kwargs, copy.deepcopy(self.parameters),
                                     self.initial_states_function, self.state_shapes, self.initial_states_function_args)
        want = validation.TIME_UNIT_US[time_unit] if date_range[0] is not None else None
        if self._generated:
            # model functions that use the time see dates when sim runs over dates (base.py:365-373): one translation each
            plug = self.device_model
            mode_sensitive = (plug is None or plug.layout.get("uses_time") or self._date_params or self._time_unit_us is not None
                              or bool(getattr(self, "_traced_tdpfs", None)))
            if mode_sensitive and (plug is None or self._time_unit_us != want):
                self._bind_generated_functor(want, date_range)
        saved = copy.deepcopy(self.parameters)
        try:
            t0, t1 = time
            n_out = int(t1 - t0 + 1)
            lib = _ffi.load()
            n_steps = int(lib.sodm_jump_n_steps(float(t0), n_out, float(tau))) if method == "tau_leap" else 0

            def tabulate():
                if method == "tau_leap":
                    return self._tabulate_tdpfs(t0, n_steps, tau, date_range[0], time_unit)
                return self._tabulate_tdpf_intervals(float(t0), float(t1), float(tdpf_resolution), date_range[0], time_unit)
            # Time-dependent parameter functions (base.py:376-380).  Translated ones run on the device with the live states and
            # the trajectory's own parameters.  The others are evaluated here, once per step for all trajectories -- legitimate
            # only for functions of the time alone: one that reads the states (guard), or is fed by a per-trajectory draw, moves
            # the model to the translated route; if the function cannot be translated either, that is an error, not a fallback.
            rng_state = (np.random.get_state(), _pyrandom.getstate())
            try:
                tdpf = tabulate()
                par, var, y0, _ = self._collect_draws(N, draw_function, draw_function_kwargs)
            except (_StatesTouched, _NeedsTracedTdpf) as exc:
                # translate for this run's time mode (a function written for dates only translates when given dates)
                before = (self._generated, self._time_unit_us, self.device_model)
                self._switch_to_translated(str(exc), want, date_range)
                np.random.set_state(rng_state[0])
                _pyrandom.setstate(rng_state[1])
                try:
                    if before == (self._generated, self._time_unit_us, self.device_model):
                        raise exc                              # nothing changed: the function in question is not translatable
                    tdpf = tabulate()
                    par, var, y0, _ = self._collect_draws(N, draw_function, draw_function_kwargs)
                except (_StatesTouched, _NeedsTracedTdpf) as exc2:
                    raise NotImplementedError(
                        f"a time-dependent parameter function of {list(self._spec.tdpf_params)} must run per trajectory ("
                        + (str(exc2) or "it reads the model states") + ") but cannot be translated for the device: "
                        f"{getattr(self, '_tdpf_reason', {})}") from None
            if "__t0__" in self._param_offsets:
                from .codegen import date_to_units
                par[self._param_offsets["__t0__"][0]] = date_to_units(date_range[0], self._time_unit_us)
            dev = engine.device(device)
            if self._spec.tdpf_params and tdpf is None:
                # constant value of a parameter that lives in the TDPF table
                flat = np.concatenate([self._flat_param(n, self.parameters[n]) for n in self._spec.tdpf_params])
                tdpf = (flat[None, :], np.zeros(max(n_steps, 1), dtype=np.int32))
            run_tau = tau if method == "tau_leap" else float(tdpf_resolution)      # SSA: width of a table interval (sodm.h)
            variables, rows = self._select_rows(states)
            host_dtype = engine.as_torch_dtype(output_dtype)
            # tau = 2^-k: every output time is a step time, counts are exact integers -> simulate in int32 and let
            # the layout kernel convert to the requested host dtype (halves the simulator's HBM writes)
            inv = 1.0 / float(tau)
            aligned = method == "tau_leap" and 0 < tau <= 1 and inv == int(inv) and (int(inv) & (int(inv) - 1)) == 0
            sim_dtype = torch.int32 if aligned else torch.float64
            if not aligned and host_dtype == torch.int32:
                raise ValueError("output_dtype='int32' needs tau-leap with tau = 2^-k (outputs are interpolated otherwise)")
            chunk = _bounded_chunk(chunk, N, len(rows) * n_out * (4 if aligned else 8))
            if seed is None:
                seed = int(np.random.randint(0, 2 ** 32)) | (int(np.random.randint(0, 2 ** 31)) << 32)
            if output == "device":
                from . import sharding
                seed = sharding.agree_on(seed)               # one ensemble across the ranks of a torch.distributed job
                return self._sim_on_device(
                    lambda buf, v, y, st, lo, n: engine.jump_run(
                        self.device_model, y, par, n, t0, n_out, run_tau, method=method, var=v, tdpf=self._to_device_tdpf(tdpf, dev),
                        compat=compat, seed=seed, traj_id0=lo, out_dtype=sim_dtype, state_sel=rows, out=buf, status=st, dev=dev),
                    var, y0, variables, len(rows), n_out, N, sim_dtype, dev, date_range, t0, t1, self._raise_on_status)
            if output != "host":
                raise ValueError("output must be 'host' or 'device'")
            host = pipeline.alloc_host_result(variables, N, n_out, host_dtype)
            pipe = pipeline.HostPipeline(variables, len(rows), n_out, chunk, sim_dtype, host_dtype, dev)
            y0_d = torch.from_numpy(np.ascontiguousarray(y0)).to(dev)
            tdpf_d = None
            if tdpf is not None:
                tdpf_d = (torch.from_numpy(np.ascontiguousarray(tdpf[0])).to(dev), torch.from_numpy(tdpf[1]).to(dev))
            keys = sorted(var)
            var_host = torch.from_numpy(np.stack([var[k] for k in keys])).pin_memory() if keys else None
            status = torch.zeros(N, dtype=torch.int32, device=dev)

            def launch(lo, n, buf, stream):
                v = (keys, var_host[:, lo:lo + n].to(dev, non_blocking=True)) if keys else None
                y = y0_d if y0_d.dim() == 1 else y0_d[:, lo:lo + n].contiguous()
                engine.jump_run(self.device_model, y, par, n, t0, n_out, run_tau, method=method, var=v, tdpf=tdpf_d,
                                compat=compat, seed=seed, traj_id0=lo, out_dtype=sim_dtype, state_sel=rows, out=buf,
                                status=status[lo:lo + n], dev=dev, stream=stream)
            pipe.run(N, launch, host)
            pipe.synchronize()
            self._raise_on_status(status)
        finally:
            self.parameters = saved
        time_name = "date" if date_range[0] is not None else "time"
        time_coord = date_range if date_range[0] is not None else np.arange(t0, t1 + 1)
        return self._to_dataset(host, variables, N, time_name, time_coord)

    @staticmethod
    def _to_device_tdpf(tdpf, dev):
        import torch
        if tdpf is None:
            return None
        return (torch.from_numpy(np.ascontiguousarray(tdpf[0])).to(dev), torch.from_numpy(tdpf[1]).to(dev))

    @staticmethod
    def _raise_on_status(status):
        from . import engine
        bits = engine.status_bits(status)
        if bits & _ffi.ST_BAD_PROB:
            raise ValueError("p < 0, p > 1 or p contains NaNs")          # NumPy's message (base.py:303)
        if bits & _ffi.ST_NONFINITE:
            raise ValueError("simulation produced non-finite propensities")
        if bits & _ffi.ST_MAX_STEPS:
            raise RuntimeError("event budget exhausted before the end of the simulation")


# ================================================================================================================
class ODE(_DeviceBackedModel):
    """Deterministic model integrated on the GPU: scipy-faithful RK45 / RK23 / DOP853 / Radau / BDF, or fixed-step Euler / RK4 when
    `tau` is given.

    Subclasses declare `states`, `parameters`, ..., the static method `integrate(t, *states, *parameters)` and
    optionally `device_model` (a registered hand-written functor); without it `integrate` is translated (codegen.py)."""
    _kind = "ode"
    _rhs_name = "integrate"

    @staticmethod
    def integrate():
        raise NotImplementedError

    def sim(self, time, N=1, draw_function=None, draw_function_kwargs={}, processes=None, method="RK45", rtol=1e-4,
            tau=None, time_unit="D", *, atol=1e-6, output_dtype="float64", output_stride=1, output_last_only=False,
            device=None, chunk=65536, output="host"):
        import torch
        from . import engine, pipeline
        validation.ode_solution_method(rtol, method, tau)
        time, date_range = validation.simulation_time(time, time_unit)
        if draw_function:
            validation.draw_function(draw_function, draw_function_kwargs, copy.deepcopy(self.parameters),
                                     self.initial_states_function, self.state_shapes, self.initial_states_function_args)
        if N != 1 and draw_function is None:
            raise ValueError(f"attempting to perform N={N} repeated simulations without using a draw function")
        if self._generated and self.time_dependent_parameters:
            # model functions see dates when sim is given dates, numbers otherwise (base.py:759-764): one translation each
            unit_us = validation.TIME_UNIT_US[time_unit] if date_range[0] is not None else None
            self._bind_generated_functor(unit_us, date_range)
        if tau:
            dev_method = "RK4" if method == "RK4" else "euler"
        elif method in ("RK45", "RK23", "DOP853", "Radau", "BDF"):
            dev_method = method          # Radau / BDF: thread-per-trajectory models of at most 32 equations (csrc/implicit.cuh)
        else:
            raise NotImplementedError(f"method '{method}' has no device implementation (available: 'RK45', 'RK23', 'DOP853', 'Radau', "
                                      "'BDF', or a fixed step `tau` for Euler / method='RK4'); there is no CPU fallback")
        if dev_method in ("Radau", "BDF"):
            plug, n_eq = self.device_model, len(self._select_rows(None)[1])
            if n_eq > 32 or plug == "packed_pfr_ode" or (plug is not None and not isinstance(plug, str) and plug.layout.get("warp")):
                raise NotImplementedError(f"method '{method}' keeps a dense Jacobian per trajectory: offered for models of at most "
                                          f"32 equations in the thread-per-trajectory layout (this model: {n_eq}); no CPU fallback")
        saved = copy.deepcopy(self.parameters)
        try:
            par, var, y0, _ = self._collect_draws(N, draw_function, draw_function_kwargs)
            if "__t0__" in self._param_offsets:                 # absolute time of t = 0, for date-valued model logic
                from .codegen import date_to_units
                par[self._param_offsets["__t0__"][0]] = date_to_units(date_range[0], self._time_unit_us)
            t0, t1 = time
            n_out = int(t1 - t0 + 1)
            dev = engine.device(device)
            lib = _ffi.load()
            n_kept = int(lib.sodm_ode_n_kept(n_out, int(output_stride), int(bool(output_last_only))))
            variables, rows = self._select_rows(None)
            host_dtype = engine.as_torch_dtype(output_dtype)
            if output == "device":
                grid = np.arange(t0, t1 + 1)
                coord = date_range if date_range[0] is not None else grid
                coord = coord[-1:] if output_last_only else coord[::int(output_stride)]

                def run(buf, v, y, st, lo, n):
                    _, nfev, st_ = engine.ode_run(self.device_model, y, par, n, t0, n_out, method=dev_method, rtol=rtol,
                                                  atol=atol, tau=tau or 0.0, var=v, model_arg=self._model_arg,
                                                  out_stride=output_stride, out_last_only=output_last_only, out=buf, dev=dev)
                    st.copy_(st_)
                    self.last_nfev = nfev

                def check(status):
                    if N and engine.status_bits(status) & (_ffi.ST_STEP_TOO_SMALL | _ffi.ST_NONFINITE):
                        raise RuntimeError("Required step size is less than spacing between numbers.")
                return self._sim_on_device(run, var, y0, variables, len(rows), n_kept, N, torch.float64, dev, date_range,
                                           t0, t1, check, time_coord=coord)
            if output != "host":
                raise ValueError("output must be 'host' or 'device'")
            host = pipeline.alloc_host_result(variables, N, n_kept, host_dtype)
            chunk = _bounded_chunk(chunk, N, len(rows) * n_kept * 8)
            pipe = pipeline.HostPipeline(variables, len(rows), n_kept, chunk, torch.float64, host_dtype, dev)
            y0_d = torch.from_numpy(np.ascontiguousarray(y0)).to(dev)
            keys = sorted(var)
            var_host = torch.from_numpy(np.stack([var[k] for k in keys])).pin_memory() if keys else None
            status = torch.zeros(N, dtype=torch.int32, device=dev)
            self.last_nfev = torch.zeros(N, dtype=torch.int32, device=dev)

            def launch(lo, n, buf, stream):
                v = (keys, var_host[:, lo:lo + n].to(dev, non_blocking=True)) if keys else None
                y = y0_d if y0_d.dim() == 1 else y0_d[:, lo:lo + n].contiguous()
                _, nfev, st = engine.ode_run(self.device_model, y, par, n, t0, n_out, method=dev_method, rtol=rtol,
                                             atol=atol, tau=tau or 0.0, var=v, model_arg=self._model_arg,
                                             out_stride=output_stride, out_last_only=output_last_only, out=buf, dev=dev,
                                             stream=stream)
                status[lo:lo + n].copy_(st, non_blocking=True)
                self.last_nfev[lo:lo + n].copy_(nfev, non_blocking=True)
            pipe.run(N, launch, host)
            pipe.synchronize()
            bits = engine.status_bits(status) if N else 0
            if bits & (_ffi.ST_STEP_TOO_SMALL | _ffi.ST_NONFINITE):
                raise RuntimeError("Required step size is less than spacing between numbers.")     # scipy's message
        finally:
            self.parameters = saved
        grid = np.arange(t0, t1 + 1)
        keep = grid[-1:] if output_last_only else grid[::int(output_stride)]
        if date_range[0] is not None:
            time_name = "date"
            time_coord = date_range[-1:] if output_last_only else date_range[::int(output_stride)]
        else:
            time_name, time_coord = "time", keep
        return self._to_dataset(host, variables, N, time_name, time_coord)
