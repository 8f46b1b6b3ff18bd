"""Build, cache and load MODEL PLUGINS (include/sodm_plugin.h): one generated model per shared library.

`codegen` turns a model's Python functions into a C++ translation unit; this module compiles it with nvcc for sm_100a
(the same flags as libsodm.so, against the same kernel headers), keeps the result under `pysodm_b200/_plugins/` keyed by
the hash of source + headers (in-tree: it travels with the repository snapshot like libsodm.so and is never looked for
anywhere else), and binds its entry points with ctypes.  Building needs nvcc but no GPU.

After a build the functor is evaluated ONCE on the host (it is `__host__ __device__`) at the model's own initial state
and parameters and compared with what the Python function returns there: a translation that silently changed the
arithmetic (an unsupported NumPy idiom traced the wrong way) fails at construction, not in the results.  This is a
self-check of one evaluation, not a CPU execution path: simulations run on the device only.
"""
import ctypes as C
import hashlib
import os
import subprocess

import numpy as np

from . import _ffi, build, codegen

PLUGIN_DIR = os.path.join(build.ROOT, "pysodm_b200", "_plugins")
_loaded = {}


class Plugin:
    """A loaded plugin: `.info` (sodm_model_info), `.jump_run(args)` / `.ode_run(args)` (ctypes), `.layout`."""

    def __init__(self, path, layout=None):
        self.path, self.layout = path, layout or {}
        lib = C.CDLL(path)
        for name, (restype, argtypes) in {
            "sodm_plugin_abi_version": (C.c_int, []),
            "sodm_plugin_last_error": (C.c_char_p, []),
            "sodm_plugin_info": (C.c_int, [C.POINTER(_ffi.ModelInfo)]),
            "sodm_plugin_jump_run": (C.c_int, [C.POINTER(_ffi.JumpArgs)]),
            "sodm_plugin_ode_run": (C.c_int, [C.POINTER(_ffi.OdeArgs)]),
        }.items():
            try:
                fn = getattr(lib, name)
            except AttributeError as exc:
                raise ImportError(f"{path} does not export `{name}` declared in include/sodm_plugin.h") from exc
            fn.restype, fn.argtypes = restype, argtypes
        if lib.sodm_plugin_abi_version() != _ffi.ABI_VERSION:
            raise ImportError(f"{path}: ABI version {lib.sodm_plugin_abi_version()} != binding {_ffi.ABI_VERSION}")
        self.lib = lib
        self.info = _ffi.ModelInfo()
        self.check(lib.sodm_plugin_info(C.byref(self.info)))
        self.name = os.path.basename(path)

    def check(self, rc):
        if rc < 0:
            raise _ffi.SodmError(rc, self.lib.sodm_plugin_last_error().decode("utf-8", "replace"))

    def jump_run(self, args):
        self.check(self.lib.sodm_plugin_jump_run(C.byref(args)))

    def ode_run(self, args):
        self.check(self.lib.sodm_plugin_ode_run(C.byref(args)))

    def __repr__(self):
        return f"<Plugin {self.name} kind={'jump' if self.info.kind == _ffi.KIND_JUMP else 'ode'} n_state={self.info.n_state}>"


def _headers():
    hs = [os.path.join(build.CSRC, f) for f in os.listdir(build.CSRC) if f.endswith((".cuh", ".h"))]
    hs += [os.path.join(build.ROOT, "include", "sodm.h"), os.path.join(build.ROOT, "include", "sodm_plugin.h")]
    return sorted(hs)


def compile_source(source, layout=None, keep_source=True):
    """Compile a generated translation unit (or return the cached library).  -> Plugin."""
    h = hashlib.sha256(source.encode())
    for p in _headers():
        with open(p, "rb") as f:
            h.update(f.read())
    key = h.hexdigest()[:20]
    if key in _loaded:
        return _loaded[key]
    os.makedirs(PLUGIN_DIR, exist_ok=True)
    so = os.path.join(PLUGIN_DIR, f"model_{key}.so")
    if not os.path.exists(so):
        cu = os.path.join(PLUGIN_DIR, f"model_{key}.cu")
        with open(cu, "w") as f:
            f.write(source)
        tmp = so + f".tmp{os.getpid()}"
        cmd = ([build.NVCC] + build.ARCH + build.COMMON + ["-I", build.CSRC, "-shared", "-o", tmp, cu, "-cudart", "static"])
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        if proc.returncode != 0:
            raise RuntimeError("nvcc failed on the generated model:\n" + " ".join(cmd) + "\n" + proc.stdout[-4000:])
        os.replace(tmp, so)                      # atomic: concurrent ranks building the same model do not collide
        if not keep_source:
            os.remove(cu)
    plug = Plugin(so, layout)
    _loaded[key] = plug
    return plug


# ---------------------------------------------------------------------------------------------------------------------
def _shaped(a, shape):
    """A returned array as the flat vector of a state of `shape` (one-element arrays stand for scalars, as in the reference)."""
    a = np.asarray(a, dtype=np.float64)
    n = int(np.prod(shape)) if shape else 1
    return a.reshape(-1) if a.size == n else np.ravel(np.broadcast_to(a, tuple(shape)))


def _flat(values, shapes):
    return np.concatenate([np.ravel(np.asarray(values[k], dtype=np.float64)) for k in shapes]) if shapes else np.zeros(0)


def build_ode(integrate, state_shapes, param_shapes, origin="", reference=None, example_states=None, example_params=None,
              **trace_kw):
    """`param_shapes` in FLAT order (model functions are called by keyword, like the reference does).  `reference(t)` ->
    the derivatives as the Python functions give them at `example_states` / `example_params` (flat), for the self-check."""
    source, n_state, n_param = codegen.ode_source(integrate, state_shapes, param_shapes, origin, **trace_kw)
    plug = compile_source(source, {"n_state": n_state, "n_param": n_param})
    if reference is None:
        return plug
    # self-check of one evaluation against the Python function
    fn = plug.lib.sodm_plugin_host_rhs
    fn.restype, fn.argtypes = None, [C.c_double, C.c_void_p, C.c_void_p, C.c_void_p]
    y = _flat(example_states, state_shapes)
    par_buf = np.concatenate([np.asarray(example_params, dtype=np.float64), np.zeros(1)])
    dy = np.empty(n_state)
    warp_fn = getattr(plug.lib, "sodm_plugin_host_rhs_warp", None)     # present when the warp-per-draw layout was generated
    if warp_fn is not None:
        warp_fn.restype, warp_fn.argtypes = None, fn.argtypes
        plug.layout["warp"] = True
    for t in (0.0, 1.5, 40.25):
        fn(t, y.ctypes.data, par_buf.ctypes.data, dy.ctypes.data)
        if warp_fn is not None:                                        # the shape templates must give what the plain functor gives
            dw = np.full(n_state, np.nan)
            warp_fn(t, y.ctypes.data, par_buf.ctypes.data, dw.ctypes.data)
            if not np.allclose(dw, dy, rtol=1e-12, atol=0.0, equal_nan=True):
                j = int(np.argmax(~np.isclose(dw, dy, rtol=1e-12, atol=0.0, equal_nan=True)))
                raise codegen.TraceError(f"warp-layout translation disagrees with the plain functor (element {j}: {dw[j]!r} vs {dy[j]!r})")
        want = reference(t)
        want = np.concatenate([_shaped(w, s) for w, s in zip(want if isinstance(want, (tuple, list)) else (want,),
                                                             state_shapes.values())])
        scale = np.maximum(np.abs(want), 1e-300)
        with np.errstate(invalid="ignore"):
            bad = ~((np.abs(dy - want) <= 1e-9 * scale + 1e-12 * np.abs(want).max(initial=0.0)) | (np.isnan(dy) & np.isnan(want)))
        if bad.any():
            j = int(np.argmax(bad))
            raise codegen.TraceError(f"generated `integrate` disagrees with the Python function at the initial state, t = {t} "
                                     f"(element {j}: {dy[j]!r} vs {want[j]!r}); the model uses a NumPy idiom the tracer "
                                     "does not reproduce")
    return plug


def build_jump(compute_rates, apply_transitionings, state_shapes, param_shapes, tdpf_params, example_states,
               example_params, origin="", date_unit_us=None, date_params=(), start_date=None, tdpf=(), model_params=None):
    """`example_params`: {name: value} of every entry of `param_shapes` except the hidden `__t0__`; in date mode
    (`date_unit_us`) the model functions see dates: `start_date` is the date of t = 0 used by the self-check.
    `tdpf`: translated time-dependent parameter functions [(name, fn, extras)] (codegen.trace_jump); `tdpf_params`: the
    parameters that are read from the per-step table instead."""
    source, layout = codegen.jump_source(compute_rates, apply_transitionings, state_shapes, param_shapes, tdpf_params, origin,
                                         date_unit_us=date_unit_us, date_params=date_params, tdpf=tdpf, model_params=model_params)
    plug = compile_source(source, layout)
    lib = plug.lib
    lib.sodm_plugin_host_propose.restype = None
    lib.sodm_plugin_host_propose.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double, C.c_int, C.c_double, C.c_void_p,
                                             C.c_void_p]
    lib.sodm_plugin_host_apply.restype = None
    lib.sodm_plugin_host_apply.argtypes = [C.c_void_p, C.c_void_p]
    lib.sodm_plugin_host_propensities.restype = None
    lib.sodm_plugin_host_propensities.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_double]
    lib.sodm_plugin_host_fire.restype = None
    lib.sodm_plugin_host_fire.argtypes = [C.c_void_p, C.c_int]
    verify_jump(plug, compute_rates, apply_transitionings, state_shapes, param_shapes, example_states, example_params,
                date_unit_us=date_unit_us, date_params=date_params, start_date=start_date, tdpf=tdpf, model_params=model_params)
    return plug


def host_propose(plug, y, par, tau, quirk=True, t=0.0):
    """(n[NDRAW], p[NDRAW]) of one tau-leap step at time t as the generated functor lists them, evaluated on the host."""
    lay = plug.layout
    y_i = np.ascontiguousarray(y, dtype=np.int32)
    par_d = np.concatenate([np.asarray(par, dtype=np.float64), np.zeros(1)])
    tab = np.ascontiguousarray(par_d[lay["tdpf_offset"]:lay["tdpf_offset"] + max(lay["tdpf_count"], 1)], dtype=np.float32)
    n = np.empty(lay["n_draw"], dtype=np.int32)
    p = np.empty(lay["n_draw"], dtype=np.float32)
    plug.lib.sodm_plugin_host_propose(y_i.ctypes.data, par_d.ctypes.data, tab.ctypes.data, float(tau), int(quirk), float(t),
                                      n.ctypes.data, p.ctypes.data)
    return n, p


def verify_jump(plug, compute_rates, apply_transitionings, state_shapes, param_shapes, states, params, tau=0.5,
                date_unit_us=None, date_params=(), start_date=None, t=3.5, tdpf=(), model_params=None):
    """One step's probabilities (at time t) and one random bookkeeping update, generated functor vs the Python functions
    (translated time-dependent parameter functions applied first, as the reference's `func` does, base.py:376-386)."""
    lay = plug.layout
    rates_fn = codegen._python_function(compute_rates)
    apply_fn = codegen._python_function(apply_transitionings)
    kw = {k: np.asarray(states[k], dtype=np.float64) for k in state_shapes}
    pars = {k: params[k] for k in param_shapes if k != "__t0__"}
    when = t
    if date_unit_us is not None:
        from datetime import timedelta
        when = start_date + timedelta(microseconds=t * date_unit_us)
    live = dict(pars)
    for name, fn, extra in tdpf:
        live[name] = codegen._python_function(fn)(when, dict(kw), pars[name], **{k: live[k] for k in extra})
    kw.update({k: live[k] for k in (model_params if model_params is not None else pars)})
    parts = []
    for k, shape in param_shapes.items():
        if k == "__t0__":
            parts.append(np.array([codegen.date_to_units(start_date, date_unit_us)]))
        elif k in date_params:
            parts.append(np.array([codegen.date_to_units(params[k], date_unit_us)]))
        else:
            parts.append(np.ravel(np.asarray(params[k], dtype=np.float64)))
    par = np.concatenate(parts) if parts else np.zeros(0)
    y = _flat(states, state_shapes)
    rates = rates_fn(when, **kw)
    n, p = host_propose(plug, y, par, tau, quirk=True, t=t)
    # expected conditional probabilities (reference law with the stay-probability quirk, base.py:336-346)
    for name, e, m, drawn, first in lay["groups"]:
        r = [float(_shaped(rates[name][i], state_shapes[name])[e]) for i in range(m)]
        pi = [1.0 - np.exp(-tau * x) for x in r]
        W = 1.0 + pi[-1]
        for i in range(drawn):
            want = pi[i] / W
            got = float(p[first + i])
            if not (abs(got - want) <= 2e-5 * max(abs(want), 1e-30) + 1e-12 or (np.isnan(got) and np.isnan(want))):
                raise codegen.TraceError(f"generated `compute_rates` disagrees with the Python function at the initial "
                                         f"state (state '{name}' element {e} exit {i}: p = {got!r} vs {want!r})")
            W -= pi[i]
    rng = np.random.default_rng(0)
    T, draws = {}, np.zeros(lay["n_draw"], dtype=np.int32)
    for name in rates:
        T[name] = []
        for i in range(len(rates[name])):
            arr = np.zeros(int(np.prod(state_shapes[name])) if state_shapes[name] else 1)
            for e in range(arr.size):
                c = lay["channel_slots"][(name, i, e)]
                if c in lay["slot_of"]:
                    arr[e] = rng.integers(0, 50)
                    draws[lay["slot_of"][c]] = int(arr[e])
            T[name].append(arr.reshape(tuple(state_shapes[name])) if state_shapes[name] else arr[0])
    want = apply_fn(when, tau, T, **kw)
    want = np.concatenate([_shaped(w, s) for w, s in zip(want, state_shapes.values())])
    y_i = np.ascontiguousarray(y, dtype=np.int32)
    plug.lib.sodm_plugin_host_apply(y_i.ctypes.data, draws.ctypes.data)
    if not np.array_equal(y_i.astype(np.float64), want):
        j = int(np.argmax(y_i != want))
        raise codegen.TraceError(f"generated `apply_transitionings` disagrees with the Python function "
                                 f"(flat state element {j}: {y_i[j]} vs {want[j]})")
