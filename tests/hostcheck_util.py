"""Loader for the TEST-ONLY CPU harness (tests/hostcheck): the kernels' __host__ __device__ code run on the CPU.

Used by the `-m "not gpu"` tests to check sampler / stepper statistics without a GPU.  Never imported by
pysodm_b200.
"""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from pysodm_b200 import _ffi, build  # noqa: E402

_lib = None


def lib():
    global _lib
    if _lib is None:
        path = build.build_hostcheck()
        _lib = C.CDLL(path)
        _lib.hc_jump_run.restype = C.c_int
        _lib.hc_jump_run.argtypes = [C.POINTER(_ffi.JumpArgs)]
        _lib.hc_binomial.restype = None
        _lib.hc_binomial.argtypes = [C.c_uint64, C.c_uint64, C.c_int32, C.c_float, C.c_int64, C.c_void_p, C.c_void_p]
        _lib.hc_binomial_words.restype = None
        _lib.hc_binomial_words.argtypes = [C.c_uint64, C.c_uint64, C.c_int32, C.c_float, C.c_int64, C.c_void_p, C.c_void_p, C.c_uint32]
        _lib.hc_philox2x32.restype = None
        _lib.hc_philox2x32.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.c_void_p]
        _lib.hc_philox.restype = None
        _lib.hc_philox.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_int64, C.c_void_p]
        _lib.hc_binomial_log_ratio_large.restype = C.c_double
        _lib.hc_binomial_log_ratio_large.argtypes = [C.c_double, C.c_double, C.c_double]
        _lib.hc_binomial_log_ratio.restype = C.c_double
        _lib.hc_binomial_log_ratio.argtypes = [C.c_double, C.c_double, C.c_double]
        _lib.hc_noise.restype = C.c_int32
        _lib.hc_noise.argtypes = [C.c_int32, C.c_double, C.c_double, C.c_uint64, C.c_uint32, C.c_int64, C.c_void_p]
        _lib.sodm_model_id.restype = C.c_int
        _lib.sodm_model_id.argtypes = [C.c_char_p]
        _lib.sodm_model_info_get.restype = C.c_int
        _lib.sodm_model_info_get.argtypes = [C.c_int, C.c_int, C.POINTER(_ffi.ModelInfo)]
        _lib.sodm_last_error.restype = C.c_char_p
    return _lib


def binomial(n, p, count, seed=1, traj=0):
    out = np.empty(count, dtype=np.int32)
    attempts = np.zeros(1, dtype=np.int64)
    lib().hc_binomial(seed, traj, int(n), float(p), count, out.ctypes.data, attempts.ctypes.data)
    return out, int(attempts[0])


def binomial_words(n, p, count, seed=1, traj=0, force_word=0):
    """The two-pass sampler's unit of work (binomial_attempt), one draw at a time."""
    out = np.empty(count, dtype=np.int32)
    attempts = np.zeros(1, dtype=np.int64)
    lib().hc_binomial_words(seed, traj, int(n), float(p), count, out.ctypes.data, attempts.ctypes.data, int(force_word))
    return out, int(attempts[0])


def noise(kind, param, x, count, seed=1, cell=0):
    out = np.empty(count, dtype=np.float64)
    status = lib().hc_noise(int(kind), float(param), float(x), seed, cell, count, out.ctypes.data)
    return out, status


def philox2x32(key, c0, c1):
    out = np.empty(2, dtype=np.uint32)
    lib().hc_philox2x32(key, c0, c1, out.ctypes.data)
    return out


def philox(seed, traj, purpose, n_blocks):
    out = np.empty(4 * n_blocks, dtype=np.uint32)
    lib().hc_philox(seed, traj, purpose, n_blocks, out.ctypes.data)
    return out.reshape(n_blocks, 4)


_DT = {np.dtype(np.int32): _ffi.I32, np.dtype(np.float32): _ffi.F32, np.dtype(np.float64): _ffi.F64}


def jump_run(model, y0, par, n_traj, t0, n_out, tau=0.5, method="tau_leap", var=None, tdpf=None,
             compat=_ffi.COMPAT_REFERENCE, seed=1718, traj_id0=0, out_dtype=np.float64, state_sel=None,
             max_events=0):
    """Host run of the jump stepper.  Returns (out[n_rows, n_out, n_traj], status[n_traj]).

    y0: [n_state] (broadcast) or [n_state, n_traj];  par: [n_param];  var: {flat_param_index: values[n_traj]};
    tdpf: (table[n_values, tdpf_count], step_index[n_steps])."""
    L = lib()
    mid = L.sodm_model_id(model.encode())
    assert mid >= 0, L.sodm_last_error()
    info = _ffi.ModelInfo()
    assert L.sodm_model_info_get(mid, 0, C.byref(info)) == 0
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    assert par.size == info.n_param, (par.size, info.n_param)
    n_rows = info.n_state if state_sel is None else len(state_sel)
    out = np.full((n_rows, n_out, n_traj), np.nan if np.dtype(out_dtype).kind == "f" else -1, dtype=out_dtype)
    status = np.zeros(n_traj, dtype=np.int32)
    a = _ffi.JumpArgs()
    a.model, a.method = mid, (_ffi.JUMP_TAU_LEAP if method == "tau_leap" else _ffi.JUMP_SSA)
    a.compat, a.out_dtype = compat, _DT[np.dtype(out_dtype)]
    a.n_traj, a.traj_id0, a.seed = n_traj, traj_id0, seed
    a.ld_in, a.ld_out = n_traj, n_traj
    a.y0, a.y0_broadcast = y0.ctypes.data, int(y0.ndim == 1)
    a.par = par.ctypes.data
    keep = [y0, par, out, status]
    if var:
        idx = np.array(sorted(var), dtype=np.int32)
        vals = np.ascontiguousarray(np.stack([np.asarray(var[i], dtype=np.float64) for i in sorted(var)]))
        a.n_var, a.var_index, a.var_values = len(idx), idx.ctypes.data, vals.ctypes.data
        keep += [idx, vals]
    if tdpf is not None:
        table = np.ascontiguousarray(np.asarray(tdpf[0], dtype=np.float64).reshape(len(tdpf[0]), -1))
        index = np.ascontiguousarray(tdpf[1], dtype=np.int32)
        a.tdpf_table, a.tdpf_index = table.ctypes.data, index.ctypes.data
        a.n_tdpf_values, a.n_tdpf_index = table.shape[0], index.size
        keep += [table, index]
    a.t0, a.n_out, a.tau = float(t0), int(n_out), float(tau)
    if state_sel is not None:
        sel = np.ascontiguousarray(state_sel, dtype=np.int32)
        a.n_sel, a.state_sel = sel.size, sel.ctypes.data
        keep.append(sel)
    a.out, a.status, a.max_events = out.ctypes.data, status.ctypes.data, max_events
    rc = L.hc_jump_run(C.byref(a))
    if rc != 0:
        raise _ffi.SodmError(rc, L.sodm_last_error().decode())
    return out, status


def lu_solve(A, b):
    """A x = b through the implicit steppers' own dense LU (n = 3, 8: compile-time-indexed path; 12: run-time path)."""
    L = lib()
    A = np.asarray(A)
    n = A.shape[0]
    cplx = np.iscomplexobj(A) or np.iscomplexobj(b)
    are = np.ascontiguousarray(A.real, dtype=np.float64)
    aim = np.ascontiguousarray(A.imag, dtype=np.float64) if cplx else None
    bre = np.ascontiguousarray(np.asarray(b).real, dtype=np.float64).copy()
    bim = np.ascontiguousarray(np.asarray(b).imag, dtype=np.float64).copy() if cplx else None
    L.hc_lu_solve.restype = C.c_int
    L.hc_lu_solve.argtypes = [C.c_int] + 4 * [C.c_void_p]
    rc = L.hc_lu_solve(n, are.ctypes.data, aim.ctypes.data if cplx else None, bre.ctypes.data, bim.ctypes.data if cplx else None)
    assert rc == 0, rc
    return bre + 1j * bim if cplx else bre


TEST_ODE_MODELS = {"test:robertson": (0, 3, 3), "test:vanderpol": (1, 2, 1), "test:chain3": (2, 3, 10), "test:chain12": (3, 12, 37)}      # name -> (id, n_state, n_param)


def ode_run(model, y0, par, n_traj, t0, n_out, method="RK45", rtol=1e-4, atol=1e-6, tau=0.0, var=None,
            model_arg=0, out_stride=1, out_last_only=False, window=0, max_steps=0):
    """Host run of the ODE integrators.  Returns (out[n_state, n_kept, n_traj], nfev[n_traj], status[n_traj]).
    window > 0: through the device thread kernel's windowed output driver (as a warp of one lane)."""
    L = lib()
    if not hasattr(L, "_ode_ready"):
        L.hc_ode_run.restype = C.c_int
        L.hc_ode_run.argtypes = [C.POINTER(_ffi.OdeArgs)]
        L._ode_ready = True
    entry = L.hc_ode_run
    if model in TEST_ODE_MODELS:                                # functors that exist in the harness only (stiff test problems)
        mid, n_state, n_param = TEST_ODE_MODELS[model]
        info = _ffi.ModelInfo(n_state=n_state, n_param=n_param)
        entry = L.hc_ode_test_model_run
        entry.restype, entry.argtypes = C.c_int, [C.POINTER(_ffi.OdeArgs)]
    else:
        mid = L.sodm_model_id(model.encode())
        assert mid >= 0, L.sodm_last_error()
        info = _ffi.ModelInfo()
        assert L.sodm_model_info_get(mid, model_arg, C.byref(info)) == 0, L.sodm_last_error()
    y0 = np.ascontiguousarray(y0, dtype=np.float64)
    par = np.ascontiguousarray(par, dtype=np.float64)
    assert par.size == info.n_param, (par.size, info.n_param)
    n_kept = 1 if out_last_only else (n_out - 1) // out_stride + 1
    out = np.full((info.n_state, n_kept, n_traj), np.nan)
    nfev = np.zeros(n_traj, dtype=np.int32)
    status = np.zeros(n_traj, dtype=np.int32)
    a = _ffi.OdeArgs()
    a.model, a.model_arg = mid, model_arg
    a.method = {"RK45": _ffi.ODE_RK45, "euler": _ffi.ODE_EULER, "RK4": _ffi.ODE_RK4, "RK23": _ffi.ODE_RK23, "DOP853": _ffi.ODE_DOP853, "Radau": _ffi.ODE_RADAU, "BDF": _ffi.ODE_BDF}[method]
    a.out_dtype = _ffi.F64
    a.n_traj, a.ld_in, a.ld_out = n_traj, n_traj, n_traj
    a.y0, a.y0_broadcast = y0.ctypes.data, int(y0.ndim == 1)
    a.par = par.ctypes.data
    keep = [y0, par]
    if var:
        idx = np.array(sorted(var), dtype=np.int32)
        vals = np.ascontiguousarray(np.stack([np.asarray(var[i], dtype=np.float64) for i in sorted(var)]))
        a.n_var, a.var_index, a.var_values = len(idx), idx.ctypes.data, vals.ctypes.data
        keep += [idx, vals]
    a.t0, a.n_out, a.out_stride, a.out_last_only = float(t0), int(n_out), int(out_stride), int(out_last_only)
    a.rtol, a.atol, a.tau = rtol, atol, tau
    a.reserved, a.max_steps = int(window), int(max_steps)
    a.out, a.nfev, a.status = out.ctypes.data, nfev.ctypes.data, status.ctypes.data
    rc = entry(C.byref(a))
    if rc != 0:
        raise _ffi.SodmError(rc, L.sodm_last_error().decode())
    return out, nfev, status
