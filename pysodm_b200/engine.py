"""Thin device layer: PyTorch owns the buffers and streams, `libsodm.so` does the work.

Every function here takes / returns `torch` CUDA tensors (or NumPy arrays that are copied to the device),
fills one argument struct of `include/sodm.h` with raw pointers and sizes and calls the C ABI through
`_ffi`.  Nothing is computed in Python or in torch; there is no CPU path.
"""
import ctypes as C

import numpy as np
import torch

from . import _ffi

_TORCH_DT = {torch.int32: _ffi.I32, torch.float32: _ffi.F32, torch.float64: _ffi.F64}
_NP_TO_TORCH = {np.dtype(np.int32): torch.int32, np.dtype(np.float32): torch.float32,
                np.dtype(np.float64): torch.float64}
_initialised = set()


def as_torch_dtype(dt):
    if isinstance(dt, torch.dtype):
        return dt
    if isinstance(dt, str):
        dt = {"i32": np.int32, "int32": np.int32, "f32": np.float32, "float32": np.float32,
              "f64": np.float64, "float64": np.float64}[dt]
    return _NP_TO_TORCH[np.dtype(dt)]


def device(index=None) -> torch.device:
    """Select (and, the first time, initialise) a CUDA device.  Raises if there is no GPU."""
    if not torch.cuda.is_available():
        raise RuntimeError("pysodm_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    if index is None:
        index = torch.cuda.current_device()
    if index not in _initialised:
        lib = _ffi.load()
        _ffi.check(lib.sodm_init(int(index)), lib)
        _initialised.add(index)
    torch.cuda.set_device(index)
    return torch.device("cuda", index)


def status_bits(status):
    """Bitwise OR of a per-trajectory status vector (SODM_ST_* flags) as a Python int -- not its maximum: a benign high bit
    on one trajectory (SODM_ST_TRUNCATED_Y0) must not hide an error bit on another.  One device-to-host read."""
    if status.numel() == 0:
        return 0
    if not status.is_cuda:                               # host vectors (tests): NumPy
        return int(np.bitwise_or.reduce(status.numpy().astype(np.int64).ravel()))
    if status.dtype != torch.int32 or not status.is_contiguous():
        raise TypeError("status must be a contiguous int32 tensor")
    lib = _ffi.load()
    out = torch.zeros(1, dtype=torch.int32, device=status.device)
    _ffi.check(lib.sodm_status_or(status.data_ptr(), int(status.numel()), out.data_ptr(), _stream_ptr(None)), lib)
    return int(out.item())


def _dev_f64(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.float64).contiguous()
    a = np.ascontiguousarray(x, dtype=np.float64)
    return torch.from_numpy(a if a.flags.writeable else a.copy()).to(dev)


def _dev_i32(x, dev):
    if isinstance(x, torch.Tensor):
        return x.to(device=dev, dtype=torch.int32).contiguous()
    return torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).to(dev)


def _resolve(model, model_arg=0):
    """-> (model id, sodm_model_info, plugin or None).  `model` is a registered functor name or a `plugin.Plugin`
    (a generated model in its own library, include/sodm_plugin.h)."""
    if isinstance(model, str):
        mid = _ffi.model_id(model)
        return mid, _ffi.model_info(mid, model_arg), None
    return 0, model.info, model


def _stream_ptr(stream):
    if stream is None:
        stream = torch.cuda.current_stream()
    return C.c_void_p(stream.cuda_stream)


def _pack_var(var, n_traj, dev):
    """{flat parameter element: values[n_traj]} -> (index[int32 host], values[n_var, n_traj] device)."""
    if not var:
        return None, None
    if isinstance(var, tuple):                      # already packed: (flat element indices, values[n_var, n_traj])
        idx = np.ascontiguousarray(var[0], dtype=np.int32)
        vals = _dev_f64(var[1], dev)
        if vals.shape != (len(idx), n_traj):
            raise ValueError(f"packed per-trajectory parameters must be [{len(idx)}, {n_traj}], got {tuple(vals.shape)}")
        return idx, vals
    keys = sorted(var)
    rows = []
    for k in keys:
        v = _dev_f64(var[k], dev).reshape(-1)
        if v.numel() != n_traj:
            raise ValueError(f"per-trajectory parameter element {k}: expected {n_traj} values, got {v.numel()}")
        rows.append(v)
    return np.asarray(keys, dtype=np.int32), torch.stack(rows).contiguous()


def jump_run(model, y0, par, n_traj, t0, n_out, tau=0.5, method="tau_leap", var=None, tdpf=None,
             compat=_ffi.COMPAT_REFERENCE, seed=0, traj_id0=0, out_dtype=torch.float64, state_sel=None,
             out=None, status=None, max_events=0, dev=None, stream=None, out_peers=None):
    """Run n_traj jump-process trajectories (`sodm_jump_run`).

    model: registered functor name.  y0: [n_state] (shared) or [n_state, n_traj].  par: [n_param] host values.
    var: {flat parameter element: per-trajectory values}.  tdpf: (table[n_values, tdpf_count], step_index[n_steps]).
    out_peers: raw device pointers (ints) of peer GPUs' buffers with the layout of `out` (`sharding.PeerEnsemble`): the kernel
    stores every output element there as well -- the gather of the shards over NVLink happens inside the simulation.
    Returns (out[n_rows, n_out, n_traj] device tensor, status[n_traj] int32 device tensor); asynchronous.
    """
    dev = dev or device()
    lib = _ffi.load()
    mid, info, plug = _resolve(model)
    out_dtype = as_torch_dtype(out_dtype)
    y0_d = _dev_f64(y0, dev)
    if y0_d.shape[0] != info.n_state or (y0_d.dim() == 2 and y0_d.shape[1] != n_traj):
        raise ValueError(f"y0 must be [{info.n_state}] or [{info.n_state}, {n_traj}], got {tuple(y0_d.shape)}")
    par_h = np.ascontiguousarray(par, dtype=np.float64)
    if par_h.size != info.n_param:
        raise ValueError(f"{model}: expected {info.n_param} flat parameter values, got {par_h.size}")
    var_idx, var_vals = _pack_var(var, n_traj, dev)
    n_rows = info.n_state if state_sel is None else len(state_sel)
    if out is None:
        out = torch.empty((n_rows, n_out, n_traj), dtype=out_dtype, device=dev)
    if status is None:
        status = torch.zeros(n_traj, dtype=torch.int32, device=dev)
    a = _ffi.JumpArgs()
    a.model = mid
    a.method = {"tau_leap": _ffi.JUMP_TAU_LEAP, "SSA": _ffi.JUMP_SSA}[method]
    a.compat, a.out_dtype = int(compat), _TORCH_DT[out.dtype]
    a.n_traj, a.traj_id0, a.seed = int(n_traj), int(traj_id0), int(seed) & 0xFFFFFFFFFFFFFFFF
    a.ld_in, a.ld_out = int(n_traj), int(out.stride(1))
    a.y0, a.y0_broadcast = y0_d.data_ptr(), int(y0_d.dim() == 1)
    a.par = par_h.ctypes.data
    if var_idx is not None:
        a.n_var, a.var_index, a.var_values = len(var_idx), var_idx.ctypes.data, var_vals.data_ptr()
    keep = [y0_d, par_h, var_idx, var_vals]
    if tdpf is not None:
        table = tdpf[0] if isinstance(tdpf[0], torch.Tensor) else np.asarray(tdpf[0], dtype=np.float64)
        table = _dev_f64(table.reshape(len(table), -1), dev)
        index = _dev_i32(tdpf[1], dev)
        a.tdpf_table, a.tdpf_index = table.data_ptr(), index.data_ptr()
        a.n_tdpf_values, a.n_tdpf_index = int(table.shape[0]), int(index.numel())
        keep += [table, index]
    a.t0, a.n_out, a.tau = float(t0), int(n_out), float(tau)
    if state_sel is not None:
        sel = np.ascontiguousarray(state_sel, dtype=np.int32)
        a.n_sel, a.state_sel = int(sel.size), sel.ctypes.data
        keep.append(sel)
    a.out, a.status, a.max_events = out.data_ptr(), status.data_ptr(), int(max_events)
    a.stream = _stream_ptr(stream)
    if out_peers:
        if len(out_peers) > _ffi.MAX_PEERS:
            raise ValueError(f"at most {_ffi.MAX_PEERS} peer destinations, got {len(out_peers)}")
        peers = (C.c_void_p * len(out_peers))(*[int(p) for p in out_peers])
        a.out_peers, a.n_out_peers = C.cast(peers, C.c_void_p), len(out_peers)
        keep.append(peers)
    if plug is not None:
        plug.jump_run(a)
    else:
        _ffi.check(lib.sodm_jump_run(C.byref(a)), lib)
    # device inputs are only read by work already queued on the stream; torch's caching allocator keeps
    # stream-ordered reuse safe on the current stream, so dropping `keep` here is fine
    return out, status


def ode_run(model, y0, par, n_traj, t0, n_out, method="RK45", rtol=1e-4, atol=1e-6, tau=0.0, var=None,
            model_arg=0, out_stride=1, out_last_only=False, out_dtype=torch.float64, out=None, dev=None,
            stream=None, max_steps=0):
    """Integrate n_traj parameter draws of an ODE model (`sodm_ode_run`).

    Returns (out[n_state, n_kept, n_traj], nfev[n_traj], status[n_traj]) device tensors; asynchronous."""
    dev = dev or device()
    lib = _ffi.load()
    mid, info, plug = _resolve(model, model_arg)
    out_dtype = as_torch_dtype(out_dtype)
    y0_d = _dev_f64(y0, dev)
    if y0_d.shape[0] != info.n_state or (y0_d.dim() == 2 and y0_d.shape[1] != n_traj):
        raise ValueError(f"y0 must be [{info.n_state}] or [{info.n_state}, {n_traj}], got {tuple(y0_d.shape)}")
    par_h = np.ascontiguousarray(par, dtype=np.float64)
    if par_h.size != info.n_param:
        raise ValueError(f"{model}: expected {info.n_param} flat parameter values, got {par_h.size}")
    var_idx, var_vals = _pack_var(var, n_traj, dev)
    n_kept = int(lib.sodm_ode_n_kept(int(n_out), int(out_stride), int(bool(out_last_only))))
    if out is None:
        out = torch.empty((info.n_state, n_kept, n_traj), dtype=out_dtype, device=dev)
    nfev = torch.zeros(n_traj, dtype=torch.int32, device=dev)
    status = torch.zeros(n_traj, dtype=torch.int32, device=dev)
    a = _ffi.OdeArgs()
    a.model, a.model_arg = mid, int(model_arg)
    a.method = {"RK45": _ffi.ODE_RK45, "euler": _ffi.ODE_EULER, "RK4": _ffi.ODE_RK4, "RK23": _ffi.ODE_RK23, "DOP853": _ffi.ODE_DOP853, "Radau": _ffi.ODE_RADAU, "BDF": _ffi.ODE_BDF}[method]
    a.out_dtype = _TORCH_DT[out.dtype]
    a.n_traj, a.ld_in, a.ld_out = int(n_traj), int(n_traj), int(out.stride(1))
    a.y0, a.y0_broadcast = y0_d.data_ptr(), int(y0_d.dim() == 1)
    a.par = par_h.ctypes.data
    if var_idx is not None:
        a.n_var, a.var_index, a.var_values = len(var_idx), var_idx.ctypes.data, var_vals.data_ptr()
    a.t0, a.n_out, a.out_stride, a.out_last_only = float(t0), int(n_out), int(out_stride), int(bool(out_last_only))
    a.rtol, a.atol, a.tau = float(rtol), float(atol), float(tau or 0.0)
    a.out, a.nfev, a.status = out.data_ptr(), nfev.data_ptr(), status.data_ptr()
    a.max_steps = int(max_steps)
    a.stream = _stream_ptr(stream)
    if plug is not None:
        plug.ode_run(a)
    else:
        _ffi.check(lib.sodm_ode_run(C.byref(a)), lib)
    return out, nfev, status


_LL_KIND = {"poisson": _ffi.LL_POISSON, "normal": _ffi.LL_NORMAL, "lognormal": _ffi.LL_LOGNORMAL,
            "negative_binomial": _ffi.LL_NEGBIN}


def loglik(kind, traj, t0, columns, t_data, ydata, ll_arg=None, weight=1.0, lp=None, status=None, dev=None,
           stream=None, column_weights=None):
    """Accumulate the log-likelihood of every trajectory against one dataset (`sodm_loglik`).

    traj: device tensor [n_row, n_out, n_traj] (any of int32/float32/float64).
    columns: list (one entry per data column) of lists of rows of `traj` that are summed into that column.
    t_data: [n_t] data times in simulation time units.  ydata: [n_t, n_col].
    ll_arg: None | scalar | [n_col] | [n_t, n_col].  column_weights: per column, the weight of each of its rows (a linear
    aggregation function folded into the column sums); None = plain sums.
    Returns lp[n_traj] (float64 device tensor, accumulated)."""
    dev = dev or device()
    lib = _ffi.load()
    n_row, n_out, n_traj = traj.shape
    ydata_h = np.ascontiguousarray(ydata, dtype=np.float64)
    if ydata_h.ndim == 1:
        ydata_h = ydata_h[:, None]
    n_t, n_col = ydata_h.shape
    if len(columns) != n_col:
        raise ValueError(f"{len(columns)} column definitions for {n_col} data columns")
    col_ptr = np.zeros(n_col + 1, dtype=np.int32)
    col_ptr[1:] = np.cumsum([len(c) for c in columns])
    col_rows = np.asarray([r for c in columns for r in c], dtype=np.int32)
    if col_rows.size and (col_rows.min() < 0 or col_rows.max() >= n_row):
        raise ValueError("column row index outside the trajectory tensor")
    t_data_h = np.ascontiguousarray(t_data, dtype=np.float64)
    if t_data_h.shape != (n_t,):
        raise ValueError("t_data must have one entry per data row")
    a = _ffi.LoglikArgs()
    a.kind = _LL_KIND[kind]
    a.traj_dtype, a.n_traj, a.ld = _TORCH_DT[traj.dtype], int(n_traj), int(traj.stride(1))
    a.traj, a.n_out, a.n_col = traj.data_ptr(), int(n_out), int(n_col)
    d_ptr, d_rows = _dev_i32(col_ptr, dev), _dev_i32(col_rows, dev)
    d_t, d_y = _dev_f64(t_data_h, dev), _dev_f64(ydata_h, dev)
    a.col_ptr, a.col_rows, a.t_data, a.ydata = d_ptr.data_ptr(), d_rows.data_ptr(), d_t.data_ptr(), d_y.data_ptr()
    d_w = None
    if column_weights is not None:
        w = np.asarray([x for c in column_weights for x in c], dtype=np.float64)
        if w.shape != col_rows.shape:
            raise ValueError("column_weights must mirror the structure of columns")
        d_w = _dev_f64(w, dev)
        a.col_weights = d_w.data_ptr()
    a.t0, a.n_t = float(t0), int(n_t)
    d_arg = None
    if ll_arg is None:
        a.arg_mode = _ffi.ARG_NONE
    else:
        arg = np.asarray(ll_arg, dtype=np.float64)
        if arg.ndim == 0 or arg.size == 1 and arg.ndim <= 1 and n_col != 1:
            a.arg_mode, arg = _ffi.ARG_SCALAR, arg.reshape(1)
        elif arg.shape == (n_col,):
            a.arg_mode = _ffi.ARG_COLUMN
        elif arg.shape == (n_t, n_col):
            a.arg_mode = _ffi.ARG_FULL
        else:
            raise ValueError(f"ll_arg of shape {arg.shape} does not broadcast to data of shape {(n_t, n_col)}")
        d_arg = _dev_f64(arg, dev)
        a.ll_arg = d_arg.data_ptr()
    a.weight = float(weight)
    if lp is None:
        lp = torch.zeros(n_traj, dtype=torch.float64, device=dev)
    if status is None:
        status = torch.zeros(n_traj, dtype=torch.int32, device=dev)
    a.lp, a.status = lp.data_ptr(), status.data_ptr()
    a.stream = _stream_ptr(stream)
    _ffi.check(lib.sodm_loglik(C.byref(a)), lib)
    return lp, status


def to_draws_major(src, r0, n_dim, dst=None, dst_dtype=torch.float64, dev=None, stream=None):
    """[row][time][traj] -> [traj][time][n_dim] for the state variable occupying rows r0..r0+n_dim-1
    (`sodm_to_draws_major`).  `dst` may be a CUDA tensor or a pinned host tensor (written over PCIe)."""
    dev = dev or device()
    lib = _ffi.load()
    n_row, n_t, n_traj = src.shape
    if dst is None:
        dst = torch.empty((n_traj, n_t, n_dim), dtype=as_torch_dtype(dst_dtype), device=dev)
    if not dst.is_contiguous() or dst.numel() != n_traj * n_t * n_dim:
        raise ValueError("dst must be contiguous with n_traj * n_t * n_dim elements")
    if not dst.is_cuda and not dst.is_pinned():
        raise ValueError("dst must be a CUDA tensor or a pinned host tensor")
    _ffi.check(lib.sodm_to_draws_major(src.data_ptr(), _TORCH_DT[src.dtype], int(src.stride(1)), int(r0), int(n_dim),
                                       int(n_t), int(n_traj), dst.data_ptr(), _TORCH_DT[dst.dtype],
                                       _stream_ptr(stream)), lib)
    return dst


# ---- ensemble post-processing (sodm_noise / sodm_draws_sum / sodm_draws_order_stats) ------------------------------
_NOISE_KINDS = {"poisson": _ffi.NOISE_POISSON, "gaussian": _ffi.NOISE_GAUSSIAN_ABS,
                "gaussian_relative": _ffi.NOISE_GAUSSIAN_REL, "negative_binomial": _ffi.NOISE_NEGBIN}


def _as_cells(ens):
    """View a [..., n_traj] CUDA tensor with contiguous leading axes as (n_cells, n_traj, ld)."""
    if not (isinstance(ens, torch.Tensor) and ens.is_cuda):
        raise TypeError("expected a CUDA tensor laid out [..., trajectories]")
    if ens.dtype not in _TORCH_DT:
        raise TypeError(f"dtype {ens.dtype} is not one of int32 / float32 / float64")
    if ens.dim() < 1 or (ens.shape[-1] > 1 and ens.stride(-1) != 1):
        raise ValueError("trajectories must be the last, unit-stride axis")
    n_traj = int(ens.shape[-1])
    n_cells = int(np.prod(ens.shape[:-1])) if ens.dim() > 1 else 1
    ld = int(ens.stride(-2)) if ens.dim() > 1 else n_traj
    expect = ld
    for ax in range(ens.dim() - 2, -1, -1):                       # leading axes must be packed with pitch ld
        if ens.shape[ax] > 1 and ens.stride(ax) != expect:
            raise ValueError("leading axes of the ensemble must be contiguous")
        expect *= int(ens.shape[ax])
    return n_cells, n_traj, ld


def noise(kind, ens, param=0.0, seed=0, traj_id0=0, out=None, out_dtype=None, dev=None, stream=None):
    """Replace every value of the ensemble `ens` [..., n_traj] by a noisy draw around it (`sodm_noise`).
    kind: 'poisson' | 'gaussian' (param = sigma) | 'gaussian_relative' | 'negative_binomial' (param = alpha).
    Returns (out, status) -- status is a one-element int32 CUDA tensor of SODM_NOISE_ST_* bits."""
    dev = dev or device()
    lib = _ffi.load()
    n_cells, n_traj, ld = _as_cells(ens)
    if out is None:
        dt = as_torch_dtype(out_dtype) if out_dtype is not None else (
            torch.float64 if kind.startswith("gaussian") and ens.dtype == torch.int32 else ens.dtype)
        out = torch.empty_strided(ens.shape, ens.stride(), dtype=dt, device=dev)
    elif out.shape != ens.shape or out.stride() != ens.stride():
        raise ValueError("out must have the shape and strides of the input ensemble")
    status = torch.zeros(1, dtype=torch.int32, device=dev)
    a = _ffi.NoiseArgs()
    a.kind, a.param = _NOISE_KINDS[kind], float(param)
    a.in_, a.in_dtype = ens.data_ptr(), _TORCH_DT[ens.dtype]
    a.out, a.out_dtype = out.data_ptr(), _TORCH_DT[out.dtype]
    a.n_cells, a.n_traj, a.ld = n_cells, n_traj, ld
    a.traj_id0, a.seed = int(traj_id0), int(seed) & 0xFFFFFFFFFFFFFFFF
    a.status, a.stream = status.data_ptr(), _stream_ptr(stream)
    _ffi.check(lib.sodm_noise(C.byref(a)), lib)
    return out, status


def draws_sum(ens, dev=None, stream=None):
    """Sum over the trajectory axis of `ens` [..., n_traj] -> float64 CUDA tensor [...] (`sodm_draws_sum`)."""
    dev = dev or device()
    lib = _ffi.load()
    n_cells, n_traj, ld = _as_cells(ens)
    out = torch.empty(ens.shape[:-1], dtype=torch.float64, device=dev)
    _ffi.check(lib.sodm_draws_sum(ens.data_ptr(), _TORCH_DT[ens.dtype], n_cells, n_traj, ld, out.data_ptr(),
                                  _stream_ptr(stream)), lib)
    return out


def draws_order_stats(ens, ranks, dev=None, stream=None):
    """The ranks[j]-th smallest value (0-based) over the trajectory axis of `ens` [..., n_traj]
    -> float64 CUDA tensor [len(ranks), ...] (`sodm_draws_order_stats`, at most MAX_RANKS ranks per launch)."""
    dev = dev or device()
    lib = _ffi.load()
    n_cells, n_traj, ld = _as_cells(ens)
    ranks = [int(r) for r in ranks]
    out = torch.empty((len(ranks),) + tuple(ens.shape[:-1]), dtype=torch.float64, device=dev)
    for lo in range(0, len(ranks), _ffi.MAX_RANKS):
        part = ranks[lo:lo + _ffi.MAX_RANKS]
        arr = (C.c_int64 * len(part))(*part)
        _ffi.check(lib.sodm_draws_order_stats(ens.data_ptr(), _TORCH_DT[ens.dtype], n_cells, n_traj, ld, arr, len(part),
                                              out[lo:lo + len(part)].data_ptr(), _stream_ptr(stream)), lib)
    return out


class SelectPhases:
    """Order statistics of an ensemble SHARDED by trajectory (`sodm_draws_select_phase`): this object owns the per-cell
    state of one search; `sharding.order_stats_sharded` drives the phases and all-reduces the tables in between."""

    def __init__(self, ens, global_ranks, dev=None):
        self.dev = dev or device()
        self.lib = _ffi.load()
        self.ens = ens
        self.n_cells, self.n_traj, self.ld = _as_cells(ens) if ens.shape[-1] > 0 else (int(np.prod(ens.shape[:-1])), 0, 1)
        self.ranks = [int(r) for r in global_ranks]
        if not 1 <= len(self.ranks) <= _ffi.MAX_RANKS:
            raise ValueError(f"between 1 and {_ffi.MAX_RANKS} ranks per search")
        J, nc = len(self.ranks), self.n_cells
        i64 = dict(dtype=torch.int64, device=self.dev)
        self.lo, self.hi = torch.empty(nc, **i64), torch.empty(nc, **i64)
        self.nan_flag = torch.zeros(nc, dtype=torch.int32, device=self.dev)
        self.settled = torch.zeros(nc, dtype=torch.int32, device=self.dev)
        self.prefix = torch.zeros((2, nc, J), **i64)
        self.remaining = torch.tensor(self.ranks, **i64).repeat(nc, 1).contiguous()
        self.hist = torch.zeros((nc, J, _ffi.SELECT_BINS), dtype=torch.int32, device=self.dev)
        self.state = _ffi.SelectState(self.lo.data_ptr(), self.hi.data_ptr(), self.nan_flag.data_ptr(), self.settled.data_ptr(),
                                      self.prefix.data_ptr(), self.remaining.data_ptr())
        self.key_bits = 64 if ens.dtype == torch.float64 else 32

    def run(self, phase, out=None, stream=None):
        _ffi.check(self.lib.sodm_draws_select_phase(
            int(phase), self.ens.data_ptr() if self.n_traj else None, _TORCH_DT[self.ens.dtype], self.n_cells, self.n_traj,
            self.ld, len(self.ranks), C.byref(self.state), self.hist.data_ptr(), out.data_ptr() if out is not None else None,
            _stream_ptr(stream)), self.lib)

    def result(self):
        out = torch.empty((len(self.ranks),) + tuple(self.ens.shape[:-1]), dtype=torch.float64, device=self.dev)
        self.run(_ffi.SELECT_DECODE, out=out)
        return out


# ---- optimiser populations (sodm_pso_move / sodm_pso_best / sodm_kde_logpdf) --------------------------------------------
# Populations are float64 CUDA tensors [n_dim, n] (structure-of-arrays, the layout of per-trajectory parameters).

def _population(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.dim() == 2 and
            (t.shape[1] <= 1 or t.stride(1) == 1)):
        raise TypeError(f"{what} must be a float64 CUDA tensor [n_dim, n] with unit stride along n")
    if not 1 <= t.shape[0] <= _ffi.POPULATION_MAX_DIM:
        raise ValueError(f"{what}: between 1 and {_ffi.POPULATION_MAX_DIM} dimensions, got {t.shape[0]}")
    return t


def _ld(t):
    """Leading stride of a population (torch reports arbitrary strides for dimensions of size 1)."""
    return int(t.stride(0)) if t.shape[0] > 1 else max(int(t.shape[1]), 1)


def pso_move(x, v, p, g, lb, ub, omega, phip, phig, seed, iteration, member_id0=0, stream=None):
    """One swarm update in place (`sodm_pso_move`): x, v, p [n_dim, n]; g [n_dim] on the device; lb, ub host values."""
    lib = _ffi.load()
    x, v, p = _population(x, "x"), _population(v, "v"), _population(p, "p")
    if not (x.shape == v.shape == p.shape and _ld(x) == _ld(v) == _ld(p)):
        raise ValueError("x, v and p must share shape and leading stride")
    lb_h, ub_h = np.ascontiguousarray(lb, dtype=np.float64), np.ascontiguousarray(ub, dtype=np.float64)
    if lb_h.shape != (x.shape[0],) or ub_h.shape != (x.shape[0],) or g.shape != (x.shape[0],) or g.dtype != torch.float64:
        raise ValueError("g, lb and ub need one float64 value per dimension")
    a = _ffi.PsoArgs()
    a.n_dim, a.n, a.ld = int(x.shape[0]), int(x.shape[1]), _ld(x)
    a.x, a.v, a.p, a.g = x.data_ptr(), v.data_ptr(), p.data_ptr(), g.data_ptr()
    a.lb, a.ub = lb_h.ctypes.data, ub_h.ctypes.data
    a.omega, a.phip, a.phig = float(omega), float(phip), float(phig)
    a.seed, a.iteration, a.member_id0 = int(seed) & 0xFFFFFFFFFFFFFFFF, int(iteration), int(member_id0)
    a.stream = _stream_ptr(stream)
    _ffi.check(lib.sodm_pso_move(C.byref(a)), lib)


def pso_best(x, fx, p, fp, feasible=None, best=None, best_index=None, stream=None):
    """Personal bests in place and the swarm's best (`sodm_pso_best`) -> (best[1 + n_dim] = (fp[i*], p[:, i*]),
    best_index[1]) device tensors; asynchronous."""
    lib = _ffi.load()
    x, p = _population(x, "x"), _population(p, "p")
    n_dim, n = x.shape
    if p.shape != x.shape or _ld(p) != _ld(x):
        raise ValueError("x and p must share shape and leading stride")
    for name, t in (("fx", fx), ("fp", fp)):
        if not (t.is_cuda and t.dtype == torch.float64 and t.shape == (n,) and t.is_contiguous()):
            raise TypeError(f"{name} must be a contiguous float64 CUDA tensor [{n}]")
    if feasible is not None and not (feasible.is_cuda and feasible.dtype in (torch.uint8, torch.bool) and
                                     feasible.shape == (n,) and feasible.is_contiguous()):
        raise TypeError(f"feasible must be a contiguous bool / uint8 CUDA tensor [{n}]")
    if best is None:
        best = torch.empty(1 + n_dim, dtype=torch.float64, device=x.device)
    if best_index is None:
        best_index = torch.empty(1, dtype=torch.int64, device=x.device)
    _ffi.check(lib.sodm_pso_best(n_dim, n, _ld(x), x.data_ptr(), fx.data_ptr(),
                                 feasible.data_ptr() if feasible is not None else None, p.data_ptr(), fp.data_ptr(),
                                 best.data_ptr(), best_index.data_ptr(), _stream_ptr(stream)), lib)
    return best, best_index


def _index_list(t, what):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.int64 and t.dim() == 1 and t.is_contiguous()):
        raise TypeError(f"{what} must be a contiguous int64 CUDA tensor")
    return t


MOVE_KINDS = {"stretch": _ffi.MOVE_STRETCH, "de": _ffi.MOVE_DE, "snooker": _ffi.MOVE_SNOOKER}


def walkers_propose_kde(x, mine, centres, chol, seed=0, iteration=0, stream=None):
    """KDE move on the device: proposals of the walkers `mine` drawn from the Gaussian KDE of the walkers `centres` with
    kernel covariance chol chol^T (`sodm_walkers_propose`, SODM_MOVE_KDE), and the log Hastings factor
    density(s) - density(q) from two pair sums over the whitened coordinates (`sodm_kde_logpdf`)
    -> (q [n_dim, n_mine], log_factor [n_mine]) device tensors; asynchronous."""
    lib = _ffi.load()
    x, mine, centres = _population(x, "x"), _index_list(mine, "mine"), _index_list(centres, "centres")
    n_dim, n_mine, K = int(x.shape[0]), int(mine.shape[0]), int(centres.shape[0])
    L = np.ascontiguousarray(chol, dtype=np.float64)
    if L.shape != (n_dim, n_dim):
        raise ValueError(f"chol must be [{n_dim}, {n_dim}]")
    q = torch.empty((n_dim, n_mine), dtype=torch.float64, device=x.device)
    white = torch.empty((n_dim, K + 2 * n_mine), dtype=torch.float64, device=x.device)
    a = _ffi.WalkerMoveArgs()
    a.kind, a.n_dim, a.n_walkers, a.ld = _ffi.MOVE_KDE, n_dim, int(x.shape[1]), _ld(x)
    a.x, a.mine, a.n_mine, a.others, a.n_others = x.data_ptr(), mine.data_ptr(), n_mine, centres.data_ptr(), K
    a.q, a.ld_q, a.log_factor = q.data_ptr(), _ld(q), None
    a.chol, a.white, a.ld_white = L.ctypes.data, white.data_ptr(), _ld(white)
    a.seed, a.iteration, a.stream = int(seed) & 0xFFFFFFFFFFFFFFFF, int(iteration), _stream_ptr(stream)
    _ffi.check(lib.sodm_walkers_propose(C.byref(a)), lib)
    if n_mine == 0:
        return q, torch.empty(0, dtype=torch.float64, device=x.device)
    dens = kde_logpdf(white[:, :K], white[:, K:], stream=stream)
    return q, dens[:n_mine] - dens[n_mine:]


def walkers_propose(kind, x, mine, others, param, sigma=0.0, seed=0, iteration=0, q=None, log_factor=None, stream=None):
    """Proposals of the walkers `mine` from the walkers `others` of the ensemble x [n_dim, n_walkers] (`sodm_walkers_propose`;
    kind 'stretch' | 'de' | 'snooker') -> (q [n_dim, n_mine], log_factor [n_mine]) device tensors; asynchronous."""
    lib = _ffi.load()
    x, mine, others = _population(x, "x"), _index_list(mine, "mine"), _index_list(others, "others")
    n_dim, n_mine = int(x.shape[0]), int(mine.shape[0])
    if q is None:
        q = torch.empty((n_dim, n_mine), dtype=torch.float64, device=x.device)
    if log_factor is None:
        log_factor = torch.empty(n_mine, dtype=torch.float64, device=x.device)
    q = _population(q, "q")
    if q.shape != (n_dim, n_mine) or log_factor.shape != (n_mine,) or log_factor.dtype != torch.float64 or not log_factor.is_contiguous():
        raise ValueError("q must be [n_dim, n_mine] and log_factor a contiguous float64 tensor [n_mine]")
    a = _ffi.WalkerMoveArgs()
    a.kind = MOVE_KINDS[kind] if isinstance(kind, str) else int(kind)
    a.n_dim, a.n_walkers, a.ld = n_dim, int(x.shape[1]), _ld(x)
    a.x, a.mine, a.n_mine, a.others, a.n_others = x.data_ptr(), mine.data_ptr(), n_mine, others.data_ptr(), int(others.shape[0])
    a.param, a.sigma = float(param), float(sigma)
    a.q, a.ld_q, a.log_factor = q.data_ptr(), _ld(q), log_factor.data_ptr()
    a.seed, a.iteration, a.stream = int(seed) & 0xFFFFFFFFFFFFFFFF, int(iteration), _stream_ptr(stream)
    _ffi.check(lib.sodm_walkers_propose(C.byref(a)), lib)
    return q, log_factor


def population_clip(x, lb, ub, penalty=None, stream=None):
    """Clip x [n_dim, n] into [lb, ub] (host values) in place (`sodm_population_clip`) -> penalty [n]: -inf where a value
    was clipped, else 0 (the reference objective's treatment of an out-of-bounds theta)."""
    lib = _ffi.load()
    x = _population(x, "x")
    lb_h, ub_h = np.ascontiguousarray(lb, dtype=np.float64), np.ascontiguousarray(ub, dtype=np.float64)
    if lb_h.shape != (x.shape[0],) or ub_h.shape != (x.shape[0],):
        raise ValueError("lb and ub need one value per dimension")
    if penalty is None:
        penalty = torch.empty(x.shape[1], dtype=torch.float64, device=x.device)
    _ffi.check(lib.sodm_population_clip(int(x.shape[0]), int(x.shape[1]), _ld(x), x.data_ptr(), lb_h.ctypes.data, ub_h.ctypes.data,
                                        penalty.data_ptr(), _stream_ptr(stream)), lib)
    return penalty


def walkers_accept(x, lp, mine, q, new_lp, log_factor, accepted, penalty=None, seed=0, iteration=0, status=None, stream=None):
    """Metropolis-Hastings test of the proposals q [n_dim, n_mine] of the walkers `mine` (`sodm_walkers_accept`): x, lp and
    accepted are updated in place; status (int32 [1]) collects SODM_WALKERS_ST_NAN."""
    lib = _ffi.load()
    x, q, mine = _population(x, "x"), _population(q, "q"), _index_list(mine, "mine")
    n_dim, n, n_mine = int(x.shape[0]), int(x.shape[1]), int(mine.shape[0])
    if q.shape != (n_dim, n_mine):
        raise ValueError("q must be [n_dim, n_mine]")
    for name, t, m in (("lp", lp, n), ("new_lp", new_lp, n_mine), ("log_factor", log_factor, n_mine), ("penalty", penalty, n_mine)):
        if t is not None and not (t.is_cuda and t.dtype == torch.float64 and t.shape == (m,) and t.is_contiguous()):
            raise TypeError(f"{name} must be a contiguous float64 CUDA tensor [{m}]")
    if not (accepted.is_cuda and accepted.dtype == torch.int64 and accepted.shape == (n,) and accepted.is_contiguous()):
        raise TypeError(f"accepted must be a contiguous int64 CUDA tensor [{n}]")
    if status is not None and not (status.is_cuda and status.dtype == torch.int32 and status.numel() >= 1):
        raise TypeError("status must be an int32 CUDA tensor")
    _ffi.check(lib.sodm_walkers_accept(n_dim, n, _ld(x), x.data_ptr(), lp.data_ptr(), mine.data_ptr(), n_mine, q.data_ptr(), _ld(q),
                                       new_lp.data_ptr(), penalty.data_ptr() if penalty is not None else None,
                                       log_factor.data_ptr(), int(seed) & 0xFFFFFFFFFFFFFFFF, int(iteration),
                                       accepted.data_ptr(), status.data_ptr() if status is not None else None,
                                       _stream_ptr(stream)), lib)


def kde_logpdf(points, queries, log_norm=0.0, out=None, stream=None):
    """log sum_k exp(-|q_j - x_k|^2 / 2) + log_norm for whitened points [n_dim, K] and queries [n_dim, M]
    (`sodm_kde_logpdf`) -> float64 CUDA tensor [M]."""
    lib = _ffi.load()
    points, queries = _population(points, "points"), _population(queries, "queries")
    if points.shape[0] != queries.shape[0]:
        raise ValueError("points and queries differ in dimension")
    if out is None:
        out = torch.empty(queries.shape[1], dtype=torch.float64, device=queries.device)
    _ffi.check(lib.sodm_kde_logpdf(int(points.shape[0]), points.data_ptr(), int(points.shape[1]), _ld(points),
                                   queries.data_ptr(), int(queries.shape[1]), _ld(queries), float(log_norm),
                                   out.data_ptr(), _stream_ptr(stream)), lib)
    return out


def gaussian_kde_logpdf(data, queries, bw_method="scott", dev=None):
    """`scipy.stats.gaussian_kde(data.T, bw_method).logpdf(queries.T)` for host arrays data[K, D], queries[M, D]: the
    D x D factorisation and the whitening on the host, the K x M pair sum on the device."""
    dev = dev or device()
    data, queries = np.asarray(data, dtype=np.float64), np.asarray(queries, dtype=np.float64)
    K, D = data.shape
    if bw_method in (None, "scott"):
        factor = K ** (-1.0 / (D + 4))
    elif bw_method == "silverman":
        factor = (K * (D + 2) / 4.0) ** (-1.0 / (D + 4))
    elif np.isscalar(bw_method):
        factor = float(bw_method)
    else:
        raise ValueError("bw_method must be 'scott', 'silverman' or a number")
    cov = np.atleast_2d(np.cov(data, rowvar=False, bias=False)) * factor ** 2
    L = np.linalg.cholesky(cov)
    from scipy.linalg import solve_triangular
    pts = solve_triangular(L, data.T, lower=True)                   # [D, K]
    qs = solve_triangular(L, queries.T, lower=True)                 # [D, M]
    log_norm = -np.log(K) - np.sum(np.log(np.diag(L))) - 0.5 * D * np.log(2.0 * np.pi)
    out = kde_logpdf(torch.from_numpy(np.ascontiguousarray(pts)).to(dev), torch.from_numpy(np.ascontiguousarray(qs)).to(dev),
                     log_norm)
    return out.cpu().numpy()
