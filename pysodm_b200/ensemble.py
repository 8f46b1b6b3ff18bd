"""An ensemble that stays in HBM: `sim(..., output='device')` returns a `DeviceEnsemble` instead of copying
`(draws, time, *dims)` arrays to the host.

The reference's workflows post-process a finished ensemble right away -- observation noise, then the mean and a few
quantiles over `draws` (`/root/reference/tutorials/influenza_1718/calibration.py:231,245-246`,
`tutorials/SIR/workflow_tutorial.py:180-185`, `tutorials/enzyme_kinetics/simulate_1D_PFR.py:128-137`).  For 10^6
influenza trajectories the ensemble is 18 GB, the post-processed result 0.1 MB: doing the reduction where the data is
turns a PCIe-bound workflow (0.33 s of DMA) into three HBM-bound kernels (`sodm_noise`, `sodm_draws_sum`,
`sodm_draws_order_stats`).

A `DeviceEnsemble` mirrors the part of `xarray.Dataset` those workflows touch: `ens[state]` (a `DeviceDataArray`),
`.sel()/.isel()` on time and state dimensions, `.mean(dim='draws')`, `.quantile(q, dim='draws')` -- which return small
HOST Datasets / DataArrays with the reference's dims and coords -- and `.to_dataset()` for the full copy.
Device layout of one state variable: tensor `[*state_dims, time, draws]` (a view of the simulator's
`[row][time][trajectory]` buffer), so trajectories are the unit-stride axis of every kernel.
"""
from collections import OrderedDict

import numpy as np
import torch

from . import engine
from .dataset import SimpleDataArray, SimpleDataset, _xr


def _bracketing_ranks(n, q):
    """np.quantile(method='linear'): virtual index (n - 1) q, neighbours floor / floor + 1 (both the last element once
    the index reaches n - 1), weight = fractional part  (numpy/lib/_function_base_impl.py `_get_indexes`, `_get_gamma`)."""
    virtual = (n - 1) * float(q)
    lo = int(np.floor(virtual))
    hi = lo + 1
    if virtual >= n - 1:
        lo = hi = n - 1
    return lo, hi, virtual - np.floor(virtual)


def _lerp(a, b, t):
    """NumPy's `_lerp`: a + (b - a) t, replaced by b - (b - a)(1 - t) where t >= 0.5."""
    diff = b - a
    out = a + diff * t
    return np.where(t >= 0.5, b - diff * (1 - t), out)


class DeviceDataArray:
    """One state variable of a device-resident ensemble: CUDA tensor `[*state_dims, time, draws]`."""

    def __init__(self, data, name, state_dims, coords, time_name, n_global=None, traj_id0=0):
        self.data, self.name = data, name
        self.traj_id0 = int(traj_id0)            # global index of this rank's first draw (keys the noise streams)
        # draws of ALL ranks when the ensemble is sharded by trajectory over a torch.distributed job (else the local count)
        self.n_global = int(n_global) if n_global is not None else int(data.shape[-1])
        self.state_dims = list(state_dims)
        self.time_name = time_name
        self.coords = OrderedDict(coords)
        self._device_dims = self.state_dims + ([time_name] if time_name else []) + ["draws"]
        if data.dim() != len(self._device_dims):
            raise ValueError(f"{name}: tensor of rank {data.dim()} for dims {self._device_dims}")

    # reference order of the dims of the full array: (draws, time, *state_dims)
    @property
    def dims(self):
        return tuple(["draws"] + ([self.time_name] if self.time_name else []) + self.state_dims)

    @property
    def n_draws(self):
        return self.n_global

    @property
    def sharded(self):
        return self.n_global != int(self.data.shape[-1])

    @property
    def shape(self):
        return tuple(self.n_global if d == "draws" else int(self.data.shape[self._device_dims.index(d)]) for d in self.dims)

    def __getitem__(self, key):
        if isinstance(key, str):
            return np.asarray(self.coords[key])
        raise TypeError("index a DeviceDataArray with .isel() / .sel(); use .values for the host copy")

    # ---- selection ---------------------------------------------------------------------------------------
    def _index(self, dim, idx):
        if dim == "draws":
            raise NotImplementedError("selecting individual draws of a device ensemble: use .values")
        if dim not in self._device_dims:
            raise KeyError(f"dimension '{dim}' not in {self.dims}")
        ax = self._device_dims.index(dim)
        scalar = np.ndim(idx) == 0 and not isinstance(idx, slice)
        key = [slice(None)] * self.data.dim()
        key[ax] = (idx if isinstance(idx, slice) or scalar else torch.as_tensor(np.asarray(idx), device=self.data.device))
        data = self.data[tuple(key)]
        coords = OrderedDict(self.coords)
        if dim in coords:
            if isinstance(idx, slice):
                coords[dim] = list(coords[dim])[idx]
            elif scalar:
                coords.pop(dim)
            else:
                labels = list(coords[dim])
                coords[dim] = [labels[i] for i in np.asarray(idx).tolist()]
        state_dims = [d for d in self.state_dims if not (scalar and d == dim)]
        time_name = None if (scalar and dim == self.time_name) else self.time_name
        return DeviceDataArray(data, self.name, state_dims, coords, time_name, self.n_global, self.traj_id0)

    def isel(self, indexers=None, **kw):
        out = self
        for dim, idx in {**(indexers or {}), **kw}.items():
            out = out._index(dim, idx)
        return out

    def sel(self, indexers=None, **kw):
        out = self
        for dim, label in {**(indexers or {}), **kw}.items():
            labels = [_norm(v) for v in out.coords[dim]]
            if isinstance(label, slice):
                lo = labels.index(_norm(label.start)) if label.start is not None else 0
                hi = labels.index(_norm(label.stop)) + 1 if label.stop is not None else len(labels)   # label slices include the stop
                idx = slice(lo, hi)
            elif isinstance(label, (list, tuple, np.ndarray)):
                idx = [labels.index(_norm(v)) for v in label]
            else:
                idx = labels.index(_norm(label))
            out = out._index(dim, idx)
        return out

    # ---- reductions over draws: small host results -------------------------------------------------------
    def _packed(self):
        d = self.data
        try:
            engine._as_cells(d)
            return d
        except ValueError:
            return d.contiguous()

    def _host_result(self, values, lead=()):
        """values: ndarray [*lead, *state_dims, time] -> reference order [*lead, time, *state_dims]."""
        n_lead = len(lead)
        n_state = len(self.state_dims)
        if self.time_name:
            values = np.moveaxis(values, -1, n_lead)
        dims = [name for name, _ in lead] + ([self.time_name] if self.time_name else []) + self.state_dims
        coords = OrderedDict(self.coords)
        for name, labels in lead:
            coords[name] = labels
        assert values.ndim == n_lead + n_state + (1 if self.time_name else 0)
        if _xr is not None:
            return _xr.DataArray(values, dims=dims, coords={k: v for k, v in coords.items() if k in dims}, name=self.name)
        return SimpleDataArray(values, dims, coords, self.name)

    def mean(self, dim="draws"):
        if dim != "draws":
            raise NotImplementedError("device ensembles reduce over 'draws'")
        from . import sharding
        s = sharding.sum_sharded(engine.draws_sum(self._packed()))
        return self._host_result(s.cpu().numpy() / self.n_draws)

    def quantile(self, q, dim="draws"):
        """`xarray.DataArray.quantile(q, dim='draws')`: linear interpolation between order statistics, bit-identical to
        np.quantile on the host copy (the device selects the exact order statistics, the interpolation is NumPy's)."""
        if dim != "draws":
            raise NotImplementedError("device ensembles reduce over 'draws'")
        qs = np.atleast_1d(np.asarray(q, dtype=np.float64))
        if np.any((qs < 0) | (qs > 1)):
            raise ValueError("Quantiles must be in the range [0, 1]")
        n = self.n_draws
        plan = [_bracketing_ranks(n, v) for v in qs]
        ranks = sorted({r for lo, hi, _ in plan for r in (lo, hi)})
        from . import sharding
        stats = sharding.order_stats_sharded(self._packed(), ranks).cpu().numpy()
        where = {r: i for i, r in enumerate(ranks)}
        out = np.stack([_lerp(stats[where[lo]], stats[where[hi]], g) for lo, hi, g in plan])
        if np.ndim(q) == 0:
            return self._host_result(out[0])
        return self._host_result(out, lead=[("quantile", list(qs))])

    # ---- full copy ---------------------------------------------------------------------------------------
    @property
    def values(self):
        """Host ndarray of the LOCAL draws in the reference's order `(draws, time, *state_dims)` (float64), through the layout kernel
        `sodm_to_draws_major` (the transpose the reference does with moveaxis + concat, base.py:1009-1010, :569-571)."""
        n_state = len(self.state_dims)
        shape = [int(x) for x in self.data.shape]
        n_dim = int(np.prod(shape[:n_state])) if n_state else 1
        n_time = shape[n_state] if self.time_name else 1
        src = self.data.reshape(n_dim, n_time, shape[-1]).contiguous()
        host = engine.to_draws_major(src, 0, n_dim).cpu().numpy()                 # (draws, time, n_dim)
        host = host.reshape([shape[-1], n_time] + shape[:n_state])
        return host if self.time_name else host[:, 0]

    def __repr__(self):
        return f"<DeviceDataArray {self.name} {dict(zip(self.dims, self.shape))} on {self.data.device}>"


def _norm(v):
    """Comparable form of a coordinate label (dates may come as str / datetime / numpy datetime64 / pandas Timestamp)."""
    if isinstance(v, str):
        try:
            return np.datetime64(v, "ns")
        except ValueError:
            return v
    if hasattr(v, "to_datetime64"):
        return v.to_datetime64().astype("datetime64[ns]")
    if isinstance(v, np.datetime64) or type(v).__name__ == "datetime":
        return np.datetime64(v, "ns")
    return v


class DeviceEnsemble:
    """All state variables of a device-resident ensemble over ONE buffer `[row][time][draws]`."""

    def __init__(self, buffer, variables, dims_per_state, coords_per_state, time_name, time_coord, traj_id0=0, n_global=None):
        self.buffer = buffer                       # CUDA tensor [n_rows, n_time, n_local_draws]
        self.n_global = int(n_global) if n_global is not None else int(buffer.shape[2])
        self.variables = list(variables)           # pipeline.Variable: name, first row, shape
        self.dims_per_state, self.coords_per_state = dims_per_state, coords_per_state
        self.time_name, self.time_coord = time_name, time_coord
        self.traj_id0 = int(traj_id0)
        self.data_vars = OrderedDict()
        n_time, n_draws = int(buffer.shape[1]), int(buffer.shape[2])
        for v in self.variables:
            view = buffer[v.r0:v.r0 + v.n_dim].view(*v.shape, n_time, n_draws)
            coords = OrderedDict([(time_name, list(time_coord))])
            for d, c in zip(dims_per_state[v.name], coords_per_state[v.name]):
                coords[d] = list(c)
            self.data_vars[v.name] = DeviceDataArray(view, v.name, dims_per_state[v.name], coords, time_name, self.n_global,
                                                     self.traj_id0)

    @property
    def n_draws(self):
        return self.n_global

    def __getitem__(self, key):
        if key in self.data_vars:
            return self.data_vars[key]
        if key == self.time_name:
            return np.asarray(self.time_coord)
        for da in self.data_vars.values():
            if key in da.coords:
                return np.asarray(da.coords[key])
        raise KeyError(key)

    def __contains__(self, key):
        return key in self.data_vars

    def __iter__(self):
        return iter(self.data_vars)

    def keys(self):
        return self.data_vars.keys()

    def like(self, buffer):
        """The same variables / coordinates over another buffer of the same shape (what a noise kernel returns)."""
        if tuple(buffer.shape) != tuple(self.buffer.shape):
            raise ValueError("buffer shape differs")
        return DeviceEnsemble(buffer, self.variables, self.dims_per_state, self.coords_per_state, self.time_name,
                              self.time_coord, self.traj_id0, self.n_global)

    def _host_dataset(self, arrays):
        if _xr is not None:
            return _xr.Dataset(arrays)
        return SimpleDataset(arrays)

    def mean(self, dim="draws"):
        if dim != "draws":
            raise NotImplementedError("device ensembles reduce over 'draws'")
        from . import sharding
        s = sharding.sum_sharded(engine.draws_sum(self.buffer)).cpu().numpy() / self.n_draws       # one launch for every variable
        out = OrderedDict()
        for v in self.variables:
            block = s[v.r0:v.r0 + v.n_dim].reshape(*v.shape, s.shape[1])
            out[v.name] = self.data_vars[v.name]._host_result(block)
        return self._host_dataset(out)

    def quantile(self, q, dim="draws"):
        return self._host_dataset(OrderedDict((k, da.quantile(q, dim)) for k, da in self.data_vars.items()))

    def to_dataset(self):
        """Full host copy with the reference's dims and coords (`base.py:969-1042`)."""
        from .dataset import make_dataset
        arrays = OrderedDict((k, da.values) for k, da in self.data_vars.items())
        n = int(self.buffer.shape[2])
        if n == 1:
            arrays = OrderedDict((k, a[0]) for k, a in arrays.items())
        return make_dataset(arrays, self.dims_per_state, self.coords_per_state, self.time_name, self.time_coord, n)

    def __repr__(self):
        return "<DeviceEnsemble " + ", ".join(repr(v) for v in self.data_vars.values()) + ">"
