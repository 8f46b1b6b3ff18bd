"""Result container: a real `xarray.Dataset` when xarray is importable, otherwise a minimal stand-in.

The reference returns `xarray.Dataset` objects (`/root/reference/src/pySODM/models/base.py:969-1042`): one
`DataArray` per state with dims `(['draws',] 'time'|'date', *state_dims)`, coordinates = the integer grid (or the
list of datetimes) and the model's coordinates; `draws` has no coordinate.  xarray is not installed in the build
image, so `SimpleDataset` implements the small subset the tutorials and the likelihood code touch
(`ds['S']`, `.values`, `.dims`, `.coords`, `.sel`, `.isel`, `.sum/.mean/.quantile(dim=...)`); it wraps the pinned
host arrays produced by the pipeline without copying them.
"""
from collections import OrderedDict

import numpy as np

try:  # pragma: no cover - depends on the environment
    import xarray as _xr
    if getattr(_xr, "__pysodm_stub__", False):
        _xr = None
except Exception:  # noqa: BLE001
    _xr = None


class SimpleDataArray:
    def __init__(self, values, dims, coords, name=None):
        self.values = np.asarray(values)
        self.dims = tuple(dims)
        self.coords = OrderedDict((d, coords[d]) for d in self.dims if d in coords)
        self.name = name
        if self.values.ndim != len(self.dims):
            raise ValueError(f"{name}: {self.values.ndim}-d values for dims {self.dims}")

    # -- basic protocol ------------------------------------------------------------------------------------
    @property
    def shape(self):
        return self.values.shape

    def __array__(self, dtype=None, copy=None):
        return np.asarray(self.values, dtype=dtype)

    def __getitem__(self, key):
        if isinstance(key, str):
            return np.asarray(self.coords[key])
        return self.values[key]

    def to_numpy(self):
        return self.values

    def _axis(self, dim):
        if dim not in self.dims:
            raise KeyError(f"dimension '{dim}' not in {self.dims}")
        return self.dims.index(dim)

    def _index_axis(self, dim, idx):
        ax = self._axis(dim)
        scalar = np.ndim(idx) == 0 and not isinstance(idx, slice)
        vals = np.take(self.values, idx, axis=ax) if not isinstance(idx, slice) else \
            self.values[(slice(None),) * ax + (idx,)]
        dims = [d for d in self.dims if not (scalar and d == dim)]
        coords = OrderedDict(self.coords)
        if dim in coords:
            c = np.asarray(coords[dim], dtype=object)[idx] if not isinstance(idx, slice) else list(coords[dim])[idx]
            if scalar:
                coords.pop(dim)
            else:
                coords[dim] = list(c)
        return SimpleDataArray(vals, dims, coords, self.name)

    def isel(self, indexers=None, **kw):
        out = self
        for dim, idx in {**(indexers or {}), **kw}.items():
            out = out._index_axis(dim, idx)
        return out

    def sel(self, indexers=None, **kw):
        out = self
        for dim, label in {**(indexers or {}), **kw}.items():
            labels = list(out.coords[dim])
            if isinstance(label, (list, tuple, np.ndarray)):
                idx = [labels.index(v) for v in label]
            else:
                idx = labels.index(label)
            out = out._index_axis(dim, idx)
        return out

    def _reduce(self, fn, dim, **kw):
        dims = [dim] if isinstance(dim, str) else list(dim if dim is not None else self.dims)
        axes = tuple(self._axis(d) for d in dims)
        vals = fn(self.values, axis=axes, **kw)
        rest = [d for d in self.dims if d not in dims]
        return vals, rest

    def sum(self, dim=None):
        vals, rest = self._reduce(np.sum, dim)
        return SimpleDataArray(vals, rest, self.coords, self.name)

    def mean(self, dim=None):
        vals, rest = self._reduce(np.mean, dim)
        return SimpleDataArray(vals, rest, self.coords, self.name)

    def quantile(self, q, dim=None):
        vals, rest = self._reduce(np.quantile, dim, q=q)
        if np.ndim(q) == 0:
            return SimpleDataArray(vals, rest, self.coords, self.name)
        coords = OrderedDict(self.coords)
        coords["quantile"] = list(np.atleast_1d(q))
        return SimpleDataArray(vals, ["quantile"] + rest, coords, self.name)

    def transpose(self, *dims):
        order = [self._axis(d) for d in dims]
        return SimpleDataArray(np.transpose(self.values, order), dims, self.coords, self.name)

    # -- arithmetic (what aggregation functions do to a state: scale, add up slices, weight along a dimension) ---------
    def _binary(self, other, fn, swap=False):
        if isinstance(other, SimpleDataArray):
            if set(other.dims) - set(self.dims):
                if set(self.dims) - set(other.dims):
                    raise ValueError(f"cannot combine dims {self.dims} with {other.dims}")
                return other._binary(self, fn, not swap)
            # broadcast `other` (a subset of this array's dims) along the missing dimensions, by NAME as xarray does
            o = other.transpose(*[d for d in self.dims if d in other.dims]).values
            o = o.reshape([self.values.shape[self._axis(d)] if d in other.dims else 1 for d in self.dims])
        else:
            o = np.asarray(other)
            if o.ndim > self.values.ndim:
                raise ValueError("cannot broadcast an array with more dimensions than the DataArray")
        vals = fn(o, self.values) if swap else fn(self.values, o)
        return SimpleDataArray(vals, self.dims, self.coords, self.name)

    def __add__(self, o): return self._binary(o, np.add)
    def __radd__(self, o): return self._binary(o, np.add, True)
    def __sub__(self, o): return self._binary(o, np.subtract)
    def __rsub__(self, o): return self._binary(o, np.subtract, True)
    def __mul__(self, o): return self._binary(o, np.multiply)
    def __rmul__(self, o): return self._binary(o, np.multiply, True)
    def __truediv__(self, o): return self._binary(o, np.divide)
    def __neg__(self): return SimpleDataArray(-self.values, self.dims, self.coords, self.name)

    def copy(self, deep=True):
        return SimpleDataArray(self.values.copy() if deep else self.values, self.dims, self.coords, self.name)

    def __repr__(self):
        return f"<SimpleDataArray {self.name} {dict(zip(self.dims, self.values.shape))}>"


class SimpleDataset:
    def __init__(self, data_vars):
        self.data_vars = OrderedDict(data_vars)

    def __getitem__(self, key):
        if key in self.data_vars:
            return self.data_vars[key]
        for da in self.data_vars.values():           # coordinate lookup: out['date'], out['time']
            if key in da.coords:
                return np.asarray(da.coords[key])
        raise KeyError(key)

    def __contains__(self, key):
        return key in self.data_vars

    def keys(self):
        return self.data_vars.keys()

    def __iter__(self):
        return iter(self.data_vars)

    @property
    def dims(self):
        out = OrderedDict()
        for da in self.data_vars.values():
            for d, n in zip(da.dims, da.values.shape):
                out[d] = n
        return out

    @property
    def coords(self):
        out = OrderedDict()
        for da in self.data_vars.values():
            out.update(da.coords)
        return out

    def copy(self):
        """Deep copy of the values (what `xarray.Dataset.copy()` + a `.values` assignment amount to in utils.py:26-32)."""
        return self._map(lambda v: SimpleDataArray(v.values.copy(), v.dims, v.coords, v.name))

    def _map(self, fn):
        return SimpleDataset(OrderedDict((k, fn(v)) for k, v in self.data_vars.items()))

    def isel(self, indexers=None, **kw):
        sel = {**(indexers or {}), **kw}
        return self._map(lambda v: v.isel({d: i for d, i in sel.items() if d in v.dims}))

    def sel(self, indexers=None, **kw):
        sel = {**(indexers or {}), **kw}
        return self._map(lambda v: v.sel({d: i for d, i in sel.items() if d in v.dims}))

    def mean(self, dim=None):
        return self._map(lambda v: v.mean(dim) if dim in v.dims else v)

    def __repr__(self):
        return "<SimpleDataset " + ", ".join(repr(v) for v in self.data_vars.values()) + ">"


def make_dataset(arrays, dims_per_state, coords_per_state, time_name, time_coord, n_draws):
    """arrays: {state: ndarray[(draws,) time, *dims]} -> Dataset with the reference's dims / coords."""
    data = OrderedDict()
    for name, values in arrays.items():
        dims = ([] if n_draws == 1 else ["draws"]) + [time_name] + list(dims_per_state[name])
        coords = {time_name: time_coord}
        for d, c in zip(dims_per_state[name], coords_per_state[name]):
            coords[d] = list(c)
        if _xr is not None:
            data[name] = _xr.DataArray(values, dims=dims, coords=coords)
        else:
            data[name] = SimpleDataArray(values, dims, coords, name)
    return _xr.Dataset(data) if _xr is not None else SimpleDataset(data)
