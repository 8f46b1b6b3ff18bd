"""Posterior samples on disk, in the reference's wire format, and as draw functions for ensembles.

The reference stores emcee chains with `_dump_sampler_to_xarray`
(`/root/reference/src/pySODM/optimization/mcmc.py:310-378`): one variable per parameter with dims
`('iteration', 'chain', '<name>_dim_0', ...)`, integer coordinates, the calibration settings as global attributes,
written with `xarray.Dataset.to_netcdf` -- the shipped `tutorials/enzyme_kinetics/data/username_SAMPLES_2025-02-05.nc` is
NetCDF-3 (64-bit offset).  The tutorials then draw parameter sets from such a file for every ensemble member
(`tutorials/enzyme_kinetics/simulate_1D_PFR.py`, `draw_fcn`).  xarray is not needed for any of this: NetCDF-3 is read and
written with `scipy.io.netcdf_file`, the format xarray itself falls back to, so files travel in both directions.
"""
from collections import OrderedDict

import numpy as np
from scipy.io import netcdf_file


class Samples(OrderedDict):
    """{parameter: ndarray [iteration, chain, *shape]} with `.attrs` (the stored calibration settings)."""

    def __init__(self, *a, attrs=None, **k):
        super().__init__(*a, **k)
        self.attrs = dict(attrs or {})

    @property
    def n_iterations(self):
        return next(iter(self.values())).shape[0]

    @property
    def n_chains(self):
        return next(iter(self.values())).shape[1]

    def flat(self, name):
        """All (iteration, chain) samples of one parameter as [n_samples, *shape]."""
        v = self[name]
        return v.reshape((-1,) + v.shape[2:])


def dump_samples(samples_np, parameter_shapes, filename=None, settings_dict=None):
    """The reference's `_dump_sampler_to_xarray`: split the raw chain `samples_np[iteration, chain, flat parameter]` into
    named parameters of the given shapes (`(1,)` = scalar, kept 2-D) and optionally write them as NetCDF-3."""
    samples_np = np.asarray(samples_np, dtype=np.float64)
    out, count = Samples(attrs=settings_dict), 0
    for name, shape in parameter_shapes.items():
        shape = tuple(shape)
        if shape != (1,):
            n = int(np.prod(shape))
            out[name] = samples_np[:, :, count:count + n].reshape(samples_np.shape[0], samples_np.shape[1], *shape)
            count += n
        else:
            out[name] = samples_np[:, :, count]
            count += 1
    if filename:
        save_samples(out, filename)
    return out


def save_samples(samples, filename):
    """Write in the reference's layout (NetCDF-3, 64-bit offset): dims iteration / chain / <name>_dim_i with integer
    coordinate variables, settings as global attributes."""
    with netcdf_file(filename, "w", version=2) as f:
        f.createDimension("iteration", samples.n_iterations)
        f.createDimension("chain", samples.n_chains)
        for dim, n in (("iteration", samples.n_iterations), ("chain", samples.n_chains)):
            v = f.createVariable(dim, "i4", (dim,))
            v[:] = np.arange(n, dtype=np.int32)
        for name, arr in samples.items():
            dims = ["iteration", "chain"]
            for i, n in enumerate(arr.shape[2:]):
                d = f"{name}_dim_{i}"
                f.createDimension(d, n)
                c = f.createVariable(d, "i4", (d,))
                c[:] = np.arange(n, dtype=np.int32)
                dims.append(d)
            v = f.createVariable(name, "f8", tuple(dims))
            v[:] = arr
        for key, value in samples.attrs.items():
            if isinstance(value, (list, tuple)):
                value = np.asarray(value)
            if isinstance(value, (bool, np.bool_)):
                value = int(value)
            setattr(f, key, value)


def load_samples(filename):
    """Read a samples file written by the reference (or by `save_samples`)."""
    with netcdf_file(filename, "r", mmap=False) as f:
        out = Samples(attrs={k: (v.decode() if isinstance(v, bytes) else v) for k, v in f._attributes.items()})
        for name, var in f.variables.items():
            if name in f.dimensions:                        # coordinate variables
                continue
            if var.dimensions[:2] != ("iteration", "chain"):
                raise ValueError(f"{filename}: variable '{name}' has dims {var.dimensions}, expected ('iteration', 'chain', ...)")
            out[name] = np.array(var.data, dtype=np.float64)
    return out


def draw_function_from_samples(samples, names=None, rng=None):
    """A VECTORISED draw function (`fn.vectorized = True`, one call for all N draws): every ensemble member gets the
    parameters of one posterior sample (iteration, chain) picked uniformly with replacement, all named parameters from
    the SAME sample -- what the tutorials' per-draw `draw_fcn` does with `np.random.choice` of an index."""
    names = list(names or samples.keys())

    def draw(parameters, N):
        gen = rng if rng is not None else np.random
        n = samples.n_iterations * samples.n_chains
        pick = gen.integers(0, n, size=N) if hasattr(gen, "integers") else gen.randint(0, n, size=N)
        for name in names:
            parameters[name] = samples.flat(name)[pick]
        return parameters
    draw.vectorized = True
    return draw
