"""Observation noise on a finished ensemble, on the GPU -- the reference's `pySODM.optimization.utils.add_*_noise`
(`/root/reference/src/pySODM/optimization/utils.py:11-95`), same names and arguments.

`output` may be
  * a `DeviceEnsemble` (from `sim(..., output='device')`): one `sodm_noise` launch over the whole buffer, the result is
    a new `DeviceEnsemble` (the input is left untouched, as the reference's `output.copy()` does);
  * a `DeviceDataArray` (`ens['Im_inc']`, possibly `.sel(...)`-ed): noise on that variable only -- a seventh of the work
    when only the observed state is plotted;
  * a host Dataset (xarray or the `SimpleDataset` stand-in): every variable is sent to the device, replaced by its noisy
    version and copied back -- the drop-in form of the reference call; there is no CPU sampler.
Random streams come from Philox keyed by (seed, draw, cell); `seed=None` takes 64 bits from NumPy's global generator,
so `np.random.seed(...)` before the call makes it reproducible like the reference's use of the global state.
"""
import numpy as np

from .. import _ffi, engine
from ..ensemble import DeviceDataArray, DeviceEnsemble


def _seed(seed):
    if seed is not None:
        return int(seed)
    return int(np.random.randint(0, 2 ** 32)) | (int(np.random.randint(0, 2 ** 31)) << 32)


def _raise_on_status(status, kind):
    bits = int(status.item())
    if bits & _ffi.NOISE_ST_BAD_MEAN:
        # NumPy's own messages for the same inputs (utils.py:30, :91)
        raise ValueError("lam < 0 or lam contains NaNs" if kind == "poisson" else "p < 0, p > 1 or p is NaN")


def _apply(output, kind, param, seed, integer_result):
    import torch
    from .. import sharding
    seed = sharding.agree_on(_seed(seed))
    if isinstance(output, DeviceEnsemble):
        buf = output.buffer
        out_dtype = buf.dtype if (integer_result or buf.dtype != torch.int32) else torch.float64
        noisy, status = engine.noise(kind, buf, param, seed=seed, traj_id0=output.traj_id0, out_dtype=out_dtype)
        _raise_on_status(status, kind)
        return output.like(noisy)
    if isinstance(output, DeviceDataArray):                   # one state variable (or a selection of it) of a device ensemble
        data = output._packed()
        out_dtype = data.dtype if (integer_result or data.dtype != torch.int32) else torch.float64
        noisy, status = engine.noise(kind, data, param, seed=seed, traj_id0=output.traj_id0, out_dtype=out_dtype)
        _raise_on_status(status, kind)
        return DeviceDataArray(noisy, output.name, output.state_dims, output.coords, output.time_name, output.n_global,
                               output.traj_id0)
    # host Dataset: variable by variable through the device
    dev = engine.device()
    new_output = output.copy()
    for i, (name, da) in enumerate(list(new_output.data_vars.items())):
        host = np.ascontiguousarray(np.asarray(da.values, dtype=np.float64))
        flat = torch.from_numpy(host.reshape(1, -1)).to(dev)
        noisy, status = engine.noise(kind, flat, param, seed=seed + i, out_dtype=torch.float64)
        _raise_on_status(status, kind)
        values = noisy.cpu().numpy().reshape(host.shape)
        if integer_result:
            values = values.astype(np.int64)                       # NumPy's samplers return int64
        new_output[name].values = values
    return new_output


def add_poisson_noise(output, seed=None):
    """Every value replaced by a Poisson draw with that mean (utils.py:11-33)."""
    return _apply(output, "poisson", 0.0, seed, True)


def add_gaussian_noise(output, sigma, relative=True, seed=None):
    """Every value replaced by a normal draw around it with standard deviation `sigma` (absolute) or `|sigma * value|`
    (relative, the default) (utils.py:35-66)."""
    if sigma < 0:
        raise ValueError("scale < 0")                               # NumPy's message
    return _apply(output, "gaussian_relative" if relative else "gaussian", float(sigma), seed, False)


def add_negative_binomial_noise(output, alpha, seed=None):
    """Every value replaced by a negative-binomial draw with that mean and variance `mean + alpha * mean^2`
    (utils.py:68-95: n = 1/alpha, p = n / (value + n))."""
    if not alpha > 0:
        raise ZeroDivisionError("float division by zero") if alpha == 0 else ValueError("n <= 0")   # as 1/alpha / NumPy
    return _apply(output, "negative_binomial", float(alpha), seed, True)
