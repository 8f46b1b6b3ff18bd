"""Oracle for the log-likelihoods and the model/data alignment (see oracle/__init__.py -- test infrastructure only).

`ll_*` follow `/root/reference/src/pySODM/optimization/objective_functions.py:550-666` (pinned bit-exactly by
`tests/golden/loglik.npz`, generated from the unmodified functions).  `align` restates what
`compute_log_likelihood` (:244-288) does through xarray -- which is not installed here, so that part of the
reference cannot be executed: sum over the model dimensions absent from the data (:250-252), linear
interpolation in time to the data's unique time points (:254), selection of the data's coordinates in
`(time, *data_axes)` order (:256-259, :271-278).
"""
import numpy as np
from scipy.special import gammaln


def ll_poisson(ymodel, ydata):
    ym = np.asarray(ymodel).astype("float64") + 1
    yd = np.asarray(ydata).astype("float64") + 1
    return -np.sum(ym) + np.sum(yd * np.log(ym)) - np.sum(gammaln(yd))


def ll_negative_binomial(ymodel, ydata, alpha):
    alpha = np.array(alpha)[np.newaxis, ...]
    ym = np.asarray(ymodel).astype("float64") + 1
    yd = np.asarray(ydata).astype("float64") + 1
    return (np.sum(yd * np.log(ym)) - np.sum((yd + 1 / alpha) * np.log(1 + alpha * ym)) + np.sum(yd * np.log(alpha))
            + np.sum(gammaln(yd + 1 / alpha)) - np.sum(gammaln(yd + 1)) - np.sum(yd.shape[0] * gammaln(1 / alpha)))


def _sigma(sigma, like):
    sigma = np.array(sigma)
    if not sigma.shape == np.shape(like):
        sigma = sigma[np.newaxis, ...]
    if len(sigma[sigma <= 0]) != 0:
        raise ValueError("the standard deviation contains values smaller than or equal to zero")
    return sigma


def ll_normal(ymodel, ydata, sigma):
    sigma = _sigma(sigma, ymodel)
    return -1 / 2 * np.sum((ydata - ymodel) ** 2 / sigma ** 2 + np.log(2 * np.pi * sigma ** 2))


def ll_lognormal(ymodel, ydata, sigma):
    sigma = _sigma(sigma, ymodel)
    return -1 / 2 * np.sum((np.log(ydata + 1) - np.log(ymodel + 1)) ** 2 / sigma ** 2 + np.log(2 * np.pi * sigma ** 2)
                           + np.log(ydata + 1))


LL = {"poisson": ll_poisson, "negative_binomial": ll_negative_binomial, "normal": ll_normal,
      "lognormal": ll_lognormal}


def align(y, t0, columns, t_data, column_weights=None):
    """y[n_row, n_out] (one trajectory on the grid t0, t0+1, ..) -> ymodel[n_t, n_col].

    columns[c] = rows summed into data column c (with column_weights[c] when a linear aggregation function was folded in,
    objective_functions.py:245-248); linear interpolation in time, NaN outside the grid."""
    y = np.asarray(y, dtype=np.float64)
    grid = t0 + np.arange(y.shape[1], dtype=np.float64)
    t_data = np.asarray(t_data, dtype=np.float64)
    out = np.empty((len(t_data), len(columns)))
    for c, rows in enumerate(columns):
        if column_weights is None:
            summed = y[list(rows)].sum(axis=0)
        else:
            summed = (np.asarray(column_weights[c], dtype=np.float64)[:, None] * y[list(rows)]).sum(axis=0)
        out[:, c] = np.interp(t_data, grid, summed, left=np.nan, right=np.nan)
    return out


def log_likelihood(kind, y, t0, columns, t_data, ydata, ll_arg=None, weight=1.0, column_weights=None):
    ymodel = align(y, t0, columns, t_data, column_weights)
    ydata = np.asarray(ydata, dtype=np.float64)
    if ydata.ndim == 1:
        ydata = ydata[:, None]
    if np.isnan(ymodel).any():
        raise ValueError("simulation output contains nan")
    if kind == "poisson":
        return weight * ll_poisson(ymodel, ydata)
    return weight * LL[kind](ymodel, ydata, ll_arg)
