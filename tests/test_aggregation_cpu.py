"""User aggregation functions of `log_posterior_probability` (reference: objective_functions.py:181-188, :245-248,
validate_aggregation_function :1181-1225; its own test: src/tests/test_calibration.py:624-668) -- CPU side: the accepted
forms and messages, and the matrix a linear aggregation function is folded into (known answers derived by hand)."""
import numpy as np
import pandas as pd
import pytest

import gen_models as gm
from oracle import loglik as oll
from pysodm_b200.objective_functions import linear_map_of, ll_negative_binomial, log_posterior_probability

DIMS, COORDS, SHAPE = ["age_groups", "spatial_units"], [["0-20", "20-120"], [0, 1, 2]], (2, 3)


def _data(by_age=True):
    t = np.arange(0, 50, 7, dtype=float)
    if by_age:
        idx = pd.MultiIndex.from_product([t, ["0-20", "20-120"]], names=["time", "age_groups"])
    else:
        idx = pd.Index(t, name="time")
    rng = np.random.default_rng(3)
    return pd.Series(rng.gamma(2.0, 40.0, size=len(idx)), index=idx, name="cases")


def test_matrix_of_linear_aggregation_functions():
    A, dims, coords, shape = linear_map_of(lambda out: out.sum(dim="spatial_units"), "I", SHAPE, DIMS, COORDS, "time")
    assert dims == ["age_groups"] and coords == [["0-20", "20-120"]] and shape == (2,)
    assert np.array_equal(A, [[1, 1, 1, 0, 0, 0], [0, 0, 0, 1, 1, 1]])
    # a weighted sum over part of a dimension, rescaled
    A, dims, coords, shape = linear_map_of(lambda out: 0.5 * out.sel(spatial_units=[0, 2]).sum(dim="spatial_units"),
                                           "I", SHAPE, DIMS, COORDS, "date")
    assert np.array_equal(A, [[0.5, 0, 0.5, 0, 0, 0], [0, 0, 0, 0.5, 0, 0.5]])
    # keeps both dimensions, mixes nothing: the identity
    A, dims, _, shape = linear_map_of(lambda out: out, "I", SHAPE, DIMS, COORDS, "time")
    assert np.array_equal(A, np.eye(6)) and dims == DIMS and shape == SHAPE
    # not linear / not pointwise in time / loses the time axis: refused, there is no host fallback
    with pytest.raises(NotImplementedError, match="not a linear map"):
        linear_map_of(lambda out: out * out, "I", SHAPE, DIMS, COORDS, "time")
    with pytest.raises(NotImplementedError, match="not a linear map"):
        linear_map_of(lambda out: out + 1.0, "I", SHAPE, DIMS, COORDS, "time")
    with pytest.raises(NotImplementedError, match="keeps the 'time' dimension"):
        linear_map_of(lambda out: out.sum(dim="time"), "I", SHAPE, DIMS, COORDS, "time")


def test_accepted_forms_and_messages_as_in_the_reference_test():
    model = gm.make_double_sir()
    data = [_data(), _data()]
    args = dict(model=model, parameter_names=["beta"], bounds=[(1e-6, 1)], data=data, states=["I", "I"],
                log_likelihood_fnc=[ll_negative_binomial, ll_negative_binomial],
                log_likelihood_fnc_args=[0.03 * np.ones(2), 0.03 * np.ones(2)], start_sim=0)

    def agg(output):
        return output.sum(dim="spatial_units")
    for ok in (agg, [agg], [agg, agg], [agg, None]):
        o = log_posterior_probability(**args, aggregation_function=ok)
        assert o._plan[0]["columns"] == [[6, 7, 8], [9, 10, 11]] and o._plan[0]["weights"] is None     # flat rows: S 0-5, I 6-11
    with pytest.raises(ValueError, match="Valid formats of aggregation functions are:"):
        log_posterior_probability(**args, aggregation_function="hello")
    with pytest.raises(ValueError, match="Valid formats of aggregation functions are:"):
        log_posterior_probability(**args, aggregation_function=["hello", agg])
    with pytest.raises(ValueError, match="number of aggregation functions must be equal to one or"):
        log_posterior_probability(**args, aggregation_function=5 * [agg])


def test_plan_known_answer():
    """columns / weights of the plan against a hand computation through the oracle's aligner."""
    model = gm.make_double_sir()

    def agg(output):                             # units 0 and 2 only, unit 2 counted twice
        return output.sel(spatial_units=0) + 2.0 * output.sel(spatial_units=2)
    o = log_posterior_probability(model, ["beta"], [(1e-6, 1)], [_data()], ["I"], [ll_negative_binomial], [0.03 * np.ones(2)],
                                  start_sim=0, aggregation_function=agg)
    plan = o._plan[0]
    assert plan["columns"] == [[6, 8], [9, 11]] and plan["weights"] == [[1.0, 2.0], [1.0, 2.0]]
    y = np.random.default_rng(0).uniform(1, 100, size=(18, 50))
    ym = oll.align(y, 0.0, plan["columns"], plan["t_data"], plan["weights"])
    I = y[6:12].reshape(2, 3, 50)
    want = (I[:, 0] + 2.0 * I[:, 2])[:, ::7].T                                   # data at t = 0, 7, .., 49: grid points
    assert np.allclose(ym, want, rtol=1e-14)
    # data without the age axis: the aggregated state's remaining dimension is summed over (:250-252)
    o2 = log_posterior_probability(model, ["beta"], [(1e-6, 1)], [_data(False)], ["I"], [ll_negative_binomial], [0.03],
                                   start_sim=0, aggregation_function=agg)
    assert o2._plan[0]["columns"] == [[6, 8, 9, 11]] and o2._plan[0]["weights"] == [[1.0, 2.0, 1.0, 2.0]]
