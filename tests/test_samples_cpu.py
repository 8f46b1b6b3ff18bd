"""Posterior-sample files in the reference's wire format (mcmc.py:310-378) and draw functions built on them."""
import os

import numpy as np
import pytest

from pysodm_b200 import samples as sm

REF_FILE = "/root/reference/tutorials/enzyme_kinetics/data/username_SAMPLES_2025-02-05.nc"


def test_dump_and_reload_round_trip(tmp_path):
    rng = np.random.default_rng(0)
    chain = rng.normal(size=(7, 12, 1 + 4 + 6))
    shapes = {"beta": (1,), "f_ud": (4,), "M": (2, 3)}
    path = str(tmp_path / "samples.nc")
    s = sm.dump_samples(chain, shapes, path, {"start_calibration": "2017-12-01", "n_chains": 12, "starting_estimate": [0.1, 0.2]})
    assert s["beta"].shape == (7, 12) and s["f_ud"].shape == (7, 12, 4) and s["M"].shape == (7, 12, 2, 3)
    assert np.array_equal(s["M"][3, 5], chain[3, 5, 5:].reshape(2, 3))
    back = sm.load_samples(path)
    assert set(back) == {"beta", "f_ud", "M"}
    for k in shapes:
        assert np.array_equal(back[k], s[k])
    assert back.attrs["start_calibration"] == "2017-12-01" and int(back.attrs["n_chains"]) == 12
    assert np.allclose(back.attrs["starting_estimate"], [0.1, 0.2])
    # the file has the reference's structure: integer coordinate variables, <name>_dim_i dimensions, 64-bit-offset NetCDF-3
    from scipy.io import netcdf_file
    with netcdf_file(path, "r", mmap=False) as f:
        assert f.version_byte == 2
        assert set(f.dimensions) == {"iteration", "chain", "f_ud_dim_0", "M_dim_0", "M_dim_1"}
        assert f.variables["M"].dimensions == ("iteration", "chain", "M_dim_0", "M_dim_1")
        assert np.array_equal(f.variables["chain"].data, np.arange(12))


def test_draw_function_takes_whole_samples():
    rng = np.random.default_rng(1)
    s = sm.dump_samples(rng.normal(size=(5, 6, 3)), {"a": (1,), "b": (2,)})
    draw = sm.draw_function_from_samples(s, rng=np.random.default_rng(2))
    assert draw.vectorized
    out = draw({"a": 0.0, "b": np.zeros(2), "c": 1.0}, N=1000)
    assert out["a"].shape == (1000,) and out["b"].shape == (1000, 2) and out["c"] == 1.0
    flat_a, flat_b = s.flat("a"), s.flat("b")
    for i in range(0, 1000, 97):                           # a and b of a draw come from the SAME posterior sample
        j = int(np.flatnonzero(flat_a == out["a"][i])[0])
        assert np.array_equal(flat_b[j], out["b"][i])
    assert len(np.unique(out["a"])) > 25


@pytest.mark.skipif(not os.path.exists(REF_FILE), reason="reference data file only exists in the build container")
def test_reads_the_reference_samples_file():
    s = sm.load_samples(REF_FILE)
    assert list(s) == ["Vf_Ks", "R_AS", "R_AW", "R_Es", "K_eq"] and s.n_iterations == 10 and s.n_chains == 90
    assert s["Vf_Ks"].shape == (10, 90) and np.isfinite(s["K_eq"]).all()
    assert int(s.attrs["discard"]) == 100 and int(s.attrs["thin"]) == 10
    assert np.allclose(s.attrs["starting_estimate"], [9.5e-04, 7.5e-01, 1.4, 5.0, 0.7])
