"""Host-side inputs restated in pysodm_b200.tutorial_data agree with what the reference's tutorial code produces."""
from datetime import date

import numpy as np

from pysodm_b200 import tutorial_data as td


def test_contact_schedule_matches_reference_tdpf(golden):
    g = golden["contact_schedule"]
    assert np.array_equal(g["table"], td.CONTACT_TABLE)
    for label in ("season", "autumn", "spring"):
        ordinal, n, tau = g[label + "_meta"]
        start = date.fromordinal(int(ordinal)).strftime("%Y-%m-%d")
        assert np.array_equal(td.influenza_schedule(start, int(n), float(tau)), g[label + "_index"]), label
    idx = td.influenza_schedule()
    # 322 steps -> 181 no-holiday-week, 92 weekend, 49 holiday-week
    assert [int((idx == k).sum()) for k in range(3)] == [181, 92, 49]


def test_initial_conditions_and_parameters(golden):
    g, o = golden["jump_single"], golden["ode"]
    assert np.array_equal(td.influenza_flat_y0(), g["flu_y0"])
    assert np.array_equal(td.sirsi_flat_y0(), g["sirsi_y0"])
    assert np.array_equal(td.pfr_flat_y0(25), o["pfr25_y0"])
    assert np.array_equal(td.pfr_flat_par(25), np.concatenate([o["pfr25_par_scalar"], o["pfr25_kL_a"], o["pfr25_D_ax"]]))
    assert td.influenza_flat_par().shape == (24,)
