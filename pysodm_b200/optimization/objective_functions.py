"""Alias of `pysodm_b200.objective_functions` at the reference's module path
(`pySODM.optimization.objective_functions`, /root/reference/src/pySODM/optimization/objective_functions.py)."""
from ..objective_functions import *  # noqa: F401,F403
from ..objective_functions import log_posterior_probability  # noqa: F401
