"""CPU oracle for the pySODM hot path -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A NumPy/SciPy restatement of the algorithms on the hot path of twallema/pySODM v0.2.8
(`/root/reference`): the tau-leap / SSA jump-process steppers, the `solve_ivp` call the ODE
runtime makes, the fixed-step Euler fallback, and the Poisson / normal / lognormal /
negative-binomial log-likelihoods with the time alignment `compute_log_likelihood` performs.
Every function cites the reference file:line it follows.

Who may import this package: `tests/`, `__graft_entry__.smoke()` and the `cpu_baseline` /
`--impl reference` legs of `bench.py` -- as the checker or as the timed CPU baseline, never as
the thing shipped.  `pysodm_b200` never imports it; the product path fails loudly when the
CUDA library is missing.

Parity pinning: the reference ships no golden vectors (its tests assert shapes, error strings
and loose inequalities only -- SURVEY.md section 4), so the oracle is pinned against OUTPUTS OF
THE REFERENCE ITSELF, generated in the build container by `tests/golden/make_golden.py`
(which imports the unmodified reference) and committed under `tests/golden/*.npz`:
  * seeded single-trajectory tau-leap / SSA runs: bit-exact (same NumPy RNG call order),
  * `solve_ivp` RK45 runs of the tutorial ODE models: bit-exact,
  * `ll_*` values: bit-exact,
  * reference ensembles (N ~ 2000) for KS tests of the batched stepper.
"""
