"""The batched optimisers reproduce the reference's pso.optimize / nelder_mead.optimize on a deterministic
objective (fixtures: seeded runs of the unmodified reference, tests/golden/make_golden.py:gen_optimizers)."""
import numpy as np

from pysodm_b200 import optimization

BOUNDS = [(-2, 2), (-3, 3), (0, 5)]


def objective(x):
    x = np.asarray(x)
    return -(np.sum((x - np.array([0.3, -1.2, 2.0])) ** 2) + 0.5 * np.sin(3 * x[0]) * x[1])


class Batched:
    """An objective with the `batch` seam of log_posterior_probability; counts how it is called."""
    expanded_bounds = BOUNDS

    def __init__(self):
        self.batch_calls, self.rows = 0, 0

    def __call__(self, x):
        return objective(x)

    def batch(self, X):
        self.batch_calls += 1
        self.rows += len(X)
        return np.array([objective(x) for x in X])


def test_pso_matches_reference(golden, capsys):
    g = golden["optimizers"]
    for seed, S, it in [(1, 20, 15), (2, 64, 40)]:
        np.random.seed(seed)
        theta, score = optimization.pso(objective, bounds=BOUNDS, swarmsize=S, max_iter=it)
        assert np.array_equal(theta, g[f"pso_{seed}_theta"]) and score == g[f"pso_{seed}_score"]
        f = Batched()
        np.random.seed(seed)
        theta, score = optimization.pso(f, swarmsize=S, max_iter=it)            # bounds from the objective, batch seam
        assert np.array_equal(theta, g[f"pso_{seed}_theta"]) and score == g[f"pso_{seed}_score"]
        assert f.batch_calls == it + 1 and f.rows == S * (it + 1)              # ONE evaluation call per iteration


def test_nelder_mead_matches_reference(golden, capsys):
    g = golden["optimizers"]
    theta, score = optimization.nelder_mead(objective, np.array([1.0, 1.0, 1.0]), [0.1, 0.1, 0.1], bounds=BOUNDS, max_iter=60)
    assert np.array_equal(theta, g["nm_theta"]) and score == g["nm_score"]
    f = Batched()
    theta, score = optimization.nelder_mead(f, [1.9, 2.5, 0.2], [0.3, 0.3, 0.3], max_iter=25)
    assert np.array_equal(theta, g["nm2_theta"]) and score == g["nm2_score"]
