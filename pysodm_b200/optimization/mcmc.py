"""Ensemble MCMC with ONE batched objective evaluation per half-ensemble update.

Next-row seam of the hot path ("emcee via vectorize=True").  The reference's
`run_EnsembleSampler` (`/root/reference/src/pySODM/optimization/mcmc.py:15-231`) hands the objective to
`emcee.EnsembleSampler` together with a process pool, which scores the proposals of half the walkers one by one
(`:172-176`).  Here the proposals of a half go through `objective.batch(Q)` -- one launch of the simulation + likelihood
kernels for all of them when the objective is a `pysodm_b200.objective_functions.log_posterior_probability` -- while the
walker bookkeeping (D <= ~10 numbers per walker) stays on the host, so 10^5 walkers cost what 10^5 PSO particles cost.

emcee is a third-party dependency of the reference (unpinned in `setup.py`, not vendored, not in this image), so the
sampler is a restatement of the published algorithms the reference's default cocktail (`mcmc.py:22-28`) names:
  stretch move        Goodman & Weare 2010, a = 2, acceptance z^(D-1) p(q)/p(s);
  differential evol.  ter Braak 2006, q = s + gamma (c_a - c_b) + eps, gamma = 2.38/sqrt(2D) or 1, eps ~ N(0, 1e-5^2);
  snooker             ter Braak & Vrugt 2008, gamma = 1.7, acceptance (|q-z|/|s-z|)^(D-1) p(q)/p(s);
  KDE move            independence proposal from a Gaussian KDE (Scott's rule) of the complementary walkers.
Every move updates the ensemble in two halves (a random split per iteration), the walkers of one half proposing from
the other, which is what keeps the parallel update in detailed balance (Foreman-Mackey et al. 2013).  Each move leaves
the target invariant (tests/test_mcmc_cpu.py samples a correlated Gaussian with every move on its own and with the
cocktail); a seeded run does not reproduce emcee's random stream.

Chains are kept in memory and checkpointed as `.npz` (the reference's HDF5 backend needs h5py, absent here); samples
are written in the reference's NetCDF layout through `pysodm_b200.samples`.  The diagnostic figures of the reference
need matplotlib and are outside the path.
"""
import datetime
import os
import sys

import numpy as np

from ..samples import dump_samples


# ---------------------------------------------------------------------------------------------------------------------
# moves: propose(s[Ns, D], c[Nc, D], rng) -> q[Ns, D], log of the Hastings factor [Ns]
# ---------------------------------------------------------------------------------------------------------------------

def _pick_distinct(rng, n_pool, n_rows, k):
    """k distinct indices into a pool of n_pool, independently for each of n_rows rows."""
    if n_pool < k:
        raise ValueError(f"a move needs at least {k} walkers in the complementary half, got {n_pool}")
    idx = np.empty((n_rows, k), dtype=np.int64)
    idx[:, 0] = rng.integers(0, n_pool, n_rows)
    for j in range(1, k):
        # draw from the n_pool - j remaining slots and step over the ones already taken (in ascending order)
        r = rng.integers(0, n_pool - j, n_rows)
        taken = np.sort(idx[:, :j], axis=1)
        for t in range(j):
            r += r >= taken[:, t]
        idx[:, j] = r
    return idx


class StretchMove:
    def __init__(self, a=2.0, **_ignored):
        self.a = float(a)

    def propose(self, s, c, rng):
        Ns, D = s.shape
        zz = ((self.a - 1.0) * rng.random(Ns) + 1.0) ** 2 / self.a
        other = c[rng.integers(0, len(c), Ns)]
        return other - (other - s) * zz[:, None], (D - 1.0) * np.log(zz)


class DEMove:
    def __init__(self, sigma=1.0e-5, gamma0=None):
        self.sigma, self.gamma0 = float(sigma), gamma0

    def propose(self, s, c, rng):
        Ns, D = s.shape
        g0 = 2.38 / np.sqrt(2.0 * D) if self.gamma0 is None else float(self.gamma0)
        ab = _pick_distinct(rng, len(c), Ns, 2)
        q = s + g0 * (c[ab[:, 0]] - c[ab[:, 1]]) + self.sigma * rng.standard_normal(Ns)[:, None]
        return q, np.zeros(Ns)


class DESnookerMove:
    def __init__(self, gammas=1.7):
        self.gammas = float(gammas)

    def propose(self, s, c, rng):
        Ns, D = s.shape
        w = _pick_distinct(rng, len(c), Ns, 3)
        z, z1, z2 = c[w[:, 0]], c[w[:, 1]], c[w[:, 2]]
        delta = s - z
        norm = np.linalg.norm(delta, axis=1)
        u = delta / norm[:, None]
        q = s + u * (self.gammas * (np.sum(u * z1, axis=1) - np.sum(u * z2, axis=1)))[:, None]
        return q, (D - 1.0) * (np.log(np.linalg.norm(q - z, axis=1)) - np.log(norm))


class KDEMove:
    """Independence proposal from a Gaussian KDE of the complementary walkers.  The two densities the Hastings factor
    needs (K kernels x 2 Ns points) are summed on the device (`sodm_kde_logpdf`; with scipy on the host they cost 10 s
    per half update at 10^5 walkers, 400x the batched objective).  `max_points` bounds the kernel count: a uniform
    subset of the complementary half, chosen without looking at the walkers being moved, is as valid a basis for an
    independence proposal as the whole half.  `logpdf(data[K, D], queries[M, D], bw_method)` replaces the device
    evaluation (the CPU tests pass scipy's `gaussian_kde`)."""

    def __init__(self, bw_method=None, max_points=4096, logpdf=None):
        self.bw_method, self.max_points, self.logpdf = bw_method, max_points, logpdf

    def bandwidth(self, K, D):
        if self.bw_method in (None, "scott"):
            return K ** (-1.0 / (D + 4))
        if self.bw_method == "silverman":
            return (K * (D + 2) / 4.0) ** (-1.0 / (D + 4))
        return float(self.bw_method)

    def factor_of(self, c):
        """Lower Cholesky factor of the kernel covariance for the centres c[K, D]."""
        return np.linalg.cholesky(np.atleast_2d(np.cov(c, rowvar=False)) * self.bandwidth(*c.shape) ** 2)

    def propose(self, s, c, rng):
        if self.max_points and len(c) > self.max_points:
            c = c[rng.choice(len(c), self.max_points, replace=False)]
        K, D = c.shape
        factor = self.bandwidth(K, D)
        L = self.factor_of(c)
        q = c[rng.integers(0, K, len(s))] + rng.standard_normal((len(s), D)) @ L.T
        logpdf = self.logpdf
        if logpdf is None:
            from ..engine import gaussian_kde_logpdf as logpdf
        dens = logpdf(c, np.concatenate([s, q]), factor)
        return q, dens[:len(s)] - dens[len(s):]


def default_moves():
    """The reference's cocktail and weights (`mcmc.py:22-28`)."""
    return [(DEMove(), 0.5 * 0.5 * 0.9), (DEMove(gamma0=1.0), 0.5 * 0.5 * 0.1), (DESnookerMove(), 0.5 * 0.5),
            (KDEMove(bw_method="scott"), 0.25), (StretchMove(), 0.25)]


# ---------------------------------------------------------------------------------------------------------------------
# integrated autocorrelation time (what `emcee.autocorr.integrated_time(chain, tol=0)` reports to the reference's
# convergence test, `visualization.py:49`, `mcmc.py:215-222`): FFT autocorrelation averaged over walkers, Sokal window c=5
# ---------------------------------------------------------------------------------------------------------------------

def _autocorr_1d(x):
    n = 1 << int(np.ceil(np.log2(max(len(x), 1))))
    f = np.fft.fft(x - np.mean(x), n=2 * n)
    acf = np.fft.ifft(f * np.conjugate(f))[:len(x)].real
    return acf / acf[0] if acf[0] != 0 else np.full(len(x), np.nan)


def integrated_time(chain, c=5.0):
    """chain[iteration, walker, D] -> tau[D]."""
    chain = np.asarray(chain, dtype=np.float64)
    n_t, n_w, n_d = chain.shape
    tau = np.empty(n_d)
    for d in range(n_d):
        f = np.zeros(n_t)
        for k in range(n_w):
            f += _autocorr_1d(chain[:, k, d])
        taus = 2.0 * np.cumsum(f / n_w) - 1.0
        m = np.arange(len(taus)) < c * taus
        tau[d] = taus[np.argmin(m) if np.any(~m) else len(taus) - 1]
    return tau


# ---------------------------------------------------------------------------------------------------------------------
# the sampler
# ---------------------------------------------------------------------------------------------------------------------

class DeviceState:
    """What the device sampler yields after an iteration: the ensemble, still on the device.  `coords` [n_walkers, D] and
    `log_prob` [n_walkers] (emcee's names) fetch host copies on demand; `x, lp = state` unpacks like the host sampler's pair."""

    def __init__(self, x, lp):
        self.x, self.lp = x, lp                          # float64 CUDA tensors [D, n_walkers], [n_walkers] (private copies)

    @property
    def coords(self):
        return self.x.T.contiguous().cpu().numpy()

    @property
    def log_prob(self):
        return self.lp.cpu().numpy()

    def __iter__(self):
        yield self.coords
        yield self.log_prob


class EnsembleSampler:
    """The part of `emcee.EnsembleSampler` the reference uses: `sample` / `run_mcmc`, `get_chain`, `get_log_prob`,
    `iteration`, `acceptance_fraction`, `get_autocorr_time`; `log_prob_fn` is scored through its `.batch` when it has one."""

    def __init__(self, nwalkers, ndim, log_prob_fn, moves=None, args=None, kwargs=None, seed=None, backend=None,
                 device_update=False):
        self.nwalkers, self.ndim = int(nwalkers), int(ndim)
        self._device_update = bool(device_update)
        if self._device_update:
            if not hasattr(log_prob_fn, "batch_device") or args or kwargs:
                raise TypeError("device_update=True needs an objective with `batch_device` (log_posterior_probability) "
                                "and no extra arguments")
            self._objective = log_prob_fn
        if self.nwalkers < 2 * self.ndim:
            raise ValueError("The number of walkers must be at least twice the number of parameters")
        args, kwargs = tuple(args or ()), dict(kwargs or {})
        if hasattr(log_prob_fn, "batch") and not args and not kwargs:
            self._score = lambda X: np.asarray(log_prob_fn.batch(np.array(X, dtype=np.float64)), dtype=np.float64)
        else:
            self._score = lambda X: np.array([log_prob_fn(np.array(x, dtype=np.float64), *args, **kwargs) for x in X],
                                             dtype=np.float64)
        moves = default_moves() if moves is None else moves
        if not isinstance(moves, (list, tuple)):
            moves = [(moves, 1.0)]
        self._moves = [m[0] if isinstance(m, (list, tuple)) else m for m in moves]
        w = np.array([m[1] if isinstance(m, (list, tuple)) else 1.0 for m in moves], dtype=np.float64)
        self._weights = w / w.sum()
        self.rng = np.random.default_rng(seed)
        self.backend = backend
        self._store = np.empty((0, self.nwalkers, self.ndim))    # capacity >= _n rows; rows [:_n] are the chain
        self._store_lp = np.empty((0, self.nwalkers))
        self._n = 0
        self._accepted_host = np.zeros(self.nwalkers)
        self._accepted_device = None                     # device sampler: the counts stay there until someone asks
        self._pending = None                             # device sampler: stored iterations not yet copied to the host
        self.n_batches = 0                               # objective launches so far
        self._half_updates = 0                           # Philox stream index of the device moves
        if backend is not None and os.path.exists(backend):
            with np.load(backend) as f:
                self._store, self._store_lp, self._accepted_host = f["chain"], f["log_prob"], f["accepted"]
            if self._store.shape[1:] != (self.nwalkers, self.ndim):
                raise ValueError(f"backend '{backend}' holds chains of shape {self._store.shape[1:]}, "
                                 f"expected {(self.nwalkers, self.ndim)}")
            self._n = self._store.shape[0]

    # -- storage ---------------------------------------------------------------------------------------------------
    @property
    def iteration(self):
        return self._n

    @property
    def _chain(self):
        self._flush()
        return self._store[:self._n]

    @property
    def _log_prob(self):
        self._flush()
        return self._store_lp[:self._n]

    def _flush(self):
        """Device sampler: bring the iterations that are still in HBM into the host chain (one bulk copy)."""
        p = self._pending
        if p is not None and p["count"]:
            import torch
            b, c = p["base"], p["count"]
            torch.from_numpy(self._store[b:b + c]).copy_(p["x"][:c])
            torch.from_numpy(self._store_lp[b:b + c]).copy_(p["lp"][:c])
            p["base"], p["count"] = b + c, 0

    @property
    def _accepted(self):
        if self._accepted_device is not None:
            self._accepted_host = self._accepted_device.cpu().numpy().astype(np.float64)
        return self._accepted_host

    @property
    def acceptance_fraction(self):
        return self._accepted / max(self.iteration, 1)

    def get_chain(self, discard=0, thin=1, flat=False):
        v = self._chain[discard + thin - 1::thin]
        return v.reshape(-1, self.ndim) if flat else v

    def get_log_prob(self, discard=0, thin=1, flat=False):
        v = self._log_prob[discard + thin - 1::thin]
        return v.reshape(-1) if flat else v

    def get_autocorr_time(self, discard=0, thin=1, c=5.0, **_ignored):
        return thin * integrated_time(self.get_chain(discard=discard, thin=thin), c=c)

    def save(self, filename=None):
        filename = filename or self.backend
        if filename:
            tmp = filename + ".tmp.npz"
            np.savez(tmp, chain=self._chain, log_prob=self._log_prob, accepted=self._accepted)
            os.replace(tmp, filename)

    # -- sampling --------------------------------------------------------------------------------------------------
    def _check_start(self, x):
        x = np.array(x, dtype=np.float64)
        if x.shape != (self.nwalkers, self.ndim):
            raise ValueError(f"incompatible input dimensions {x.shape}, expected {(self.nwalkers, self.ndim)}")
        if not np.all(np.isfinite(x)):
            raise ValueError("Initial state has non-finite coordinates")
        if np.linalg.cond(np.atleast_2d(np.cov(x, rowvar=False))) > 1e8:
            raise ValueError("Initial state has a large condition number. Make sure that your walkers are linearly "
                             "independent for the best performance")
        return x

    def sample(self, initial_state, iterations=1, store=True, progress=False, log_prob0=None, **_ignored):
        """Generator: one ensemble update (two half updates, two batched evaluations) per item."""
        x = self._check_start(initial_state)
        if self._device_update:                              # ranks of a torch.distributed job hold replicas of the ensemble
            from .. import sharding
            x = sharding.same_on_every_rank(x)
            self.rng = np.random.default_rng(sharding.agree_on(int(self.rng.integers(0, 2 ** 63))))
        if log_prob0 is None:
            lp = sharding.batch_sharded(self._objective, x) if self._device_update else self._score(x)
            self.n_batches += 1
        else:
            lp = np.array(log_prob0, dtype=np.float64)
        if np.any(np.isnan(lp)):
            raise ValueError("The initial log_prob was NaN")
        if store and self._store.shape[0] < self._n + iterations:
            self._store = np.concatenate([self._chain, np.empty((iterations, self.nwalkers, self.ndim))])
            self._store_lp = np.concatenate([self._log_prob, np.empty((iterations, self.nwalkers))])
        half = self.nwalkers // 2
        if self._device_update:
            yield from self._sample_on_device(x, lp, iterations, store, progress)
            return
        for it in range(iterations):
            move = self._moves[self.rng.choice(len(self._moves), p=self._weights)]
            order = self.rng.permutation(self.nwalkers)
            for mine, others in ((order[:half], order[half:]), (order[half:], order[:half])):
                q, log_factor = move.propose(x[mine], x[others], self.rng)
                new_lp = self._score(q)                     # may clip q in place, like the reference's objective
                self.n_batches += 1
                if np.any(np.isnan(new_lp)):
                    raise ValueError("Probability function returned NaN")
                with np.errstate(invalid="ignore"):
                    log_ratio = log_factor + new_lp - lp[mine]
                log_ratio[~np.isfinite(new_lp)] = -np.inf
                ok = log_ratio > np.log(self.rng.random(len(mine)))
                x[mine[ok]], lp[mine[ok]] = q[ok], new_lp[ok]
                self._accepted[mine[ok]] += 1
            if store:
                self._store[self._n], self._store_lp[self._n] = x, lp
                self._n += 1
            if progress:
                print(f"\riteration {it + 1}/{iterations}", end="" if it + 1 < iterations else "\n", file=sys.stderr)
            yield x, lp

    def _sample_on_device(self, x, lp, iterations, store, progress):
        """The ensemble lives on the device as float64 [D, n_walkers] (the layout `batch_device` reads): per half update one
        `sodm_walkers_propose` (stretch / DE / snooker / KDE draws; the KDE densities through `sodm_kde_logpdf`),
        `sodm_population_clip` (the reference objective's clip + -inf), the simulation and likelihood kernels, and
        `sodm_walkers_accept`.  Per iteration the host reads 4 bytes (the NaN flag); the stored chain stays in HBM and is copied
        in bulk when it is read (`get_chain`, `save`), when its buffer fills, and at the end; each iteration yields a
        `DeviceState`.  Random numbers: Philox
        keyed by (seed, walker, half-update counter), so a run does not depend on the launch geometry.  Under
        `torch.distributed` every rank holds a replica of the ensemble (D x n doubles), proposes and accepts redundantly
        from the same streams, and scores only its contiguous range of each half (`sharding.batch_device_sharded`: one
        all-gather of n/2 log-posteriors per half update)."""
        import torch
        from .. import engine, sharding
        dev = engine.device()
        obj = self._objective
        lb, ub = np.array([b[0] for b in obj.expanded_bounds], dtype=np.float64), np.array([b[1] for b in obj.expanded_bounds], dtype=np.float64)
        n, D, half = self.nwalkers, self.ndim, self.nwalkers // 2
        x_d = torch.from_numpy(np.ascontiguousarray(x.T)).to(dev)
        lp_d = torch.from_numpy(np.ascontiguousarray(lp)).to(dev)
        acc_d = torch.from_numpy(self._accepted.astype(np.int64)).to(dev)
        self._accepted_device = acc_d
        status = torch.zeros(1, dtype=torch.int32, device=dev)
        single = sharding.world()[1] == 1                    # ranks score their range of the proposals under one agreed seed
        gen = torch.Generator(device=dev)
        gen.manual_seed(int(self.rng.integers(0, 2 ** 63)))
        seed = int(self.rng.integers(0, 2 ** 63))
        if store:
            # the chain stays in HBM (at most ~4 GB of it) and reaches the host in bulk: when the buffer is full, when
            # someone reads the chain (get_chain, save, the convergence test of run_EnsembleSampler), and at the end
            cap = max(1, min(iterations, int(4e9 // (n * (D + 1) * 8))))
            self._pending = {"x": torch.empty((cap, n, D), dtype=torch.float64, device=dev),
                             "lp": torch.empty((cap, n), dtype=torch.float64, device=dev), "base": self._n, "count": 0, "cap": cap}
        try:
            yield from self._device_iterations(iterations, store, progress, x_d, lp_d, acc_d, status, gen, seed, single, lb, ub)
        finally:
            self._flush()
            self._pending = None

    def _device_iterations(self, iterations, store, progress, x_d, lp_d, acc_d, status, gen, seed, single, lb, ub):
        import torch
        from .. import engine, sharding
        dev, obj = x_d.device, self._objective
        n, D, half = self.nwalkers, self.ndim, self.nwalkers // 2
        for it in range(iterations):
            move = self._moves[self.rng.choice(len(self._moves), p=self._weights)]
            order = torch.randperm(n, device=dev, generator=gen)
            for mine, others in ((order[:half], order[half:]), (order[half:], order[:half])):
                mine, others = mine.contiguous(), others.contiguous()
                k = self._half_updates
                self._half_updates += 1
                if isinstance(move, StretchMove):
                    q, lf = engine.walkers_propose("stretch", x_d, mine, others, move.a, 0.0, seed, k)
                elif isinstance(move, DEMove):
                    q, lf = engine.walkers_propose("de", x_d, mine, others, 0.0 if move.gamma0 is None else move.gamma0,
                                                   move.sigma, seed, k)
                elif isinstance(move, DESnookerMove):
                    q, lf = engine.walkers_propose("snooker", x_d, mine, others, move.gammas, 0.0, seed, k)
                elif isinstance(move, KDEMove) and move.logpdf is None:
                    centres = others[:move.max_points] if move.max_points else others    # `others` is in random order
                    q, lf = engine.walkers_propose_kde(x_d, mine, centres, move.factor_of(x_d[:, centres].cpu().numpy().T),
                                                       seed, k)
                else:                                           # host proposal (user moves) from a copy of the ensemble
                    xh = x_d.cpu().numpy().T
                    qh, lfh = move.propose(xh[mine.cpu().numpy()], xh[others.cpu().numpy()], self.rng)
                    q = torch.from_numpy(np.ascontiguousarray(qh.T)).to(dev)
                    lf = torch.from_numpy(np.ascontiguousarray(lfh, dtype=np.float64)).to(dev)
                penalty = engine.population_clip(q, lb, ub)
                sim_seed = int(self.rng.integers(0, 2 ** 63))   # drawn in every mode: the host stream does not depend on the ranks
                new_lp = sharding.batch_device_sharded(obj, q, seed=None if single else sim_seed)
                self.n_batches += 1
                engine.walkers_accept(x_d, lp_d, mine, q, new_lp, lf, acc_d, penalty=penalty, seed=seed, iteration=k, status=status)
            if store:
                p = self._pending
                if p["count"] == p["cap"]:
                    self._flush()
                p["x"][p["count"]].copy_(x_d.T)              # [n, D], the chain's layout
                p["lp"][p["count"]].copy_(lp_d)
                p["count"] += 1
                self._n += 1
            if int(status.item()):                           # the iteration's only read from the device: 4 bytes
                raise ValueError("Probability function returned NaN")
            if progress:
                print(f"\riteration {it + 1}/{iterations}", end="" if it + 1 < iterations else "\n", file=sys.stderr)
            yield DeviceState(x_d.clone(), lp_d.clone())

    def run_mcmc(self, initial_state, nsteps, **kwargs):
        out = None
        for out in self.sample(initial_state, iterations=nsteps, **kwargs):
            pass
        return out


# ---------------------------------------------------------------------------------------------------------------------
# the reference's front end
# ---------------------------------------------------------------------------------------------------------------------

def _validate_settings_dict(settings_dict):
    """NetCDF-exportable values only (reference `mcmc.py:380-412`): int/float/bool/str, homogeneous non-empty lists
    of those, 1-D integer or float arrays."""
    for key, value in settings_dict.items():
        if isinstance(value, list):
            ok = bool(value) and type(value[0]) in (int, float, str, bool) and all(type(v) is type(value[0]) or
                                                                                 isinstance(v, type(value[0])) for v in value)
            why = "lists must be homogeneous and contain only int, float, bool or str."
        elif isinstance(value, np.ndarray):
            ok = value.ndim == 1 and value.dtype.kind in "if"
            why = "numpy arrays must be 1D and contain only int or float."
        else:
            ok = isinstance(value, (int, float, bool, str))
            why = ("valid types are: (scalar values) int, float, bool, str. (arrays) homogeneous lists containing int, "
                   "float, str, bool. 1D numpy arrays containing int or float.")
        if not ok:
            raise TypeError(f"Invalid datatype in `settings_dict` for key: '{key}' (type: {type(value)}). {why}")


def _directory(path):
    if not path:
        return os.getcwd() + "/"
    path = os.path.join(os.getcwd(), path if path.endswith("/") else path + "/")
    os.makedirs(path, exist_ok=True)
    return path


def run_EnsembleSampler(pos, max_n, identifier, objective_function, objective_function_args=None,
                        objective_function_kwargs=None, moves=None, fig_path=None, samples_path=None, print_n=10,
                        discard=0, thin=1, backend=None, processes=1, progress=True, settings_dict={}, seed=None,
                        device_update=False):
    """The reference's `run_EnsembleSampler` (`mcmc.py:15-231`): set up the sampler (fresh, or continued from `backend`),
    run at most `max_n` iterations, every `print_n` iterations write the samples file and test convergence (chain longer
    than 50 integrated autocorrelation times AND their mean changed by < 3 %).  `processes` and `fig_path` are accepted
    for compatibility (the evaluation is batched; figures are not drawn); `device_update=True` keeps the ensemble on the
    device (`EnsembleSampler._sample_on_device`).  Returns (sampler, samples)."""
    pos = np.asarray(pos, dtype=np.float64)
    samples_path = _directory(samples_path)
    run_date = str(datetime.date.today())
    nwalkers, ndim = pos.shape
    settings_dict = dict(settings_dict)
    settings_dict.update({"discard": discard, "thin": thin})
    _validate_settings_dict(settings_dict)
    print("\nMarkov-Chain Monte-Carlo sampling\n=================================\n")
    if not backend:
        fn_backend = samples_path + f"{identifier}_BACKEND_{run_date}.npz"
        if os.path.exists(fn_backend):
            os.remove(fn_backend)
        print(f"Created new backend: '{fn_backend}'\n\nStarting new run\n----------------\n")
    else:
        fn_backend = backend if os.path.exists(backend) else os.path.splitext(backend)[0] + ".npz"
        if not os.path.exists(fn_backend):
            raise FileNotFoundError("backend not found.")
        print(f"Found existing backend:'{fn_backend}'\n\nContinuing run\n--------------\n")
    sampler = EnsembleSampler(nwalkers, ndim, objective_function, moves=moves, args=objective_function_args,
                              kwargs=objective_function_kwargs, seed=seed, backend=fn_backend,
                              device_update=device_update)
    if backend:
        pos = sampler.get_chain()[-1]
    fn_samples = samples_path + f"{identifier}_SAMPLES_{run_date}.nc"
    print(f"Parameters: {ndim}\nMarkov chains: {nwalkers}\nBatched objective evaluation: {nwalkers // 2} walkers per call")
    print(f"Iterations: {max_n}" + (f" (found {sampler.iteration} previous iterations)" if backend else ""))
    print(f"Automatically checking convergence every {print_n} iterations")
    print(f"Saving samples every {print_n} iterations\nSamples: {fn_samples}\n")
    sys.stdout.flush()

    iteration_0, old_tau, samples, done = sampler.iteration, np.inf, None, 0
    for _ in sampler.sample(pos, iterations=max_n, progress=progress, store=True):
        done += 1
        if not ((sampler.iteration % print_n == 0) or (done == max_n)):
            continue
        sampler.save()
        samples = dump_samples(sampler.get_chain(discard=discard, thin=thin), objective_function.parameter_shapes,
                               filename=fn_samples, settings_dict=settings_dict)
        with np.errstate(invalid="ignore", divide="ignore"):
            tau = integrated_time(sampler.get_chain())
            converged = np.all(np.max(tau) * 50.0 < sampler.iteration)
            converged &= np.all(np.abs(np.mean(old_tau) - np.mean(tau)) / np.mean(tau) < 0.03)
        old_tau = tau
        if converged:
            print("Convergence: The chain is longer than 50 times the intergrated autocorrelation time.\n")
            break
        print("Non-convergence: The chain is shorter than 50 times the integrated autocorrelation time.\n")
        sys.stdout.flush()
    return sampler, samples


def perturbate_theta(theta, pert, multiplier=2, bounds=None, verbose=None):
    """Starting positions of `multiplier * len(theta)` chains: the estimate, clipped so that the perturbed values stay
    inside the bounds, times `1 + pert * U(-1, 1)`; redrawn while the matrix is singular (reference `mcmc.py:234-308`,
    same use of NumPy's global generator).  Returns (ndim, nwalkers, pos)."""
    if len(theta) != len(pert):
        raise Exception('The parameter value array "theta" must have the same length as the perturbation value array "pert".')
    if len(bounds) != len(theta):
        raise Exception("If bounds is not None, it must contain a tuple for every parameter in theta")
    theta = np.array(theta)
    low = [b[0] / (1 - p) for b, p in zip(bounds, pert)]
    high = [b[1] / (1 + p) for b, p in zip(bounds, pert)]
    ndim = len(theta)
    nwalkers = ndim * multiplier
    theta = np.clip(theta, low, high)
    for attempt in range(21):
        pos = theta + theta * pert * np.random.uniform(low=-1, high=1, size=(nwalkers, ndim))
        if np.linalg.cond(pos) != np.inf:
            break
        if verbose:
            print("Condition number too high, recalculating perturbations. Perhaps one or more of the bounds is zero?")
            sys.stdout.flush()
    else:
        raise Exception("Attempted 20 times to perturb parameter values but the condition number remains too large.")
    if verbose:
        print("Total number of markov chains: " + str(nwalkers) + "\n")
        sys.stdout.flush()
    return ndim, nwalkers, pos
