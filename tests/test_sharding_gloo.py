"""The N > 1 path on CPU: world_size-2 `gloo` job exercising the shard arithmetic and the final gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pysodm_b200 import sharding


def test_shard_ranges_partition_exactly():
    for n in (0, 1, 7, 10 ** 6, 10 ** 6 + 3):
        for g in (1, 2, 4, 8):
            ranges = [sharding.shard_range(n, r, g) for r in range(g)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(ranges, ranges[1:]))
            sizes = [hi - lo for lo, hi in ranges]
            assert max(sizes) - min(sizes) <= 1


class _FakeObjective:
    """Stands in for log_posterior_probability.batch (which needs a GPU): lp = f(theta) row-wise."""

    def batch(self, thetas):
        return (np.asarray(thetas) ** 2).sum(axis=1)


    def batch_device(self, thetas, seed=None, member_id0=0):
        """[D, P] -> lp[P]; a stochastic model's term depends on (seed, GLOBAL member index) only."""
        member = torch.arange(member_id0, member_id0 + thetas.shape[1], dtype=torch.float64)
        return (thetas ** 2).sum(dim=0) + 1e-3 * member + 1e-9 * float(seed % 1000)


def _worker(rank, port, n, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=2)
    try:
        rng = np.random.default_rng(0)
        thetas = rng.normal(size=(n, 3))
        lp = sharding.batch_sharded(_FakeObjective(), thetas)
        lo, hi = sharding.shard_range(n)
        local = torch.arange(lo, hi, dtype=torch.int64)[:, None] * torch.ones(1, 4, dtype=torch.int64)
        full = sharding.all_gather_concat(local, n)
        cols = sharding.all_gather_concat(local.T.contiguous(), n, axis=1)
        # device-resident population replicated on the ranks: each rank scores its columns, one agreed seed
        np.random.seed(100 + rank)                            # the ranks' own generators disagree
        lp_dev = sharding.batch_device_sharded(_FakeObjective(), torch.from_numpy(np.ascontiguousarray(thetas.T)))
        np.save(os.path.join(out_dir, f"lpdev{rank}.npy"), lp_dev.numpy())
        start = sharding.same_on_every_rank(np.full(3, float(rank)))
        assert np.array_equal(start, np.zeros(3))
        np.save(os.path.join(out_dir, f"lp{rank}.npy"), lp)
        np.save(os.path.join(out_dir, f"full{rank}.npy"), full.numpy())
        np.save(os.path.join(out_dir, f"cols{rank}.npy"), cols.numpy())
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n", [10, 11])
def test_two_rank_gloo_gather(tmp_path, n):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_worker, args=(port, n, str(tmp_path)), nprocs=2, join=True)
    want = (np.random.default_rng(0).normal(size=(n, 3)) ** 2).sum(axis=1)
    for r in range(2):
        np.testing.assert_allclose(np.load(tmp_path / f"lp{r}.npy"), want)
        lpdev = np.load(tmp_path / f"lpdev{r}.npy")
        seed_term = lpdev - want - 1e-3 * np.arange(n)        # one seed for every member, the same on both ranks
        assert np.allclose(seed_term, seed_term[0], atol=1e-12) and 0 <= seed_term[0] < 1e-6 + 1e-12
        assert np.array_equal(lpdev, np.load(tmp_path / "lpdev0.npy"))
        full = np.load(tmp_path / f"full{r}.npy")
        assert full.shape == (n, 4) and (full[:, 0] == np.arange(n)).all()
        cols = np.load(tmp_path / f"cols{r}.npy")
        assert cols.shape == (4, n) and (cols[0] == np.arange(n)).all()


# ---- sharded order statistics: the exchange protocol on CPU -------------------------------------------------------------
class _HostPhases:
    """NumPy stand-in for `engine.SelectPhases` (the CUDA phases need a GPU): same state, same four phases, same tables
    -- TEST INFRASTRUCTURE that lets the world-size-2 gloo job exercise `sharding.order_stats_sharded`'s protocol (what
    is all-reduced, with which operator, in which order) without a device.  int32 ensembles only."""
    BINS, DIGIT = 2048, 11

    def __init__(self, ens, global_ranks, dev=None):
        self.x = ens.numpy().astype(np.int64).reshape(-1, ens.shape[-1])
        self.shape = tuple(ens.shape[:-1])
        nc, J = self.x.shape[0], len(global_ranks)
        self.lo, self.hi = torch.zeros(nc, dtype=torch.int64), torch.zeros(nc, dtype=torch.int64)
        self.nan_flag = torch.zeros(nc, dtype=torch.int32)
        self.settled = np.zeros(nc, dtype=np.int64)
        self.prefix = np.zeros((nc, J), dtype=np.int64)
        self.remaining = np.tile(np.asarray(global_ranks, dtype=np.int64), (nc, 1))
        self.hist = torch.zeros((nc, J, self.BINS), dtype=torch.int32)
        self.key_bits = 32

    def _bits(self):
        rng = (self.hi - self.lo).numpy()
        return np.array([int(r).bit_length() for r in rng])

    def run(self, phase, out=None, stream=None):
        from pysodm_b200 import _ffi
        nc, J = self.prefix.shape
        if phase == _ffi.SELECT_MINMAX:
            if self.x.shape[1]:
                self.lo[:] = torch.from_numpy(self.x.min(axis=1))
                self.hi[:] = torch.from_numpy(self.x.max(axis=1))
            else:
                self.lo[:], self.hi[:] = torch.iinfo(torch.int64).max, torch.iinfo(torch.int64).min
            return
        total = self._bits()
        for c in range(nc):
            if self.settled[c] >= total[c]:
                continue
            bits = min(self.DIGIT, total[c] - self.settled[c])
            shift = total[c] - self.settled[c] - bits
            key = self.x[c] - int(self.lo[c])
            digit = (key >> shift) & ((1 << bits) - 1)
            high = key >> (shift + bits)
            if phase == _ffi.SELECT_HISTOGRAM:
                self.hist[c] = 0
                taken = np.zeros(key.shape, bool)
                for j in range(J):
                    hit = (np.ones(key.shape, bool) if self.settled[c] == 0 else (high == self.prefix[c, j])) & ~taken
                    if self.settled[c] == 0 and j > 0:
                        break
                    self.hist[c, j] = torch.from_numpy(np.bincount(digit[hit], minlength=self.BINS).astype(np.int32))
                    taken |= hit
            elif phase == _ffi.SELECT_PICK:
                new = self.prefix[c].copy()
                for j in range(J):
                    slot = 0 if self.settled[c] == 0 else int(np.flatnonzero(self.prefix[c] == self.prefix[c, j])[0])
                    cum = np.cumsum(self.hist[c, slot].numpy().astype(np.int64))
                    b = int(np.searchsorted(cum, self.remaining[c, j], side="right"))
                    self.remaining[c, j] -= cum[b - 1] if b else 0
                    new[j] = ((self.prefix[c, j] << bits) if self.settled[c] else 0) | b
                self.prefix[c] = new
                self.settled[c] += bits

    def result(self):
        key = self.lo.numpy()[:, None] + np.where(self.settled[:, None] > 0, self.prefix, 0)
        return torch.from_numpy(key.T.astype(np.float64).reshape((self.prefix.shape[1],) + self.shape))


def _select_worker(rank, port, out_dir):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=2)
    try:
        from pysodm_b200 import engine
        engine.SelectPhases = _HostPhases                     # no GPU in this job: the phases run in NumPy
        rng = np.random.default_rng(5)
        n = 4001
        full = np.stack([rng.integers(6_000_000, 7_400_000, n), rng.integers(0, 5, n), np.full(n, 17),
                         rng.integers(-2 ** 31, 2 ** 31 - 1, n), rng.integers(0, 3000, n)]).astype(np.int32).reshape(5, 1, n)
        lo, hi = sharding.shard_range(n)
        ranks = [0, 100, n // 2, n - 2, n - 1, 7]             # six ranks: two searches of at most four
        got = sharding.order_stats_sharded(torch.from_numpy(full[..., lo:hi].copy()), ranks)
        s = sharding.sum_sharded(torch.from_numpy(full[..., lo:hi].astype(np.float64).sum(axis=-1)))
        np.save(os.path.join(out_dir, f"os{rank}.npy"), got.numpy())
        np.save(os.path.join(out_dir, f"sum{rank}.npy"), s.numpy())
        np.save(os.path.join(out_dir, "full.npy"), full)
        assert sharding.agree_on(1000 + rank) == 1000
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_order_statistics_protocol(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    mp.spawn(_select_worker, args=(port, str(tmp_path)), nprocs=2, join=True)
    full = np.load(tmp_path / "full.npy")
    ranks = [0, 100, 2000, 3999, 4000, 7]
    want = np.moveaxis(np.sort(full, axis=-1)[..., ranks], -1, 0).astype(np.float64)
    for r in range(2):
        assert np.array_equal(np.load(tmp_path / f"os{r}.npy"), want)
        assert np.array_equal(np.load(tmp_path / f"sum{r}.npy"), full.astype(np.float64).sum(axis=-1))
