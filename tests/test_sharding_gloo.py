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
        full = np.load(tmp_path / f"full{r}.npy")
        assert full.shape == (n, 4) and (full[:, 0] == np.arange(n)).all()
        cols = np.load(tmp_path / f"cols{r}.npy")
        assert cols.shape == (4, n) and (cols[0] == np.arange(n)).all()
