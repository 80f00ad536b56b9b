"""world_size-2 gloo test of the only cross-rank exchange on the path: merging rnd statistics."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _stats_numpy(r, mode, max_rnd):
    r = r.astype(np.float64)
    keep = np.isfinite(r) if mode == 0 else (r < max_rnd if mode == 1 else np.ones_like(r, bool))
    k = r[keep]
    mx = (-k).max() if k.size else -np.inf
    return np.array([k.size, k.sum(), (k * k).sum(), mx, np.exp(-k - mx).sum() if k.size else 0.0, r.size, 0, 0])


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from sde_sampler_b200.dist import combine_stats, shard_range

    rng = np.random.default_rng(0)
    full = (rng.standard_normal(4096) * 3 + 100).astype(np.float32)
    full[5] = np.inf
    full[4000] = 2e8
    lo, hi = shard_range(full.size, rank, world)
    local = torch.from_numpy(_stats_numpy(full[lo:hi], 1, 1e8))
    merged = combine_stats(local, dist.group.WORLD)
    q.put((rank, merged.numpy().tolist(), (lo, hi)))
    dist.barrier()
    dist.destroy_process_group()


def test_merge_stats_two_ranks():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    rng = np.random.default_rng(0)
    full = (rng.standard_normal(4096) * 3 + 100).astype(np.float32)
    full[5] = np.inf
    full[4000] = 2e8
    want = _stats_numpy(full, 1, 1e8)
    assert sorted(r[2] for r in res) == [(0, 2048), (2048, 4096)]
    for _, got, _ in res:
        np.testing.assert_allclose(got[:6], want[:6], rtol=1e-12)
    # the quantities the loss / log-Z are built from
    n, s1, s2, mx, se = want[:5]
    var = (s2 - s1 * s1 / n) / (n - 1)
    kept = full[full < 1e8].astype(np.float64)
    assert var == pytest.approx(kept.var(ddof=1), rel=1e-9)
    assert np.log(se / n) + mx == pytest.approx(np.log(np.mean(np.exp(-kept - (-kept).max()))) + (-kept).max(), rel=1e-12)


def test_merge_handles_empty_rank():
    from sde_sampler_b200.dist import merge_stats

    a = torch.tensor([3.0, 6.0, 14.0, -1.0, 1.2, 4.0, 0, 0], dtype=torch.float64)
    b = torch.tensor([0.0, 0.0, 0.0, -float("inf"), 0.0, 2.0, 0, 0], dtype=torch.float64)
    m = merge_stats(torch.stack([a, b]))
    assert m[0] == 3 and m[3] == -1.0 and m[4] == pytest.approx(1.2) and m[5] == 6
