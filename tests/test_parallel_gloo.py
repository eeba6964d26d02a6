"""Multi-rank host logic on CPU (gloo, world_size 2): sharding, the all-gather of the scaler
partials in rank order and the variable-length row gather used at inference."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import idelucs_oracle as orc
    from idelucs_b200 import featurise as ft
    from idelucs_b200 import parallel
    rng = np.random.default_rng(0)
    X = (rng.random((5000, 64)) * 1e-3).astype(np.float32)
    X[:, 7] = X[0, 7]
    lengths = rng.integers(100, 5000, size=5000)
    ranges = parallel.shard_ranges(lengths, world)
    lo, hi = ranges[rank]
    parts, ns = orc.colstats_partials(X[lo:hi])
    gp, gn = ft.gather_partials(torch.from_numpy(parts), torch.from_numpy(ns))
    mean, var = orc.merge_partials(gp.numpy(), gn.numpy())
    rows = parallel.all_gather_rows(torch.from_numpy(X[lo:hi]), [h - l for l, h in ranges])
    # inference tail (idelucs/models.py:164-171): predictions, probabilities and latent rows of every rank, rank order
    yp = torch.arange(lo, hi, dtype=torch.int64) % 7
    pr = torch.from_numpy(X[lo:hi, 0].copy())
    lat = torch.from_numpy(X[lo:hi].copy())
    g_y, g_p, g_l = parallel.all_gather_ragged([yp, pr, lat])
    tail_ok = (torch.equal(g_y, torch.arange(0, 5000, dtype=torch.int64) % 7) and torch.equal(g_p, torch.from_numpy(X[:, 0].copy()))
               and torch.equal(g_l, torch.from_numpy(X)))
    ok = tail_ok and (np.allclose(mean, X.astype(np.float64).mean(axis=0), rtol=1e-13) and
          np.allclose(var, X.astype(np.float64).var(axis=0), rtol=1e-9, atol=1e-30) and var[7] == 0.0 and
          torch.equal(rows, torch.from_numpy(X)) and float(gn.sum()) == 5000.0)
    t = torch.tensor([1.0 if ok else 0.0])
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        ret["ok"] = bool(t.item() == 1.0)
        ret["ranges"] = ranges
    dist.destroy_process_group()


def test_two_rank_scaler_and_gather():
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(2, port, ret), nprocs=2, join=True)
    assert ret["ok"]
    (a, b), (c, d) = ret["ranges"]
    assert a == 0 and b == c and d == 5000 and 0 < b < 5000


def test_shard_ranges_properties():
    sys.path.insert(0, ROOT)
    from idelucs_b200 import parallel
    rng = np.random.default_rng(1)
    for world in (1, 2, 4, 8):
        for n in (0, 1, 7, 1000):
            lengths = rng.integers(1, 100000, size=n)
            r = parallel.shard_ranges(lengths, world)
            assert len(r) == world and r[0][0] == 0 and r[-1][1] == n
            assert all(r[i][1] == r[i + 1][0] for i in range(world - 1))
            if n >= 1000:
                cost = [int(lengths[lo:hi].sum()) for lo, hi in r]
                assert max(cost) < 1.5 * sum(cost) / world
