"""world_size-2 gloo test (CPU) of the data-parallel path's host logic: the row-pair
sharding rule and the sum-allreduce contract.  Each rank computes the oracle gradient of
its shard (eps rows taken from the single threefry stream by the shard's row indices),
scales by n_local / B_global, allreduces, and must reproduce the full-batch gradient."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import datasets as ods
from oracle import jax_random as jr
from oracle import vae_step as vs
from vae_training_b200 import dp


def test_shard_rows_cover_and_pair():
    for B, W in [(100, 1), (100, 2), (100, 3), (65536, 8), (10, 4), (6, 8)]:
        got = []
        for r in range(W):
            lo, hi = dp.shard_rows(B, r, W)
            assert 0 <= lo <= hi <= B // 2
            got.append(dp.shard_row_indices(B, lo, hi))
        allrows = np.sort(np.concatenate(got))
        assert allrows.tolist() == list(range(B))
        sizes = [len(g) for g in got]
        assert max(sizes) - min(sizes) <= 2
    with pytest.raises(Exception):
        dp.shard_rows(101, 0, 2)


def test_pair_rows_share_threefry_blocks():
    """Row b and row b + B/2 are produced by the same blocks (the reason for the rule)."""
    B, C = 10, 7
    n = B * C
    half = n // 2
    for b in range(B // 2):
        elems = np.arange(b * C, (b + 1) * C)
        partner = elems + half
        assert (partner // C == b + B // 2).all() and (partner % C == elems % C).all()


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        B = 20
        d = vs.VaeDesc(B, 6, 4, [12], [12], epsilon0=-2.0)
        flat = vs.init_params(d, 3)
        ds = ods.SphereDataset(69, 3, 3)
        x = ds.get_batch(B)
        z = jr.normal(jr.PRNGKey(5), (B, 10))
        lo, hi = dp.shard_rows(B, rank, world)
        idx = dp.shard_row_indices(B, lo, hi)
        dl = vs.VaeDesc(len(idx), 6, 4, [12], [12], epsilon0=-2.0)
        loss, g = vs.loss_and_grad(dl, flat.astype(np.float64), x[idx], z[idx, :4], z[idx, 4:], dtype=np.float64)
        buf = torch.from_numpy(np.concatenate([g, [loss]]) * (len(idx) / B))
        dp.allreduce_sum(buf)
        lf, gf = vs.loss_and_grad(d, flat.astype(np.float64), x, z[:, :4], z[:, 4:], dtype=np.float64)
        err = float(np.max(np.abs(buf.numpy()[:-1] - gf)) / np.max(np.abs(gf)))
        lerr = abs(float(buf[-1]) - lf) / abs(lf)
        out[rank] = (err, lerr)
    finally:
        dist.destroy_process_group()


def test_gloo_world2_sharded_gradient_equals_full():
    world = 2
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, _free_port(), out), nprocs=world, join=True)
    assert len(out) == world
    for r in range(world):
        err, lerr = out[r]
        assert err <= 1e-12 and lerr <= 1e-12
