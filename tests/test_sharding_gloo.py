"""CPU: the N>1 host logic -- hash partition of the interactions and the delta all-reduce of
the replicated table -- with two gloo ranks."""
import os
import socket

import numpy as np
import pytest
import scipy.sparse as sp
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import helpers as H
from lightfm_b200 import sharding


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


@pytest.mark.parametrize("axis", ("item", "user"))
@pytest.mark.parametrize("world", (2, 4))
def test_partition_is_exact_cover(axis, world):
    inter = H.synthetic_interactions(300, 200, 5000, 3)
    w = np.random.default_rng(0).random(inter.nnz).astype(np.float32)
    seen = []
    for rank in range(world):
        local, lw, pos, smap = sharding.partition(inter, w, axis, rank, world)
        assert pos.has_sorted_indices and pos.shape == local.shape
        assert len(lw) == local.nnz
        if axis == "item":
            g_rows, g_cols = local.row, smap.global_ids[local.col]
            assert local.shape == (300, smap.n_local)
        else:
            g_rows, g_cols = smap.global_ids[local.row], local.col
            assert local.shape == (smap.n_local, 200)
        seen.append(np.stack([g_rows.astype(np.int64) * 200 + g_cols, np.round(lw * 1e6).astype(np.int64)]))
        # every owned id maps back to itself
        assert np.array_equal(smap.local_of[smap.global_ids], np.arange(smap.n_local))
    allk = np.concatenate(seen, axis=1)
    want = np.stack([inter.row.astype(np.int64) * 200 + inter.col, np.round(w * 1e6).astype(np.int64)])
    assert allk.shape == want.shape
    assert np.array_equal(allk[:, np.argsort(allk[0])], want[:, np.argsort(want[0])])


def test_shards_are_balanced():
    owner = sharding.shard_of(np.arange(1_000_000), 8)
    counts = np.bincount(owner, minlength=8)
    assert counts.min() > 0.97 * 125000 and counts.max() < 1.03 * 125000


def test_slice_and_merge_state_roundtrip():
    arrays = H.init_arrays(np.random.RandomState(0), 50, 40, 8)
    merged = {k: np.zeros_like(v) for k, v in arrays.items()}
    for rank in range(3):
        smap = sharding.ShardMap(50, rank, 3)
        local = sharding.slice_state(arrays, "item", smap)
        assert local["item_embeddings"].shape == (smap.n_local, 8)
        assert local["user_embeddings"].shape == (40, 8)
        sharding.merge_state(merged, local, "item", smap)
    for k in arrays:
        assert np.array_equal(merged[k], arrays[k]), k


def _worker(rank, world, port, out_dir):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rs = np.random.RandomState(0)
    base = torch.from_numpy(rs.rand(64, 8).astype(np.float32))
    acc0 = torch.ones(64, 8)
    w = base.clone()
    g = acc0.clone()
    # each rank applies its own "epoch": a rank-dependent delta on a rank-dependent row subset
    rows = torch.arange(rank, 64, world)
    w[rows] += 0.01 * (rank + 1)
    g[rows] += 0.5 * (rank + 1)
    sharding.allreduce_deltas([w, g], [base, acc0])
    np.save(os.path.join(out_dir, "w%d.npy" % rank), w.numpy())
    np.save(os.path.join(out_dir, "g%d.npy" % rank), g.numpy())
    dist.destroy_process_group()


def test_allreduce_deltas_world2(tmp_path):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    rs = np.random.RandomState(0)
    want_w = rs.rand(64, 8).astype(np.float32)
    want_g = np.ones((64, 8), np.float32)
    for rank in range(world):
        want_w[rank::world] += np.float32(0.01 * (rank + 1))
        want_g[rank::world] += np.float32(0.5 * (rank + 1))
    for rank in range(world):
        assert np.allclose(np.load(tmp_path / ("w%d.npy" % rank)), want_w, atol=1e-6)
        assert np.allclose(np.load(tmp_path / ("g%d.npy" % rank)), want_g, atol=1e-6)
