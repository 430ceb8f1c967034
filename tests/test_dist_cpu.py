"""CPU tests of the multi-GPU host logic with world_size-2 gloo process groups: sequence sharding,
row sharding, the all-reduce of k-means partials (combined result must equal the single-process Lloyd step), and
the tilemap gather."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tiler_b200 import dist as tdist


def test_shard_sequences_balanced_and_complete():
    counts = [30, 75, 12, 75, 40, 8, 60, 33]
    for world in (1, 2, 4, 8):
        shards = tdist.shard_sequences(counts, world)
        assert sorted(i for s in shards for i in s) == list(range(len(counts)))
        loads = [sum(counts[i] for i in s) for s in shards]
        assert max(loads) - min(loads) <= max(counts)
    assert tdist.shard_sequences([5], 4) == [[0], [], [], []]


def test_shard_rows_partition():
    for n in (0, 1, 7, 4194304):
        for world in (1, 2, 3, 8):
            edges = [tdist.shard_rows(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(edges[i][1] == edges[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _partials(x, cent, labels):
    """numpy stand-in for tm_kmeans_partial_step (the GPU kernel is covered by the -m gpu tests)."""
    d = ((x[:, None, :] - cent[None, :, :]) ** 2).sum(-1)
    new = d.argmin(1).astype(np.int32)
    k = cent.shape[0]
    sums = np.zeros_like(cent)
    np.add.at(sums, new, x)
    counts = np.bincount(new, minlength=k).astype(np.int64)
    return new, sums, counts, int((new != labels).sum()), float(d.min(1).sum())


def _worker(rank, world, port, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    rng = np.random.default_rng(3)
    x = rng.normal(0, 10, size=(1000, 6))
    cent = x[:5].copy()
    lo, hi = tdist.shard_rows(len(x), rank, world)
    labels = np.full(hi - lo, -1, np.int32)
    for _ in range(3):
        labels, sums, counts, changed, inertia = _partials(x[lo:hi], cent, labels)
        s, c, changed, inertia = tdist.allreduce_partials(torch.from_numpy(sums), torch.from_numpy(counts), changed, inertia)
        cent = np.where(c.numpy()[:, None] > 0, s.numpy() / np.maximum(c.numpy()[:, None], 1), cent)
    shards = tdist.shard_sequences([3, 5, 2, 4], world)
    local = {i: np.full(4, i, np.int32) for i in shards[rank]}
    merged = tdist.gather_tilemaps(local, shards)
    # per-frame rows sharded like PredictMotion in TilingEncoder.encode(sharded=True): 7 frames x 3 tiles over 2 ranks
    flo, fhi = tdist.shard_rows(7, rank, world)
    rows = tdist.gather_rows(np.arange(7 * 3, dtype=np.float32).reshape(7, 3)[flo:fhi], 7)
    assert tdist.world_info() == (rank, world) and np.array_equal(rows, np.arange(21, dtype=np.float32).reshape(7, 3))
    q.put((rank, cent, changed, inertia, sorted(merged)))
    dist.destroy_process_group()


def test_gloo_world2_kmeans_allreduce_matches_single_process():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    rng = np.random.default_rng(3)
    x = rng.normal(0, 10, size=(1000, 6))
    cent = x[:5].copy()
    labels = np.full(len(x), -1, np.int32)
    for _ in range(3):
        labels, sums, counts, changed, inertia = _partials(x, cent, labels)
        cent = np.where(counts[:, None] > 0, sums / np.maximum(counts[:, None], 1), cent)
    for rank, c, ch, inr, keys in res:
        assert np.allclose(c, cent, rtol=1e-12, atol=1e-12)
        assert ch == changed and abs(inr - inertia) <= 1e-9 * inertia
        assert keys == [0, 1, 2, 3]
