"""world_size-2 gloo test of the N>1 path: points are partitioned round-robin, each rank computes its own
rows, one all_gather returns them in point order on every rank."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _costs(n_points):
    """A sweep-like cost profile: iteration counts peak around the phase transition."""
    x = np.linspace(0.0, 4.0, n_points)
    return 140.0 + 500.0 * np.exp(-((x - 3.0) / 0.4) ** 2)


def _worker(rank, world, port, n_points, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tensor-networks-simple-update_b200"))
    import torch.distributed as dist

    from tnsu_b200.sharding import gather_points, partition

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    for tag, costs in (("", None), ("c", _costs(n_points))):
        idx = partition(n_points, world, rank, costs)
        local = np.stack([idx * 1.5, -idx.astype(float), np.full(len(idx), rank)], axis=1)  # stand-in for (E, Mx, rank)
        full = gather_points(local, n_points, world, rank, costs=costs)
        np.save(os.path.join(out_dir, f"r{rank}{tag}.npy"), full)
    dist.destroy_process_group()


def test_gather_two_ranks(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_points, world = 11, 2
    mp.spawn(_worker, args=(world, port, n_points, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "r0.npy"), np.load(tmp_path / "r1.npy")
    assert np.array_equal(a, b)
    assert np.array_equal(a[:, 0], np.arange(n_points) * 1.5)
    assert np.array_equal(a[:, 1], -np.arange(n_points, dtype=float))
    assert np.array_equal(a[:, 2], np.arange(n_points) % world)


def test_cost_sorted_partition_balances_and_covers(tmp_path):
    """Cost-sorted placement: every point lands on exactly one rank, batches are equal in size, each rank's
    list is longest first, and the predicted load is far better balanced than contiguous blocks and at least as good
    as round-robin; the gathered result is in point order on both ranks of the gloo test above."""
    from tnsu_b200.sharding import partition, predict_costs

    n_points, world = 256, 8
    costs = _costs(n_points)
    parts = [partition(n_points, world, r, costs) for r in range(world)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(n_points))
    assert {len(p) for p in parts} == {n_points // world}
    for p in parts:
        assert np.all(np.diff(costs[p]) <= 0)
    load = np.array([costs[p].sum() for p in parts])
    rr = np.array([costs[partition(n_points, world, r)].sum() for r in range(world)])
    blocks = np.array([costs[r * 32:(r + 1) * 32].sum() for r in range(world)])
    assert load.max() / load.mean() < 1.01 and load.max() <= rr.max() + 1e-9 and blocks.max() / blocks.mean() > 1.5
    # the pilot: costs of a fine sweep interpolated from a coarse one
    fine = predict_costs(np.linspace(0, 4, 64), _costs(64), np.linspace(0, 4, 1024))
    assert fine.shape == (1024,) and abs(fine.max() - _costs(1024).max()) < 25.0


def test_gather_two_ranks_cost_sorted(tmp_path):
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    n_points, world = 11, 2
    mp.spawn(_worker, args=(world, port, n_points, str(tmp_path)), nprocs=world, join=True)
    a, b = np.load(tmp_path / "r0c.npy"), np.load(tmp_path / "r1c.npy")
    assert np.array_equal(a, b)
    assert np.array_equal(a[:, 0], np.arange(n_points) * 1.5)
    assert sorted(a[:, 2].tolist()).count(0.0) == 6 and sorted(a[:, 2].tolist()).count(1.0) == 5
