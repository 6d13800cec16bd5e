"""world_size-2 gloo test of the N>1 path: points are partitioned round-robin, each rank computes its own
rows, one all_gather returns them in point order on every rank."""
import os
import socket
import sys

import numpy as np
import torch.multiprocessing as mp

from conftest import ROOT


def _worker(rank, world, port, n_points, out_dir):
    sys.path.insert(0, os.path.join(ROOT, "tensor-networks-simple-update_b200"))
    import torch.distributed as dist

    from tnsu_b200.sharding import gather_points, partition

    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    idx = partition(n_points, world, rank)
    local = np.stack([idx * 1.5, -idx.astype(float), np.full(len(idx), rank)], axis=1)  # stand-in for (E, Mx, rank)
    full = gather_points(local, n_points, world, rank)
    np.save(os.path.join(out_dir, f"r{rank}.npy"), full)
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
