"""Sweep-point sharding over the GPUs of one box.

Independent points (fields h_k, couplings, angles, seeds) never exchange data: each rank runs its own
`SimpleUpdateBatch` on its GPU and the ONLY collective is one all_gather of the per-point results
(a few doubles per point) at the end.  `torch.distributed` is plumbing (NCCL on GPUs, gloo in the CPU
tests); a single network does not shard at these sizes ("replicas only").
"""
from __future__ import annotations

import numpy as np


def partition(n_points: int, world_size: int, rank: int) -> np.ndarray:
    """Point indices of `rank`: round-robin, so neighbouring points of a sweep (similar cost: iteration counts
    vary smoothly with the swept parameter) spread evenly over the ranks."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    return np.arange(rank, n_points, world_size)


def gather_points(local_values: np.ndarray, n_points: int, world_size: int, rank: int, device=None) -> np.ndarray:
    """All-gather per-point rows computed under `partition` back into point order.

    local_values: [n_local, k] float64 rows for partition(n_points, world_size, rank).  Returns
    [n_points, k] on every rank.  With world_size == 1 no process group is needed."""
    local_values = np.asarray(local_values, dtype=np.float64)
    if local_values.ndim == 1:
        local_values = local_values[:, None]
    k = local_values.shape[1]
    if world_size == 1:
        return local_values.copy()
    import torch
    import torch.distributed as dist

    n_max = (n_points + world_size - 1) // world_size
    buf = torch.zeros((n_max, k), dtype=torch.float64, device=device if device is not None else "cpu")
    buf[: local_values.shape[0]] = torch.from_numpy(local_values).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world_size)]
    dist.all_gather(out, buf)
    full = np.zeros((n_points, k), dtype=np.float64)
    for r in range(world_size):
        idx = partition(n_points, world_size, r)
        full[idx] = out[r][: len(idx)].cpu().numpy()
    return full
