"""Sweep-point sharding over the GPUs of one box.

Independent points (fields h_k, couplings, angles, seeds) never exchange data: each rank runs its own
`SimpleUpdateBatch` on its GPU and the ONLY collective is one all_gather of the per-point results
(a few doubles per point) at the end.  `torch.distributed` is plumbing (NCCL on GPUs, gloo in the CPU
tests); a single network does not shard at these sizes ("replicas only").
"""
from __future__ import annotations

import numpy as np


def partition(n_points: int, world_size: int, rank: int, costs=None) -> np.ndarray:
    """Point indices of `rank`.

    Without `costs`: round-robin, so neighbouring points of a sweep (similar cost: iteration counts vary smoothly
    with the swept parameter) spread evenly over the ranks.  With `costs` (predicted sweeps per point — iteration
    counts vary 139-644 across a TFI sweep): cost-sorted placement — the points are sorted by
    predicted cost, longest first, and dealt to the rank with the least accumulated cost so far (longest processing
    time first); every rank evaluates the same deterministic assignment, and a rank's list comes back longest first,
    which is also the order in which its GPU should start them."""
    if not (0 <= rank < world_size):
        raise ValueError(f"rank {rank} outside world of {world_size}")
    if costs is None:
        return np.arange(rank, n_points, world_size)
    costs = np.asarray(costs, dtype=np.float64)
    if costs.shape != (n_points,):
        raise ValueError(f"costs must have one entry per point ({n_points}), got shape {costs.shape}")
    order = np.argsort(-costs, kind="stable")
    load = np.zeros(world_size)
    count = np.zeros(world_size, dtype=np.int64)
    cap = (n_points + world_size - 1) // world_size  # keep the per-rank batches equal in size (gather buffer, engine batch)
    mine = []
    for k in order:
        open_ranks = np.nonzero(count < cap)[0]
        r = int(open_ranks[np.argmin(load[open_ranks])])
        load[r] += costs[k]
        count[r] += 1
        if r == rank:
            mine.append(int(k))
    return np.asarray(mine, dtype=np.int64)


def predict_costs(known_x, known_cost, query_x) -> np.ndarray:
    """Predicted cost of the points `query_x` of a sweep from a coarser sweep of the same parameter (linear
    interpolation of e.g. the sweeps each point needed): the pilot a cost-sorted partition of a larger sweep uses."""
    kx, kc = np.asarray(known_x, dtype=np.float64), np.asarray(known_cost, dtype=np.float64)
    o = np.argsort(kx)
    return np.interp(np.asarray(query_x, dtype=np.float64), kx[o], kc[o])


def gather_points(local_values: np.ndarray, n_points: int, world_size: int, rank: int, device=None, costs=None) -> np.ndarray:
    """All-gather per-point rows computed under `partition` back into point order.

    local_values: [n_local, k] float64 rows for partition(n_points, world_size, rank, costs).  Returns
    [n_points, k] on every rank.  With world_size == 1 no process group is needed."""
    local_values = np.asarray(local_values, dtype=np.float64)
    if local_values.ndim == 1:
        local_values = local_values[:, None]
    k = local_values.shape[1]
    if world_size == 1:
        full = np.zeros((n_points, k), dtype=np.float64)
        full[partition(n_points, 1, 0, costs)] = local_values
        return full
    import torch
    import torch.distributed as dist

    n_max = (n_points + world_size - 1) // world_size
    buf = torch.zeros((n_max, k), dtype=torch.float64, device=device if device is not None else "cpu")
    buf[: local_values.shape[0]] = torch.from_numpy(local_values).to(buf.device)
    out = [torch.zeros_like(buf) for _ in range(world_size)]
    dist.all_gather(out, buf)
    full = np.zeros((n_points, k), dtype=np.float64)
    for r in range(world_size):
        idx = partition(n_points, world_size, r, costs)
        full[idx] = out[r][: len(idx)].cpu().numpy()
    return full
