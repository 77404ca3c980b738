"""Multi-GPU plumbing: one process per GPU, independent replicas.

Requests never cross GPUs: the router (rank 0, device-resident state) assigns each admitted request to a
replica, the assignment vector is broadcast, every rank serves its own share.  The only data-path
collective is the start-up weight broadcast (models.broadcast_weights).  Works with NCCL (cuda tensors)
and gloo (cpu tensors; used by the CPU tests).
"""
from __future__ import annotations

from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.distributed as dist


def assignments_from_decisions(decisions: Sequence[Tuple[int, int, int, int]], replica_of: Sequence[int]) -> np.ndarray:
    """(status, deployment, group, chain_pos) per request -> replica index per request (-1 = rejected)."""
    out = np.full(len(decisions), -1, dtype=np.int32)
    for i, (status, dep, _g, _p) in enumerate(decisions):
        if status == 0 and dep >= 0:
            out[i] = replica_of[dep]
    return out


def scatter_assignments(decisions, n_req: int, world: int, rank: int, device, replica_of: Sequence[int],
                        group=None) -> np.ndarray:
    """Rank 0 passes the K1 decisions (others pass None); returns the request indices this rank serves."""
    t = torch.empty(n_req, dtype=torch.int32, device=device)
    if rank == 0:
        t.copy_(torch.from_numpy(assignments_from_decisions(decisions, replica_of)))
    if world > 1:
        dist.broadcast(t, src=0, group=group)
    a = t.cpu().numpy()
    return np.nonzero(a == rank)[0]


def gather_done(n_mine: int, world: int, rank: int, device, group=None) -> int:
    """Completion barrier: total number of requests finished across ranks."""
    t = torch.tensor([n_mine], dtype=torch.int64, device=device)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return int(t.item())
