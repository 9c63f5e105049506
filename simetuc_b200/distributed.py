"""Multi-GPU plumbing: one process per GPU (`torchrun`), independent ODE systems dealt
round-robin, no collective on the hot path - only a final gather of populations / emission
curves.  Works with the `nccl` backend on GPUs and `gloo` on CPUs (tests)."""
import os
from typing import List, Sequence

import numpy as np


def world():
    """`(rank, world_size, local_rank)` from the torchrun environment (1 process -> 0, 1, 0)."""
    return (int(os.environ.get('RANK', '0')), int(os.environ.get('WORLD_SIZE', '1')),
            int(os.environ.get('LOCAL_RANK', '0')))


def shard_round_robin(n_systems: int, world_size: int, rank: int) -> List[int]:
    """Indices of the systems solved by `rank`: `system_id % world_size == rank`."""
    return list(range(rank, n_systems, world_size))


def unshard_round_robin(per_rank: Sequence[np.ndarray], n_systems: int) -> np.ndarray:
    """Inverse of `shard_round_robin` for gathered per-rank result arrays `[n_local, ...]`."""
    world_size = len(per_rank)
    first = next(a for a in per_rank if len(a))
    out = np.empty((n_systems,) + tuple(first.shape[1:]), dtype=first.dtype)
    for rank, arr in enumerate(per_rank):
        out[rank::world_size] = arr
    return out


def gather_results(local: np.ndarray, dst: int = 0):
    """Gather one array per rank on `dst` (list of arrays there, None elsewhere).  The arrays may
    differ in their first dimension."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [np.asarray(local)]
    rank, size = dist.get_rank(), dist.get_world_size()
    device = torch.device('cuda', torch.cuda.current_device()) \
        if dist.get_backend() == 'nccl' else torch.device('cpu')
    local = np.ascontiguousarray(local, dtype=np.float64)
    n_local = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    counts = [torch.zeros_like(n_local) for _ in range(size)]
    dist.all_gather(counts, n_local)
    n_max = int(max(int(c.item()) for c in counts))
    padded = torch.zeros((n_max,) + local.shape[1:], dtype=torch.float64, device=device)
    if local.shape[0]:
        padded[:local.shape[0]] = torch.from_numpy(local).to(device)
    bufs = [torch.zeros_like(padded) for _ in range(size)]
    dist.all_gather(bufs, padded)
    if rank != dst:
        return None
    return [b[:int(c.item())].cpu().numpy() for b, c in zip(bufs, counts)]


def max_over_ranks(values: Sequence[float]) -> List[float]:
    """Element-wise maximum over ranks (timings are reported as the slowest rank's)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return [float(v) for v in values]
    device = torch.device('cuda', torch.cuda.current_device()) \
        if dist.get_backend() == 'nccl' else torch.device('cpu')
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(v) for v in t.tolist()]
