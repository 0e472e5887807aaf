"""Multi-GPU host logic: one process per GPU, environments sharded by index range, no data-path collective.

The reference's only parallelism is ``hyperparam_tuning.py:42-91``: one OS process per (hyper-parameter point,
seed), no communication, results written to separate directories.  Here the (point, seed) grid is the environment
axis; each rank owns a contiguous slice of it and the only collectives are the end-of-run reductions of timing
(MAX over ranks) and work counters (SUM) -- ``torch.distributed`` with NCCL on the GPU box, gloo in the CPU tests.
"""
from __future__ import annotations

import itertools
import os
from typing import Dict, Optional, Sequence, Tuple

import numpy as np


def world() -> Tuple[int, int, int]:
    """(rank, world_size, local_rank) from the torchrun environment."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_total: int, rank: int, world_size: int) -> Tuple[int, int]:
    """[lo, hi) of the global environment indices owned by ``rank``: contiguous, sizes differ by at most one."""
    if not (0 <= rank < world_size):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def grid_points(hyperparams: Dict[str, Sequence[float]], seeds: Sequence[int]) -> Dict[str, np.ndarray]:
    """hyperparam_tuning.py:42-48: ``product(*hyperparams.values())`` x ``random_seeds`` as per-environment arrays
    (point-major, seed-minor, the order in which the reference launches its processes)."""
    names = list(hyperparams)
    rows = [(*pt, sd) for pt in itertools.product(*[hyperparams[n] for n in names]) for sd in seeds]
    arr = np.array(rows, dtype=np.float64).reshape(len(rows), len(names) + 1)
    out = {n: arr[:, i].copy() for i, n in enumerate(names)}
    out["seeds"] = arr[:, -1].astype(np.uint64)
    return out


def shard_grid(grid: Dict[str, np.ndarray], rank: int, world_size: int) -> Dict[str, np.ndarray]:
    n = len(next(iter(grid.values())))
    lo, hi = shard_range(n, rank, world_size)
    return {k: v[lo:hi] for k, v in grid.items()}


def reduce_run(dist, times_ms: Sequence[float], counts: Sequence[int], device=None):
    """End-of-run reduction: element-wise MAX of the timings and SUM of the integer work counters over all ranks.
    ``dist`` is ``torch.distributed`` (initialised) or None for a single process."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return [float(x) for x in times_ms], [int(x) for x in counts]
    import torch
    t = torch.tensor(list(times_ms), dtype=torch.float64, device=device)
    c = torch.tensor(list(counts), dtype=torch.int64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dist.all_reduce(c, op=dist.ReduceOp.SUM)
    return [float(x) for x in t.tolist()], [int(x) for x in c.tolist()]


def gather_metrics(dist, local: np.ndarray, device=None) -> Optional[np.ndarray]:
    """Concatenate per-environment metric rows (first axis = local environments) on rank 0; None elsewhere."""
    if dist is None or not dist.is_initialized() or dist.get_world_size() == 1:
        return local
    import torch
    ws, rank = dist.get_world_size(), dist.get_rank()
    n = torch.tensor([local.shape[0]], dtype=torch.int64, device=device)
    sizes = [torch.zeros_like(n) for _ in range(ws)]
    dist.all_gather(sizes, n)
    sizes = [int(s.item()) for s in sizes]
    m = max(sizes)
    pad = np.zeros((m,) + local.shape[1:], local.dtype)
    pad[:local.shape[0]] = local
    t = torch.from_numpy(np.ascontiguousarray(pad).view(np.uint8).reshape(-1)).to(device) if device is not None else \
        torch.from_numpy(np.ascontiguousarray(pad).view(np.uint8).reshape(-1))
    bufs = [torch.zeros_like(t) for _ in range(ws)]
    dist.all_gather(bufs, t)
    if rank != 0:
        return None
    parts = [b.cpu().numpy().view(local.dtype).reshape((m,) + local.shape[1:])[:sizes[i]] for i, b in enumerate(bufs)]
    return np.concatenate(parts, axis=0)
