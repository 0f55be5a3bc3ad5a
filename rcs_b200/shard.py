"""Sharding of a state batch over the ranks of one node (one process per GPU).

States are independent and the model is tiny and replicated, so the batch is cut into
contiguous slices and there is no data-path collective; the only exchange is one all-gather
of compact per-rank summary statistics, or of compact per-state results of equal
slices (configs[4]: the hand-tip frames of every state) — NCCL over NVLink on the GPU box, gloo
in the CPU tests.
"""
from __future__ import annotations

from typing import Dict, Tuple

import numpy as np

STAT_FIELDS = ("count", "sum", "sum_sq", "min", "max")


def slice_of(rank: int, world: int, n_global: int) -> Tuple[int, int]:
    """[first, first + count) of `rank`: contiguous, sizes differ by at most one state."""
    if not (0 <= rank < world):
        raise ValueError(f"rank {rank} outside [0, {world})")
    base, extra = divmod(n_global, world)
    first = rank * base + min(rank, extra)
    return first, base + (1 if rank < extra else 0)


def local_stats(values) -> "np.ndarray | object":
    """Summary statistics of a 1-D per-state result (numpy array or torch tensor):
    [count, sum, sum of squares, min, max] as float64, on the array's device."""
    if type(values).__module__.startswith("torch"):
        import torch

        if values.is_cuda and values.dtype == torch.float64 and values.dim() == 1 and values.numel() > 0:
            from .api import summary_stats

            return summary_stats(values)          # one librcsb kernel instead of five torch ops
        v = values.reshape(-1).to(torch.float64)
        if v.numel() == 0:
            return torch.tensor([0.0, 0.0, 0.0, float("inf"), float("-inf")], dtype=torch.float64,
                                device=v.device)
        return torch.stack([torch.tensor(float(v.numel()), dtype=torch.float64, device=v.device),
                            v.sum(), (v * v).sum(), v.min(), v.max()])
    v = np.asarray(values, dtype=np.float64).reshape(-1)
    if v.size == 0:
        return np.array([0.0, 0.0, 0.0, np.inf, -np.inf])
    return np.array([float(v.size), v.sum(), (v * v).sum(), v.min(), v.max()])


def merge_stats(gathered) -> Dict[str, float]:
    """Combines the gathered [world x 5] statistics into those of the whole batch."""
    g = gathered.detach().cpu().numpy() if hasattr(gathered, "detach") else np.asarray(gathered)
    g = g.reshape(-1, len(STAT_FIELDS))
    count = float(g[:, 0].sum())
    mean = float(g[:, 1].sum() / count) if count > 0 else 0.0
    var = float(g[:, 2].sum() / count - mean * mean) if count > 0 else 0.0
    return {"count": count, "mean": mean, "std": float(np.sqrt(max(var, 0.0))),
            "min": float(g[:, 3].min()), "max": float(g[:, 4].max())}


def all_gather_stats(stats, out=None, async_op=False):
    """One all_gather_into_tensor of the per-rank statistics over the default process group
    (NCCL on GPU tensors, gloo on CPU tensors). Without an initialised group: a copy.

    async_op=True returns (gathered, work): the collective is enqueued behind the work already
    on the current stream and overlaps whatever the caller launches next; work.wait() before
    reading `gathered`."""
    import torch
    import torch.distributed as dist

    t = stats if isinstance(stats, torch.Tensor) else torch.from_numpy(np.asarray(stats, dtype=np.float64))
    if not (dist.is_available() and dist.is_initialized()):
        g = t.reshape(1, -1).clone()
        return (g, None) if async_op else g
    world = dist.get_world_size()
    if out is None:
        out = torch.empty(world * t.numel(), dtype=t.dtype, device=t.device)
    work = dist.all_gather_into_tensor(out, t.contiguous(), async_op=async_op)
    g = out.reshape(world, -1)
    return (g, work) if async_op else g


def all_gather_results(local, out=None, async_op=False):
    """All-gather of compact per-state results [n_local x ...] of equally sized slices into
    [world * n_local x ...], rank r's rows at [r * n_local, (r + 1) * n_local): with
    slice_of() that is the order of the global state sequence. Same calling convention as
    all_gather_stats."""
    import torch
    import torch.distributed as dist

    t = local if isinstance(local, torch.Tensor) else torch.from_numpy(np.ascontiguousarray(local))
    if not (dist.is_available() and dist.is_initialized()):
        g = t.clone()
        return (g, None) if async_op else g
    world = dist.get_world_size()
    if out is None:
        out = torch.empty((world * t.shape[0],) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    work = dist.all_gather_into_tensor(out, t.contiguous(), async_op=async_op)
    return (out, work) if async_op else out
