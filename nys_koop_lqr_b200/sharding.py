"""Host-side logic of the sample-sharded fit: who owns which samples, how the landmark set is assembled,
and the single data-path collective.  Backend-agnostic (NCCL on the GPUs, gloo in the CPU tests)."""
from __future__ import annotations

import numpy as np


def shard_bounds(n_total: int, world: int, rank: int):
    """Contiguous block partition; the first n_total % world ranks hold one extra sample. Returns (offset, n_local)."""
    base, extra = divmod(int(n_total), int(world))
    n_local = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, n_local


def global_layout(n_local: int, group=None, device="cpu"):
    """All ranks learn (n_total, my_offset) from the local counts with one small all_reduce."""
    import torch
    import torch.distributed as dist
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    counts = torch.zeros(world, dtype=torch.int64, device=device)
    counts[rank] = int(n_local)
    dist.all_reduce(counts, group=group)
    counts = counts.cpu().numpy()
    return int(counts.sum()), int(counts[:rank].sum())


def assemble_landmarks(idx, offset, n_local, rows_of, d, group=None, device="cpu"):
    """Z (m,d) = Y_global[idx]: every rank contributes the rows it owns (rows_of(local_indices) -> (k,d) tensor on
    `device`), the rest arrives through one all_reduce of a zero-filled buffer.  idx must be identical on all ranks
    (drawn with the reference's np.random.choice over the global sample index, regressors.py:129-132)."""
    import torch
    import torch.distributed as dist
    idx = np.asarray(idx)
    Z = torch.zeros(len(idx), d, dtype=torch.float64, device=device)
    mine = np.nonzero((idx >= offset) & (idx < offset + n_local))[0]
    if mine.size:
        Z[torch.as_tensor(mine, device=device)] = rows_of(idx[mine] - offset).to(device=device, dtype=torch.float64)
    dist.all_reduce(Z, group=group)
    return Z


def allreduce_grams(flat, group=None):
    """The only collective on the data path: float64 sum of the packed Grams [Gxx|Gyx|Gyy|Gxu|Gyu|Guu|GYy]."""
    import torch.distributed as dist
    dist.all_reduce(flat, group=group)
    return flat
