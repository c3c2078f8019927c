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


HEAD_COEFF = 48.0     # m^3-flop equivalents of the landmark-only stage (see head_samples); tools/multi_gpu_round.sh sweeps it


def head_samples(m: int, d: int, p: int) -> int:
    """How many samples of fused lift+Gram work the landmark-only stage of a fit is worth (K_zz, Cholesky, 17 Newton-Schulz
    iterations of the symmetric square root, S and S^-1: ~56 m^3 flop on the row-major GEMM at ~82% of the DMMA rate, plus the
    packing of the three m x m results for the broadcast, against 4m^2+6md+4mp flop per sample at ~90%).  Calibrated on the
    measurement at m=4096, d=192.  Round 1: 143 ms against 467 k samples/s = 67 k samples (coefficient 64).  Round 2 (TMA-fed GEMM):
    125 ms against 474 k samples/s = 59 k samples (coefficient 56), then 107 ms with the spectrum estimate of the symmetric
    square root = 50 k samples (coefficient 48; profiles/r02_fit_distributed_phases_2gpu.log).  Overshooting is cheap (the surplus is spread over the other ranks), undershooting
    is paid in full."""
    per_sample = 4.0 * m * m + 6.0 * m * d + 4.0 * m * p
    return int(HEAD_COEFF * m ** 3 / per_sample * (0.90 / 0.82))


def balanced_bounds(n_total: int, world: int, rank: int, head: int = 0, head_rank: int = 0):
    """Contiguous block partition in which `head_rank`'s block is `head` samples shorter than the others' (it also owns
    the landmark-only stage, see KoopmanNystromRegressor.fit_distributed): n_r = (n + head) / world for the others.
    Falls back to the even partition when the head's block would be empty.  Returns (offset, n_local)."""
    n_total, world, head = int(n_total), int(world), int(head)
    if world == 1 or head <= 0:
        return shard_bounds(n_total, world, rank)
    big = (n_total + head) // world
    if big - head < 1:
        return shard_bounds(n_total, world, rank)
    sizes = [big] * world
    sizes[head_rank] = big - head
    rem = n_total - sum(sizes)                       # 0 <= rem < world: hand the leftovers to the non-head ranks
    r = 0
    while rem > 0:
        if r != head_rank:
            sizes[r] += 1
            rem -= 1
        r = (r + 1) % world
    return int(sum(sizes[:rank])), int(sizes[rank])


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


# ---- column-sharded solve (nk_solve_abc_part) --------------------------------------------------------------------
def solve_row_ranges(n_rows: int, world: int, rank: int):
    """Rows of G^T ((m+p) x m) or C^T (m x d) that `rank` solves: equal slabs of ceil(n_rows / world) rows (the gather buffer
    is world * slab rows, zero-padded at the end), the last ranks may get fewer or none.  Returns (slab, start, count)."""
    slab = -(-int(n_rows) // int(world))
    start = min(rank * slab, n_rows)
    return slab, start, max(0, min(slab, n_rows - rank * slab))


def gather_rows(all_buf, mine, group=None):
    """Assembles the row slabs every rank computed into `all_buf` (world * slab rows; `mine` is this rank's slab, a view of
    it).  NCCL: one all_gather; other backends (gloo in the tests): an all_reduce of the zero-filled buffer -- every row is
    written by exactly one rank, so the sum is the concatenation."""
    import torch.distributed as dist
    if dist.get_world_size(group) == 1:
        return all_buf
    if dist.get_backend(group) == "nccl":
        dist.all_gather_into_tensor(all_buf, mine.clone(), group=group)
    else:
        dist.all_reduce(all_buf, group=group)
    return all_buf


# ---- cross-validation sweep over several GPUs ("CV sweep") ----------------------------------------------
def kfold_bounds(n_total: int, n_splits: int):
    """sklearn KFold(n_splits) without shuffling over the GLOBAL sample index: contiguous blocks, the first
    n_total % n_splits one sample longer.  [(start, stop), ...]"""
    sizes = np.full(n_splits, n_total // n_splits, dtype=np.int64)
    sizes[: n_total % n_splits] += 1
    stops = np.cumsum(sizes)
    return [(int(e - s), int(e)) for s, e in zip(sizes, stops)]


def fold_local_ranges(n_total: int, n_splits: int, offset: int, n_local: int):
    """Intersection of every global fold with this rank's contiguous block, in LOCAL row indices.
    [(lo, hi), ...] with lo == hi where the rank holds nothing of that fold."""
    out = []
    for s, e in kfold_bounds(n_total, n_splits):
        lo, hi = max(s, offset), min(e, offset + n_local)
        out.append((lo - offset, hi - offset) if hi > lo else (0, 0))
    return out


def cv_tasks(n_splits: int, n_gammas: int, world: int):
    """The (fold, gamma-slice) solve tasks of one kernel and who runs them.  Every (fold, gamma) appears exactly once.
    The gamma axis is cut into as many contiguous groups as it takes for the task count to be a multiple of the world size
    (capped by the number of gammas): 5 folds on 8 ranks -> 8 groups -> 40 tasks, 5 per rank.
    Returns [(rank, fold, g_lo, g_hi), ...]."""
    from math import gcd
    groups = max(1, min(n_gammas, world // gcd(world, n_splits)))
    cuts = [round(i * n_gammas / groups) for i in range(groups + 1)]
    tasks, t = [], 0
    for f in range(n_splits):
        for gi in range(groups):
            if cuts[gi + 1] > cuts[gi]:
                tasks.append((t % world, f, cuts[gi], cuts[gi + 1]))
                t += 1
    return tasks


def allreduce_sum(t, group=None):
    import torch.distributed as dist
    dist.all_reduce(t, group=group)
    return t
