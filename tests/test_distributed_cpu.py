"""CPU, world_size=2 over gloo: host logic of the sample-sharded fit (layout, landmark assembly, Gram allreduce)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import nk_oracle as O


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, n, d, p, m, out):
    from nys_koop_lqr_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        Xs, U, Y = O.synthetic(n, d, p, seed=4)
        off, n_local = sharding.shard_bounds(n, world, rank)
        n_total, off2 = sharding.global_layout(n_local)
        assert (n_total, off2) == (n, off)
        np.random.seed(0)                                          # same global RNG state on every rank
        idx = np.random.choice(np.arange(0, n_total), size=m, replace=False)
        Yl = Y[off:off + n_local]
        Z = sharding.assemble_landmarks(idx, off, n_local, lambda loc: torch.from_numpy(Yl[loc]), d).numpy()
        assert np.array_equal(Z, Y[idx]), "assembled landmarks differ from the single-process draw"
        ls = np.full(d, 3.0)
        G = O.grams(Xs[off:off + n_local], Yl, U[off:off + n_local], Z, O.RBF, ls)
        names = ("Gxx", "Gyx", "Gyy", "Gxu", "Gyu", "Guu", "GYy")
        flat = torch.from_numpy(np.concatenate([G[k].ravel() for k in names]))
        sharding.allreduce_grams(flat)
        if rank == 0:
            full = O.grams(Xs, Y, U, Z, O.RBF, ls)
            ref = np.concatenate([full[k].ravel() for k in names])
            out.put(float(np.linalg.norm(flat.numpy() - ref) / np.linalg.norm(ref)))
    finally:
        dist.destroy_process_group()


def test_shard_bounds_cover_everything():
    from nys_koop_lqr_b200.sharding import shard_bounds
    for n, w in ((10, 3), (7, 8), (10_000_000, 8), (5, 1)):
        spans = [shard_bounds(n, w, r) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2


def test_sharded_grams_and_landmarks_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, 1501, 6, 2, 40, out)) for r in range(2)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(timeout=120)
        assert p_.exitcode == 0
    err = out.get(timeout=5)
    assert err <= 1e-13, f"shard-sum invariance violated: {err:.2e}"


def test_cv_fold_ranges_and_tasks_cover_everything():
    """Host logic of the multi-GPU CV sweep: global unshuffled folds cut by the rank blocks; every (fold, gamma) solved once."""
    from nys_koop_lqr_b200 import sharding
    from sklearn.model_selection import KFold
    for n, k, w in ((1003, 5, 2), (2_000_000, 5, 8), (17, 3, 4), (100, 5, 1)):
        want = [(int(te[0]), int(te[-1]) + 1) for _, te in KFold(k).split(np.zeros((n, 1)))] if n <= 5000 else sharding.kfold_bounds(n, k)
        assert sharding.kfold_bounds(n, k) == want
        covered = np.zeros((k, n), dtype=np.int8) if n <= 5000 else None
        total = np.zeros(k, dtype=np.int64)
        for r in range(w):
            off, nl = sharding.shard_bounds(n, w, r)
            for f, (lo, hi) in enumerate(sharding.fold_local_ranges(n, k, off, nl)):
                assert 0 <= lo <= hi <= nl
                total[f] += hi - lo
                if covered is not None:
                    covered[f, off + lo:off + hi] += 1
        assert list(total) == [e - s for s, e in want]
        if covered is not None:
            for f, (s, e) in enumerate(want):
                assert covered[f, s:e].min() == 1 and covered[f].sum() == e - s
    for k, g, w in ((5, 16, 8), (5, 16, 1), (5, 3, 8), (3, 16, 2), (5, 1, 4)):
        tasks = sharding.cv_tasks(k, g, w)
        seen = np.zeros((k, g), dtype=int)
        load = np.zeros(w)
        for r, f, g0, g1 in tasks:
            assert 0 <= r < w and g1 > g0
            seen[f, g0:g1] += 1
            load[r] += g1 - g0
        assert (seen == 1).all()
        if k * g >= w:
            assert load.max() - load.min() <= max(1, np.ceil(g / max(1, w)))   # round-robin keeps the ranks balanced


def _gather_worker(rank, world, port, out):
    from nys_koop_lqr_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        k, g = 5, 6
        W = torch.zeros(k, g, 3, 4, dtype=torch.float64)
        for r, f, g0, g1 in sharding.cv_tasks(k, g, world):
            if r == rank:
                W[f, g0:g1] = torch.arange(f * 100 + g0, f * 100 + g1, dtype=torch.float64).view(-1, 1, 1)
        sharding.allreduce_sum(W)                                  # exchange = sum of disjointly written slices
        want = (torch.arange(k).view(-1, 1) * 100 + torch.arange(g).view(1, -1)).double().view(k, g, 1, 1).expand(k, g, 3, 4)
        if rank == 0:
            out.put(bool(torch.equal(W, want)))
    finally:
        dist.destroy_process_group()


def test_cv_weight_exchange_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_gather_worker, args=(r, 2, port, out)) for r in range(2)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(timeout=120)
        assert p_.exitcode == 0
    assert out.get(timeout=5)


def test_balanced_bounds_give_the_head_rank_a_shorter_block():
    from nys_koop_lqr_b200 import sharding
    h = sharding.head_samples(4096, 192, 6)
    assert 40_000 < h < 80_000                       # measured on B200: 107 ms of landmark-only work ~ 50 k samples of Gram work
    for n, w, head, hr in ((10_000_000, 8, h, 0), (10_000_000, 2, h, 0), (1003, 4, 100, 2), (50, 4, 1000, 0), (7, 1, 3, 0)):
        spans = [sharding.balanced_bounds(n, w, r, head, hr) for r in range(w)]
        assert spans[0][0] == 0 and sum(c for _, c in spans) == n
        for (o1, c1), (o2, _) in zip(spans, spans[1:]):
            assert o1 + c1 == o2 and c1 >= 0
        if w > 1 and (n + head) // w - head >= 1:
            others = [c for r, (_, c) in enumerate(spans) if r != hr]
            assert max(others) - min(others) <= 1
            assert abs((spans[hr][1] + head) - others[0]) <= 1      # equal total work
        else:
            assert spans == [sharding.shard_bounds(n, w, r) for r in range(w)]


def test_solve_row_ranges_cover_every_column_once():
    """Column-sharded solve (nk_solve_abc_part): every row of G^T / C^T belongs to exactly one rank, slabs are equal, the tail is padding."""
    from nys_koop_lqr_b200 import sharding
    for rows, world in ((4102, 8), (4102, 2), (106, 8), (7, 8), (3, 4), (1, 1), (4096, 3)):
        seen = np.zeros(rows, dtype=int)
        for r in range(world):
            slab, start, count = sharding.solve_row_ranges(rows, world, r)
            assert slab * world >= rows and 0 <= count <= slab and start + count <= rows
            assert count == 0 or start == r * slab
            seen[start:start + count] += 1
        assert (seen == 1).all()


def _solve_shard_worker(rank, world, port, out):
    """The host side of the sharded solve with the oracle standing in for nk_solve_abc_part: each rank solves its columns,
    gather_rows assembles G^T and C^T, and the result equals the one-process solve; a failure on one rank raises on both."""
    from nys_koop_lqr_b200 import sharding
    from nys_koop_lqr_b200.regressors import KoopmanNystromRegressor as K
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d, p, m = 700, 4, 2, 37
        Xs, U, Y = O.synthetic(n, d, p, seed=8)
        Z = Y[np.random.default_rng(0).choice(n, m, replace=False)]
        ls = np.full(d, 2.5)
        G = O.grams(Xs, Y, U, Z, O.RBF, ls)
        Kzz = O.kernel_matrix(Z, Z, O.RBF, ls)
        A, B, C, W = O.solve_abc(G, Kzz, 1e-2 * n)
        GT_full, CT_full = np.hstack((A, B)).T, C.T
        N1 = m + p
        g_slab, g0, gc = sharding.solve_row_ranges(N1, world, rank)
        c_slab, c0, cc = sharding.solve_row_ranges(m, world, rank)
        GT_all = torch.zeros(world * g_slab, m, dtype=torch.float64)
        CT_all = torch.zeros(world * c_slab, d, dtype=torch.float64)
        GT_all[rank * g_slab:rank * g_slab + gc] = torch.from_numpy(GT_full[g0:g0 + gc])
        CT_all[rank * c_slab:rank * c_slab + cc] = torch.from_numpy(CT_full[c0:c0 + cc])
        sharding.gather_rows(GT_all, GT_all[rank * g_slab:(rank + 1) * g_slab])
        sharding.gather_rows(CT_all, CT_all[rank * c_slab:(rank + 1) * c_slab])
        ok = np.array_equal(GT_all[:N1].numpy(), GT_full) and np.array_equal(CT_all[:m].numpy(), CT_full)
        ok = ok and float(GT_all[N1:].abs().sum()) == 0.0
        # collective error agreement: rank 1 fails locally, BOTH ranks must raise (nobody is left waiting in a collective)
        raised = False
        try:
            K._agree(None, "cpu", (lambda: 1 / 0) if rank == 1 else (lambda: 7), "stage")
        except ZeroDivisionError:
            raised = rank == 1
        except Exception as exc:  # noqa: BLE001
            raised = rank == 0 and "another rank" in str(exc)
        assert K._agree(None, "cpu", lambda: rank, "stage") == rank
        out.put((rank, bool(ok), bool(raised)))
    finally:
        dist.destroy_process_group()


def test_sharded_solve_gather_and_error_agreement_world2():
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_solve_shard_worker, args=(r, 2, port, out)) for r in range(2)]
    for p_ in procs:
        p_.start()
    for p_ in procs:
        p_.join(timeout=120)
        assert p_.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(2))
    assert got == [(0, True, True), (1, True, True)]
