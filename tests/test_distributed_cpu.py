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
