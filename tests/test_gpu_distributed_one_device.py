"""GPU, world_size = 2 on ONE device (gloo transport, both ranks on cuda:0): the sample-sharded fit end to end -- shared landmark
draw, fused-kernel Gram pass per shard, allreduce, head-rank landmark stage + broadcast, column-sharded solve
(nk_solve_abc_part), gather, finish, lazy download -- against the single-process fit and the oracle.  NCCL refuses two ranks on
one GPU, so the 2-GPU NCCL test (test_gpu_cv_distributed.py) is skipped on a one-GPU box; this one is not: everything except the
transport is the same code."""
import os
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, out, distinct):
    import torch.distributed as dist
    import regressors as R
    from nys_koop_lqr_b200 import sharding
    from oracle import nk_oracle as O
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(0)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n, d, p, m = 6000, 24, 3, 300
        Xs, U, Y = O.synthetic(n, d, p, seed=21)
        X = np.hstack((Xs, U))
        ls = np.resize([4.0, 5.0, 6.0], d)
        off, nl = sharding.balanced_bounds(n, world, rank, head=500)
        np.random.seed(100 + rank)                                   # DIFFERENT numpy RNG states: rank 0's draw must win on both
        reg = R.KoopmanNystromRegressor(p, kernel=R.ThreeDimensionalKernel(4.0, 5.0, 6.0, d), gamma=1e-2, m=m)
        if distinct:
            rng = np.random.default_rng(5)
            reg.nystrom_centers_output = np.ascontiguousarray(Y[rng.choice(n, m, replace=False)].T)
            reg.nystrom_centers_input = np.ascontiguousarray(Xs[rng.choice(n, m, replace=False)].T)
        reg.fit_distributed(torch.from_numpy(X[off:off + nl]).cuda(), torch.from_numpy(Y[off:off + nl]).cuda())
        assert reg.__dict__["_A"] is None and reg.__dict__["_pending"], "results must stay on the device until read"
        Z = reg.nystrom_centers_output.T
        np.random.seed(100)
        if not distinct:
            assert np.array_equal(Z, Y[np.random.choice(np.arange(0, n), size=m, replace=False)]), "landmarks are not rank 0's draw"
        single = R.KoopmanNystromRegressor(p, kernel=reg.kernel, gamma=1e-2, m=m)
        single.nystrom_centers_output = reg.nystrom_centers_output.copy()
        single.nystrom_centers_input = reg.nystrom_centers_input.copy() if distinct else None
        single.fit(X, Y)
        e_single = max(O.relerr(getattr(reg, k), getattr(single, k)) for k in ("A", "B", "C", "weights"))
        e_oracle = None
        if not distinct:
            want = O.fit(X, Y, p, O.RBF, ls, 1e-2, Z=Z)
            e_oracle = max(O.relerr(getattr(reg, k), want[w]) for k, w in (("A", "A"), ("B", "B"), ("C", "C"), ("weights", "W")))
        yh = reg.predict(X[:50])
        e_pred = O.relerr(yh, single.predict(X[:50]))
        out.put((rank, e_single, e_oracle, e_pred))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("distinct", [False, True])
def test_fit_distributed_two_ranks_on_one_gpu(distinct):
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out, distinct)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=600)
        assert p.exitcode == 0
    for rank, e_single, e_oracle, e_pred in sorted(out.get(timeout=5) for _ in range(2)):
        assert e_single <= 1e-9, (rank, e_single)           # same kernels, different shard / column split: summation order only
        assert e_oracle is None or e_oracle <= 1e-9, (rank, e_oracle)
        assert e_pred <= 1e-9
