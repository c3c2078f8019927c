"""GPU, world_size=2 over NCCL: the multi-GPU CV sweep and sharded fit give the single-GPU results (needs >= 2 GPUs)."""
import os
import pathlib
import socket

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
FX = pathlib.Path(__file__).parent / "golden" / "cv" / "synthetic_rbf_cv.npz"


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import regressors as R
    from nys_koop_lqr_b200 import sharding
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        fx = np.load(FX)
        holders = []
        for ls in fx["ls"]:
            h = R.ThreeDimensionalKernel(1, 1, 1, ls.size)
            h.kernel.length_scale = ls.reshape(1, -1)
            holders.append(h)
        n = fx["X"].shape[0]
        off, nl = sharding.shard_bounds(n, world, rank)
        reg = R.KoopmanNystromRegressor(int(fx["n_inputs"]), kernel=holders[0], gamma=1e-3, m=int(fx["m"]))
        reg.nystrom_centers_output = fx["Z"].copy()
        res = reg.fit_cv_distributed(fx["X"][off:off + nl], fx["Y"][off:off + nl], holders, list(fx["gammas"]), n_splits=int(fx["n_splits"]))
        split = np.stack([res[f"split{k}_test_score"] for k in range(int(fx["n_splits"]))], axis=1)
        rel = float((np.abs(split - fx["split_test_score"]) / np.abs(fx["split_test_score"])).max())
        errA = float(np.linalg.norm(reg.A - fx["best_A"]) / np.linalg.norm(fx["best_A"]))
        out.put((rank, rel, reg.best_index_ == int(fx["best_index"]), errA))
    finally:
        dist.destroy_process_group()


def test_fit_cv_distributed_world2_matches_gridsearchcv_golden():
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, out)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    got = sorted(out.get(timeout=5) for _ in range(2))
    for rank, rel, best_ok, errA in got:
        assert rel <= 1e-8 and best_ok and errA <= 1e-7, (rank, rel, best_ok, errA)


def test_two_devices_in_one_process():
    """One process, two handles (one per GPU): every kernel's opt-in shared-memory attribute is configured per device."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from nys_koop_lqr_b200.engine import Engine
    from oracle import nk_oracle as O
    rng = np.random.default_rng(0)
    n, d, p, m = 700, 5, 2, 40
    Xs, U, Y = O.synthetic(n, d, p, seed=3)
    Z = Y[rng.choice(n, m, replace=False)]
    ls = np.full(d, 2.0)
    ref = O.grams(Xs, Y, U, Z, O.RBF, ls)
    before = torch.cuda.current_device()
    for devno in (0, 1):
        eng = Engine.get(devno)
        # the library runs on its handle's device and must leave the CALLER's current device alone (it used to leave the
        # handle's device selected: the next `.cuda()` of the caller then landed on the wrong GPU)
        assert torch.cuda.current_device() == before
        with torch.cuda.device(devno):
            t = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device=f"cuda:{devno}")
            G = eng.grams(t(np.hstack((Xs, U))), t(Y), t(Z), t(1.0 / ls), O.RBF, p)
            assert O.relerr(G["Gyx"].cpu().numpy(), ref["Gyx"]) <= 1e-12
            K = eng.kzz(t(Z), t(1.0 / ls), O.RBF)
            K.diagonal().add_(1e-6)
            S, Sinv = eng.sym_sqrt(K)
            assert float((S @ S - K).norm() / K.norm()) <= 1e-12
            z0 = t(rng.standard_normal((3, m)))
            out = eng.rollout(t(rng.standard_normal((m, m)) * 0.1), None, t(rng.standard_normal((d, m))), z0, None, Ytrue=t(rng.standard_normal((1, 3, d))))
            assert out["Yhat"].shape == (1, 3, d)
        assert torch.cuda.current_device() == before
    # a handle of another device driven WITHOUT a torch device guard around the calls: still no change of the current device
    eng1 = Engine.get(1)
    t1 = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64, device="cuda:1")
    K1 = eng1.kzz(t1(Z), t1(1.0 / ls), O.RBF)
    assert K1.device.index == 1 and torch.cuda.current_device() == before
    assert O.relerr(K1.cpu().numpy(), O.kernel_matrix(Z, Z, O.RBF, ls)) <= 1e-13
