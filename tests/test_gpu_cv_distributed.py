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
