#!/usr/bin/env python
"""bench_cv.py -- BASELINE.json configs[4]: synthetic CV sweep, 16 lengthscales x 16 regularisation values at m=8192
landmarks, n=2e6 samples, 5 folds, then a 10^5-trajectory open-loop rollout of the refitted model (T=101).

    python bench_cv.py                       # full configuration (about 8 minutes on one B200)
    python bench_cv.py --n 200000 --m 4096 --kernels 2 --traj 10000      # reduced, for a quick look
    torchrun --nproc-per-node 8 bench_cv.py                               # samples and trajectories sharded over the GPUs

Not the driver's headline bench (that is bench.py, configs[3]); this script measures the second synthetic configuration
through the drop-in estimator's batched search (`fit_cv`) and `Engine.rollout`, and prints ONE JSON line.
Phase times are host wall-clock around device synchronisation (the sweep is many launches); the rollout is timed with
CUDA events.  Algorithmic flops: Gram pass F = 4m^2+6md+4mp per sample and kernel; per (kernel, fold, gamma)
two Cholesky factorisations (m^3/3 + (m+p)^3/3); rollout 2m^2+2mp+2dm per trajectory-step.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


# ----------------------------------------------------------------------------------------------------------
# stdout carries exactly ONE line (the JSON result): everything else any library writes to file descriptor 1 (NCCL prints an
# "NCCL version ..." line there when NCCL_DEBUG is VERSION or WARN, torchrun banners, ...) is sent to stderr instead
# ----------------------------------------------------------------------------------------------------------
_RESULT_OUT = None


def protect_stdout():
    global _RESULT_OUT
    if _RESULT_OUT is None:
        sys.stdout.flush()
        _RESULT_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line):
    out = _RESULT_OUT or sys.stdout
    out.write(json.dumps(line) + "\n")
    out.flush()


def parse(argv=None):
    ap = argparse.ArgumentParser()
    ap.add_argument("--samples", "--n", dest="n", type=int, default=2_000_000)
    ap.add_argument("--landmarks", "--m", dest="m", type=int, default=8192)
    ap.add_argument("--d", type=int, default=192)
    ap.add_argument("--p", type=int, default=6)
    ap.add_argument("--kernels", type=int, default=16)
    ap.add_argument("--gammas", type=int, default=16)
    ap.add_argument("--folds", type=int, default=5)
    ap.add_argument("--traj", type=int, default=100_000)
    ap.add_argument("--T", type=int, default=101)
    return ap.parse_args(argv)


def main():
    args = parse()
    protect_stdout()
    line = run(args)
    if line is not None:
        emit(line)


def run(args, standalone=True):
    """One pass of configs[4] at the sizes in `args`; returns the result dict on rank 0 (None elsewhere).  `standalone=False`:
    called from bench.py inside an already initialised single-GPU process (no process-group handling)."""
    import torch
    import regressors as R
    from nys_koop_lqr_b200.engine import Engine
    if not torch.cuda.is_available():
        raise SystemExit("bench_cv.py: no CUDA device (no CPU fallback)")
    import torch.distributed as dist
    from nys_koop_lqr_b200 import sharding
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    dev = torch.device("cuda", local_rank)
    torch.cuda.set_device(dev)
    if not standalone:
        world, rank = 1, 0
    if world > 1:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # lands on stderr (protect_stdout)
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine.get(local_rank)
    n_total, m, d, p = args.n, args.m, args.d, args.p
    off, n = sharding.shard_bounds(n_total, world, rank)          # this rank's contiguous block of the samples
    g = torch.Generator(device=dev); g.manual_seed(1234)
    M = torch.randn(d, d, dtype=torch.float64, device=dev, generator=g) * (0.9 / d ** 0.5)
    Bu = 0.1 * torch.randn(d, p, dtype=torch.float64, device=dev, generator=g)
    g.manual_seed(5000 + rank)
    X = torch.empty(n, d + p, dtype=torch.float64, device=dev)
    Y = torch.empty(n, d, dtype=torch.float64, device=dev)
    for s in range(0, n, 1 << 19):
        e = min(n, s + (1 << 19))
        X[s:e].normal_(generator=g)
        Y[s:e] = torch.tanh(X[s:e, :d] @ M.T) + X[s:e, d:] @ Bu.T
    # grid of SURVEY 8(d): lengthscales 10^(0.5..1.5), gammas 10^(-6..-2.25) step 0.25 (benchmark_lqr_hjb.py:56)
    ls = np.logspace(0.5, 1.5, args.kernels)
    kernels = [R.ThreeDimensionalKernel(l, l, l, d) for l in ls]
    gammas = [float(10.0 ** e) for e in (-6.0 + 0.25 * np.arange(args.gammas))]
    np.random.seed(0)
    reg = R.KoopmanNystromRegressor(p, kernel=kernels[0], gamma=gammas[0], m=m)
    reg.cv_profile = True
    l0 = eng.launch_count()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
    barrier(); t0 = time.perf_counter()
    if world > 1:
        res = reg.fit_cv_distributed(X, Y, kernels, gammas, n_splits=args.folds, refit=False)
    else:
        res = reg.fit_cv(X, Y, kernels, gammas, n_splits=args.folds, refit=False)
    barrier(); t_cv = time.perf_counter() - t0
    prof = dict(reg.cv_profile_) if getattr(reg, "cv_profile_", None) else {"gram_s": float("nan"), "weights_s": float("nan"), "score_s": float("nan")}
    t0 = time.perf_counter()
    reg.kernel, reg.gamma = reg.best_params_["kernel"], reg.best_params_["gamma"]
    if world > 1:
        reg.fit_distributed(X, Y)
    else:
        reg.fit(X, Y)
    barrier(); t_refit = time.perf_counter() - t0
    launches_cv = eng.launch_count() - l0

    # ---- rollout of the refitted model over many trajectories (replicas only: trajectories are independent) ----
    nb_total, T = args.traj, args.T
    nb = sharding.shard_bounds(nb_total, world, rank)[1]         # trajectories are independent: replicas only, no collective
    n = n_total
    dv = reg._device_state(d)
    A, B, C = (torch.from_numpy(np.ascontiguousarray(a)).to(dev) for a in (reg.A, reg.B, reg.C))
    x0 = X[:nb, :d].contiguous()
    U = torch.randn(T - 1, nb, p, dtype=torch.float64, device=dev, generator=g)
    Ytrue = torch.randn(T, nb, d, dtype=torch.float64, device=dev, generator=g)
    barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    l1 = eng.launch_count()
    e0.record()
    Z0 = eng.lift(dv["Z"], dv["inv_ls"], dv["kind"], dv["Sinv"], x0, transposed=True)
    e1.record()
    out = eng.rollout(A, B, C, Z0, U, Ytrue=Ytrue, return_traj=False)
    e2.record(); torch.cuda.synchronize()
    ms_lift, ms_roll = e0.elapsed_time(e1), e1.elapsed_time(e2)
    if world > 1:                                                 # device time, max over ranks
        tt = torch.tensor([ms_lift, ms_roll], dtype=torch.float64, device=dev)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        ms_lift, ms_roll = float(tt[0]), float(tt[1])
    nb = nb_total
    fl_roll = nb * ((T - 1) * (2.0 * m * m + 2.0 * m * p) + T * 2.0 * d * m)
    rmse = torch.sqrt(out["sq_err"] / (d * T))

    F = 4.0 * m * m + 6.0 * m * d + 4.0 * m * p
    n_solves = args.kernels * args.gammas * args.folds
    chol_flops = n_solves * (m ** 3 / 3.0 + (m + p) ** 3 / 3.0)
    peak = eng.probe_dmma_tflops(200.0) * world
    line = {
        "config": {"workload": f"synthetic CV sweep {args.kernels} lengthscales x {args.gammas} gamma at m={m}, n={n}, d={d}, p={p}, "
                               f"{args.folds} folds + {nb}-trajectory rollout T={T} (BASELINE.json configs[4])",
                   "landmarks": "one shared landmark set for all candidates and folds (stated deviation from GridSearchCV's per-clone redraw)"},
        "cv_seconds": t_cv, "phase_seconds": prof, "refit_seconds": t_refit,
        "candidates": args.kernels * args.gammas, "fits_equivalent": n_solves,
        "cv_fits_per_s": n_solves / t_cv,
        "gram_pass": {"samples": n * args.kernels, "seconds": prof["gram_s"], "samples_per_s": n * args.kernels / prof["gram_s"],
                      "tflops": F * n * args.kernels / prof["gram_s"] * 1e-12, "frac_of_dmma_peak": F * n * args.kernels / prof["gram_s"] * 1e-12 / peak},
        "batched_solves": {"systems": 2 * n_solves, "seconds": prof["weights_s"], "cholesky_tflops": chol_flops / prof["weights_s"] * 1e-12},
        "scoring": {"seconds": prof["score_s"], "predictions": n * args.kernels * args.gammas},
        "rollout": {"trajectories": nb, "T": T, "lift_ms": ms_lift, "rollout_ms": ms_roll, "trajectory_steps_per_s": nb * (T - 1) / (ms_roll * 1e-3),
                    "tflops": fl_roll / (ms_roll * 1e-3) * 1e-12, "frac_of_dmma_peak": fl_roll / (ms_roll * 1e-3) * 1e-12 / peak,
                    "rmse_finite": bool(torch.isfinite(rmse).all().item())},
        "dmma_peak_tflops": peak, "gpu_launches": int(launches_cv + (eng.launch_count() - l1)),
        "best": {"index": reg.best_index_, "lengthscale": float(ls[kernels.index(reg.best_params_["kernel"])]), "gamma": reg.best_params_["gamma"],
                 "score": reg.best_score_, "nan_candidates": int(np.isnan(res["mean_test_score"]).sum())},
        "dtype": "f64", "data": "synthetic", "n_gpus": world,
    }
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    eng.release_scratch()
    return line if rank == 0 else None


if __name__ == "__main__":
    main()
