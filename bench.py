#!/usr/bin/env python
"""bench.py -- Nystrom-Koopman fit throughput on B200 (BASELINE.json metric, config "synthetic Koopman fit n=1e7,
d=192, m=4096, FP64").

    python bench.py --gpus 1 --steps K --warmup W            # this repo's arm (default N=1, K=2, W=3)
    python bench.py --impl reference --steps K --warmup W    # CPU arm: the reference's path restated (oracle), host cores
    torchrun ... bench.py --gpus N ...                       # N>1: samples sharded over ranks (strong scaling), one allreduce

A step is ONE complete fit of the named workload through the drop-in estimator
(regressors.KoopmanNystromRegressor.fit: landmark matrices, fused kernel-lift + Grams, [allreduce], the two regularised
solves, A/B/C/weights back on the host).  `value` times it with the samples already resident in HBM; `e2e` times the
same call with the samples in pinned HOST memory (streamed up inside the timed region).  Prints one JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "Nystrom-Koopman fit samples/s (n=1e7, m=4096)"
UNIT = "samples/s"


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


def algorithmic_flops_per_sample(m, d, p):
    """lifts 4md, two symmetric Grams m^2 each, cross Gram 2m^2, 2x2mp control products, 2md reconstruction."""
    return 4.0 * m * m + 6.0 * m * d + 4.0 * m * p


# ----------------------------------------------------------------------------------------------------------
# clocks / throttle sampling during the timed region
# ----------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        self.index = index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def __enter__(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "200"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
        return self

    def __exit__(self, *a):
        if self.p is not None:
            self.p.terminate()
            try:
                self.p.wait(timeout=5)
            except subprocess.TimeoutExpired:
                self.p.kill()

    def summary(self):
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.f.read().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1]))
            except ValueError:
                continue
            for nm, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        os.unlink(self.f.name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        return {"sm_mhz": float(np.median(sm)), "sm_max_mhz": float(max(mx)), "reasons": sorted(reasons), "samples": len(sm)}


# ----------------------------------------------------------------------------------------------------------
# CPU arm: the reference's own path (restated in oracle/, same scipy/numpy calls) on a bounded sample
# ----------------------------------------------------------------------------------------------------------
def cpu_rate(n_sample, m, d, p, seed=0):
    """Times the n-proportional stage of the reference fit (kernel lift via scipy cdist, as sklearn's kernels do, plus the
    Gram dgemms of regressors.py:141-142,147,151,153,162,164) on n_sample samples, with every host thread: OpenBLAS threads
    for the dgemms and a thread pool over sample strips for cdist (the reference itself runs cdist on ONE thread, so this
    is the faster of the two).  Returns (samples/s, seconds)."""
    from oracle import nk_oracle as O
    Xs, U, Y = O.synthetic(max(n_sample, m), d, p, seed=seed)
    np.random.seed(0)
    Z = O.draw_landmarks(Y, m)
    Xs, U, Y = Xs[:n_sample], U[:n_sample], Y[:n_sample]
    # all host threads for the BLAS part, also when a launcher (torchrun) exported OMP_NUM_THREADS=1 before numpy was loaded
    try:
        from threadpoolctl import threadpool_limits
        ctx = threadpool_limits(limits=os.cpu_count())
    except Exception:  # noqa: BLE001
        import contextlib
        ctx = contextlib.nullcontext()
    with ctx:
        t0 = time.perf_counter()
        O.grams(Xs, Y, U, Z, O.RBF, np.full(d, 10.0), chunk=n_sample, threads=os.cpu_count() or 1)
        dt = time.perf_counter() - t0
    return n_sample / dt, dt


def unmodified_reference_fit(d, p, m_ref=1024, sizes=(4000, 20000)):
    m_ref = int(os.environ.get("NK_BENCH_REF_M", m_ref))                      # tests shrink it
    if os.environ.get("NK_BENCH_REF_N"):
        sizes = tuple(int(v) for v in os.environ["NK_BENCH_REF_N"].split(","))
    """SURVEY 8(d): the UNMODIFIED reference (baseline/_ref/regressors.py, installed from /root/reference by
    tools/stage_reference.sh) timed on the host: KoopmanNystromRegressor.fit at n in `sizes`, fitted to T(n) = T0 + n / r.
    m = 4096 costs ~420 s per fit (two scipy sqrtm of 138 s each, SURVEY 3.1), so the in-bench fit runs at m_ref landmarks and the
    m = 4096 figures of SURVEY 3.1 are quoted beside it.  Returns None when baseline/_ref is not on this box."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    path = os.path.join(ref_dir, "regressors.py")
    if not os.path.exists(path):
        return None
    import importlib.util
    from oracle import nk_oracle as O
    spec = importlib.util.spec_from_file_location("_nk_unmodified_reference", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    times = []
    for n in sizes:
        Xs, U, Y = O.synthetic(n, d, p, seed=7)
        np.random.seed(0)
        reg = mod.KoopmanNystromRegressor(p, kernel=mod.ThreeDimensionalKernel(10, 10, 10, d), gamma=1e-4, m=m_ref)
        t0 = time.perf_counter()
        reg.fit(np.hstack((Xs, U)), Y)
        times.append(time.perf_counter() - t0)
    (n0, n1), (t0_, t1_) = sizes, times
    r = (n1 - n0) / max(t1_ - t0_, 1e-9)
    T0 = t0_ - n0 / r
    return {"m": m_ref, "d": d, "p": p, "n": list(sizes), "fit_s": [round(t, 3) for t in times], "r_samples_per_s": r, "T0_s": T0,
            "naive_samples_per_s": n1 / t1_,
            "note": "unmodified baseline/_ref/regressors.py::KoopmanNystromRegressor.fit, T(n) = T0 + n/r; at m=4096 the same code measured "
                    "T0 ~ 365 s and r ~ 365 samples/s on 8 vCPU (SURVEY 3.1: 424 s at n=20000) -- too long to repeat inside the bench"}


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    m, d, p = args.m, args.d, args.p
    n_sample = args.cpu_sample
    times = []
    for i in range(args.warmup + args.steps):
        r, dt = cpu_rate(n_sample, m, d, p, seed=i)
        if i >= args.warmup:
            times.append(dt)
    ms = 1e3 * float(np.mean(times))
    value = n_sample / (ms * 1e-3)
    unmod = None if args.no_ref_fit else unmodified_reference_fit(d, p)     # kind stays "port": the per-step value is the port's
    sample = (f"{n_sample} of the {args.n} samples per step, m={m}, d={d}: the n-proportional stage of the reference fit (cdist kernel lift "
              f"+ 7 Gram dgemms, regressors.py:141-164) restated in oracle/nk_oracle.py with the same scipy/numpy calls, cdist spread over "
              f"{os.cpu_count()} threads (the reference runs it on one).  The reference's n-independent stage (2 sqrtm + 2 lstsq + 2 solve, "
              f"T0 ~ 365 s at m=4096 in SURVEY 3.1) is excluded from the per-step time: at n=1e7 that overstates the CPU rate by "
              f"T0 / (n / r) ~ 365 / {args.n / value:.0f} s.  The unmodified reference cannot run n=1e7 (3 x 327 GB of m x n matrices); its own "
              f"fit is timed once per run at m=1024 in `unmodified_reference`")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": workload_config(args, args.gpus),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port", "sample": sample,
                         "unmodified_reference": unmod},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


def workload_config(args, n_gpus):
    return {
        "workload": f"synthetic Koopman fit n={args.n} samples, d={args.d}, p={args.p}, m={args.m} landmarks, FP64 (BASELINE.json configs[3])",
        "n": args.n, "d": args.d, "p": args.p, "m": args.m, "kernel": "RBF length_scale=10 (ThreeDimensionalKernel(10,10,10,d))", "gamma": args.gamma,
        "generator": "x~N(0,I), u~N(0,I), y=tanh(x M^T)+u Bu^T (SURVEY 8d), generated on device per shard",
        "parallelism": (f"sample-sharded x{n_gpus} (rank 0's block shorter by the landmark-only stage it owns), one allreduce of the Grams, "
                        f"one broadcast of S, S^-1") if n_gpus > 1 else "single GPU",
        "l2": "inputs (31.2 GB) are far larger than L2; no flush needed between steps",
    }


# ----------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------
def generate_shard(torch, dev, n_local, d, p, seed, pinned):
    """SURVEY 8(d) generator on the device, in blocks. Returns device X (n,d+p), Y (n,d) and, if asked, pinned host copies."""
    g = torch.Generator(device=dev)
    g.manual_seed(1234)
    M = torch.randn(d, d, dtype=torch.float64, device=dev, generator=g) * (0.9 / d ** 0.5)
    Bu = 0.1 * torch.randn(d, p, dtype=torch.float64, device=dev, generator=g)
    g.manual_seed(seed)
    X = torch.empty(n_local, d + p, dtype=torch.float64, device=dev)
    Y = torch.empty(n_local, d, dtype=torch.float64, device=dev)
    blk = 1 << 19
    for s in range(0, n_local, blk):
        e = min(n_local, s + blk)
        X[s:e].normal_(generator=g)
        Y[s:e] = torch.tanh(X[s:e, :d] @ M.T) + X[s:e, d:] @ Bu.T
    Xh = Yh = None
    if pinned:
        Xh = torch.empty(n_local, d + p, dtype=torch.float64, pin_memory=True)
        Yh = torch.empty(n_local, d, dtype=torch.float64, pin_memory=True)
        Xh.copy_(X); Yh.copy_(Y)
    return X, Y, Xh, Yh


def parity_check(args, eng, R, kernel, world, rank, dist, torch):
    """ONE numpy-generated prefix (oracle/nk_oracle.py::synthetic) is fed to both implementations before the timed
    region, at the bench's own (m, d, p, gamma) -- the same fused-kernel instantiation the timed fit runs.  The seven Grams are
    compared element-wise with the oracle's (scipy cdist + dgemm); A / B / C / weights with the oracle's eigh + Cholesky
    statement of regressors.py:147-169.  `floor` is how far the ORACLE's own A / B / C move when its Grams are summed in another
    chunk order; the A/B/C gate is max(1e-9, 0.5 eps cond(inner_term)) (cond ~ 5e7 at gamma = 1e-4: every float64 statement of the
    solve, the reference's own scipy sequence included, moves by ~1e-9 there -- tests/test_gpu_headline_parity.py gates 1e-9 flat
    at gamma = 1e-3).  Under torchrun the estimator
    path is fit_distributed on the sharded prefix (allreduce and sharded solve included); rank 0 runs the oracle."""
    from nys_koop_lqr_b200 import sharding
    from oracle import nk_oracle as O
    try:        # all host threads for the oracle's BLAS, also when torchrun exported OMP_NUM_THREADS=1 before numpy was loaded
        from threadpoolctl import threadpool_limits
        threadpool_limits(limits=os.cpu_count())
    except Exception:  # noqa: BLE001
        pass
    m, d, p = args.m, args.d, args.p
    n = max(int(args.parity_samples), m)
    Xs, U, Y = O.synthetic(n, d, p, seed=2024)
    np.random.seed(0)
    Z = O.draw_landmarks(Y, m)
    X = np.hstack((Xs, U))
    reg = R.KoopmanNystromRegressor(p, kernel=kernel, gamma=args.gamma, m=m)
    reg.nystrom_centers_output = np.ascontiguousarray(Z.T)
    t0 = time.perf_counter()
    if world > 1:
        off, cnt = sharding.shard_bounds(n, world, rank)
        reg.fit_distributed(torch.from_numpy(X[off:off + cnt]).to(eng.tdev), torch.from_numpy(Y[off:off + cnt]).to(eng.tdev))
    else:
        reg.fit(X, Y)
    gpu_s = time.perf_counter() - t0
    out = None
    if rank == 0:
        dev_ = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(eng.tdev)
        G = eng.grams(dev_(X), dev_(Y), dev_(Z), dev_(np.full(d, 0.1)), 0, p)
        summ = eng.gram_plan(m, d, p)
        t0 = time.perf_counter()
        ls = np.full(d, 10.0)
        thr = os.cpu_count() or 1
        ref = O.grams(Xs, Y, U, Z, O.RBF, ls, chunk=n, threads=thr)
        ref2 = O.grams(Xs, Y, U, Z, O.RBF, ls, chunk=1024, threads=thr)
        Kzz = O.kernel_matrix(Z, Z, O.RBF, ls)
        want = O.solve_abc(ref, Kzz, args.gamma * n, solver="chol")
        want2 = O.solve_abc(ref2, Kzz, args.gamma * n, solver="chol")
        floor = max(O.relerr(a, b) for a, b in zip(want2, want))
        errs = {k: O.relerr(G[k].cpu().numpy(), ref[k]) for k in ("Gxx", "Gyx", "Gyy", "Gxu", "Gyu", "Guu", "GYy")}
        abc = {k: O.relerr(g, w) for k, g, w in zip("ABCW", (reg.A, reg.B, reg.C, reg.weights), want)}
        ev = np.linalg.eigvalsh(np.block([[ref["Gxx"] + args.gamma * n * (Kzz + 1e-6 * np.eye(m)), ref["Gxu"]],
                                          [ref["Gxu"].T, ref["Guu"] + args.gamma * n * np.eye(p)]]))
        cond = float(ev[-1] / ev[0])
        gate_abc = max(1e-9, 0.5 * np.finfo(float).eps * cond)
        out = {"n_prefix": n, "G": max(errs.values()), "A": abc["A"], "B": abc["B"], "C": abc["C"], "W": abc["W"],
               "gate_G": 1e-12, "gate_ABC": gate_abc, "cond_inner_term": cond, "oracle_floor": floor,
               "ok": bool(max(errs.values()) <= 1e-12 and max(abc.values()) <= gate_abc),
               "gram_kernel_nslots": summ["nslots"], "grams": errs, "oracle_s": round(time.perf_counter() - t0, 1), "gpu_fit_s": round(gpu_s, 3),
               "against": "oracle/nk_oracle.py (grams: scipy cdist + dgemm; solve_abc(solver='chol')) on the same numpy prefix, "
                          "reference regressors.py:141-167"}
    if world > 1:
        dist.barrier()
    return out


def run_gpu_arm(args):
    import torch
    import torch.distributed as dist
    from nys_koop_lqr_b200.engine import Engine
    import regressors as R

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this arm has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    distributed = world > 1
    if distributed:
        os.environ.setdefault("NCCL_DEBUG", "WARN")   # its output (incl. the version banner) lands on stderr: see protect_stdout
        dist.init_process_group("nccl", device_id=dev)
    eng = Engine.get(local_rank)
    n, d, p, m = args.n, args.d, args.p, args.m
    from nys_koop_lqr_b200 import sharding
    # rank 0 also owns the landmark-only stage of the fit (square root of K_mm): its block is shorter by that much work
    head = sharding.head_samples(m, d, p) if world > 1 else 0
    counts = [sharding.balanced_bounds(n, world, r, head)[1] for r in range(world)]
    n_local = counts[rank]

    # ---- data (untimed) ----
    X, Y, Xh, Yh = generate_shard(torch, dev, n_local, d, p, seed=1000 + rank, pinned=not args.no_e2e)
    # landmarks: the reference's draw over the GLOBAL index (regressors.py:129-132), same RNG state on every rank
    np.random.seed(0)
    idx = np.random.choice(np.arange(0, n), size=m, replace=False)
    off = int(np.sum(counts[:rank]))
    Zbuf = torch.zeros(m, d, dtype=torch.float64, device=dev)
    mine = np.nonzero((idx >= off) & (idx < off + n_local))[0]
    if mine.size:
        Zbuf[torch.as_tensor(mine, device=dev)] = Y[torch.as_tensor(idx[mine] - off, device=dev)]
    if distributed:
        dist.all_reduce(Zbuf)
    centers = np.ascontiguousarray(Zbuf.cpu().numpy().T)          # (d, m) like the reference attribute
    kernel = R.ThreeDimensionalKernel(10, 10, 10, d)

    def one_fit(Xin, Yin):
        reg = R.KoopmanNystromRegressor(p, kernel=kernel, gamma=args.gamma, m=m)
        reg.nystrom_centers_output = centers.copy()               # fresh object: nothing cached from earlier steps
        if distributed:
            reg.fit_distributed(Xin, Yin)
            if rank == 0:                                          # results are read on the host once per node: rank 0
                t0_ = time.perf_counter()
                _ = (reg.A, reg.B, reg.C, reg.weights)             # downloads them INSIDE the timed region; the other ranks keep theirs on the device
                if getattr(reg, "profile_", None) is not None:
                    reg.profile_["download_ms"] = (time.perf_counter() - t0_) * 1e3
            if getattr(reg, "profile_", None):
                print(f"[bench rank {rank}] fit_distributed phases (ms): " + json.dumps({k: round(v, 1) for k, v in reg.profile_.items()}), file=sys.stderr)
        else:
            reg.fit(Xin, Yin)
        return reg

    def barrier():
        if distributed:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(Xin, Yin, steps, warmup, collect=False):
        reg = None
        for _ in range(warmup):
            reg = one_fit(Xin, Yin)       # same reference pattern as the timed loop (the previous estimator dies when the next one is bound):
        barrier()                         # the pool of page-locked result buffers reaches its steady size during warm-up
        if collect:
            eng.gram_events = []
        l0 = eng.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            reg = one_fit(Xin, Yin)
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1) / steps
        if distributed:
            t = torch.tensor([ms], dtype=torch.float64, device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        launches = eng.launch_count() - l0
        return ms, launches, reg

    parity = None if args.no_parity else parity_check(args, eng, R, kernel, world, rank, dist, torch)
    peak_tflops = eng.probe_dmma_tflops(300.0)                    # FP64 tensor roofline denominator, measured here
    # second, independent denominator: the vendor library's FP64 GEMM on this GPU, this run (8192^3, best of 5)
    peak_dgemm = None
    try:
        a_ = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        b_ = torch.randn(8192, 8192, dtype=torch.float64, device=dev)
        best = float("inf")
        for i in range(6):
            t0_, t1_ = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            t0_.record(); c_ = a_ @ b_; t1_.record(); torch.cuda.synchronize()
            if i:
                best = min(best, t0_.elapsed_time(t1_))
        peak_dgemm = 2.0 * 8192.0 ** 3 / (best * 1e-3) * 1e-12
        del a_, b_, c_
        torch.cuda.empty_cache()
    except Exception as exc:  # noqa: BLE001
        print(f"bench.py: torch.matmul FP64 probe failed: {exc}", file=sys.stderr)
    with ClockSampler(local_rank) as cs:
        ms_step, launches, reg = timed(X, Y, args.steps, args.warmup, collect=True)
    clocks = cs.summary()
    # dominant kernel: the fused lift+Gram kernel, CUDA events on its launching stream
    ev = eng.gram_events or []
    eng.gram_events = None
    k_ms = [a.elapsed_time(b) for a, b, _ in ev]
    k_n = [c for _, _, c in ev]
    F = algorithmic_flops_per_sample(m, d, p)
    kernel_ms = float(np.mean(k_ms)) if k_ms else None
    achieved = (F * float(np.mean(k_n)) / (kernel_ms * 1e-3) * 1e-12) if k_ms else None

    # ---- end to end: samples in pinned host memory, H2D inside the timed region, results read back ----
    e2e = None
    if not args.no_e2e:
        e2e_steps = max(1, min(args.steps, args.e2e_steps))
        ms_e2e, _, reg_e = timed(Xh, Yh, e2e_steps, 1)
        out_bytes = (m * m + m * p + d * m + d * (m + p)) * 8
        e2e = {"value": n / (ms_e2e * 1e-3), "unit": UNIT, "h2d_bytes_per_step": int(n_local * (2 * d + p) * 8 + m * d * 8),
               "d2h_bytes_per_step": int(out_bytes), "ms_per_step": ms_e2e, "steps": e2e_steps,
               "note": "per-rank bytes; samples streamed from pinned host memory in 262144-row blocks on a copy stream overlapped with the fused kernel"}

    # ---- BASELINE.json configs[4] at reduced size, untimed by the headline: a driver-visible record of the CV sweep + rollout ----
    extra = None
    if world == 1 and not args.no_config5:
        try:
            del X, Y, Xh, Yh
            torch.cuda.empty_cache()
            import bench_cv
            c5 = bench_cv.run(bench_cv.parse(["--n", "200000", "--m", "8192", "--kernels", "2", "--gammas", "16", "--traj", "10000"]), standalone=False)
            extra = {"config5_small": {k: c5[k] for k in ("config", "cv_seconds", "phase_seconds", "refit_seconds", "cv_fits_per_s", "gram_pass",
                                                          "batched_solves", "scoring", "rollout", "best")},
                     "note": "BASELINE.json configs[4] (16 x 16 sweep at m=8192, n=2e6, 1e5 trajectories) scaled to 2 lengthscales x 16 gamma, n=2e5, "
                             "1e4 trajectories so that it fits the default run; full size: python bench_cv.py (profiles/)"}
        except Exception as exc:  # noqa: BLE001
            extra = {"config5_small": None, "error": repr(exc)}
    if rank == 0:
        cpu = None
        if not args.no_cpu and world == 1:           # the CPU baseline is reported at N=1 only
            r, dt = cpu_rate(args.cpu_sample, m, d, p)
            cpu = {"value": r, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                   "sample": f"{args.cpu_sample} samples at m={m}, d={d}: n-proportional stage of the reference fit (scipy cdist lift + Gram dgemms) "
                             f"via oracle/nk_oracle.py with cdist on {os.cpu_count()} threads, {dt:.1f} s; the n-independent stage (~365 s at m=4096 in "
                             f"the reference) is excluded; `bench.py --impl reference` also times the unmodified baseline/_ref fit"}
        traffic = None
        tnote = "no ncu capture recorded yet"
        try:
            with open(os.path.join(ROOT, "profiles", "gram_kernel_dram.json")) as f:
                prof = json.load(f)
            traffic = float(prof["dram_bytes_per_sample"]) * float(np.mean(k_n)) if k_n else None
            tnote = prof.get("note", "")
        except (OSError, KeyError, ValueError):
            pass
        line = {
            "metric": METRIC, "value": n / (ms_step * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(args, world),
            "roofline": {"bound": "tensor", "achieved": achieved, "peak": peak_tflops, "unit": "TFLOP/s",
                         "frac": (achieved / peak_tflops) if achieved else None, "traffic": traffic,
                         "kernel": "nk::gram_kernel (fused kernel lift + Gram, FP64 DMMA)", "kernel_ms": kernel_ms,
                         "algorithmic_flops_per_sample": F, "samples_per_launch": float(np.mean(k_n)) if k_n else None,
                         "peak_source": "register-only DMMA.8x8x4 issue-rate probe (nk_probe_dmma_tflops) run on this GPU just before the timed region; "
                                        "MEASURED_PEAKS.json has no FP64 entry; peak_dgemm = torch.matmul FP64 8192^3 (cuBLAS) measured in this run",
                         "peak_dgemm": peak_dgemm, "frac_dgemm": (achieved / peak_dgemm) if (achieved and peak_dgemm) else None,
                         "whole_fit_frac": F * n / world / (ms_step * 1e-3) * 1e-12 / peak_tflops, "traffic_note": tnote,
                         "peak_nominal": 40.0, "frac_nominal": (achieved / 40.0) if achieved else None,
                         "nominal_note": "NVIDIA's B200 FP64 tensor figure (40 TFLOP/s); the DMMA issue rate measured on this pool is 37.1",
                         "gpu_launches_note": "gpu_launches = kernels of libnkb200.so launched per step (one fit), counted by the handle"},
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": int(round(launches / max(args.steps, 1))), "clocks": clocks, "parity": parity, "extra": extra,
            "model_check": {"A_shape": list(reg.A.shape), "A_fro": float(np.linalg.norm(reg.A)), "finite": bool(np.isfinite(reg.A).all() and np.isfinite(reg.C).all())},
        }
        emit(line)
    if distributed:
        dist.barrier()
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--samples", "--n-samples", dest="n", type=int, default=10_000_000)
    ap.add_argument("--m", type=int, default=4096)
    ap.add_argument("--d", type=int, default=192)
    ap.add_argument("--p", type=int, default=6)
    ap.add_argument("--gamma", type=float, default=1e-4)
    ap.add_argument("--cpu-sample", type=int, default=8192)   # ~7-20 s of host work per pass (single-threaded cdist dominates)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-ref-fit", action="store_true", help="reference arm: skip the one unmodified baseline/_ref fit")
    ap.add_argument("--no-config5", action="store_true", help="skip the reduced configs[4] (CV sweep + rollout) side measurement")
    ap.add_argument("--no-parity", action="store_true", help="skip the oracle parity check on a numpy prefix before the timed region")
    ap.add_argument("--parity-samples", type=int, default=8192)
    args = ap.parse_args()
    protect_stdout()
    if args.impl == "reference":
        return run_reference_arm(args)
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if world != args.gpus and world > 1:
        print(f"bench.py: WORLD_SIZE={world} but --gpus {args.gpus}", file=sys.stderr)
    return run_gpu_arm(args)


if __name__ == "__main__":
    sys.exit(main())
