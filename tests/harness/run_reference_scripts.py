#!/usr/bin/env python
"""Runs the reference's three experiment scripts' own functions against (a) the unmodified reference estimator on the
host cores and (b) this repo's drop-in `regressors` module on the B200, same seeds, and reports parity of
(A, B, C), the Riccati gain, the open-loop forecast RMSE and the closed-loop trajectories, next to the reference's own
self-floor (same landmarks, permuted samples).

The scripts are not modified: each is loaded as a module (its `__main__` block does not run) with shim modules for
`control` (dlqr -> scipy DARE; python-control is not installed) and `matplotlib` (no-op), and its hard-coded globals
(`n_inputs`, `n_states`, `dynamical_system`) are set the way its `__main__` sets them.  The reference tree is read
from NK_REFERENCE_PATH (default: /root/reference, else baseline/_ref staged by tools/stage_reference.sh -- git-ignored).

    python tests/harness/run_reference_scripts.py [--quick] [--out profiles/r01_reference_scripts_parity.md]
"""
from __future__ import annotations

import argparse
import importlib.util
import json
import os
import pathlib
import sys
import time
import types

import numpy as np
import scipy.linalg

ROOT = pathlib.Path(__file__).resolve().parents[2]


def find_reference() -> pathlib.Path:
    for c in (os.environ.get("NK_REFERENCE_PATH"), "/root/reference", str(ROOT / "baseline" / "_ref")):
        if c and (pathlib.Path(c) / "regressors.py").exists():
            return pathlib.Path(c)
    raise SystemExit("reference tree not found (set NK_REFERENCE_PATH or run tools/stage_reference.sh)")


REF = find_reference()
os.environ["NK_REFERENCE_PATH"] = str(REF)


# ------------------------------------------------------------------------------------------- shims
def install_shims():
    ctl = types.ModuleType("control")

    def dlqr(A, B, Q, R):
        P = scipy.linalg.solve_discrete_are(A, B, Q, R)
        K = np.linalg.solve(B.T @ P @ B + R, B.T @ P @ A)
        return K, P, np.linalg.eigvals(A - B @ K)
    ctl.dlqr = dlqr
    sys.modules["control"] = ctl

    class _Noop:
        def __getattr__(self, name):
            return _Noop()

        def __call__(self, *a, **k):
            return _Noop()

        def __iter__(self):
            return iter(())
    def _attr(name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Noop()
    mpl = types.ModuleType("matplotlib")
    mpl.use = lambda *a, **k: None
    mpl.__getattr__ = _attr
    plt = types.ModuleType("matplotlib.pyplot")
    plt.__getattr__ = _attr
    mpl.pyplot = plt
    sys.modules["matplotlib"] = mpl
    sys.modules["matplotlib.pyplot"] = plt
    for sub in ("animation", "cm", "colors", "ticker"):
        m = types.ModuleType(f"matplotlib.{sub}")
        m.__getattr__ = _attr
        sys.modules[f"matplotlib.{sub}"] = m
        setattr(mpl, sub, m)


def load_regressors(which: str):
    """'ref' -> the unmodified reference module; 'b200' -> this repo's drop-in."""
    if which == "ref":
        spec = importlib.util.spec_from_file_location("regressors", REF / "regressors.py")
    else:
        spec = importlib.util.spec_from_file_location("regressors", ROOT / "regressors.py")
    mod = importlib.util.module_from_spec(spec)
    sys.modules["regressors"] = mod
    spec.loader.exec_module(mod)
    return mod


def load_script(name: str, which: str):
    """Imports benchmark_*.py as a module with `regressors` bound to the chosen implementation."""
    if str(ROOT) not in sys.path:
        sys.path.insert(0, str(ROOT))
    if str(REF) not in sys.path:
        sys.path.append(str(REF))          # dynamical_systems.py
    regs = load_regressors(which)
    spec = importlib.util.spec_from_file_location(f"{name}_{which}", REF / f"{name}.py")
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod._regs = regs
    return mod


def relerr(a, b):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-300))


class Report:
    def __init__(self):
        self.rows = []
        self.timing = []

    def add(self, script, config, quantity, err, floor=None, note=""):
        self.rows.append(dict(script=script, config=config, quantity=quantity, err=err, floor=floor, note=note))
        fl = f" (reference self-floor {floor:.1e})" if floor is not None else ""
        print(f"[{script}] {config}: {quantity}: {err:.2e}{fl} {note}", flush=True)


def fit_pair(mods, make, X, Y, seed):
    """Fits the same configuration with both implementations (same RNG state) + a permuted-sample refit of the
    reference with the same landmarks (self-floor)."""
    out = {}
    for which, mod in mods.items():
        np.random.seed(seed)
        reg = make(mod)
        t0 = time.perf_counter()
        reg.fit(X.T, Y.T)
        out[which] = reg
        out[which + "_s"] = time.perf_counter() - t0
    ref = out["ref"]
    assert np.array_equal(np.asarray(out["b200"].nystrom_centers_output), np.asarray(ref.nystrom_centers_output)), "landmark draw differs"
    perm = np.random.default_rng(0).permutation(X.shape[1])
    reg2 = make(mods["ref"])
    reg2.nystrom_centers_output = ref.nystrom_centers_output
    reg2.nystrom_centers_input = ref.nystrom_centers_output
    reg2.fit(X.T[perm], Y.T[perm])
    out["floor"] = {k: relerr(getattr(reg2, k), getattr(ref, k)) for k in ("A", "B", "C")}
    out["ref_perm"] = reg2
    return out


def compare_model(rep, script, config, pair):
    for k in ("A", "B", "C"):
        rep.add(script, config, k, relerr(getattr(pair["b200"], k), getattr(pair["ref"], k)), pair["floor"][k])
    rep.timing.append(dict(script=script, config=config, ref_fit_s=pair["ref_s"], b200_fit_s=pair["b200_s"]))


def solver_study(rep, script, config, pair, X, Y, kind, ls, gamma):
    """Where the B200 result sits far above the reference's self-floor: separate Gram error from solver-algorithm error.
    (1) the GPU Grams pushed through the reference's own scipy sqrtm/solve/lstsq sequence (oracle, verification only) must land
    on the reference's floor; (2) both answers against a high-precision solve of the same float64 system (Cholesky + iterative
    refinement with long-double residuals): the SPD-Cholesky dense stage is the more accurate of the two."""
    import torch
    from oracle import nk_oracle as O
    from nys_koop_lqr_b200.engine import Engine
    eng = Engine.get()
    ref, b2 = pair["ref"], pair["b200"]
    d = Y.shape[0]
    p = X.shape[0] - d
    Z = np.ascontiguousarray(np.asarray(ref.nystrom_centers_output).T)
    n = X.shape[1]
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()
    G = eng.grams(t(X.T), t(Y.T), t(Z), t(1.0 / np.asarray(ls, dtype=float)), kind, p)
    Gh = {k: v.cpu().numpy() for k, v in G.items() if k != "_flat"}
    Kzz = O.kernel_matrix(Z, Z, kind, ls)
    A1, B1, C1, _ = O.solve_abc(Gh, Kzz, gamma * n, solver="reference")
    rep.add(script, config, "A: GPU Grams -> reference's scipy sqrtm/solve/lstsq (verification mode)", relerr(A1, ref.A), pair["floor"]["A"])
    rep.add(script, config, "C: GPU Grams -> reference's scipy sqrtm/solve/lstsq (verification mode)", relerr(C1, ref.C), pair["floor"]["C"])
    # high-precision solve of inner * sol = right
    m = Z.shape[0]
    Kmm = Kzz + 1e-6 * np.eye(m)
    w, V = np.linalg.eigh(Kmm)
    Sinv = (V / np.sqrt(w)) @ V.T
    inner = np.block([[Gh["Gxx"] + gamma * n * Kmm, Gh["Gxu"]], [Gh["Gxu"].T, Gh["Guu"] + gamma * n * np.eye(p)]])
    right = scipy.linalg.block_diag(Kzz @ Sinv, np.eye(p))
    left = Sinv @ np.hstack((Gh["Gyx"], Gh["Gyu"]))
    cf = scipy.linalg.cho_factor(inner, lower=True)
    sol = scipy.linalg.cho_solve(cf, right)
    iL, rL = inner.astype(np.longdouble), right.astype(np.longdouble)
    for _ in range(40):
        res = (rL - iL @ sol.astype(np.longdouble)).astype(np.float64)
        sol = sol + scipy.linalg.cho_solve(cf, res)
    A_hp = (left @ sol)[:, :m]
    rep.add(script, config, "A vs high-precision solve: reference (lstsq/gelsd)", relerr(ref.A, A_hp), note=f"cond(inner)={np.linalg.cond(inner):.1e}")
    rep.add(script, config, "A vs high-precision solve: B200 (Cholesky)", relerr(b2.A, A_hp))


def gain_of(mod, reg, qscale):
    Q = qscale * reg.C.T @ reg.C
    Q = (Q + Q.T) / 2
    return mod.control.dlqr(reg.A, reg.B, Q, np.eye(reg.B.shape[1]))[0]


# ------------------------------------------------------------------------------------------- Duffing
def run_classic(rep, quick):
    mods = {w: load_script("benchmark_lqr_classic", w) for w in ("ref", "b200")}
    csv = lambda f: np.loadtxt(REF / "duffing" / f, delimiter=",")
    X = np.hstack((csv("duffing_x_forced.csv"), csv("duffing_x_unforced.csv")))
    U = np.hstack((csv("duffing_u_forced.csv").reshape(1, -1), np.zeros((1, csv("duffing_x_unforced.csv").shape[1]))))
    X = np.vstack((X, U))
    Y = np.hstack((csv("duffing_y_forced.csv"), csv("duffing_y_unforced.csv")))
    params = dict(Ts=0.01, name="duffing", n_states=2, n_inputs=1, radius_sampling=1.0, angle_sampling=2, input_lb=[-1], input_ub=[1])
    golden = np.loadtxt(REF / "duffing" / "all_rmses_nystrom_double_dataset.csv")
    ms_all = np.around(np.logspace(1, 2.3, num=20)).astype(int)
    for w, mod in mods.items():
        mod.dynamical_system = mod.DuffingOscillator(**params)
        mod.n_inputs, mod.n_states = 1, 2
    seeds = range(1 if quick else 3)
    m_idx = [0, 2, 5] if quick else [0, 1, 2, 3, 5, 8]
    for seed in seeds:
        trajs = {}
        for w, mod in mods.items():
            np.random.seed(seed)
            trajs[w] = mod.simulate_true_system(mod.dynamical_system, 2)
        traj, ctrl = trajs["ref"]
        for j in m_idx:
            m = int(ms_all[j])
            make = lambda mod, m=m: mod.KoopmanNystromRegressor(1, kernel=mod.KernelWrapper([1, 1]), gamma=1e-6, m=m)
            pair = fit_pair(mods, make, X, Y, seed * 1000 + j)
            cfg = f"open-loop seed={seed} m={m}"
            compare_model(rep, "classic", cfg, pair)
            r = {w: mods[w].validate_dyn_sys(pair[w], traj, ctrl) for w in ("ref", "b200")}
            r_floor = mods["ref"].validate_dyn_sys(pair["ref_perm"], traj, ctrl)
            rep.add("classic", cfg, "forecast RMSE %", abs(r["b200"] - r["ref"]) / r["ref"], abs(r_floor - r["ref"]) / r["ref"],
                    note=f"ref {r['ref']:.8g} b200 {r['b200']:.8g}")
        # golden CSV protocol (two draws, first used) for the m=10 column
        np.random.seed(seed)
        idx = np.random.choice(np.arange(0, X.shape[1]), size=10, replace=False)
        np.random.choice(np.arange(0, X.shape[1]), size=10, replace=False)
        reg = mods["b200"].KoopmanNystromRegressor(1, kernel=mods["b200"].KernelWrapper([1, 1]), gamma=1e-6, m=10)
        reg.nystrom_centers_output = Y[:, idx]
        reg.fit(X.T, Y.T)
        got = mods["b200"].validate_dyn_sys(reg, traj, ctrl)
        rep.add("classic", f"golden G1 seed={seed} m=10", "forecast RMSE % vs all_rmses_nystrom_double_dataset.csv",
                abs(got - golden[seed, 0]) / golden[seed, 0], note=f"csv {golden[seed, 0]:.8g} b200 {got:.8g}")
    # LQR branch (benchmark_lqr_classic.py:265-292), m = 20
    for seed in seeds:
        make = lambda mod: mod.KoopmanNystromRegressor(1, kernel=mod.KernelWrapper([1, 1]), gamma=1e-6, m=20)
        pair = fit_pair(mods, make, X, Y, seed)
        cfg = f"LQR seed={seed} m=20"
        compare_model(rep, "classic", cfg, pair)
        K = {w: gain_of(mods[w], pair[w], 1.0) for w in ("ref", "b200")}
        Kf = gain_of(mods["ref"], pair["ref_perm"], 1.0)
        rep.add("classic", cfg, "Riccati gain K", relerr(K["b200"], K["ref"]), relerr(Kf, K["ref"]))
        steps = 60 if quick else 300
        init, refp = np.array([-0.5, 0.0]).reshape(-1, 1), np.zeros((2, 1))
        cl = {w: mods[w].lqr_control(steps, refp, init, pair[w], K[w]) for w in ("ref", "b200")}
        rep.add("classic", cfg, f"closed-loop x1 over {steps} steps (lift per step)", relerr(cl["b200"][0], cl["ref"][0]))


# ------------------------------------------------------------------------------------------- cloth
def run_cloth(rep, quick):
    mods = {w: load_script("benchmark_lqr_cloth", w) for w in ("ref", "b200")}
    p = REF / "8x8_cloth_swing_xyz"
    n_trajs = 50
    trajs = [np.loadtxt(p / f"state_samples_cloth_swing_{i}.csv", delimiter=",").T for i in range(n_trajs)]
    ctrls = [np.loadtxt(p / f"input_samples_cloth_swing_{i}.csv", delimiter=",")[:, :6].T for i in range(n_trajs)]
    all_trajs, all_controls = trajs[10:], ctrls[10:]
    for mod in mods.values():
        mod.n_states, mod.n_inputs = 192, 6
    ms_all = np.logspace(1.0, 2.6, num=20, dtype=int)
    # open-loop branch (benchmark_lqr_cloth.py:168-207), seed 0
    seed = 0
    np.random.seed(seed)
    idx = np.arange(0, n_trajs - 10)
    np.random.shuffle(idx)
    train, test = idx[:30], idx[30:]
    X, Y = mods["ref"].create_data_matrices(all_trajs, all_controls, train)
    traj, ctrl = all_trajs[test[0]], all_controls[test[0]]
    for j in ([0, 4, 8] if quick else [0, 2, 4, 6, 8, 12]):
        m = int(ms_all[j])
        make = lambda mod, m=m: mod.KoopmanNystromRegressor(6, kernel=mod.ThreeDimensionalKernel(10, 10, 10, 192), gamma=1e-7, m=m)
        pair = fit_pair(mods, make, X, Y, 100 + j)
        cfg = f"open-loop seed=0 m={m}"
        compare_model(rep, "cloth", cfg, pair)
        r = {w: mods[w].validate_dyn_sys(pair[w], traj, ctrl) for w in ("ref", "b200")}
        r_floor = mods["ref"].validate_dyn_sys(pair["ref_perm"], traj, ctrl)
        rep.add("cloth", cfg, "forecast RMSE", abs(r["b200"] - r["ref"]) / r["ref"], abs(r_floor - r["ref"]) / r["ref"],
                note=f"ref {r['ref']:.8g} b200 {r['b200']:.8g}")
    # LQR branch (benchmark_lqr_cloth.py:212-270), m = 100, training set = trajectories 0..29 of the 40
    X, Y = mods["ref"].create_data_matrices(all_trajs, all_controls, np.arange(0, 30))
    make = lambda mod: mod.KoopmanNystromRegressor(6, kernel=mod.ThreeDimensionalKernel(10, 10, 10, 192), gamma=1e-7, m=100)
    pair = fit_pair(mods, make, X, Y, 0)
    compare_model(rep, "cloth", "LQR seed=0 m=100", pair)
    solver_study(rep, "cloth", "LQR seed=0 m=100", pair, X, Y, 0, np.full(192, 10.0), 1e-7)
    K = {w: gain_of(mods[w], pair[w], 0.0075) for w in ("ref", "b200")}
    Kf = gain_of(mods["ref"], pair["ref_perm"], 0.0075)
    rep.add("cloth", "LQR seed=0 m=100", "Riccati gain K", relerr(K["b200"], K["ref"]), relerr(Kf, K["ref"]))
    # stored golden regressor (G3) has the same landmarks for seed 0
    init = all_trajs[0][:, 0].reshape(-1, 1)
    reference = init + 0.05
    cl = {w: mods[w].lqr_control(60, reference, init, pair[w], K[w]) for w in ("ref", "b200")}
    rep.add("cloth", "LQR seed=0 m=100", "closed-loop lifted rollout x over 60 steps", relerr(cl["b200"][0], cl["ref"][0]),
            relerr(mods["ref"].lqr_control(60, reference, init, pair["ref_perm"], Kf)[0], cl["ref"][0]))
    # the batched device closed loop (nk_closed_loop) against the script's own numpy loop on the SAME fitted model and gain:
    # isolates the kernel from the solver-algorithm gap above
    states, ctrls = pair["b200"].closed_loop(K["b200"], init, reference, 60)
    xs_script = cl["b200"][0]                                   # (64, 61): x coordinates incl. the initial state
    rep.add("cloth", "LQR seed=0 m=100", "device closed loop (nk_closed_loop) vs the script's loop, same model: x over 60 steps",
            relerr(states[0::3, :], xs_script[:, 1:]))


# ------------------------------------------------------------------------------------------- HJB
def run_hjb(rep, quick):
    mods = {w: load_script("benchmark_lqr_hjb", w) for w in ("ref", "b200")}
    params = dict(Ts=0.01, name="hjb", n_states=1, n_inputs=1, state_lb=-1.0, state_ub=1.0, input_lb=[-1], input_ub=[1])
    for mod in mods.values():
        mod.dynamical_system = mod.HJB(**params)
        mod.n_inputs, mod.n_states = 1, 1
    np.random.seed(0)
    X, Y = mods["ref"].generate_dataset(mods["ref"].dynamical_system, 20, int(2 // 0.01))
    np.random.seed(1)
    traj, ctrl = mods["ref"].simulate_true_system(mods["ref"].dynamical_system, 2)
    for gamma in ((1e-3,) if quick else (1e-6, 1e-3)):
        make = lambda mod, g=gamma: mod.KoopmanNystromRegressor(1, kernel=mod.KernelWrapper([1.0]), gamma=g, m=100)
        pair = fit_pair(mods, make, X, Y, 0)
        cfg = f"m=100 l=1 gamma={gamma:g}"
        compare_model(rep, "hjb", cfg, pair)
        if gamma < 1e-4:
            solver_study(rep, "hjb", cfg, pair, X, Y, 1, [1.0], gamma)
        r = {w: mods[w].validate_dyn_sys(pair[w], traj, ctrl) for w in ("ref", "b200")}
        r_floor = mods["ref"].validate_dyn_sys(pair["ref_perm"], traj, ctrl)
        rep.add("hjb", cfg, "forecast RMSE %", abs(r["b200"] - r["ref"]) / r["ref"], abs(r_floor - r["ref"]) / r["ref"],
                note=f"ref {r['ref']:.8g} b200 {r['b200']:.8g}")
        K = {w: gain_of(mods[w], pair[w], 1.0) for w in ("ref", "b200")}
        Kf = gain_of(mods["ref"], pair["ref_perm"], 1.0)
        rep.add("hjb", cfg, "Riccati gain K", relerr(K["b200"], K["ref"]), relerr(Kf, K["ref"]))
        steps = 40 if quick else 200
        init, refp = np.array([0.9]).reshape(-1, 1), np.zeros((1, 1))
        cl = {w: mods[w].lqr_control(steps, refp, init, pair[w], K[w]) for w in ("ref", "b200")}
        rep.add("hjb", cfg, f"closed-loop state over {steps} steps", relerr(cl["b200"][0], cl["ref"][0]))
    if not quick:
        # CV path (benchmark_lqr_hjb.py:47-71): GridSearchCV -> clone -> fit -> predict, n_jobs forced to 1, reduced grid
        from sklearn.model_selection import GridSearchCV
        scores = {}
        for w, mod in mods.items():
            np.random.seed(0)
            clf = GridSearchCV(mod.KoopmanNystromRegressor(1), {"kernel": [mod.KernelWrapper([1.0])], "gamma": [1e-5, 1e-3], "m": [100]},
                               scoring="neg_root_mean_squared_error", n_jobs=1)
            t0 = time.perf_counter()
            clf.fit(X.T, Y.T)
            scores[w] = (clf.cv_results_["mean_test_score"], time.perf_counter() - t0)
        rep.add("hjb", "GridSearchCV 2 candidates x 5 folds, m=100", "mean_test_score", relerr(scores["b200"][0], scores["ref"][0]),
                note=f"ref {scores['ref'][1]:.1f} s, b200 {scores['b200'][1]:.1f} s")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--quick", action="store_true")
    ap.add_argument("--out", default=str(ROOT / "gpurun_out" / "reference_scripts_parity.md"))
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    import torch  # noqa: F401  (import before the shims: torch inspects every module in sys.modules)
    install_shims()
    rep = Report()
    for name, fn in (("classic", run_classic), ("cloth", run_cloth), ("hjb", run_hjb)):
        if args.only and name not in args.only:
            continue
        t0 = time.perf_counter()
        fn(rep, args.quick)
        print(f"== {name} done in {time.perf_counter() - t0:.1f} s", flush=True)
    out = pathlib.Path(args.out)
    out.parent.mkdir(parents=True, exist_ok=True)
    with open(out, "w") as f:
        f.write("# Reference scripts: B200 drop-in vs unmodified reference (same seeds)\n\n")
        f.write("Produced by `tests/harness/run_reference_scripts.py` on the GPU box: the scripts' own functions (`validate_dyn_sys`, `lqr_control`, "
                "`create_data_matrices`, `generate_dataset`, `simulate_true_system`) run unmodified against both `regressors` modules. "
                "`err` = relative error of the B200 result w.r.t. the reference; `self-floor` = how far the reference moves from itself "
                "when its training samples are permuted (same landmarks) -- where cond(inner_term) is large that is the meaningful yardstick (SURVEY 8c).\n\n")
        f.write("| script | configuration | quantity | err | reference self-floor | note |\n|---|---|---|---:|---:|---|\n")
        for r in rep.rows:
            fl = f"{r['floor']:.1e}" if r["floor"] is not None else ""
            f.write(f"| {r['script']} | {r['config']} | {r['quantity']} | {r['err']:.1e} | {fl} | {r['note']} |\n")
        f.write("\n## Fit wall time (host-to-host, includes upload and result download)\n\n| script | configuration | reference fit (s) | B200 fit (s) |\n|---|---|---:|---:|\n")
        for t in rep.timing:
            f.write(f"| {t['script']} | {t['config']} | {t['ref_fit_s']:.3f} | {t['b200_fit_s']:.3f} |\n")
    with open(out.with_suffix(".json"), "w") as f:
        json.dump(dict(rows=rep.rows, timing=rep.timing), f, indent=1)
    print("wrote", out)


if __name__ == "__main__":
    main()
