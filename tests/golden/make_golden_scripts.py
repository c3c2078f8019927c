"""Generates tests/golden/scripts/*.npz: the LQR / open-loop configurations of the reference's three experiment scripts
(BASELINE.json configs[0..2]), produced by the scripts' OWN functions driving the UNMODIFIED reference estimator in this
container (tests/harness/run_reference_scripts.py loads the scripts as modules with `control` / `matplotlib` shims).

    hjb_m100_g1e-3, hjb_m100_g1e-6   benchmark_lqr_hjb.py:166-168 dataset (20 x 199 samples), m=100 (:184), Matern l=1
    duffing_lqr_m20                  benchmark_lqr_classic.py:174-178 dataset (n=69 900, from tests/golden/g1), m=20 (:265)
    cloth_lqr_m100                   benchmark_lqr_cloth.py:212-270 (training trajectories 0..29 of the 40, from tests/golden/g2), m=100

Each fixture holds: the landmarks of the reference's RNG call, the reference's A, B, C, its self-floor (same landmarks,
samples permuted), the Riccati gain (scipy DARE as the control.dlqr stand-in) and its floor, cond(inner_term), the open-loop
forecast RMSE of validate_dyn_sys, the script's closed loop (hjb / classic: lqr_control on the true RK system with a lift per
step; cloth: the lifted-model loop), and -- the SURVEY 8c protocol for the ill-conditioned configurations -- A from a
HIGH-PRECISION evaluation of the same formulas on the same float64 Grams (square root from a 50-digit eigen-decomposition,
solves refined with long-double residuals, long-double products), with the distances of the reference (sqrtm / lstsq) and of
plain float64 eigh + Cholesky to it.

Run:  python tests/golden/make_golden_scripts.py      (needs /root/reference; the fixtures are committed)
"""
import pathlib
import sys

import numpy as np
import scipy.linalg

HERE = pathlib.Path(__file__).resolve().parent
ROOT = HERE.parents[1]
sys.path.insert(0, str(ROOT))
sys.path.insert(0, str(ROOT / "tests" / "harness"))
import run_reference_scripts as H          # noqa: E402  (script loader + shims; REF = /root/reference)
from oracle import nk_oracle as O          # noqa: E402

OUT = HERE / "scripts"


def exact_sqrt_pair(Kmm, digits=50):
    """S = K_mm^(1/2) and S^-1 from a 50-digit symmetric eigen-decomposition (mpmath), rounded to long double.  float64 eigh /
    sqrtm / polar roots all carry ~eps * cond(K_mm) relative error in S^-1 (1e-8 on the Matern configurations), which the product
    S^-1 [Gyx|Gyu] inner^-1 ... amplifies ~100 x; a "truth" must not share that error with any of the candidates."""
    import mpmath as mp
    mp.mp.dps = digits
    m = Kmm.shape[0]
    E, Q = mp.eigsy(mp.matrix(Kmm.tolist()))
    ld = lambda M: np.array([[np.longdouble(mp.nstr(M[i, j], 25)) for j in range(M.cols)] for i in range(M.rows)], dtype=np.longdouble)
    Ql = ld(Q)
    w = np.array([np.longdouble(mp.nstr(E[i], 25)) for i in range(m)], dtype=np.longdouble)
    return (Ql * np.sqrt(w)) @ Ql.T, (Ql / np.sqrt(w)) @ Ql.T


def hp_truth(G, Kzz, gamma_n, refinements=60):
    """[A|B] and C from the SAME float64 Grams and kernel matrix, everything downstream in extended precision: S, S^-1 from a
    50-digit eigen-decomposition, the two solves by Cholesky + iterative refinement with long-double residuals and a long-double
    solution, the products in long double.  `*_chol`: the same formulas in plain float64 with an eigh root (the CPU statement of
    what the GPU dense stage computes)."""
    m, p = Kzz.shape[0], G["Guu"].shape[0]
    L = np.longdouble
    Kmm = Kzz + 1e-6 * np.eye(m)
    S_x, Sinv_x = exact_sqrt_pair(Kmm)
    w, V = np.linalg.eigh(Kmm)
    S, Sinv = (V * np.sqrt(w)) @ V.T, (V / np.sqrt(w)) @ V.T
    inner = np.block([[G["Gxx"] + gamma_n * Kmm, G["Gxu"]], [G["Gxu"].T, G["Guu"] + gamma_n * np.eye(p)]])
    cross = np.hstack((G["Gyx"], G["Gyu"]))

    def refined(M, R_exact):
        cf = scipy.linalg.cho_factor(M, lower=True)
        x = scipy.linalg.cho_solve(cf, R_exact.astype(np.float64)).astype(L)
        ML = M.astype(L)
        for _ in range(refinements):
            x = x + scipy.linalg.cho_solve(cf, (R_exact - ML @ x).astype(np.float64)).astype(L)
        return x
    right_x = np.zeros((m + p, m + p), dtype=L)
    right_x[:m, :m] = Kzz.astype(L) @ Sinv_x
    right_x[m:, m:] = np.eye(p)
    G_hp = ((Sinv_x @ cross.astype(L)) @ refined(inner, right_x)).astype(np.float64)
    C_hp = (G["GYy"].astype(L) @ refined(gamma_n * Kmm + G["Gyy"], S_x)).astype(np.float64)
    # plain float64 statement (eigh root, Cholesky, reference association order)
    right = scipy.linalg.block_diag(Kzz @ Sinv, np.eye(p))
    G_chol = (Sinv @ cross) @ scipy.linalg.cho_solve(scipy.linalg.cho_factor(inner, lower=True), right)
    C_chol = G["GYy"] @ scipy.linalg.cho_solve(scipy.linalg.cho_factor(gamma_n * Kmm + G["Gyy"], lower=True), S)
    return dict(G_hp=G_hp, G_chol=G_chol, C_hp=C_hp, C_chol=C_chol, cond_inner=np.linalg.cond(inner),
                sinv_eigh_vs_exact=O.relerr(Sinv, Sinv_x.astype(np.float64)), cond_kmm=float(w[-1] / w[0]))


def fit_reference(mod, make, X, Y, seed):
    """X (d+p, n), Y (d, n) column samples as the scripts hold them.  Returns the fitted reference estimator + permuted refit."""
    np.random.seed(seed)
    reg = make(mod)
    reg.fit(X.T, Y.T)
    perm = np.random.default_rng(0).permutation(X.shape[1])
    reg2 = make(mod)
    reg2.nystrom_centers_output = reg.nystrom_centers_output
    reg2.nystrom_centers_input = reg.nystrom_centers_output
    reg2.fit(X.T[perm], Y.T[perm])
    return reg, reg2


def common(reg, reg2, X, Y, kind, ls, gamma, qscale, mod):
    d = Y.shape[0]
    p = X.shape[0] - d
    n = X.shape[1]
    Z = np.ascontiguousarray(np.asarray(reg.nystrom_centers_output).T)
    G = O.grams(X[:d].T, Y.T, X[d:].T, Z, kind, ls, chunk=n)
    Kzz = O.kernel_matrix(Z, Z, kind, ls)
    hp = hp_truth(G, Kzz, gamma * n)
    m = Z.shape[0]
    K = H.gain_of(mod, reg, qscale)
    K2 = H.gain_of(mod, reg2, qscale)
    out = dict(Z=np.asarray(reg.nystrom_centers_output), A=reg.A, B=reg.B, C=reg.C, W=reg.weights, kind=kind, ls=np.asarray(ls, dtype=float),
               gamma=gamma, m=m, n_inputs=p, qscale=qscale,
               floor_A=O.relerr(reg2.A, reg.A), floor_B=O.relerr(reg2.B, reg.B), floor_C=O.relerr(reg2.C, reg.C),
               K_lqr=K, floor_K=O.relerr(K2, K), cond_inner=hp["cond_inner"],
               A_hp=hp["G_hp"][:, :m], B_hp=hp["G_hp"][:, m:], C_hp=hp["C_hp"],
               ref_vs_hp_A=O.relerr(reg.A, hp["G_hp"][:, :m]), chol_vs_hp_A=O.relerr(hp["G_chol"][:, :m], hp["G_hp"][:, :m]),
               ref_vs_hp_C=O.relerr(reg.C, hp["C_hp"]), chol_vs_hp_C=O.relerr(hp["C_chol"], hp["C_hp"]),
               sinv_eigh_vs_exact=hp["sinv_eigh_vs_exact"], cond_kmm=hp["cond_kmm"])
    return out


def describe(name, out):
    print(f"{name}: m={int(out['m'])} cond(inner)={float(out['cond_inner']):.1e} floor A/B/C {float(out['floor_A']):.1e}/{float(out['floor_B']):.1e}/"
          f"{float(out['floor_C']):.1e} K {float(out['floor_K']):.1e}; vs HP truth: reference A {float(out['ref_vs_hp_A']):.1e} C {float(out['ref_vs_hp_C']):.1e}, "
          f"float64 Cholesky A {float(out['chol_vs_hp_A']):.1e} C {float(out['chol_vs_hp_C']):.1e}; cond(K_mm) {float(out['cond_kmm']):.1e}, "
          f"eigh S^-1 vs exact {float(out['sinv_eigh_vs_exact']):.1e}", flush=True)


def make_hjb():
    mod = H.load_script("benchmark_lqr_hjb", "ref")
    params = dict(Ts=0.01, name="hjb", n_states=1, n_inputs=1, state_lb=-1.0, state_ub=1.0, input_lb=[-1], input_ub=[1])
    mod.dynamical_system = mod.HJB(**params)
    mod.n_inputs, mod.n_states = 1, 1
    np.random.seed(0)
    X, Y = mod.generate_dataset(mod.dynamical_system, 20, int(2 // 0.01))           # benchmark_lqr_hjb.py:166-168
    np.random.seed(1)
    traj, ctrl = mod.simulate_true_system(mod.dynamical_system, 2)
    for gamma in (1e-3, 1e-6):
        make = lambda md, g=gamma: md.KoopmanNystromRegressor(1, kernel=md.KernelWrapper([1.0]), gamma=g, m=100)
        reg, reg2 = fit_reference(mod, make, X, Y, 0)
        out = common(reg, reg2, X, Y, O.MATERN52, [1.0], gamma, 1.0, mod)
        out.update(X=X.T.copy(), Y=Y.T.copy(), traj=traj, ctrl=ctrl, rmse_percent=mod.validate_dyn_sys(reg, traj, ctrl),
                   rmse_percent_floor=abs(mod.validate_dyn_sys(reg2, traj, ctrl) - mod.validate_dyn_sys(reg, traj, ctrl)))
        init, refp, steps = np.array([0.9]).reshape(-1, 1), np.zeros((1, 1)), 200
        xs, us = mod.lqr_control(steps, refp, init, reg, out["K_lqr"])             # benchmark_lqr_hjb.py:74-97 (true system, lift per step)
        out.update(cl_init=init, cl_ref=refp, cl_steps=steps, cl_x=np.asarray(xs), cl_u=np.asarray(us))
        name = f"hjb_m100_g{gamma:g}".replace("0.001", "1e-3").replace("1e-06", "1e-6")
        np.savez_compressed(OUT / f"{name}.npz", **out)
        describe(name, out)


def make_duffing():
    mod = H.load_script("benchmark_lqr_classic", "ref")
    params = dict(Ts=0.01, name="duffing", n_states=2, n_inputs=1, radius_sampling=1.0, angle_sampling=2, input_lb=[-1], input_ub=[1])
    mod.dynamical_system = mod.DuffingOscillator(**params)
    mod.n_inputs, mod.n_states = 1, 2
    g1 = np.load(HERE / "g1" / "duffing_g1.npz")
    X, Y = g1["X"].T.copy(), g1["Y"].T.copy()                                       # the script's dataset (benchmark_lqr_classic.py:174-178)
    make = lambda md: md.KoopmanNystromRegressor(1, kernel=md.KernelWrapper([1, 1]), gamma=1e-6, m=20)   # LQR branch, :265
    reg, reg2 = fit_reference(mod, make, X, Y, 0)
    out = common(reg, reg2, X, Y, O.MATERN52, [1.0, 1.0], 1e-6, 1.0, mod)
    np.random.seed(0)
    traj, ctrl = mod.simulate_true_system(mod.dynamical_system, 2)
    out.update(dataset="tests/golden/g1/duffing_g1.npz (X, Y)", traj=traj, ctrl=ctrl, rmse_percent=mod.validate_dyn_sys(reg, traj, ctrl),
               rmse_percent_floor=abs(mod.validate_dyn_sys(reg2, traj, ctrl) - mod.validate_dyn_sys(reg, traj, ctrl)))
    init, refp, steps = np.array([-0.5, 0.0]).reshape(-1, 1), np.zeros((2, 1)), 300
    xs, ys, us = mod.lqr_control(steps, refp, init, reg, out["K_lqr"])              # benchmark_lqr_classic.py:67-89
    out.update(cl_init=init, cl_ref=refp, cl_steps=steps, cl_x=np.asarray(xs), cl_y=np.asarray(ys), cl_u=np.asarray(us))
    np.savez_compressed(OUT / "duffing_lqr_m20.npz", **out)
    describe("duffing_lqr_m20", out)


def make_cloth():
    mod = H.load_script("benchmark_lqr_cloth", "ref")
    mod.n_states, mod.n_inputs = 192, 6
    g2 = np.load(HERE / "g2" / "cloth_g2.npz")
    trajs = g2["traj_q"].astype(np.float64) / 10.0 ** g2["traj_k"].astype(np.float64)      # the 40 trajectories the script keeps (10..49)
    ctrls = g2["ctrl_q"].astype(np.float64) / 10.0 ** g2["ctrl_k"].astype(np.float64)
    all_trajs, all_controls = [t for t in trajs], [c for c in ctrls]
    X, Y = mod.create_data_matrices(all_trajs, all_controls, np.arange(0, 30))            # benchmark_lqr_cloth.py:218-220
    make = lambda md: md.KoopmanNystromRegressor(6, kernel=md.ThreeDimensionalKernel(10, 10, 10, 192), gamma=1e-7, m=100)
    reg, reg2 = fit_reference(mod, make, X, Y, 0)
    out = common(reg, reg2, X, Y, O.RBF, np.full(192, 10.0), 1e-7, 0.0075, mod)           # Q = 0.0075 C'C (:239)
    traj, ctrl = all_trajs[35], all_controls[35]
    out.update(dataset="tests/golden/g2/cloth_g2.npz (trajectories 0..29 -> create_data_matrices)", test_traj=35,
               rmse_cloth=mod.validate_dyn_sys(reg, traj, ctrl),
               rmse_cloth_floor=abs(mod.validate_dyn_sys(reg2, traj, ctrl) - mod.validate_dyn_sys(reg, traj, ctrl)))
    init = all_trajs[0][:, 0].reshape(-1, 1)
    reference = init + 0.05
    x_s, y_s, z_s, final_us = mod.lqr_control(60, reference, init, reg, out["K_lqr"])      # benchmark_lqr_cloth.py:69-104 (lifted model)
    x2 = mod.lqr_control(60, reference, init, reg2, H.gain_of(mod, reg2, 0.0075))[0]
    out.update(cl_init=init, cl_ref=reference, cl_steps=60, cl_x=x_s, cl_y=y_s, cl_z=z_s, cl_u=final_us, cl_floor=O.relerr(x2, x_s))
    np.savez_compressed(OUT / "cloth_lqr_m100.npz", **out)
    describe("cloth_lqr_m100", out)


def main():
    import torch  # noqa: F401  (imported before the shims: torch inspects sys.modules)
    OUT.mkdir(exist_ok=True)
    H.install_shims()
    make_hjb()
    make_duffing()
    make_cloth()


if __name__ == "__main__":
    main()
