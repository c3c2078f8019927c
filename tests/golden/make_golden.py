"""Generates tests/golden/*.npz by running the UNMODIFIED reference (imported from /root/reference) in this container.

Each fixture holds the inputs (so the GPU box needs neither the reference nor its data directories), the landmarks
that were injected, and the reference's outputs: A, B, C, weights, lift() on query points, predict() on query rows,
the open-loop forecast of validate_dyn_sys (trajectory + both RMSE definitions), the DARE gain, plus the
reference's own SELF-FLOOR for that configuration (same landmarks, sample order permuted) -- the yardstick for the
parity tolerance where cond(inner_term) makes 1e-9 unattainable.

Run:  python tests/golden/make_golden.py      (needs /root/reference; the fixtures are committed)
"""
import pathlib
import sys

import numpy as np
import scipy.linalg

REF = pathlib.Path("/root/reference")
HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(REF))                 # the reference's `regressors` must win over this repo's drop-in of the same name
import regressors as ref                     # noqa: E402  the reference itself
import dynamical_systems as refsys           # noqa: E402
from oracle import nk_oracle as O            # noqa: E402
assert pathlib.Path(ref.__file__).parent == REF, "golden vectors must come from the unmodified reference"


def reference_rollout(reg, true_traj, controls):
    """benchmark_lqr_cloth.py:18-36 loop, driven through the reference estimator's own lift/A/B/C."""
    z = reg.lift(true_traj[:, 0].reshape(-1, 1))
    sim = reg.C @ z
    for i in range(true_traj.shape[1] - 1):
        z = reg.A @ z + reg.B @ controls[:, i].reshape(-1, 1)
        sim = np.hstack((sim, reg.C @ z))
    return sim


def run_case(name, X, Y, n_inputs, kernel, kind, ls, gamma, m, seed, Xq, traj=None, ctrl=None, q_scale=1.0, distinct_input=False):
    """X (n, d+p), Y (n, d) rows.  distinct_input: inject input landmarks drawn from the CURRENT states (regressors.py:133-134
    only aliases the output landmarks when nystrom_centers_input is unset)."""
    n = X.shape[0]
    np.random.seed(seed)
    idx = np.random.choice(np.arange(0, n), size=m, replace=False)       # regressors.py:130
    Zc = Y.T[:, idx]
    Zin = Zc
    if distinct_input:
        Zin = np.ascontiguousarray(X[np.random.choice(np.arange(0, n), size=m, replace=False), :Y.shape[1]].T)
    reg = ref.KoopmanNystromRegressor(n_inputs, kernel=kernel, gamma=gamma, m=m)
    reg.nystrom_centers_output = Zc
    reg.nystrom_centers_input = Zin
    reg.fit(X, Y)
    # self-floor: same landmarks, permuted samples
    perm = np.random.default_rng(1).permutation(n)
    reg2 = ref.KoopmanNystromRegressor(n_inputs, kernel=kernel, gamma=gamma, m=m)
    reg2.nystrom_centers_output = Zc
    reg2.nystrom_centers_input = Zin
    reg2.fit(X[perm], Y[perm])
    floor = dict(A=O.relerr(reg2.A, reg.A), B=O.relerr(reg2.B, reg.B), C=O.relerr(reg2.C, reg.C))
    d = Y.shape[1]
    out = dict(X=X, Y=Y, n_inputs=n_inputs, kind=kind, ls=np.asarray(ls, dtype=float), gamma=gamma, m=m, seed=seed, Z=Zc,
               A=reg.A, B=reg.B, C=reg.C, W=reg.weights, floor_A=floor["A"], floor_B=floor["B"], floor_C=floor["C"],
               Xq=Xq, lift_q=reg.lift(Xq[:, :d].T), predict_q=reg.predict(Xq))
    if distinct_input:
        out["Z_in"] = Zin
    # conditioning of the first system (regressors.py:151)
    G = O.grams(X[:, :d], Y, X[:, d:], Zin.T, kind, ls)
    Kzz = O.kernel_matrix(Zin.T, Zin.T, kind, ls)
    inner = np.block([[G["Gxx"] + gamma * n * (Kzz + 1e-6 * np.eye(m)), G["Gxu"]], [G["Gxu"].T, G["Guu"] + gamma * n * np.eye(n_inputs)]])
    out["cond_inner"] = np.linalg.cond(inner)
    if traj is not None:
        sim = reference_rollout(reg, traj, ctrl)
        out.update(traj=traj, ctrl=ctrl, sim=sim, rmse_cloth=O.rmse_cloth(traj, sim), rmse_percent=O.rmse_percent(traj, sim))
    # DARE gain as the scripts form it (Q = q_scale * C'C, R = I), scipy as the control.dlqr stand-in (golden G4 pins it)
    Q = q_scale * reg.C.T @ reg.C
    Q = (Q + Q.T) / 2
    try:
        K, _ = O.dlqr(reg.A, reg.B, Q, np.eye(n_inputs))
        out["K_lqr"] = K
        Q2 = q_scale * reg2.C.T @ reg2.C
        K2, _ = O.dlqr(reg2.A, reg2.B, (Q2 + Q2.T) / 2, np.eye(n_inputs))
        out["floor_K"] = O.relerr(K2, K)          # the reference's own reproducibility of the gain (permuted samples)
        print(f"  [{name}] gain self-floor {out['floor_K']:.1e}")
    except Exception as ex:  # noqa: BLE001
        print(f"  [{name}] DARE failed: {ex}")
    np.savez_compressed(HERE / f"{name}.npz", **out)
    print(f"{name}: n={n} d={d} p={n_inputs} m={m} cond(inner)={out['cond_inner']:.2e} self-floor A={floor['A']:.1e} C={floor['C']:.1e}"
          + (f" rmse%={out['rmse_percent']:.6g}" if traj is not None else ""))


def main():
    rng = np.random.default_rng(0)
    only = sys.argv[1:]
    # ---- synthetic (SURVEY 8d family), small ----
    Xs, U, Y = O.synthetic(1500, d=12, p=2, seed=3)
    X = np.hstack((Xs, U))
    run_case("synthetic_rbf", X, Y, 2, ref.ThreeDimensionalKernel(3.0, 4.0, 5.0, 12), O.RBF, np.resize([3.0, 4.0, 5.0], 12), 1e-4, 64, 0,
             X[:50], traj=None)
    run_case("synthetic_rbf_g1e-2", X, Y, 2, ref.ThreeDimensionalKernel(3.0, 4.0, 5.0, 12), O.RBF, np.resize([3.0, 4.0, 5.0], 12), 1e-2, 64, 0,
             X[:50], traj=None)
    # distinct input / output landmark sets (a caller may inject nystrom_centers_input, regressors.py:133-134)
    run_case("synthetic_rbf_distinct_centers", X, Y, 2, ref.ThreeDimensionalKernel(3.0, 4.0, 5.0, 12), O.RBF, np.resize([3.0, 4.0, 5.0], 12), 1e-2, 64,
             0, X[:50], traj=None, distinct_input=True)
    if only == ["distinct"]:
        return
    # ---- Duffing (benchmark_lqr_classic.py:174-178 data), every 20th sample; Matern-5/2 l=[1,1], gamma=1e-6 (G6/G1) ----
    dx = np.hstack((np.loadtxt(REF / "duffing/duffing_x_forced.csv", delimiter=","), np.loadtxt(REF / "duffing/duffing_x_unforced.csv", delimiter=",")))
    du = np.hstack((np.loadtxt(REF / "duffing/duffing_u_forced.csv", delimiter=",").reshape(1, -1), np.zeros((1, np.loadtxt(REF / "duffing/duffing_x_unforced.csv", delimiter=",").shape[1]))))
    dy = np.hstack((np.loadtxt(REF / "duffing/duffing_y_forced.csv", delimiter=","), np.loadtxt(REF / "duffing/duffing_y_unforced.csv", delimiter=",")))
    Xd = np.vstack((dx, du)).T[::20].copy()
    Yd = dy.T[::20].copy()
    # test trajectory exactly as benchmark_lqr_classic.py:122-133 with seed 0
    sysd = refsys.DuffingOscillator(Ts=0.01, name="duffing", n_states=2, n_inputs=1, radius_sampling=1.0, angle_sampling=2, input_lb=[-1], input_ub=[1])
    np.random.seed(0)
    length = np.sqrt(np.random.uniform(0, 1.0)); angle = np.pi * np.random.uniform(0, 2)
    st = np.array([length * np.cos(angle), length * np.sin(angle)]).reshape(-1, 1)
    times = np.linspace(0, 2, 100)
    us = 1.0 * scipy.signal.square(2 * np.pi * 10 / 3 * times)
    traj = st.copy()
    for u in us:
        st = sysd.update_SOM(st, u)
        traj = np.hstack((traj, st.reshape(-1, 1)))
    for m in (10, 20):
        run_case(f"duffing_m{m}", Xd, Yd, 1, ref.KernelWrapper([1, 1]), O.MATERN52, [1.0, 1.0], 1e-6, m, 0, Xd[:40], traj, us.reshape(1, -1))
    # ---- HJB (benchmark_lqr_hjb.py:110-126 generator, 20 x 199 samples), Matern l=1 ----
    sysh = refsys.HJB(Ts=0.01, name="hjb", n_states=1, n_inputs=1, state_lb=-1.0, state_ub=1.0, input_lb=[-1], input_ub=[1])
    np.random.seed(0)
    rows_x, rows_y = [], []
    for _ in range(20):
        x = np.random.uniform(-1.0, 1.0)
        for _ in range(199):
            u = np.random.uniform([-1], [1]).reshape(1, 1)
            xn = sysh.update_SOM(np.array([x]).reshape(-1, 1), u).reshape(-1, 1)
            rows_x.append([float(x), float(u[0, 0])]); rows_y.append([float(xn[0, 0])])
            x = float(xn[0, 0])
    Xh, Yh = np.array(rows_x), np.array(rows_y)
    np.random.seed(1)
    s0 = np.random.uniform(-1.0, 1.0)
    us_h = 2 * np.linspace(0, 2, 100)
    sth = np.array(s0).reshape(-1, 1); trajh = sth.copy()
    for u in us_h:
        sth = sysh.update_SOM(sth, u); trajh = np.hstack((trajh, sth.reshape(-1, 1)))
    run_case("hjb_m30_g1e-3", Xh, Yh, 1, ref.KernelWrapper([1.0]), O.MATERN52, [1.0], 1e-3, 30, 0, Xh[:40], trajh, us_h.reshape(1, -1))
    # ---- cloth (benchmark_lqr_cloth.py:149-156 loaders): 3 training trajectories + 1 test, RBF l=10, gamma=1e-7 (G3) ----
    p = REF / "8x8_cloth_swing_xyz"
    trajs = [np.loadtxt(p / f"state_samples_cloth_swing_{i}.csv", delimiter=",").T for i in (10, 11, 12, 13)]
    ctrls = [np.loadtxt(p / f"input_samples_cloth_swing_{i}.csv", delimiter=",")[:, :6].T for i in (10, 11, 12, 13)]
    Xc = np.hstack([np.vstack((t[:, :-1], c[:, :-1])) for t, c in zip(trajs[:3], ctrls[:3])]).T
    Yc = np.hstack([t[:, 1:] for t in trajs[:3]]).T
    run_case("cloth_m20", Xc, Yc, 6, ref.ThreeDimensionalKernel(10, 10, 10, 192), O.RBF, np.full(192, 10.0), 1e-7, 20, 0, Xc[:30],
             trajs[3], ctrls[3], q_scale=0.005)


if __name__ == "__main__":
    main()
