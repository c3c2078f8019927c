"""Generates tests/golden/g2/cloth_g2.npz: the inputs and the reference's OWN published numbers for golden vector G2
(SURVEY.md section 4): `8x8_cloth_swing_xyz/sim_results/nystrom/data/all_rmses_nystrom_cloth_swing_angle.csv`, open-loop
forecast RMSE of the cloth model over 20 rows (2 seeds x 10 test trajectories) x 20 landmark counts, written by the
`validate_sys_id` branch of benchmark_lqr_cloth.py:168-207 (RBF length scale 10 on all 192 coordinates, gamma = 1e-7; the
kernel parameters are the ones stored in the pickled regressors, golden G3).

Protocol reproduced here (and confirmed against the CSV with the oracle before the fixture is written):
  np.random.seed(seed); random.seed(seed); shuffle(arange(40)) -> 30 training / 10 test trajectories (:172-176);
  for each test trajectory, for each m in logspace(1, 2.6, 20, dtype=int): a fresh estimator is fitted; the published file
  was produced by a `fit` that drew TWO `np.random.choice(arange(n), m, replace=False)` per fit and used the first as both
  centre sets -- the RNG stream runs on across m and test trajectories.
Stored: the 40 non-validation trajectories / controls (the CSV files hold 5-6 significant decimal digits; they are stored
as exact integer mantissas q and decimal exponents k and decoded as q / 10^k with one correctly-rounded division, which reproduces np.loadtxt bit for bit --
asserted below), per seed the shuffled trajectory order, per (seed, test trajectory, m) the landmark indices, and the
CSV's entries for m = 10, 12, 14, 17.

Run:  python tests/golden/make_golden_g2.py      (needs /root/reference; the fixture is committed)
"""
import pathlib
import random
import sys

import numpy as np

REF = pathlib.Path("/root/reference/8x8_cloth_swing_xyz")
HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
from oracle import nk_oracle as O   # noqa: E402

N_INPUTS, N_TRAJS, N_VAL, N_TRAIN = 6, 50, 10, 30
MS_ALL = np.logspace(1.0, 2.6, num=20, dtype=int)       # benchmark_lqr_cloth.py:134
MS_KEPT = (10, 12, 14, 17)
SEEDS = (0, 1)


def encode(a):
    """exact decimal coding of %.5g-formatted data: per element an integer mantissa q and a decimal exponent k with
    q / 10^k == a bit for bit (q and 10^k are exact doubles, the division is correctly rounded -- like the text parse)."""
    a = np.asarray(a, dtype=np.float64)
    q = np.zeros(a.shape, dtype=np.int64)
    k = np.zeros(a.shape, dtype=np.int8)
    done = a == 0.0
    mag = np.floor(np.log10(np.abs(np.where(done, 1.0, a)))).astype(int)
    for extra in (4, 5, 6, 3):
        kk = np.clip(extra - mag, 0, 22)
        qq = np.rint(a * 10.0 ** kk)
        ok = ~done & (qq / 10.0 ** kk == a) & (np.abs(qq) < 2 ** 31)
        q[ok], k[ok] = qq[ok].astype(np.int64), kk[ok]
        done |= ok
    if not done.all():
        raise SystemExit("trajectory data is not short decimal")
    return q.astype(np.int32), k


def decode(q, k):
    return q.astype(np.float64) / 10.0 ** k.astype(np.float64)


def data_matrices(trajs, ctrls, indices):
    """benchmark_lqr_cloth.py:116-130 (create_data_matrices), rows = samples."""
    S = np.hstack([trajs[i][:, :-1] for i in indices])
    Nx = np.hstack([trajs[i][:, 1:] for i in indices])
    U = np.hstack([ctrls[i][:, :-1] for i in indices])
    return np.vstack((S, U)).T.copy(), Nx.T.copy()


def main():
    trajs, ctrls = [], []
    for i in range(N_VAL, N_TRAJS):                       # :149-156 (the first 10 are the hyper-parameter validation set)
        trajs.append(np.loadtxt(REF / f"state_samples_cloth_swing_{i}.csv", delimiter=",").T)
        ctrls.append(np.loadtxt(REF / f"input_samples_cloth_swing_{i}.csv", delimiter=",")[:, :N_INPUTS].T)
    T = np.stack(trajs)        # (40, 192, 102)
    U = np.stack(ctrls)        # (40, 6, 102)
    Tq, Tk = encode(T)
    Uq, Uk = encode(U)
    assert np.array_equal(decode(Tq, Tk), T) and np.array_equal(decode(Uq, Uk), U)
    csv = np.loadtxt(REF / "sim_results" / "nystrom" / "data" / "all_rmses_nystrom_cloth_swing_angle.csv")
    cols = [int(np.where(MS_ALL == m)[0][0]) for m in MS_KEPT]
    ls = np.full(192, 10.0)
    order, idx, want = [], {m: [] for m in MS_KEPT}, []
    worst = 0.0
    for seed in SEEDS:
        np.random.seed(seed)
        random.seed(seed)
        ti = np.arange(0, N_TRAJS - N_VAL)
        np.random.shuffle(ti)
        order.append(ti.copy())
        X, Y = data_matrices(trajs, ctrls, ti[:N_TRAIN])
        for i, te in enumerate(ti[N_TRAIN:]):
            for m in MS_ALL:
                first = np.random.choice(np.arange(0, X.shape[0]), size=m, replace=False)
                np.random.choice(np.arange(0, X.shape[0]), size=m, replace=False)
                if int(m) in MS_KEPT:
                    idx[int(m)].append(first)
                    Z = Y[first]
                    fit = O.fit(X, Y, N_INPUTS, O.RBF, ls, 1e-7, Z=Z)
                    z0 = O.lift(Z, trajs[te][:, :1], O.RBF, ls)[:, 0]
                    sim = O.rollout(fit["A"], fit["B"], fit["C"], z0, ctrls[te][:, :-1])
                    got = np.sqrt(np.mean((trajs[te] - sim) ** 2))           # benchmark_lqr_cloth.py:34
                    w = csv[seed * 10 + i, list(MS_ALL).index(m)]
                    worst = max(worst, abs(got - w) / w)
            want.append(csv[seed * 10 + i, cols])
    print(f"oracle vs published CSV over {len(want)} rows x {len(MS_KEPT)} columns: worst relative deviation {worst:.2e}")
    assert worst < 5e-7
    out = dict(traj_q=Tq, traj_k=Tk, ctrl_q=Uq, ctrl_k=Uk, order=np.stack(order), seeds=np.array(SEEDS), ms=np.array(MS_KEPT),
               want=np.array(want).reshape(len(SEEDS), 10, len(MS_KEPT)))
    for m in MS_KEPT:
        out[f"idx{m}"] = np.stack(idx[m]).reshape(len(SEEDS), 10, m).astype(np.int32)
    (HERE / "g2").mkdir(exist_ok=True)
    np.savez_compressed(HERE / "g2" / "cloth_g2.npz", **out)
    print("wrote", HERE / "g2" / "cloth_g2.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
