"""Generates tests/golden/g1/duffing_g1.npz: the inputs and the reference's OWN published numbers for golden vector G1
(SURVEY.md section 4): `duffing/all_rmses_nystrom_double_dataset.csv`, open-loop forecast RMSE % over 200 seeds x 20 landmark
counts, written by benchmark_lqr_classic.py:254.  Stored: the Duffing dataset exactly as benchmark_lqr_classic.py:174-178
assembles it (69 900 samples), and for the first seeds the test trajectory / controls of `simulate_true_system`
(benchmark_lqr_classic.py:122-133, produced with the reference's own `dynamical_systems.DuffingOscillator`), the landmark
indices of the documented RNG protocol (seed; first of two `np.random.choice` draws), and the CSV's m=10,12,14 entries.

Run:  python tests/golden/make_golden_g1.py      (needs /root/reference; the fixture is committed)
"""
import pathlib
import sys

import numpy as np
import scipy.signal

REF = pathlib.Path("/root/reference")
HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(REF))
import dynamical_systems as refsys   # noqa: E402  the reference's simulator

N_SEEDS = 12


def main():
    ld = lambda f: np.loadtxt(REF / "duffing" / f, delimiter=",")
    xf, xu, yf, yu, uf = ld("duffing_x_forced.csv"), ld("duffing_x_unforced.csv"), ld("duffing_y_forced.csv"), ld("duffing_y_unforced.csv"), ld("duffing_u_forced.csv")
    X = np.vstack((np.hstack((xf, xu)), np.hstack((uf.reshape(1, -1), np.zeros((1, xu.shape[1])))))).T.copy()    # (n, 3) = [x | u]
    Y = np.hstack((yf, yu)).T.copy()                                                                              # (n, 2)
    csv = np.loadtxt(REF / "duffing" / "all_rmses_nystrom_double_dataset.csv")
    ms = np.around(np.logspace(1, 2.3, 20)).astype(int)       # benchmark_lqr_classic.py:179
    cols = [int(np.where(ms == m)[0][0]) for m in (10, 12, 14)]
    sysd = refsys.DuffingOscillator(Ts=0.01, name="duffing", n_states=2, n_inputs=1, radius_sampling=1.0, angle_sampling=2, input_lb=[-1], input_ub=[1])
    trajs, idxs = [], []
    us = 1.0 * scipy.signal.square(2 * np.pi * 10 / 3 * np.linspace(0, 2, 100))
    for seed in range(N_SEEDS):
        np.random.seed(seed)
        length = np.sqrt(np.random.uniform(0, 1.0)); angle = np.pi * np.random.uniform(0, 2)
        st = np.array([length * np.cos(angle), length * np.sin(angle)]).reshape(-1, 1)
        traj = st.copy()
        for u in us:
            st = sysd.update_SOM(st, u)
            traj = np.hstack((traj, st.reshape(-1, 1)))
        trajs.append(traj)
        # landmark protocol of the CSV: reseed, then for each m in order two draws, the first one used
        np.random.seed(seed)
        per_m = {}
        for m in ms:
            first = np.random.choice(np.arange(0, X.shape[0]), size=m, replace=False)
            np.random.choice(np.arange(0, X.shape[0]), size=m, replace=False)
            if m in (10, 12, 14):
                per_m[int(m)] = first
        idxs.append(per_m)
    out = dict(X=X, Y=Y, controls=us.reshape(1, -1), trajs=np.stack(trajs), ms=np.array([10, 12, 14]),
               want=csv[:N_SEEDS][:, cols], idx10=np.stack([i[10] for i in idxs]), idx12=np.stack([i[12] for i in idxs]),
               idx14=np.stack([i[14] for i in idxs]))
    (HERE / "g1").mkdir(exist_ok=True)
    np.savez_compressed(HERE / "g1" / "duffing_g1.npz", **out)
    print("wrote", HERE / "g1" / "duffing_g1.npz", {k: np.asarray(v).shape for k, v in out.items()})


if __name__ == "__main__":
    main()
