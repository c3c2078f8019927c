"""Generates tests/golden/g3/: golden vectors G3 / G4 -- a fitted estimator PICKLED BY THE REFERENCE
(`8x8_cloth_swing_xyz/sim_results/nystrom/data/regressor_seed_0.npy`, written by benchmark_lqr_cloth.py:266-267; m=100,
RBF l=10, gamma=1e-7) and the LQR gain the reference exported for it (`K_lqr_seed_0.csv`, benchmark_lqr_cloth.py:262-265).

The pickle is a data artefact of the reference (numpy arrays + kernel parameters, 557 KB) and is copied verbatim: it is
what the MATLAB closed loop loads (cloth_simulator/closed_loop_lqr_control.m:155-171), and the drop-in must be able to load
it too.  `cloth_g3.npz` holds what the UNMODIFIED reference computes from that object in this container:
  lift (regressors.py:171-178) of 24 cloth states, predict (regressors.py:48-55) of 24 augmented rows, the open-loop
  simulation of validate_dyn_sys (benchmark_lqr_cloth.py:18-36) on one trajectory, plus the gain file.
Inputs are taken from the G2 fixture (tests/golden/g2/cloth_g2.npz), so nothing else needs to travel.

Run:  python tests/golden/make_golden_g3.py      (needs /root/reference; the fixture is committed)
"""
import pathlib
import pickle
import shutil
import sys

import numpy as np

REF = pathlib.Path("/root/reference")
DATA = REF / "8x8_cloth_swing_xyz" / "sim_results" / "nystrom" / "data"
HERE = pathlib.Path(__file__).resolve().parent


def main():
    (HERE / "g3").mkdir(exist_ok=True)
    shutil.copyfile(DATA / "regressor_seed_0.npy", HERE / "g3" / "regressor_seed_0.npy")
    sys.path.insert(0, str(REF))
    import regressors as ref_regressors      # noqa: F401  the reference module: the pickle resolves its classes here
    with open(DATA / "regressor_seed_0.npy", "rb") as f:
        reg = pickle.load(f)
    assert type(reg).__module__ == "regressors" and ref_regressors.__file__.startswith(str(REF))
    g2 = np.load(HERE / "g2" / "cloth_g2.npz")
    trajs = g2["traj_q"].astype(np.float64) / 10.0 ** g2["traj_k"].astype(np.float64)
    ctrls = g2["ctrl_q"].astype(np.float64) / 10.0 ** g2["ctrl_k"].astype(np.float64)
    tr = 5
    states = np.ascontiguousarray(trajs[tr][:, ::4][:, :24])                      # (192, 24) column samples
    lifted = reg.lift(states)                                                      # reference lift: (100, 24)
    X_aug = np.vstack((states, ctrls[tr][:, ::4][:, :24])).T.copy()                # (24, 198) rows [x | u]
    pred = reg.predict(X_aug)                                                      # reference predict: (24, 192)
    # validate_dyn_sys (benchmark_lqr_cloth.py:18-36) on trajectory `tr`, with the reference's own lift and loop
    traj, controls = trajs[tr], ctrls[tr]
    z = reg.lift(traj[:, 0].reshape(-1, 1))
    sim = reg.C @ z
    for i in range(traj.shape[1] - 1):
        z = reg.A @ z + reg.B @ controls[:, i].reshape(-1, 1)
        sim = np.hstack((sim, reg.C @ z))
    rmse = np.sqrt(np.mean((traj - sim) ** 2))
    out = dict(traj_index=np.array(tr), states=states, lifted=lifted, X_aug=X_aug, pred=pred, sim=sim, rmse=np.array(rmse),
               K_lqr=np.loadtxt(DATA / "K_lqr_seed_0.csv"), A=reg.A, B=reg.B, C=reg.C)
    np.savez_compressed(HERE / "g3" / "cloth_g3.npz", **out)
    print("wrote", HERE / "g3", {k: np.asarray(v).shape for k, v in out.items()}, "rmse", rmse)


if __name__ == "__main__":
    main()
