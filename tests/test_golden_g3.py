"""Golden vectors G3 / G4 and the pickle interchange of SURVEY 8f row 3: an estimator fitted and
PICKLED BY THE REFERENCE (benchmark_lqr_cloth.py:266-267; committed copy tests/golden/g3/regressor_seed_0.npy) is loaded
with the drop-in `regressors` module, exactly as benchmark_lqr_cloth.py / closed_loop_lqr_control.m:155-171 would, and must
reproduce what the unmodified reference computes from the same object (tests/golden/make_golden_g3.py): lift
(regressors.py:171-178), predict (:48-55), the open-loop simulation of validate_dyn_sys (benchmark_lqr_cloth.py:18-36),
and the exported LQR gain K_lqr_seed_0.csv (G4: Q = 0.005 C'C, R = I, rows permuted as benchmark_lqr_cloth.py:263).
"""
import pathlib
import pickle

import numpy as np
import pytest

from oracle import nk_oracle as O

G3 = pathlib.Path(__file__).parent / "golden" / "g3"


def load_pickle():
    import regressors as R     # the drop-in module (repo root): the pickle's `regressors.KoopmanNystromRegressor` resolves here
    assert "nys_koop_lqr_b200" in R.KoopmanNystromRegressor.__module__
    with open(G3 / "regressor_seed_0.npy", "rb") as f:
        reg = pickle.load(f)
    assert isinstance(reg, R.KoopmanNystromRegressor) and isinstance(reg.kernel, R.ThreeDimensionalKernel)
    return reg


def test_reference_pickle_loads_into_the_dropin():
    fx = np.load(G3 / "cloth_g3.npz")
    reg = load_pickle()
    assert reg.m == 100 and reg.n_inputs == 6 and reg.gamma == 1e-7 and reg.jitter == 1e-6
    for name in ("A", "B", "C"):
        assert isinstance(getattr(reg, name), np.ndarray) and np.array_equal(getattr(reg, name), fx[name])
    assert reg.nystrom_centers_output.shape == (192, 100)
    from nys_koop_lqr_b200.regressors import kernel_spec
    kind, ls = kernel_spec(reg.kernel, 192)
    assert kind == 0 and np.array_equal(ls, np.full(192, 10.0))
    # and it pickles back (the device cache is never part of the state)
    again = pickle.loads(pickle.dumps(reg))
    assert np.array_equal(again.A, reg.A) and "_dev" not in again.__dict__


def test_oracle_matches_reference_on_the_pickled_model():
    """G4 gain (host DARE, scipy standing in for control.dlqr) and the oracle's lift / rollout against the reference's."""
    fx = np.load(G3 / "cloth_g3.npz")
    A, B, C = fx["A"], fx["B"], fx["C"]
    Q = 0.005 * C.T @ C
    K, _ = O.dlqr(A, B, (Q + Q.T) / 2, np.eye(6))
    assert O.relerr(K[[0, 3, 1, 4, 2, 5], :], fx["K_lqr"]) <= 1e-8
    reg = load_pickle()
    Z = np.ascontiguousarray(reg.nystrom_centers_output.T)
    ls = np.full(192, 10.0)
    assert O.relerr(O.lift(Z, fx["states"], O.RBF, ls), fx["lifted"]) <= 1e-9
    g2 = np.load(G3.parent / "g2" / "cloth_g2.npz")
    tr = int(fx["traj_index"])
    ctrl = g2["ctrl_q"][tr].astype(np.float64) / 10.0 ** g2["ctrl_k"][tr].astype(np.float64)
    x0 = g2["traj_q"][tr][:, :1].astype(np.float64) / 10.0 ** g2["traj_k"][tr][:, :1].astype(np.float64)
    z0 = O.lift(Z, x0, O.RBF, ls)[:, 0]
    assert O.relerr(O.rollout(A, B, C, z0, ctrl[:, :-1]), fx["sim"]) <= 1e-9


@pytest.mark.gpu
def test_gpu_lift_predict_forecast_of_the_reference_pickle(engine):
    fx = np.load(G3 / "cloth_g3.npz")
    reg = load_pickle()
    lifted = reg.lift(fx["states"])                                  # GPU: S^-1 k(Z, x) with S from nk_sym_sqrt
    assert lifted.shape == (100, 24) and O.relerr(lifted, fx["lifted"]) <= 1e-9
    pred = reg.predict(fx["X_aug"])
    assert pred.shape == (24, 192) and O.relerr(pred, fx["pred"]) <= 1e-9
    g2 = np.load(G3.parent / "g2" / "cloth_g2.npz")
    tr = int(fx["traj_index"])
    traj = g2["traj_q"][tr].astype(np.float64) / 10.0 ** g2["traj_k"][tr].astype(np.float64)
    ctrl = g2["ctrl_q"][tr].astype(np.float64) / 10.0 ** g2["ctrl_k"][tr].astype(np.float64)
    sim, rmse, _ = reg.forecast(traj[:, 0], ctrl[:, :-1], true_trajectories=traj)
    assert O.relerr(sim, fx["sim"]) <= 1e-9
    assert abs(rmse - float(fx["rmse"])) <= 1e-9 * float(fx["rmse"])
    # closed loop of the exported gain on the lifted model (benchmark_lqr_cloth.py:69-104), against the oracle's loop
    K = fx["K_lqr"][[0, 2, 4, 1, 3, 5], :]                           # undo the export permutation [0,3,1,4,2,5]
    states, controls = reg.closed_loop(K, traj[:, 0], traj[:, 50], 20)
    want_x, want_u = O.closed_loop(fx["A"], fx["B"], fx["C"], K, fx["lifted"][:, 0], O.lift(
        np.ascontiguousarray(reg.nystrom_centers_output.T), traj[:, 50:51], O.RBF, np.full(192, 10.0))[:, 0], 20)
    assert O.relerr(states, want_x) <= 1e-8 and O.relerr(controls, want_u) <= 1e-8
