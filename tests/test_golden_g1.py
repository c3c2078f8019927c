"""Golden vector G1: the reference's own published result file
`duffing/all_rmses_nystrom_double_dataset.csv` (open-loop forecast RMSE %, benchmark_lqr_classic.py:179-254), reproduced
from the committed fixture tests/golden/g1/duffing_g1.npz (dataset, test trajectories from the reference's simulator,
landmark indices of the documented RNG protocol, CSV entries; tests/golden/make_golden_g1.py).

Whole path: fit (n=69 900, Matern-5/2, gamma=1e-6) -> lift -> 100-step rollout -> RMSE %.  North-star bar: identical to 6
significant digits (m=10); m=12 / m=14 are held to 5e-6 (the reference itself only reproduces those columns to ~1e-6
across machines).  CPU: the oracle.  GPU: the drop-in estimator through the C ABI.
"""
import pathlib

import numpy as np
import pytest

from oracle import nk_oracle as O

FX = pathlib.Path(__file__).parent / "golden" / "g1" / "duffing_g1.npz"
TOL = {10: 1e-6, 12: 5e-6, 14: 5e-6}


@pytest.mark.parametrize("m", [10, 14])
def test_oracle_reproduces_published_rmse_csv(m):
    fx = np.load(FX)
    X, Y = fx["X"], fx["Y"]
    col = list(fx["ms"]).index(m)
    for seed in range(4):
        Z = Y[fx[f"idx{m}"][seed]]
        fit = O.fit(X, Y, 1, O.MATERN52, [1.0, 1.0], 1e-6, Z=Z)
        traj = fx["trajs"][seed]
        z0 = O.lift(Z, traj[:, :1], O.MATERN52, [1.0, 1.0])[:, 0]
        got = O.rmse_percent(traj, O.rollout(fit["A"], fit["B"], fit["C"], z0, fx["controls"]))
        want = fx["want"][seed, col]
        assert abs(got - want) <= TOL[m] * want, (seed, got, want)


@pytest.mark.gpu
@pytest.mark.parametrize("m", [10, 12, 14])
def test_estimator_reproduces_published_rmse_csv(engine, m):
    import regressors as R
    fx = np.load(FX)
    X, Y = fx["X"], fx["Y"]
    col = list(fx["ms"]).index(m)
    worst = 0.0
    for seed in range(fx["trajs"].shape[0]):
        reg = R.KoopmanNystromRegressor(1, kernel=R.KernelWrapper([1, 1]), gamma=1e-6, m=m)
        reg.nystrom_centers_output = np.ascontiguousarray(Y[fx[f"idx{m}"][seed]].T)
        reg.fit(X, Y)
        traj = fx["trajs"][seed]
        _, _, pct = reg.forecast(traj[:, 0], fx["controls"], true_trajectories=traj)
        want = fx["want"][seed, col]
        worst = max(worst, abs(pct - want) / want)
        assert abs(pct - want) <= TOL[m] * want, (m, seed, pct, want)
    print(f"G1 m={m}: worst relative deviation from the published CSV over {fx['trajs'].shape[0]} seeds: {worst:.2e}")
