"""GPU: `control.dlqr` on the device (nys_koop_lqr_b200/dare.py through ``EngineOps``: m x m products on nk_gemm) against the gains
the reference's call sequence produced for the scripts' LQR configurations (tests/golden/scripts/*.npz, key K_lqr), against scipy
(``O.dlqr``) on a model fitted by the drop-in estimator, and against the equation itself at a size scipy would need minutes for.
Same tolerances as tests/test_dare_cpu.py, which runs the same iteration code through the torch statement of the ops.
"""
import pathlib

import numpy as np
import pytest

from oracle import nk_oracle as O

pytestmark = pytest.mark.gpu

SCRIPTS = sorted((pathlib.Path(__file__).parent / "golden" / "scripts").glob("*.npz"))


@pytest.mark.parametrize("path", SCRIPTS, ids=[p.stem for p in SCRIPTS])
def test_device_gain_matches_the_reference_call_sequence(engine, path):
    from nys_koop_lqr_b200 import dare
    fx = np.load(path, allow_pickle=True)
    A, B, C = fx["A"], fx["B"], fx["C"]
    Q = float(fx["qscale"]) * C.T @ C
    Q = (Q + Q.T) / 2
    before = engine.launch_count()
    K, P, E = dare.dlqr(A, B, Q, np.eye(B.shape[1]), eigenvalues=True)
    assert O.relerr(K, fx["K_lqr"]) <= 1e-8
    assert dare.dlqr.last_info["residual"] <= 1e-12
    assert np.max(np.abs(E)) < 1.0
    if A.shape[0] >= dare.EngineOps.min_gemm_dim:
        assert engine.launch_count() > before              # the m-sized products ran on this library's GEMM


def test_estimator_lqr_gain_matches_scipy_on_its_own_model(engine):
    import regressors as R
    n, d, p, m = 3000, 6, 2, 96
    Xs, U, Y = O.synthetic(n, d, p, seed=5)
    reg = R.KoopmanNystromRegressor(p, kernel=R.ThreeDimensionalKernel(3.0, 3.0, 3.0, d), gamma=1e-4, m=m)
    np.random.seed(5)
    reg.fit(np.hstack((Xs, U)), Y)
    K = reg.lqr_gain(q_scale=0.0075)                        # the cloth script's Q = 0.0075 C'C, R = I
    Q = 0.0075 * reg.C.T @ reg.C
    K0, P0 = O.dlqr(reg.A, reg.B, (Q + Q.T) / 2, np.eye(p))
    assert K.shape == (p, m) and O.relerr(K, K0) <= 1e-9
    assert reg.lqr_info_["residual"] <= 1e-12
    Qx, Rx = np.eye(m) * 0.3, np.diag([0.5, 2.0])          # explicit weights
    K2, P2, info = reg.lqr_gain(Q=Qx, R=Rx, return_all=True)
    K20, P20 = O.dlqr(reg.A, reg.B, Qx, Rx)
    assert O.relerr(K2, K20) <= 1e-9 and O.relerr(P2, P20) <= 1e-9
    # the gain feeds the lifted closed loop (benchmark_lqr_cloth.py:80-84) like a control.dlqr gain does
    xs, us = reg.closed_loop(K, Xs[0], Xs[1], 5)
    assert xs.shape == (d, 5) and np.isfinite(xs).all() and np.isfinite(us).all()


def test_large_model_solves_the_equation_and_agrees_with_the_torch_statement(engine):
    import torch
    from nys_koop_lqr_b200 import dare
    rng = np.random.default_rng(11)
    m, p, d = 768, 6, 192
    A = rng.standard_normal((m, m)) / np.sqrt(m) * 1.02     # spectral radius just above one, like the fitted models
    B = rng.standard_normal((m, p))
    C = rng.standard_normal((d, m)) / np.sqrt(m)
    Q = C.T @ C
    Q = (Q + Q.T) / 2
    R = np.eye(p)
    K, P, _ = dare.dlqr(A, B, Q, R)
    info = dict(dare.dlqr.last_info)
    assert info["residual"] <= 5e-11 and info["iterations"] <= 40      # 1.8e-12 through the torch statement on the CPU (a Ginibre A: P is large)
    Kc, Pc, _ = dare.dlqr(A, B, Q, R, ops=dare.TorchOps(), device="cpu")
    assert O.relerr(K, Kc) <= 1e-9 and O.relerr(P, Pc) <= 1e-9
    assert np.max(np.abs(np.linalg.eigvals(A - B @ K))) < 1.0
