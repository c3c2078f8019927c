"""CPU: the doubling solver of nys_koop_lqr_b200/dare.py (`control.dlqr` call sites benchmark_lqr_cloth.py:262,
_classic.py:288, _hjb.py:293,356), run through its torch statement of the ops (``TorchOps``) -- the SAME iteration code the GPU
runs through ``EngineOps`` (tests/test_gpu_dare.py) -- against

  * the gains the reference's own call sequence produced for the three scripts' LQR configurations (fixtures
    tests/golden/scripts/*.npz, key K_lqr: scipy ``solve_discrete_are`` + K = (B'PB+R)^-1 B'PA, the oracle SURVEY 8c names);
  * scipy on models fitted by the oracle, stable and unstable;
  * the equation itself (relative DARE residual) and the stability of the closed loop.

Tolerances: K relative Frobenius error <= 1e-8 against scipy on the script fixtures (the worst case, cloth m=100 with a closed-loop
spectral radius of 0.99994, sits at 2e-9; scipy's own DARE residual there is 2e-12, the doubling solver's 2e-14), <= 1e-10 on the
well-conditioned models; residual <= 1e-12.
"""
import pathlib

import numpy as np
import pytest
import scipy.linalg
import torch

from nys_koop_lqr_b200 import dare
from oracle import nk_oracle as O

SCRIPTS = sorted((pathlib.Path(__file__).parent / "golden" / "scripts").glob("*.npz"))


def _cpu_dlqr(A, B, Q, R, **kw):
    return dare.dlqr(A, B, Q, R, ops=dare.TorchOps(), device="cpu", **kw)


@pytest.mark.parametrize("path", SCRIPTS, ids=[p.stem for p in SCRIPTS])
def test_gain_matches_the_reference_call_sequence_on_the_script_models(path):
    fx = np.load(path, allow_pickle=True)
    A, B, C = fx["A"], fx["B"], fx["C"]
    Q = float(fx["qscale"]) * C.T @ C
    Q = (Q + Q.T) / 2
    R = np.eye(B.shape[1])
    K, P, E = _cpu_dlqr(A, B, Q, R, eigenvalues=True)
    assert O.relerr(K, fx["K_lqr"]) <= 1e-8
    info = dare.dlqr.last_info
    assert info["residual"] <= 1e-12 and info["iterations"] <= 40
    assert np.max(np.abs(E)) < 1.0                         # the stabilising solution
    assert O.relerr(P, P.T) <= 1e-15 and np.linalg.eigvalsh(P).min() >= -1e-9 * np.linalg.norm(P, 2)


@pytest.mark.parametrize("n,d,p,m,gamma,ls", [(1500, 6, 2, 120, 1e-4, 3.0), (2000, 192, 6, 160, 1e-5, 12.0)])
def test_gain_matches_scipy_on_oracle_fitted_models(n, d, p, m, gamma, ls):
    Xs, U, Y = O.synthetic(n, d, p, seed=1)
    np.random.seed(0)
    Z = O.draw_landmarks(Y, m)
    f = O.fit(np.hstack((Xs, U)), Y, p, O.RBF, np.full(d, ls), gamma, Z=Z)
    Q = f["C"].T @ f["C"]
    Q = (Q + Q.T) / 2
    R = np.diag(np.linspace(0.5, 2.0, p))                  # not the identity
    K0, P0 = O.dlqr(f["A"], f["B"], Q, R)
    K, P, _ = _cpu_dlqr(f["A"], f["B"], Q, R)
    assert O.relerr(K, K0) <= 1e-10 and O.relerr(P, P0) <= 1e-10
    assert dare.dlqr.last_info["residual"] <= 1e-13


def test_unstable_open_loop_and_single_input():
    rng = np.random.default_rng(4)
    m = 12                                                   # (at m = 40 with ONE input cond(P) is 1e6 and scipy itself stops at a 9e-11 residual)
    A = rng.standard_normal((m, m)) / np.sqrt(m) * 1.3       # spectral radius ~1.3
    B = rng.standard_normal((m, 1))
    Q = np.eye(m)
    R = np.array([[2.0]])
    assert np.max(np.abs(np.linalg.eigvals(A))) > 1.0
    K, P, E = _cpu_dlqr(A, B.reshape(-1), Q, 2.0, eigenvalues=True)   # B as a vector, R as a scalar: accepted like control.dlqr
    P0 = scipy.linalg.solve_discrete_are(A, B, Q, R)
    assert O.relerr(P, P0) <= 1e-10
    assert O.relerr(K, np.linalg.solve(R + B.T @ P0 @ B, B.T @ P0 @ A)) <= 1e-10
    assert np.max(np.abs(E)) < 1.0


def test_not_stabilisable_is_an_error_not_a_wrong_gain():
    A = np.diag([1.5, 0.5])
    B = np.array([[0.0], [1.0]])                            # the unstable mode cannot be reached
    with pytest.raises(dare.NkError):
        _cpu_dlqr(A, B, np.eye(2), np.eye(1))
    with pytest.raises(ValueError):
        dare.solve_dare(torch.eye(3, dtype=torch.float64), torch.ones(3, 1, dtype=torch.float64),
                        torch.eye(2, dtype=torch.float64), torch.eye(1, dtype=torch.float64), dare.TorchOps())


def test_engine_ops_only_override_the_products_and_the_spd_solve():
    """The GPU ops differ from the torch statement in exactly two methods; the iteration, the LU and the bookkeeping are shared,
    which is what makes the CPU tests above a check of the code the GPU runs."""
    own = {k for k, v in vars(dare.EngineOps).items() if callable(v) and not k.startswith("__")}
    assert own == {"mm", "spd_solve"}
    assert issubclass(dare.EngineOps, dare.TorchOps)
