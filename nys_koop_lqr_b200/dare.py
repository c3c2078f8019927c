"""Discrete-time algebraic Riccati equation and LQR gain on the device: ``control.dlqr(A, B, Q, R)`` of the reference's
scripts (benchmark_lqr_cloth.py:262, benchmark_lqr_classic.py:288, benchmark_lqr_hjb.py:293,356) for lifted models whose
dimension m makes the host solver (python-control -> scipy ``solve_discrete_are``: a QZ decomposition of a 2m x 2m pencil,
14 s at m = 500, hours at m = 4096) by far the slowest step after the fit.  SURVEY.md section 8(f), row 4.

Method: the structure-preserving doubling algorithm for

    P = A' P A - A' P B (R + B' P B)^-1 B' P A + Q,        K = (R + B' P B)^-1 B' P A

with G_0 = B R^-1 B', H_0 = Q, A_0 = A and, per step (W = I + G_k H_k, nonsingular because G_k, H_k are positive semidefinite),

    A_{k+1} = A_k W^-1 A_k,    G_{k+1} = G_k + A_k W^-1 G_k A_k',    H_{k+1} = H_k + A_k' H_k W^-1 A_k,

H_k -> P quadratically (k steps cover 2^k steps of the Riccati recursion; 15-27 steps on the scripts' models, whose open-loop
spectral radius is 1.0000-1.0015).  Cost per step: one LU factorisation of W with 2m right-hand sides and five m x m x m
products.  The products -- 10/13 of the flops -- run on this repo's FP64 tensor-core GEMM (``nk_gemm``); the LU solve is
``torch.linalg`` (cuSOLVER getrf/getrs): the one library call, the Nystrom path has no general (non-symmetric) solver to
reuse.  The p x p systems (R, R + B'PB; p = 1..6) are negligible.

The iteration is written once against a small ``ops`` interface so that the same code is checked on the CPU
(``TorchOps``, tests/test_dare_cpu.py, against scipy on the reference-generated script fixtures) and runs on the GPU
(``EngineOps``).  This module is an ADDITIVE API: the scripts keep working with ``control.dlqr`` on the host.
"""
from __future__ import annotations

import torch

from ._lib import NkError


class TorchOps:
    """Dense products and solves through torch on whatever device the operands live on: the statement of the ops that the CPU test
    suite runs.  Never chosen implicitly -- every function below takes ``ops`` explicitly, and ``dlqr`` / ``reg.lqr_gain`` build
    ``EngineOps`` (which raises without the library and a GPU)."""

    def mm(self, A, B, ta=False, tb=False):
        return (A.T if ta else A) @ (B.T if tb else B)

    def lu(self, W):
        return torch.linalg.lu_factor(W)

    def lu_solve(self, fact, Y):
        return torch.linalg.lu_solve(fact[0], fact[1], Y)

    def spd_solve(self, M, Y):
        return torch.cholesky_solve(Y, torch.linalg.cholesky(M))


class EngineOps(TorchOps):
    """The same ops on a B200: m-sized products through ``nk_gemm`` (DMMA), SPD solves through ``nk_potrf`` +
    ``nk_trsm_lower``; the LU stays ``torch.linalg`` (cuSOLVER).  Anything with a dimension below ``min_gemm_dim`` (the p-sized
    algebra of the gain, p = 1..6 in the scripts: launch-bound, no flops) uses the torch statement."""

    min_gemm_dim = 32

    def __init__(self, engine):
        self.eng = engine

    def mm(self, A, B, ta=False, tb=False):
        M, K = (A.shape[1], A.shape[0]) if ta else A.shape
        N = B.shape[0] if tb else B.shape[1]
        if min(M, N, K) < self.min_gemm_dim:
            return super().mm(A, B, ta, tb)
        return self.eng.gemm(A.contiguous(), B.contiguous(), transa=ta, transb=tb)

    def spd_solve(self, M, Y):
        if M.shape[0] < self.min_gemm_dim:                  # p x p with p = 1..6 in every script: one tiny library call
            return super().spd_solve(M, Y)
        L = self.eng.potrf(M.contiguous().clone())          # raises NkError (NK_E_NOT_SPD) if R + B'PB is not positive definite
        X = Y.contiguous().clone()
        self.eng.trsm_lower(L, X, trans=False)
        self.eng.trsm_lower(L, X, trans=True)
        return X


def _sym(M):
    return (M + M.T) * 0.5


def solve_dare(A, B, Q, R, ops, tol=1e-13, max_iter=64):
    """Stabilising solution P of the DARE by doubling.  A (m,m), B (m,p), Q (m,m) symmetric PSD, R (p,p) symmetric PD: float64
    torch tensors on one device.  Returns (P, info) with info = dict(iterations, delta (last relative change of H),
    residual (relative DARE residual of the returned P)).  Raises NkError if the iteration has not settled after ``max_iter``
    doublings or produced non-finite values ((A, B) not stabilisable / (A, Q) not detectable)."""
    m, p = B.shape
    if A.shape != (m, m) or Q.shape != (m, m) or R.shape != (p, p):
        raise ValueError("solve_dare: A (m,m), B (m,p), Q (m,m), R (p,p) expected")
    eye = torch.eye(m, dtype=A.dtype, device=A.device)
    RinvBt = ops.spd_solve(_sym(R), B.T.contiguous())                # (p, m)
    G = _sym(ops.mm(B, RinvBt))
    H = _sym(Q)
    Ak = A.contiguous().clone()
    delta, it = float("inf"), 0
    for it in range(1, int(max_iter) + 1):
        fact = ops.lu(eye + ops.mm(G, H))
        T1 = ops.lu_solve(fact, Ak)                                  # W^-1 A_k
        T2 = ops.lu_solve(fact, G)                                   # W^-1 G_k (symmetric)
        HT1 = ops.mm(H, T1)
        A_next = ops.mm(Ak, T1)
        G_next = _sym(G + ops.mm(ops.mm(Ak, T2), Ak, tb=True))
        H_next = _sym(H + ops.mm(Ak, HT1, ta=True))
        num, den = float(torch.linalg.norm(H_next - H)), float(torch.linalg.norm(H_next))
        if not (num == num and den == den and den != float("inf")):
            raise NkError("solve_dare: the doubling iteration produced non-finite values (is (A, B) stabilisable?)")
        delta = num / max(den, 1e-300)
        Ak, G, H = A_next, G_next, H_next
        if delta <= tol:
            break
    else:
        raise NkError(f"solve_dare: no convergence after {max_iter} doublings (last relative change {delta:.3e}); "
                      "(A, B) must be stabilisable and (A, Q) detectable")
    return H, dict(iterations=it, delta=delta, residual=dare_residual(A, B, Q, R, H, ops))


def gain_from_solution(A, B, R, P, ops):
    """K = (R + B' P B)^-1 B' P A, (p, m): the first return value of control.dlqr."""
    PB = ops.mm(P, B)                                                # (m, p)
    M = _sym(R + ops.mm(B, PB, ta=True))
    return ops.spd_solve(M, ops.mm(PB, A, ta=True))


def dare_residual(A, B, Q, R, P, ops):
    """|| A'PA - A'PB (R + B'PB)^-1 B'PA + Q - P ||_F / || P ||_F."""
    PA = ops.mm(P, A)
    K = gain_from_solution(A, B, R, P, ops)
    res = ops.mm(A, PA, ta=True) - ops.mm(ops.mm(PA, B, ta=True), K) + _sym(Q) - P       # (PA)'B = A'PB, P symmetric
    return float(torch.linalg.norm(res)) / max(float(torch.linalg.norm(P)), 1e-300)


def dlqr(A, B, Q, R, eigenvalues=False, ops=None, device=None):
    """``control.dlqr(A, B, Q, R)`` -> (K, S, E) with numpy in and out, computed on the device (``device`` default: the engine's
    GPU; raises without one -- no CPU fallback in the product path; tests pass ``ops=TorchOps(), device='cpu'``).
    E, the closed-loop eigenvalues of A - B K, is only computed when asked for (numpy on the host: none of the scripts' loops
    needs it, benchmark_lqr_hjb.py only prints it) and is None otherwise."""
    import numpy as np
    if ops is None:
        from .engine import Engine
        eng = Engine.get()
        ops, device = EngineOps(eng), eng.tdev
    to = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(device)
    Ad, Bd, Qd, Rd = to(A), to(np.asarray(B, dtype=np.float64).reshape(np.asarray(A).shape[0], -1)), to(Q), to(np.atleast_2d(R))
    P, info = solve_dare(Ad, Bd, Qd, Rd, ops=ops)
    K = gain_from_solution(Ad, Bd, Rd, P, ops)
    Kh, Ph = K.cpu().numpy(), P.cpu().numpy()
    E = np.linalg.eigvals(np.asarray(A, dtype=np.float64) - np.asarray(B, dtype=np.float64).reshape(Kh.shape[1], -1) @ Kh) if eigenvalues else None
    dlqr.last_info = info
    return Kh, Ph, E
