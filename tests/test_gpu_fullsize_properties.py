"""GPU: size-independent properties of the fused lift+Gram kernel at the BASELINE shapes (m=4096, d=192, p=6), where the
CPU oracle cannot run.  The checksums are computed through an INDEPENDENT device path (kernel rows from nk_kernel_cross in
blocks, reduced with plain tensor sums), never through the fused kernel:

    1' Gxx 1 = sum_s (1' phi_x(s))^2      1' Gyx 1 = sum_s (1' phi_y(s)) (1' phi_x(s))      1' Gyy 1 = sum_s (1' phi_y(s))^2
    1' Gxu   = sum_s (1' phi_x(s)) u_s    1' Gyu   = sum_s (1' phi_y(s)) u_s                GYy 1    = sum_s y_s (1' phi_y(s))
    Guu = U'U,   trace(Gxx) = sum_s |phi_x(s)|^2,   shard-sum invariance (two halves == whole), bit-identical reruns.

n defaults to 1e6 so that the suite stays fast; NK_FULLSIZE=1 runs the full n=1e7 of BASELINE.json configs[3].
"""
import os

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_gram_checksums_at_baseline_shapes(engine):
    n = 10_000_000 if os.environ.get("NK_FULLSIZE") == "1" else 1_000_000
    m, d, p = 4096, 192, 6
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(7)
    M = torch.randn(d, d, dtype=torch.float64, device=dev, generator=g) * (0.9 / d ** 0.5)
    Bu = 0.1 * torch.randn(d, p, dtype=torch.float64, device=dev, generator=g)
    X = torch.empty(n, d + p, dtype=torch.float64, device=dev)
    Y = torch.empty(n, d, dtype=torch.float64, device=dev)
    for s in range(0, n, 1 << 19):
        e = min(n, s + (1 << 19))
        X[s:e].normal_(generator=g)
        Y[s:e] = torch.tanh(X[s:e, :d] @ M.T) + X[s:e, d:] @ Bu.T
    Z = Y[torch.randperm(n, device=dev, generator=g)[:m]].contiguous()
    il = torch.full((d,), 0.1, dtype=torch.float64, device=dev)
    G = engine.grams(X, Y, Z, il, 0, p)
    # independent path: kernel rows in blocks
    rx = torch.empty(n, dtype=torch.float64, device=dev)
    ry = torch.empty(n, dtype=torch.float64, device=dev)
    tr_xx = torch.zeros((), dtype=torch.float64, device=dev)
    blk = 1 << 16
    for s in range(0, n, blk):
        e = min(n, s + blk)
        Kx = engine.kernel_cross(Z, X[s:e, :d], il, 0)         # (m, N)
        rx[s:e] = Kx.sum(dim=0)
        tr_xx += (Kx * Kx).sum()
        ry[s:e] = engine.kernel_cross(Z, Y[s:e], il, 0).sum(dim=0)
    U = X[:, d:]
    rel = lambda a, b: float((a - b).abs().max() / b.abs().max())
    assert rel(G["Gxx"].sum(), (rx * rx).sum()) <= 1e-11
    assert rel(G["Gyx"].sum(), (ry * rx).sum()) <= 1e-11
    assert rel(G["Gyy"].sum(), (ry * ry).sum()) <= 1e-11
    assert rel(G["Gxx"].diagonal().sum(), tr_xx) <= 1e-11
    assert rel(G["Gxu"].sum(dim=0), rx @ U) <= 1e-10
    assert rel(G["Gyu"].sum(dim=0), ry @ U) <= 1e-10
    assert rel(G["GYy"].sum(dim=1), Y.T @ ry) <= 1e-11
    assert rel(G["Guu"], U.T @ U) <= 1e-12
    assert torch.equal(G["Gxx"], G["Gxx"].T) and torch.equal(G["Gyy"], G["Gyy"].T)
    # shard-sum invariance and determinism at this size
    h = n // 2 + 77
    engine.gram_begin(Z, il, 0, p)
    engine.gram_update(X[:h], Y[:h])
    engine.gram_update(X[h:], Y[h:])
    G2 = engine.gram_finalize()
    assert rel(G2["_flat"], G["_flat"]) <= 1e-13
    G3 = engine.grams(X, Y, Z, il, 0, p)
    assert torch.equal(G3["_flat"], G["_flat"])


def test_rollout_linearity_at_large_m(engine):
    """The rollout is linear in (z0, U): rollout(z0a + z0b, Ua + Ub) == rollout(z0a, Ua) + rollout(z0b, Ub), at m=4096 with
    20 000 trajectories (the CPU loop would need hours); plus agreement with a plain float64 tensor recurrence on a slice."""
    m, p, d, nb, T = 4096, 6, 192, 20_000, 6
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(3)
    A = torch.randn(m, m, dtype=torch.float64, device=dev, generator=g) * (0.9 / m ** 0.5)
    B = torch.randn(m, p, dtype=torch.float64, device=dev, generator=g)
    C = torch.randn(d, m, dtype=torch.float64, device=dev, generator=g) / m ** 0.5
    za, zb = (torch.randn(nb, m, dtype=torch.float64, device=dev, generator=g) for _ in range(2))
    Ua, Ub = (torch.randn(T - 1, nb, p, dtype=torch.float64, device=dev, generator=g) for _ in range(2))
    ya = engine.rollout(A, B, C, za, Ua)["Yhat"]
    yb = engine.rollout(A, B, C, zb, Ub)["Yhat"]
    yab = engine.rollout(A, B, C, za + zb, Ua + Ub, return_final=True)
    err = float((yab["Yhat"] - (ya + yb)).norm() / yab["Yhat"].norm())
    assert err <= 1e-13, err
    # a slice against the plain recurrence (torch float64 matmul as an independent device path)
    z = (za + zb)[:64].clone()
    for i in range(T):
        yi = z @ C.T
        assert float((yab["Yhat"][i, :64] - yi).norm() / yi.norm()) <= 1e-12
        if i < T - 1:
            z = z @ A.T + (Ua + Ub)[i, :64] @ B.T
    assert float((yab["Zfinal"][:64] - z).norm() / z.norm()) <= 1e-12


def test_dense_stage_residuals_at_large_m(engine):
    """Blocked Cholesky (two-level, 9 panels), triangular solves and the symmetric square root at m=4096+37 on a real kernel
    matrix: residual identities instead of a CPU comparison (LAPACK would take minutes here)."""
    m, d = 4096 + 37, 24
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(11)
    Z = torch.randn(m, d, dtype=torch.float64, device=dev, generator=g)
    il = torch.full((d,), 1.0 / 3.0, dtype=torch.float64, device=dev)
    K = engine.kzz(Z, il, 0)
    K.diagonal().add_(1e-6)
    nrm = float(K.norm())
    L = engine.potrf(K.clone())
    assert float((L @ L.T - K).norm()) <= 1e-14 * nrm * 10
    assert float(torch.triu(L, 1).abs().max()) == 0.0
    # triangular solves: L (L^-1 B) == B, L^T (L^-T B) == B
    B = torch.randn(m, 130, dtype=torch.float64, device=dev, generator=g)
    Y1 = engine.trsm_lower(L, B.clone(), trans=False)
    assert float((L @ Y1 - B).norm() / B.norm()) <= 1e-10
    Y2 = engine.trsm_lower(L, B.clone(), trans=True)
    assert float((L.T @ Y2 - B).norm() / B.norm()) <= 1e-10
    S, Sinv = engine.sym_sqrt(K, 1e-6)
    assert float((S @ S - K).norm()) <= 1e-12 * nrm
    assert float((S - S.T).abs().max()) == 0.0 and float((Sinv - Sinv.T).abs().max()) == 0.0
    eye = torch.eye(m, dtype=torch.float64, device=dev)
    assert float((Sinv @ K @ Sinv - eye).norm() / eye.norm()) <= 1e-8       # S^-1 K S^-1 = I, amplified by cond(K) ~ 1e9 * eps
