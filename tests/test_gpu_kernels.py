"""GPU parity of every C-ABI entry point against the CPU oracle (oracle/nk_oracle.py) on seeded inputs.

Tolerances (relative Frobenius): kernel matrices / Grams 1e-12 (north star: Gram-level gate);
dense building blocks 1e-11..1e-12 scaled by conditioning as stated per test.
"""
import numpy as np
import pytest
import torch

from oracle import nk_oracle as O

pytestmark = pytest.mark.gpu


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()


def host(t):
    return t.detach().cpu().numpy()


def make_problem(n, d, p, m, seed, scale=1.0):
    rng = np.random.default_rng(seed)
    Xs = rng.standard_normal((n, d)) * scale
    U = rng.standard_normal((n, p))
    M = rng.standard_normal((d, d)) * 0.9 / np.sqrt(d)
    Bu = 0.1 * rng.standard_normal((d, p))
    Y = np.tanh(Xs @ M.T) * scale + (U @ Bu.T if p else 0.0)
    Z = Y[rng.choice(n, m, replace=False)]
    return Xs, U, Y, Z


GRAM_CASES = [
    # n, d, p, m, kind, ls, chunk
    (808, 192, 6, 100, O.RBF, 10.0, 0),         # cloth-like CV fold (ragged n, m not a tile multiple)
    (3030, 192, 6, 57, O.RBF, 10.0, 256),       # cloth full, odd m, smaller chunk -> 12 chunks
    (5000, 2, 1, 20, O.MATERN52, 1.0, 0),       # duffing-like
    (3980, 1, 1, 100, O.MATERN52, 0.5, 0),      # hjb-like d=1
    (1500, 7, 0, 10, O.RBF, 2.0, 128),          # no controls
    (2100, 33, 3, 260, O.MATERN52, 4.0, 512),   # three landmark blocks
    (127, 5, 2, 127, O.RBF, 1.5, 0),            # n < one tile, m == n
]


@pytest.mark.parametrize("n,d,p,m,kind,ls,chunk", GRAM_CASES)
def test_gram_parity(engine, n, d, p, m, kind, ls, chunk):
    Xs, U, Y, Z = make_problem(n, d, p, m, seed=n + m)
    if kind == O.RBF and d > 3:
        lsv = np.array([ls, ls * 1.5, ls * 0.75] * (d // 3 + 1))[:d]   # anisotropic, cycling like ThreeDimensionalKernel
    else:
        lsv = np.full(d, ls)
    ref = O.grams(Xs, Y, U, Z, kind, lsv)
    Xa = dev(np.hstack((Xs, U)))
    G = engine.grams(Xa, dev(Y), dev(Z), dev(1.0 / lsv), kind, p, chunk)
    torch.cuda.synchronize()
    for k in ("Gxx", "Gyx", "Gyy", "Gxu", "Gyu", "Guu", "GYy"):
        if ref[k].size == 0:
            continue
        err = O.relerr(host(G[k]), ref[k])
        assert err <= 1e-12, f"{k}: rel err {err:.3e}"
    # exact symmetry of the symmetric Grams (mirrored reads of one accumulator)
    assert torch.equal(G["Gxx"], G["Gxx"].T) and torch.equal(G["Gyy"], G["Gyy"].T)


@pytest.mark.parametrize("n,d,p,m,chunk", [(20000, 2, 1, 20, 128), (20000, 6, 2, 300, 128), (30000, 3, 1, 600, 256), (9000, 4, 1, 1000, 128)])
def test_gram_many_chunks_in_flight(engine, n, d, p, m, chunk):
    """Few landmarks: one chunk's items cannot fill the GPU, so nk_gram_begin keeps several chunks in flight (chunk c in buffer
    slot c % S, up to 16 slots; S = 2 again from m ~ 1000).  Every slot is reused many times here (n / chunk >> S): parity with
    the oracle, and bit-identical reruns (per-tile chunk order is fixed by the tile's version counter whatever the slot count)."""
    Xs, U, Y, Z = make_problem(n, d, p, m, seed=m)
    lsv = np.full(d, 1.5)
    ref = O.grams(Xs, Y, U, Z, O.MATERN52, lsv)
    Xa, Yd, Zd, il = dev(np.hstack((Xs, U))), dev(Y), dev(Z), dev(1.0 / lsv)
    G = engine.grams(Xa, Yd, Zd, il, O.MATERN52, p, chunk)
    Gb = engine.grams(Xa, Yd, Zd, il, O.MATERN52, p, chunk)
    torch.cuda.synchronize()
    assert torch.equal(G["_flat"], Gb["_flat"]), "not run-to-run deterministic with several chunks in flight"
    for k in ("Gxx", "Gyx", "Gyy", "Gxu", "Gyu", "Guu", "GYy"):
        assert O.relerr(host(G[k]), ref[k]) <= 1e-12, k


def test_gram_streaming_and_determinism(engine):
    """update() over sample blocks == one shot (shard-sum invariance, <=1e-13), and reruns are bit-identical."""
    n, d, p, m = 4000, 24, 2, 150
    Xs, U, Y, Z = make_problem(n, d, p, m, seed=5)
    lsv = np.full(d, 3.0)
    Xa, Yd, Zd, il = dev(np.hstack((Xs, U))), dev(Y), dev(Z), dev(1.0 / lsv)
    G1 = engine.grams(Xa, Yd, Zd, il, O.RBF, p)
    G1b = engine.grams(Xa, Yd, Zd, il, O.RBF, p)
    torch.cuda.synchronize()
    assert torch.equal(G1["_flat"], G1b["_flat"]), "fused Gram engine is not run-to-run deterministic"
    engine.gram_begin(Zd, il, O.RBF, p)
    for s, e in ((0, 1111), (1111, 1111), (1111, 3000), (3000, 4000)):
        if e > s:
            engine.gram_update(Xa[s:e], Yd[s:e])
    G2 = engine.gram_finalize()
    torch.cuda.synchronize()
    for k in ("Gxx", "Gyx", "Gyy", "Gxu", "Gyu", "Guu", "GYy"):
        assert O.relerr(host(G2[k]), host(G1[k])) <= 1e-13, k


@pytest.mark.parametrize("m,d,kind,ls", [(100, 192, O.RBF, 10.0), (37, 2, O.MATERN52, 1.0), (300, 1, O.MATERN52, 0.3), (129, 6, O.RBF, 2.0)])
def test_kzz_and_cross(engine, m, d, kind, ls):
    rng = np.random.default_rng(m)
    Z = rng.standard_normal((m, d))
    Z[3] = Z[1]                         # coincident landmarks (r = 0 off the diagonal)
    X = rng.standard_normal((211, d))
    X[0] = Z[2]
    lsv = np.full(d, ls)
    K = host(engine.kzz(dev(Z), dev(1.0 / lsv), kind))
    ref = O.kernel_matrix(Z, Z, kind, lsv)
    assert np.all(np.diag(K) == 1.0)
    assert np.array_equal(K, K.T)
    assert np.max(np.abs(K - ref)) <= 2e-14
    Kx = host(engine.kernel_cross(dev(Z), dev(X), dev(1.0 / lsv), kind))
    assert np.max(np.abs(Kx - O.kernel_matrix(Z, X, kind, lsv))) <= 2e-14


@pytest.mark.parametrize("M,N,K,ta,tb", [(128, 128, 128, 0, 1), (100, 57, 33, 0, 0), (300, 260, 515, 1, 0), (7, 1000, 129, 1, 1), (513, 131, 16, 0, 1)])
def test_gemm(engine, M, N, K, ta, tb):
    rng = np.random.default_rng(M + N + K)
    A = rng.standard_normal((K, M) if ta else (M, K))
    B = rng.standard_normal((N, K) if tb else (K, N))
    C0 = rng.standard_normal((M, N))
    out = dev(C0)
    engine.gemm(dev(A), dev(B), bool(ta), bool(tb), alpha=0.7, beta=-0.3, out=out)
    ref = 0.7 * (A.T if ta else A) @ (B.T if tb else B) - 0.3 * C0
    assert O.relerr(host(out), ref) <= 1e-14


@pytest.mark.parametrize("n", [1, 20, 128, 129, 300, 1000])
def test_potrf_trsm(engine, n):
    rng = np.random.default_rng(n)
    Q = rng.standard_normal((n, n + 5))
    A = Q @ Q.T + n * 1e-2 * np.eye(n)
    L = host(engine.potrf(dev(A)))
    Lref = np.linalg.cholesky(A)
    assert np.all(np.triu(L, 1) == 0.0)
    assert O.relerr(L, Lref) <= 1e-12
    B = rng.standard_normal((n, 37))
    X0 = host(engine.trsm_lower(dev(Lref), dev(B), trans=False))
    X1 = host(engine.trsm_lower(dev(Lref), dev(B), trans=True))
    import scipy.linalg
    assert O.relerr(X0, scipy.linalg.solve_triangular(Lref, B, lower=True)) <= 1e-11
    assert O.relerr(X1, scipy.linalg.solve_triangular(Lref.T, B, lower=False)) <= 1e-11


def test_potrf_not_spd(engine):
    from nys_koop_lqr_b200.engine import NkError
    A = np.eye(200)
    A[150, 150] = -1.0
    with pytest.raises(NkError):
        engine.potrf(dev(A))


@pytest.mark.parametrize("m,d,kind,ls", [(100, 192, O.RBF, 10.0), (20, 2, O.MATERN52, 1.0), (128, 1, O.MATERN52, 1.0), (1, 3, O.RBF, 1.0), (129, 8, O.RBF, 3.0),
                                         (300, 8, O.RBF, 3.0), (513, 192, O.RBF, 10.0),
                                         (1100, 192, O.RBF, 10.0), (1200, 2, O.MATERN52, 1.0), (1500, 6, O.RBF, 1.0)])
def test_sym_sqrt(engine, m, d, kind, ls):
    """S = sqrtm(K_mm): same distance to the eigh root as scipy's sqrtm has (the oracle's own floor).  From m = 1024 the
    Newton-Schulz schedule starts from an inverse-iteration ESTIMATE of lambda_min (nk_dense.cu) instead of the jitter bound:
    well-conditioned (RBF l=10: fewer iterations than the 17 of the bound), ill-conditioned (Matern in 2-D: lambda_min ~ jitter)
    and clustered (RBF l=1 in 6-D: K_mm close to the identity) spectra must all land on the same floor."""
    Xs, U, Y, Z = make_problem(3000, d, 1, m, seed=m)
    lsv = np.full(d, ls)
    Kmm = O.kernel_matrix(Z, Z, kind, lsv) + 1e-6 * np.eye(m)
    S, Sinv = engine.sym_sqrt(dev(Kmm))
    S, Sinv = host(S), host(Sinv)
    print(f"m={m}: {engine.last_sqrt_iters} Newton-Schulz iterations")
    w, V = np.linalg.eigh(Kmm)
    S0 = (V * np.sqrt(w)) @ V.T
    cond = w[-1] / w[0]
    assert np.array_equal(S, S.T) and np.array_equal(Sinv, Sinv.T)
    assert O.relerr(S @ S, Kmm) <= (1e-13 if m <= 1000 else 3e-13)
    assert O.relerr(S, S0) <= 1e-15 * max(10.0, cond ** 0.5) * 10
    assert np.linalg.norm(Sinv @ S - np.eye(m)) / np.sqrt(m) <= 1e-15 * max(10.0, cond ** 0.5) * 100


SOLVE_CASES = [
    (4000, 16, 2, 64, O.RBF, 4.0, 1e-4),
    (3000, 192, 6, 100, O.RBF, 10.0, 1e-2),
    (5000, 2, 1, 10, O.MATERN52, 1.0, 1e-6),
    (3000, 8, 0, 130, O.RBF, 3.0, 1e-3),
]


@pytest.mark.parametrize("n,d,p,m,kind,ls,gamma", SOLVE_CASES)
def test_solve_abc_vs_oracle(engine, n, d, p, m, kind, ls, gamma):
    """Dense stage on oracle Grams: vs the Cholesky statement <=1e-9*, vs the reference-order statement (sqrtm/solve/lstsq)."""
    Xs, U, Y, Z = make_problem(n, d, p, m, seed=m + d)
    lsv = np.full(d, ls)
    G = O.grams(Xs, Y, U, Z, kind, lsv)
    Kzz = O.kernel_matrix(Z, Z, kind, lsv)
    A0, B0, C0, W0 = O.solve_abc(G, Kzz, gamma * n, solver="chol")
    A1, B1, C1, W1 = O.solve_abc(G, Kzz, gamma * n, solver="reference")
    Gd = {k: dev(v) for k, v in G.items()}
    Kd = dev(Kzz)
    S, Sinv = engine.sym_sqrt(Kd + 1e-6 * torch.eye(m, dtype=torch.float64, device="cuda"))
    A, B, C, W = engine.solve_abc(Gd, Kd, S, Sinv, gamma * n)
    floor = max(O.relerr(A0, A1), 1e-12)    # distance between the two CPU statements = conditioning floor of this case
    for name, got, r0, r1 in (("A", A, A0, A1), ("B", B, B0, B1), ("C", C, C0, C1), ("W", W, W0, W1)):
        if r0.size == 0:
            continue
        e0, e1 = O.relerr(host(got), r0), O.relerr(host(got), r1)
        assert e1 <= max(1e-9, 20 * floor), f"{name}: vs reference-order oracle {e1:.2e} (floor {floor:.2e})"
        assert e0 <= max(1e-9, 20 * floor), f"{name}: vs cholesky oracle {e0:.2e}"


def test_lift_predict_rollout(engine):
    n, d, p, m = 3000, 12, 2, 80
    Xs, U, Y, Z = make_problem(n, d, p, m, seed=11)
    lsv = np.linspace(2.0, 4.0, d)
    fit = O.fit(np.hstack((Xs, U)), Y, p, O.RBF, lsv, 1e-4, Z=Z, solver="chol")
    Zd, il = dev(Z), dev(1.0 / lsv)
    Kmm = dev(fit["Kzz"] + 1e-6 * np.eye(m))
    S, Sinv = engine.sym_sqrt(Kmm)
    Xq = Xs[:333]
    phi = host(engine.lift(Zd, il, O.RBF, Sinv, dev(Xq)))
    phi_ref = O.lift(Z, Xq.T, O.RBF, lsv)
    assert O.relerr(phi, phi_ref) <= 1e-11
    phiT = host(engine.lift(Zd, il, O.RBF, Sinv, dev(Xq), transposed=True))
    assert np.array_equal(phiT.T, phi)
    Xaq = np.hstack((Xs, U))[:333]
    yh = host(engine.predict(Zd, il, O.RBF, Sinv, dev(fit["W"]), dev(Xaq), p))
    assert O.relerr(yh, O.predict(fit["W"], Z, Xaq, p, O.RBF, lsv)) <= 1e-11
    # rollout: 9 trajectories, T = 40, against the reference loop
    nb, T = 9, 40
    rng = np.random.default_rng(0)
    z0 = phi_ref[:, :nb].T.copy()
    Uc = rng.standard_normal((T - 1, nb, p))
    Yt = rng.standard_normal((T, nb, d))
    res = engine.rollout(dev(fit["A"]), dev(fit["B"]), dev(fit["C"]), dev(z0), dev(Uc), Ytrue=dev(Yt), return_final=True)
    Yh = host(res["Yhat"])
    for b in range(nb):
        sim = O.rollout(fit["A"], fit["B"], fit["C"], z0[b], Uc[:, b, :].T)     # (d, T)
        assert O.relerr(Yh[:, b, :].T, sim) <= 1e-12
        se, ss = host(res["sq_err"])[b], host(res["sq_sim"])[b]
        true = Yt[:, b, :].T
        assert abs(np.sqrt(se / true.size) - O.rmse_cloth(true, sim)) <= 1e-12 * O.rmse_cloth(true, sim)
        assert abs(np.sqrt(se) / np.sqrt(ss) * 100 - O.rmse_percent(true, sim)) <= 1e-12 * O.rmse_percent(true, sim)


def test_predict_more_rows_than_a_grid_dimension(engine):
    """N > 65535 rows (regression: the control-column copy used the row count as grid.y)."""
    rng = np.random.default_rng(0)
    N, d, p, m = 70001, 2, 1, 20
    Z = rng.standard_normal((m, d))
    ls = np.array([1.0, 2.0])
    Xa = rng.standard_normal((N, d + p))
    W = rng.standard_normal((d, m + p))
    Kmm = O.kernel_matrix(Z, Z, O.MATERN52, ls) + 1e-6 * np.eye(m)
    w, V = np.linalg.eigh(Kmm)
    Sinv = (V / np.sqrt(w)) @ V.T
    got = host(engine.predict(dev(Z), dev(1.0 / ls), O.MATERN52, dev(Sinv), dev(W), dev(Xa), p))
    want = (W @ np.vstack((Sinv @ O.kernel_matrix(Z, Xa[:, :d], O.MATERN52, ls), Xa[:, d:].T))).T
    assert O.relerr(got, want) <= 1e-10


@pytest.mark.parametrize("m,p,d,nb,T", [(57, 1, 2, 300, 7), (131, 0, 5, 129, 4), (260, 3, 192, 1, 12), (10, 2, 1, 1000, 1), (128, 6, 64, 257, 3)])
def test_rollout_shapes(engine, m, p, d, nb, T):
    """Packed persistent GEMM rollout on shapes that are not tile multiples (odd m, no controls, one trajectory, T = 1),
    element-wise against the reference loop (benchmark_lqr_cloth.py:29-32) and the returned final lifted state."""
    rng = np.random.default_rng(m + nb)
    A = rng.standard_normal((m, m)) * (0.9 / np.sqrt(m))
    B = rng.standard_normal((m, p))
    Cm = rng.standard_normal((d, m))
    z0 = rng.standard_normal((nb, m))
    Uc = rng.standard_normal((max(T - 1, 0), nb, p))
    Ud = dev(Uc) if (p and T > 1) else (torch.zeros(T - 1, nb, p, dtype=torch.float64).cuda() if T > 1 else None)
    if T == 1:
        res = engine.rollout(dev(A), dev(B) if p else None, dev(Cm), dev(z0), None, Ytrue=dev(rng.standard_normal((1, nb, d))), return_final=True)
    else:
        res = engine.rollout(dev(A), dev(B) if p else None, dev(Cm), dev(z0), Ud, return_final=True)
    Yh = host(res["Yhat"])
    assert Yh.shape == (T, nb, d)
    for b in range(0, nb, max(1, nb // 7)):
        sim = O.rollout(A, B, Cm, z0[b], Uc[:, b, :].T if T > 1 else np.zeros((p, 0)))
        assert O.relerr(Yh[:, b, :].T, sim) <= 1e-12
    z = z0.copy()
    for i in range(T - 1):
        z = z @ A.T + (Uc[i] @ B.T if p else 0.0)
    assert O.relerr(host(res["Zfinal"]), z) <= 1e-12


@pytest.mark.parametrize("m,p,d,nb,steps", [(80, 2, 12, 5, 30), (131, 6, 192, 130, 6), (20, 1, 2, 1, 60)])
def test_closed_loop_against_reference_loop(engine, m, p, d, nb, steps):
    """Batched lifted closed loop vs the reference's lqr_control loop body (benchmark_lqr_cloth.py:80-84), same operation order."""
    rng = np.random.default_rng(m * 7 + nb)
    A = rng.standard_normal((m, m)) * (0.6 / np.sqrt(m))
    B = rng.standard_normal((m, p)) * 0.3
    Cm = rng.standard_normal((d, m))
    K, _ = O.dlqr(A, B, np.eye(m), np.eye(p))        # any stabilising gain; DARE stays on the host
    z0, zr = rng.standard_normal((nb, m)), rng.standard_normal((nb, m))
    Xs, Us, Zf = engine.closed_loop(dev(A), dev(B), dev(Cm), dev(K), dev(z0), dev(zr), steps, return_final=True)
    Xs, Us = host(Xs), host(Us)
    for b in range(0, nb, max(1, nb // 5)):
        xs, us = O.closed_loop(A, B, Cm, K, z0[b], zr[b], steps)
        assert O.relerr(Xs[:, b, :].T, xs) <= 1e-11
        assert O.relerr(Us[:, b, :].T, us) <= 1e-11


def test_kernel_function_accuracy(engine):
    """The device kernel function (nk_kernel_function: what every lift epilogue evaluates) against numpy on 2e6 exponents spanning
    the whole range, including the denormal tail and exact underflow: <= 2 ulp for normal results, absolute 1e-320 below."""
    rng = np.random.default_rng(0)
    e = -np.concatenate([rng.uniform(0, 40, 1_000_000), 10.0 ** rng.uniform(-12, 2.9, 900_000), rng.uniform(700, 760, 100_000),
                         [0.0, 1e-300, 708.0, 745.0, 746.0, 2000.0, 1e300]])
    got = host(engine.kernel_function(dev(e), O.RBF))
    want = np.exp(e)
    normal = want > 1e-300
    ulp = np.abs(got[normal] - want[normal]) / np.spacing(want[normal])
    assert ulp.max() <= 2.0, ulp.max()
    assert np.abs(got[~normal] - want[~normal]).max() <= 1e-300 * 1e-15 + 1e-320
    assert got[-7] == 1.0 and got[-2] == 0.0 and got[-1] == 0.0
    # Matern-5/2: (1 + a + a^2/3) exp(-a), a = sqrt(5) r, exponent = -r^2/2
    r = rng.uniform(0, 30, 200_000)
    gm = host(engine.kernel_function(dev(-0.5 * r * r), O.MATERN52))
    a = np.sqrt(5.0) * r
    wm = (1.0 + a + a * a / 3.0) * np.exp(-a)
    assert (np.abs(gm - wm) / wm).max() <= 1e-14        # r is recovered from -r^2/2: a few ulp of a in the exponent


def test_release_scratch(engine):
    """Workspaces can be dropped and are rebuilt on demand; an accumulation in progress is discarded loudly."""
    from nys_koop_lqr_b200._lib import NkError
    Xs, U, Y, Z = make_problem(500, 4, 1, 30, seed=1)
    lsv = np.full(4, 2.0)
    Xa, Yd, Zd, il = dev(np.hstack((Xs, U))), dev(Y), dev(Z), dev(1.0 / lsv)
    G1 = engine.grams(Xa, Yd, Zd, il, O.RBF, 1)
    engine.gram_begin(Zd, il, O.RBF, 1)
    engine.release_scratch()
    with pytest.raises(NkError):
        engine.gram_update(Xa, Yd)
    G2 = engine.grams(Xa, Yd, Zd, il, O.RBF, 1)
    assert torch.equal(G1["_flat"], G2["_flat"])
