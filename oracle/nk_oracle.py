"""CPU oracle for the Nystrom-Koopman hot path  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module.  The product path (``nys_koop_lqr_b200``) never does: it calls the
sm_100a kernels through the C-ABI and raises if the extension is missing.

What it is: a numpy/scipy float64 restatement of ``KoopmanNystromRegressor.fit / lift / predict``
(reference ``regressors.py:114-178``, ``48-55``) and of the open-loop rollout ``validate_dyn_sys``
(``benchmark_lqr_cloth.py:18-36``, ``benchmark_lqr_classic.py:23-41``, ``benchmark_lqr_hjb.py:23-44``),
factored the way the CUDA path is factored: kernel lift -> seven data-sample Grams -> landmark
matrices -> two regularised solves -> (A, B, C, W).

Third-party arithmetic the reference delegates to (not vendored under /root/reference, versions
unpinned by its requirements.txt -- "parity unpinned" for these):
  * scikit-learn ``RBF.__call__`` / ``Matern.__call__`` (nu=2.5) -> ``scipy.spatial.distance.cdist``;
    restated in ``kernel_matrix`` from the published formulas
    (RBF: exp(-0.5*sum(((x-c)/l)^2)); Matern-5/2: (1+a+a^2/3)exp(-a), a=sqrt(5)*||(x-c)/l||).
  * ``scipy.linalg.sqrtm / solve(assume_a='her') / lstsq`` -- ``solver='reference'`` below calls exactly
    these in the reference's order; ``solver='chol'`` is the SPD-Cholesky statement the GPU implements.
  * ``control.dlqr`` -> ``scipy.linalg.solve_discrete_are`` + K=(B'PB+R)^-1 B'PA (golden G4 pins it).

Pinning: PINNED.  tests/test_oracle_vs_reference.py checks this file against the reference imported from
/root/reference (when mounted); tests/test_oracle_golden.py and tests/test_oracle_cv_golden.py against fixtures the
reference (and sklearn's GridSearchCV driving it) generated (tests/golden/make_golden*.py); tests/test_golden_g1.py,
test_golden_g2.py and test_golden_g3.py against the reference's OWN published artefacts: the Duffing
and cloth forecast-RMSE result files (G1, G2: 6 significant digits), a regressor pickled by the reference and its exported
LQR gain (G3, G4).  Only the third-party library versions are unpinned (see above).
"""
from __future__ import annotations

import numpy as np
import scipy.linalg
from scipy.spatial.distance import cdist

RBF, MATERN52 = 0, 1
JITTER = 1e-6  # regressors.py:120


# --------------------------------------------------------------------------------------------
# kernel lift  (regressors.py:139,141-144 -> sklearn kernels.py RBF / Matern nu=2.5)
# --------------------------------------------------------------------------------------------
def kernel_matrix(A_rows, B_rows, kind, length_scale):
    """k(A, B) with rows = points, via scipy cdist exactly as sklearn's RBF / Matern __call__ do."""
    A_rows = np.atleast_2d(np.asarray(A_rows, dtype=np.float64))
    B_rows = np.atleast_2d(np.asarray(B_rows, dtype=np.float64))
    ls = np.broadcast_to(np.asarray(length_scale, dtype=np.float64).reshape(-1), (A_rows.shape[1],)) \
        if np.size(length_scale) in (1, A_rows.shape[1]) else None
    if ls is None:
        raise ValueError("length_scale must be scalar or have one entry per state dimension")
    # sklearn: dists = cdist(X / length_scale, Y / length_scale, metric='sqeuclidean' | 'euclidean') -- the same scipy
    # C loop the reference ends up in (direct differences, single-threaded), so timing this port times that path
    out = cdist(A_rows / ls, B_rows / ls, metric="sqeuclidean")
    if kind == RBF:
        return np.exp(-0.5 * out)
    if kind == MATERN52:
        a = np.sqrt(out) * np.sqrt(5.0)
        return (1.0 + a + a * a / 3.0) * np.exp(-a)
    raise NotImplementedError(f"kernel kind {kind}")


# --------------------------------------------------------------------------------------------
# the seven data-sample Grams  (regressors.py:147,151,153,162,164)
# --------------------------------------------------------------------------------------------
def grams(Xs, Y, U, Z, kind, length_scale, chunk=8192, threads=1, Z_in=None):
    """Xs (n,d) states, Y (n,d) next states, U (n,p) controls, Z (m,d) landmarks.

    Returns dict: Gxx = Phi_x Phi_x' (m,m), Gyx = Phi_y Phi_x' (m,m), Gyy (m,m), Gxu = Phi_x U (m,p),
    Gyu (m,p), Guu (p,p), GYy = Y' Phi_y' (d,m); Phi_x = k(Z, Xs) (m,n), Phi_y = k(Z, Y).
    Z_in: distinct input landmarks (regressors.py:133-134, 142): Phi_x = k(Z_in, Xs); default Z_in = Z.
    Chunked over samples so n*m never has to exist (the reference materialises it; the sums are the same).
    threads > 1: the kernel lifts of a chunk (scipy cdist releases the GIL) are computed by a thread pool over column
    strips -- same arithmetic per entry, so the result is bit-identical to threads=1; only the wall time changes (the
    reference itself runs cdist on one thread; the CPU baseline of bench.py says which it timed).
    """
    n, d = Xs.shape
    m, p = Z.shape[0], U.shape[1]
    G = dict(Gxx=np.zeros((m, m)), Gyx=np.zeros((m, m)), Gyy=np.zeros((m, m)), Gxu=np.zeros((m, p)),
             Gyu=np.zeros((m, p)), Guu=np.zeros((p, p)), GYy=np.zeros((d, m)))
    pool = None
    if threads and threads > 1:
        from concurrent.futures import ThreadPoolExecutor
        pool = ThreadPoolExecutor(int(threads))

    def lifted(rows, Zc=None):
        Zc = Z if Zc is None else Zc
        if pool is None or rows.shape[0] < 2 * threads:
            return kernel_matrix(Zc, rows, kind, length_scale)
        cuts = np.linspace(0, rows.shape[0], int(threads) + 1).astype(int)
        parts = list(pool.map(lambda i: kernel_matrix(Zc, rows[cuts[i]:cuts[i + 1]], kind, length_scale), range(int(threads))))
        return np.hstack(parts)
    try:
        for s in range(0, n, chunk):
            e = min(n, s + chunk)
            Px = lifted(Xs[s:e], Z_in)
            Py = lifted(Y[s:e])
            Uc, Yc = U[s:e], Y[s:e]
            G["Gxx"] += Px @ Px.T
            G["Gyx"] += Py @ Px.T
            G["Gyy"] += Py @ Py.T
            G["Gxu"] += Px @ Uc
            G["Gyu"] += Py @ Uc
            G["Guu"] += Uc.T @ Uc
            G["GYy"] += Yc.T @ Py.T
    finally:
        if pool is not None:
            pool.shutdown()
    return G


# --------------------------------------------------------------------------------------------
# landmark matrices  (regressors.py:139-140,143-144,163)
# --------------------------------------------------------------------------------------------
def landmark_matrices(Z, kind, length_scale, sqrt="eigh"):
    """K_zz = k(Z,Z) (no jitter, :144), K_mm = K_zz + 1e-6 I (:139,143), S = K_mm^(1/2) symmetric, S^-1."""
    Kzz = kernel_matrix(Z, Z, kind, length_scale)
    Kmm = Kzz + JITTER * np.eye(Z.shape[0])
    if sqrt == "sqrtm":
        S = scipy.linalg.sqrtm(Kmm).real
        Sinv = None
    else:
        w, V = np.linalg.eigh(Kmm)
        S = (V * np.sqrt(w)) @ V.T
        Sinv = (V / np.sqrt(w)) @ V.T
    return Kzz, Kmm, S, Sinv


# --------------------------------------------------------------------------------------------
# dense stage: Grams -> (A, B, C, W)   (regressors.py:147-169)
# --------------------------------------------------------------------------------------------
def solve_abc(G, Kzz, gamma_n, solver="chol", Kzz_in=None, Kio=None):
    """solver='reference': scipy sqrtm / solve(assume_a='her') / lstsq in the reference's call order.
    solver='chol': eigh root + Cholesky solves (what the sm_100a dense stage computes).
    Kzz_in = k(Z_in, Z_in), Kio = k(Z_in, Z_out) (regressors.py:143-144) when the input landmarks differ; default both = Kzz."""
    m = Kzz.shape[0]
    p = G["Guu"].shape[0]
    Kmm = Kzz + JITTER * np.eye(m)
    Kmm_in = Kmm if Kzz_in is None else Kzz_in + JITTER * np.eye(m)
    Kio = Kzz if Kio is None else Kio
    inner = np.empty((m + p, m + p))
    inner[:m, :m] = G["Gxx"] + gamma_n * Kmm_in
    inner[:m, m:] = G["Gxu"]
    inner[m:, :m] = G["Gxu"].T
    inner[m:, m:] = G["Guu"] + gamma_n * np.eye(p)
    cross = np.hstack((G["Gyx"], G["Gyu"]))  # K_mn_out @ K_mn_in.T  (:153)
    inner_rec = gamma_n * Kmm + G["Gyy"]      # (:162)
    if solver == "reference":
        S = scipy.linalg.sqrtm(Kmm).real
        right = scipy.linalg.block_diag(scipy.linalg.solve(S, Kio.T, assume_a="her").T, np.eye(p))
        left = scipy.linalg.solve(S, cross, assume_a="her")
        sol = scipy.linalg.lstsq(inner, right)[0]
        Gls = left @ sol
        sol_rec = scipy.linalg.lstsq(inner_rec, scipy.linalg.sqrtm(Kmm).real)[0]
        C = G["GYy"] @ sol_rec
    elif solver == "chol":
        w, V = np.linalg.eigh(Kmm)
        S = (V * np.sqrt(w)) @ V.T
        Sinv = (V / np.sqrt(w)) @ V.T
        right = scipy.linalg.block_diag(Kio @ Sinv, np.eye(p))
        left = Sinv @ cross
        sol = scipy.linalg.cho_solve(scipy.linalg.cho_factor(inner, lower=True), right)
        Gls = left @ sol
        sol_rec = scipy.linalg.cho_solve(scipy.linalg.cho_factor(inner_rec, lower=True), S)
        C = G["GYy"] @ sol_rec
    else:
        raise ValueError(solver)
    A, B = Gls[:, :m], Gls[:, m:]
    W = C @ Gls
    return A, B, C, W


def draw_landmarks(Y, m):
    """regressors.py:129-132 -- global legacy RNG, one draw, landmarks are NEXT-state samples. Y is (n,d)."""
    idx = np.random.choice(np.arange(0, Y.shape[0]), size=m, replace=False)
    return Y[idx]


def fit(X_aug, Y, n_inputs, kind, length_scale, gamma, m=None, Z=None, solver="chol", Z_in=None):
    """X_aug (n, d+p) with the p controls LAST (regressors.py:122-126), Y (n,d). Returns dict.  Z_in: distinct input landmarks."""
    X_aug = np.asarray(X_aug, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    n = X_aug.shape[0]
    d = X_aug.shape[1] - n_inputs
    if Z is None:
        Z = draw_landmarks(Y, m)
    G = grams(X_aug[:, :d], Y, X_aug[:, d:], Z, kind, length_scale, Z_in=Z_in)
    Kzz = kernel_matrix(Z, Z, kind, length_scale)
    Kzz_in = Kio = None
    if Z_in is not None:
        Kzz_in, Kio = kernel_matrix(Z_in, Z_in, kind, length_scale), kernel_matrix(Z_in, Z, kind, length_scale)
    A, B, C, W = solve_abc(G, Kzz, gamma * n, solver=solver, Kzz_in=Kzz_in, Kio=Kio)
    return dict(A=A, B=B, C=C, W=W, Z=Z, G=G, Kzz=Kzz)


def lift(Z, Xcols, kind, length_scale, solver="chol"):
    """regressors.py:171-178: phi = S^-1 k(Z, X); Xcols is (d, N) column-samples; returns (m, N)."""
    Kzz = kernel_matrix(Z, Z, kind, length_scale)
    Kmm = Kzz + JITTER * np.eye(Z.shape[0])
    Kmn = kernel_matrix(Z, np.asarray(Xcols).T, kind, length_scale)
    if solver == "reference":
        return scipy.linalg.solve(scipy.linalg.sqrtm(Kmm).real, Kmn, assume_a="her")
    w, V = np.linalg.eigh(Kmm)
    return (V / np.sqrt(w)) @ (V.T @ Kmn)


def predict(W, Z, X_aug, n_inputs, kind, length_scale, solver="chol"):
    """regressors.py:48-55: (W @ [phi(x); u]).T for X_aug (N, d+p) -> (N, d)."""
    X_aug = np.asarray(X_aug, dtype=np.float64)
    d = X_aug.shape[1] - n_inputs
    phi = lift(Z, X_aug[:, :d].T, kind, length_scale, solver)
    return (W @ np.vstack((phi, X_aug[:, d:].T))).T


# --------------------------------------------------------------------------------------------
# open-loop rollout  (benchmark_lqr_cloth.py:18-36 and copies)
# --------------------------------------------------------------------------------------------
def rollout(A, B, C, z0, controls):
    """z0 (m,), controls (p, T-1) -> simulated states (d, T): y_0 = C z0; z_{i+1} = A z_i + B u_i."""
    z = np.asarray(z0, dtype=np.float64).reshape(-1, 1)
    T = controls.shape[1] + 1
    out = np.empty((C.shape[0], T))
    out[:, 0:1] = C @ z
    for i in range(T - 1):
        z = A @ z + B @ controls[:, i].reshape(-1, 1)
        out[:, i + 1:i + 2] = C @ z
    return out


def closed_loop(A, B, C, K, z0, zref, num_steps):
    """benchmark_lqr_cloth.py:80-84 (lqr_control loop body) for one trajectory: u = K (phi_ref - phi); state = C phi;
    phi <- A phi + B u.  z0, zref (m,).  Returns states (d, num_steps), controls (p, num_steps)."""
    z = np.asarray(z0, dtype=np.float64).reshape(-1, 1)
    zr = np.asarray(zref, dtype=np.float64).reshape(-1, 1)
    xs, us = [], []
    for _ in range(num_steps):
        u = K @ (zr - z)
        us.append(u[:, 0])
        xs.append((C @ z)[:, 0])
        z = A @ z + B @ u
    return np.array(xs).T, np.array(us).T


def rmse_cloth(true_traj, sim):
    """benchmark_lqr_cloth.py:34"""
    return float(np.sqrt(np.mean(np.square(true_traj - sim))))


def rmse_percent(true_traj, sim):
    """benchmark_lqr_classic.py:39 / benchmark_lqr_hjb.py:42 (denominator is the SIMULATED trajectory)."""
    return float(np.sqrt(np.sum(np.square(true_traj - sim))) / np.sqrt(np.sum(np.square(sim))) * 100)


# --------------------------------------------------------------------------------------------
# hyper-parameter search  (learn_hyperparams: benchmark_lqr_cloth.py:39-66, _classic.py:44-64, _hjb.py:47-71)
# --------------------------------------------------------------------------------------------
def kfold_bounds(n, n_splits=5):
    """sklearn KFold(n_splits) without shuffling (GridSearchCV's default cv for a regressor): contiguous blocks, the
    first n % n_splits of them one sample longer.  Returns [(start, stop), ...]."""
    sizes = np.full(n_splits, n // n_splits, dtype=int)
    sizes[: n % n_splits] += 1
    stops = np.cumsum(sizes)
    return [(int(e - s), int(e)) for s, e in zip(sizes, stops)]


def neg_rmse(Y_true, Y_pred):
    """sklearn 'neg_root_mean_squared_error' with multioutput='uniform_average': mean over outputs of per-output RMSE."""
    return -float(np.mean(np.sqrt(np.mean(np.square(Y_true - Y_pred), axis=0))))


def cv_scores(X_aug, Y, n_inputs, kinds_ls, gammas, Z, n_splits=5, solver="chol"):
    """GridSearchCV restated with a fixed landmark set Z (m,d): for every kernel (kind, length_scale), gamma and fold,
    fit on the training rows (gamma_n = gamma * n_train, regressors.py:127) and score `predict` on the held-out block.
    Returns scores[kernel, gamma, fold]."""
    X_aug = np.asarray(X_aug, dtype=np.float64)
    Y = np.asarray(Y, dtype=np.float64)
    n = X_aug.shape[0]
    out = np.empty((len(kinds_ls), len(gammas), n_splits))
    for ki, (kind, ls) in enumerate(kinds_ls):
        for fi, (s, e) in enumerate(kfold_bounds(n, n_splits)):
            tr = np.r_[0:s, e:n]
            for gi, gamma in enumerate(gammas):
                f = fit(X_aug[tr], Y[tr], n_inputs, kind, ls, gamma, Z=Z, solver=solver)
                out[ki, gi, fi] = neg_rmse(Y[s:e], predict(f["W"], Z, X_aug[s:e], n_inputs, kind, ls, solver))
    return out


def cv_weights(G, Kzz, gamma_n):
    """Prediction weights in kernel-matrix coordinates, Wk (d, m+p), such that regressors.py:48-55 reads
    Yhat = Wk [k(Z,x); u].  With weights = C G_ls (regressors.py:167) and lift = S^-1 k(Z,x) (:171-178) the symmetric square
    roots cancel:  Wk = [V_phi Kzz Kmm^-1 | V_u],  V = GYy inner_rec^-1 [Gyx|Gyu] inner^-1  (inner / inner_rec of :151,:162),
    so that large landmark counts need no eigen-decomposition.  tests/test_oracle_golden.py::test_cv_weights_identity checks this
    algebra against `solve_abc` + `lift`."""
    m = Kzz.shape[0]
    p = G["Guu"].shape[0]
    Kmm = Kzz + JITTER * np.eye(m)
    inner = np.empty((m + p, m + p))
    inner[:m, :m] = G["Gxx"] + gamma_n * Kmm
    inner[:m, m:] = G["Gxu"]
    inner[m:, :m] = G["Gxu"].T
    inner[m:, m:] = G["Guu"] + gamma_n * np.eye(p)
    cross = np.hstack((G["Gyx"], G["Gyu"]))
    inner_rec = gamma_n * Kmm + G["Gyy"]
    Ta = scipy.linalg.cho_solve(scipy.linalg.cho_factor(inner_rec, lower=True), G["GYy"].T).T      # GYy inner_rec^-1
    V = scipy.linalg.cho_solve(scipy.linalg.cho_factor(inner, lower=True), (Ta @ cross).T).T      # ... cross inner^-1
    Wphi = scipy.linalg.cho_solve(scipy.linalg.cho_factor(Kmm, lower=True), (V[:, :m] @ Kzz).T).T
    return np.hstack((Wphi, V[:, m:]))


def dlqr(A, B, Q, R):
    """control.dlqr stand-in (benchmark_lqr_cloth.py:262): DARE + K = (B'PB+R)^-1 B'PA."""
    P = scipy.linalg.solve_discrete_are(A, B, Q, R)
    K = np.linalg.solve(B.T @ P @ B + R, B.T @ P @ A)
    return K, P


# --------------------------------------------------------------------------------------------
# synthetic generator of SURVEY.md section 8(d) / BASELINE.md section 3 (numpy statement; the device
# generator in the product reproduces the same family, parity runs use this one on both sides)
# --------------------------------------------------------------------------------------------
def synthetic(n, d=192, p=6, seed=0):
    rng = np.random.default_rng(seed)
    Xs = rng.standard_normal((n, d))
    U = rng.standard_normal((n, p))
    M = rng.standard_normal((d, d)) * 0.9 / np.sqrt(d)
    Bu = 0.1 * rng.standard_normal((d, p))
    Y = np.tanh(Xs @ M.T) + U @ Bu.T
    return Xs, U, Y


def relerr(a, b):
    return float(np.linalg.norm(np.asarray(a) - np.asarray(b)) / max(np.linalg.norm(np.asarray(b)), 1e-300))
