"""GPU edge cases of the estimator path (through the drop-in class and the C ABI) against the oracle: degenerate sizes
(one landmark, one state dimension, no controls, fewer samples than one tile), landmarks that coincide (K_mm singular but for
the 1e-6 jitter of regressors.py:139), samples that coincide with landmarks (r = 0 inside the lift), sample counts around the
128-sample strip and 512-sample chunk edges, and an empty sample block inside a streamed accumulation.
"""
import numpy as np
import pytest

from oracle import nk_oracle as O

pytestmark = pytest.mark.gpu


def _fit_pair(n, d, p, m, kind, ls, gamma, seed=0, Z=None):
    import regressors as R
    Xs, U, Y = O.synthetic(n, d, p, seed=seed)
    X = np.hstack((Xs, U)) if p else Xs
    lsv = np.full(d, float(ls))
    if Z is None:
        np.random.seed(seed)
        Z = O.draw_landmarks(Y, m)
    holder = R.ThreeDimensionalKernel(ls, ls, ls, d) if kind == O.RBF else R.KernelWrapper(list(lsv))
    reg = R.KoopmanNystromRegressor(p, kernel=holder, gamma=gamma, m=m)
    reg.nystrom_centers_output = np.ascontiguousarray(Z.T)
    reg.fit(X, Y)
    want = O.fit(X, Y, p, kind, lsv, gamma, Z=Z)
    return reg, want, (X, Y, Z, lsv)


@pytest.mark.parametrize("n,d,p,m,kind", [
    (3, 1, 0, 1, O.RBF),            # one landmark, one state, no controls, three samples
    (2, 2, 1, 2, O.MATERN52),       # every sample is a landmark
    (5, 192, 6, 3, O.RBF),          # cloth-shaped rows, far fewer samples than one 128-sample strip
    (127, 3, 2, 7, O.MATERN52), (128, 3, 2, 7, O.RBF), (129, 3, 2, 7, O.RBF),        # strip edges
    (511, 4, 1, 9, O.RBF), (512, 4, 1, 9, O.MATERN52), (513, 4, 1, 9, O.RBF), (1025, 4, 1, 9, O.RBF),   # chunk edges
])
def test_degenerate_and_edge_sizes(engine, n, d, p, m, kind):
    reg, want, _ = _fit_pair(n, d, p, m, kind, 2.0, 1e-2)
    assert reg.A.shape == (m, m) and reg.B.shape == (m, p) and reg.C.shape == (d, m) and reg.weights.shape == (d, m + p)
    for k, got in (("A", reg.A), ("B", reg.B), ("C", reg.C), ("W", reg.weights)):
        if got.size:
            assert O.relerr(got, want[k]) <= 1e-9, (k, O.relerr(got, want[k]))


def test_coincident_landmarks_are_carried_by_the_jitter(engine):
    """Two identical landmark columns: K_zz is singular, K_mm = K_zz + 1e-6 I is not (regressors.py:139); S, S^-1 and the lift
    still match the reference construction (eigenvalue 1e-6 => the comparison is held to 1e-6 relative)."""
    n, d, p, m = 600, 3, 1, 12
    Xs, U, Y = O.synthetic(n, d, p, seed=3)
    np.random.seed(3)
    Z = O.draw_landmarks(Y, m)
    Z[5] = Z[2]
    reg, want, (X, Y, Z, lsv) = _fit_pair(n, d, p, m, O.RBF, 1.5, 1e-3, seed=3, Z=Z)
    phi = reg.lift(Xs[:9].T)
    assert np.isfinite(phi).all() and O.relerr(phi, O.lift(Z, Xs[:9].T, O.RBF, lsv)) <= 1e-6
    # the one-step predictor W [phi(x); u] is well defined even though the lifted coordinates are not: compare predictions
    Xq = X[:50]
    assert O.relerr(reg.predict(Xq), (want["W"] @ np.vstack((O.lift(Z, Xq[:, :d].T, O.RBF, lsv), Xq[:, d:].T))).T) <= 1e-6


def test_samples_equal_to_landmarks_have_unit_kernel(engine):
    """x == z gives r = 0 inside the fused lift (norm expansion): the kernel value must be 1 to rounding, for both kernels."""
    import torch
    d, m = 5, 40
    rng = np.random.default_rng(0)
    Z = rng.standard_normal((m, d)) * 3.0 + 7.0           # off-centre landmarks: the expansion cancels large norms
    for kind, tol in ((O.RBF, 1e-12), (O.MATERN52, 1e-12)):   # Matern-5/2 is 1 - 5 r^2 / 6 + O(r^3): no linear term in r
        X = np.hstack((Z, np.zeros((m, 1))))
        Zd = torch.from_numpy(Z).cuda()
        il = torch.full((d,), 1.0 / 2.0, dtype=torch.float64).cuda()
        G = engine.grams(torch.from_numpy(X).cuda(), Zd, Zd, il, kind, 1)
        K = O.kernel_matrix(Z, Z, kind, np.full(d, 2.0))
        assert O.relerr(G["Gxx"].cpu().numpy(), K @ K.T) <= tol
        Kzz = engine.kzz(Zd, il, kind).cpu().numpy()
        assert np.array_equal(np.diag(Kzz), np.ones(m))   # direct-difference form: the diagonal is exactly 1


def test_empty_block_inside_a_streamed_accumulation(engine):
    import torch
    n, d, p, m = 700, 6, 2, 20
    Xs, U, Y = O.synthetic(n, d, p, seed=5)
    X = torch.from_numpy(np.hstack((Xs, U))).cuda()
    Yd = torch.from_numpy(Y).cuda()
    np.random.seed(5)
    Zd = torch.from_numpy(O.draw_landmarks(Y, m)).cuda()
    il = torch.full((d,), 0.5, dtype=torch.float64).cuda()
    whole = engine.grams(X, Yd, Zd, il, O.RBF, p)["_flat"].clone()
    engine.gram_begin(Zd, il, O.RBF, p)
    engine.gram_update(X[:300], Yd[:300])
    engine.gram_update(X[300:300], Yd[300:300])          # empty block: a no-op
    engine.gram_update(X[300:], Yd[300:])
    parts = engine.gram_finalize()["_flat"]
    assert O.relerr(parts.cpu().numpy(), whole.cpu().numpy()) <= 1e-13


def test_indefinite_regularised_system_is_retried_with_a_diagonal_shift(engine):
    """Where the reference's lstsq (gelsd, rcond = eps) would truncate a numerically singular inner_term, the Cholesky path retries once
    with a Tikhonov shift of 64 eps max-diag, warns, and records the shift instead of failing the fit."""
    import warnings
    import regressors as R
    rng = np.random.default_rng(0)
    d, p, m, n = 3, 1, 16, 20
    x = rng.standard_normal((1, d + p))
    X = np.repeat(x, n, axis=0)                         # n copies of ONE sample: G_xx has rank 1, and gamma = 0 adds nothing
    Y = np.repeat(rng.standard_normal((1, d)), n, axis=0)
    reg = R.KoopmanNystromRegressor(p, kernel=R.KernelWrapper([1.0] * d), gamma=0.0, m=m)
    reg.nystrom_centers_output = rng.standard_normal((d, m))
    with warnings.catch_warnings(record=True) as caught:
        warnings.simplefilter("always")
        reg.fit(X, Y)
    assert reg.spd_shift_ > 0.0 and any("diagonal shift" in str(w.message) for w in caught)
    assert np.isfinite(reg.A).all() and np.isfinite(reg.B).all() and np.isfinite(reg.C).all()
    # a healthy system is not touched
    Xs, U, Yh = O.synthetic(400, d, p, seed=2)
    reg2 = R.KoopmanNystromRegressor(p, kernel=R.KernelWrapper([1.0] * d), gamma=1e-3, m=m)
    reg2.nystrom_centers_output = reg.nystrom_centers_output
    reg2.fit(np.hstack((Xs, U)), Yh)
    assert reg2.spd_shift_ == 0.0


def test_lift_with_injected_centres_and_no_m(engine):
    """`lift` only needs the centres and the kernel (regressors.py:171-178 never reads self.m): an estimator built with m=None whose
    centres were assigned by hand lifts and, given weights, predicts -- also after a pickle round trip (device cache is not pickled)."""
    import pickle
    import regressors as R
    rng = np.random.default_rng(3)
    d, p, m, N = 4, 2, 11, 37
    Zc = rng.standard_normal((d, m))
    lsv = np.array([0.9, 1.1, 1.3, 0.7])
    reg = R.KoopmanNystromRegressor(p, kernel=R.KernelWrapper(list(lsv)), gamma=1e-3)
    assert reg.m is None
    reg.nystrom_centers_output = Zc
    reg.nystrom_centers_input = Zc
    X = rng.standard_normal((d, N))
    got = reg.lift(X)
    want = O.lift(np.ascontiguousarray(Zc.T), X, O.MATERN52, lsv)
    assert got.shape == (m, N)
    assert O.relerr(got, want) <= 1e-9
    reg.weights = rng.standard_normal((d, m + p))
    Xa = np.hstack((X.T, rng.standard_normal((N, p))))
    pred = reg.predict(Xa)
    assert O.relerr(pred, (reg.weights @ np.vstack((got, Xa[:, d:].T))).T) <= 1e-12
    clone = pickle.loads(pickle.dumps(reg))
    assert O.relerr(clone.predict(Xa), pred) <= 1e-14
