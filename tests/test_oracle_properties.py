"""CPU: size-independent properties of the oracle restatement -- the same identities the full-size GPU tests
(tests/test_gpu_fullsize_properties.py) and the sample-sharded fit rely on, checked here on the checker itself.

  * the seven Grams are sums over samples: additive over sample blocks (what one allreduce of per-rank Grams assumes) and
    invariant under a permutation of the samples up to summation order;
  * the fit only sees the Grams: fitting from the summed block Grams equals fitting from all samples;
  * a closed loop with K = 0 is the open-loop rollout with zero controls, and feeding the closed loop's own controls to the
    rollout reproduces its states (benchmark_lqr_cloth.py:18-36 against :80-84);
  * prediction weights are W = C [A | B] (regressors.py:166-169), so predict(x, u) = C (A phi(x) + B u).
"""
import numpy as np
import pytest

from oracle import nk_oracle as O


def _problem(n=600, d=5, p=2, m=24, seed=0, kind=O.RBF):
    Xs, U, Y = O.synthetic(n, d, p, seed=seed)
    np.random.seed(seed)
    Z = O.draw_landmarks(Y, m)
    ls = np.linspace(1.5, 2.5, d)
    return Xs, U, Y, Z, ls, kind


@pytest.mark.parametrize("kind", [O.RBF, O.MATERN52])
def test_grams_are_additive_over_sample_blocks(kind):
    Xs, U, Y, Z, ls, _ = _problem(kind=kind)
    whole = O.grams(Xs, Y, U, Z, kind, ls)
    cuts = [0, 1, 130, 131, 400, Xs.shape[0]]           # ragged blocks, one of a single sample
    parts = [O.grams(Xs[a:b], Y[a:b], U[a:b], Z, kind, ls) for a, b in zip(cuts[:-1], cuts[1:])]
    for k in whole:
        total = sum(part[k] for part in parts)
        assert O.relerr(total, whole[k]) <= 1e-13, k


def test_grams_do_not_depend_on_sample_order_or_chunking():
    Xs, U, Y, Z, ls, kind = _problem(seed=1)
    ref = O.grams(Xs, Y, U, Z, kind, ls, chunk=8192)
    perm = np.random.default_rng(5).permutation(Xs.shape[0])
    shuffled = O.grams(Xs[perm], Y[perm], U[perm], Z, kind, ls, chunk=97)
    threaded = O.grams(Xs, Y, U, Z, kind, ls, chunk=8192, threads=3)
    for k in ref:
        assert O.relerr(shuffled[k], ref[k]) <= 1e-13, k
        assert np.array_equal(threaded[k], ref[k]), k      # threads only split the lift by columns: bit-identical
    # exactly symmetric where the definition is
    assert O.relerr(ref["Gxx"], ref["Gxx"].T) <= 1e-15 and O.relerr(ref["Gyy"], ref["Gyy"].T) <= 1e-15


def test_fit_from_summed_block_grams_equals_fit_from_all_samples():
    Xs, U, Y, Z, ls, kind = _problem(n=900, seed=2)
    gamma = 1e-3
    want = O.fit(np.hstack((Xs, U)), Y, U.shape[1], kind, ls, gamma, Z=Z)
    halves = [O.grams(Xs[a:b], Y[a:b], U[a:b], Z, kind, ls) for a, b in ((0, 333), (333, 900))]
    G = {k: halves[0][k] + halves[1][k] for k in halves[0]}
    Kzz = O.kernel_matrix(Z, Z, kind, ls)
    A, B, C, W = O.solve_abc(G, Kzz, gamma * 900)
    for got, key in ((A, "A"), (B, "B"), (C, "C"), (W, "W")):
        assert O.relerr(got, want[key]) <= 1e-9, key


def test_weights_are_reconstruction_times_dynamics_and_predict_uses_them():
    Xs, U, Y, Z, ls, kind = _problem(seed=3, kind=O.MATERN52)
    p = U.shape[1]
    X_aug = np.hstack((Xs, U))
    f = O.fit(X_aug, Y, p, kind, ls, 1e-2, Z=Z)
    assert O.relerr(f["W"], f["C"] @ np.hstack((f["A"], f["B"]))) <= 1e-12
    phi = O.lift(Z, Xs[:50].T, kind, ls)
    want = (f["C"] @ (f["A"] @ phi + f["B"] @ U[:50].T)).T
    assert O.relerr(O.predict(f["W"], Z, X_aug[:50], p, kind, ls), want) <= 1e-10


def test_closed_loop_and_rollout_are_the_same_recursion():
    rng = np.random.default_rng(7)
    m, p, d, T = 12, 2, 4, 15
    A = rng.standard_normal((m, m)) * 0.2
    B = rng.standard_normal((m, p))
    C = rng.standard_normal((d, m))
    K = rng.standard_normal((p, m)) * 0.1
    z0, zr = rng.standard_normal(m), rng.standard_normal(m)
    # K = 0: the closed loop is the autonomous rollout
    xs0, us0 = O.closed_loop(A, B, C, np.zeros((p, m)), z0, zr, T)
    assert np.all(us0 == 0.0)
    assert O.relerr(xs0, O.rollout(A, B, C, z0, np.zeros((p, T - 1)))) <= 1e-14
    # the closed loop's own controls, replayed open loop, give its states
    xs, us = O.closed_loop(A, B, C, K, z0, zr, T)
    assert O.relerr(xs, O.rollout(A, B, C, z0, us[:, :T - 1])) <= 1e-13
    # the reference point is a fixed point of the controller (u = 0 there)
    _, us_ref = O.closed_loop(A, B, C, K, zr, zr, 1)
    assert np.all(us_ref == 0.0)


def test_kfold_bounds_are_sklearn_unshuffled_kfold():
    from sklearn.model_selection import KFold
    for n, k in ((10, 5), (11, 5), (103, 4), (7, 7)):
        want = [(int(te[0]), int(te[-1]) + 1) for _, te in KFold(k).split(np.zeros((n, 1)))]
        assert [tuple(map(int, b)) for b in O.kfold_bounds(n, k)] == want
