"""GPU: the batched hyper-parameter search (fit_cv -> nk_gram_*, nk_axpy, nk_cv_weights, nk_cv_score through the C ABI)
against (1) cv_results_ of scikit-learn's GridSearchCV driving the unmodified reference (tests/golden/cv/*.npz) and
(2) the CPU oracle on a seeded case that is not a tile multiple anywhere.  Split scores: 1e-8 relative."""
import pathlib

import numpy as np
import pytest
import torch

from oracle import nk_oracle as O

pytestmark = pytest.mark.gpu
CV_GOLDEN = sorted(pathlib.Path(__file__).parent.glob("golden/cv/*.npz"))


def holders_for(fx):
    import regressors as R
    out = []
    for kind, ls in zip(fx["kinds"], fx["ls"]):
        if int(kind) == O.RBF:
            h = R.ThreeDimensionalKernel(1, 1, 1, ls.size)
            h.kernel.length_scale = ls.reshape(1, -1)
        else:
            h = R.KernelWrapper(list(ls))
        out.append(h)
    return out


@pytest.mark.parametrize("path", CV_GOLDEN, ids=[p.stem for p in CV_GOLDEN])
def test_fit_cv_matches_gridsearchcv_golden(engine, path):
    import regressors as R
    fx = np.load(path)
    holders = holders_for(fx)
    reg = R.KoopmanNystromRegressor(int(fx["n_inputs"]), kernel=holders[0], gamma=float(fx["gammas"][0]), m=int(fx["m"]))
    reg.nystrom_centers_output = fx["Z"].copy()
    res = reg.fit_cv(fx["X"], fx["Y"], holders, list(fx["gammas"]), n_splits=int(fx["n_splits"]))
    # same candidate order as sklearn's ParameterGrid
    assert [holders.index(pp["kernel"]) for pp in res["params"]] == list(fx["cand_kernel_index"])
    assert np.array_equal(res["param_gamma"], fx["cand_gamma"])
    split = np.stack([res[f"split{k}_test_score"] for k in range(int(fx["n_splits"]))], axis=1)
    rel = np.abs(split - fx["split_test_score"]) / np.abs(fx["split_test_score"])
    assert rel.max() <= 1e-8, f"split scores differ by {rel.max():.2e}"
    assert np.abs(res["mean_test_score"] - fx["mean_test_score"]).max() <= 1e-8 * np.abs(fx["mean_test_score"]).max()
    assert np.abs(res["std_test_score"] - fx["std_test_score"]).max() <= 1e-8
    assert np.array_equal(res["rank_test_score"], fx["rank_test_score"])
    assert reg.best_index_ == int(fx["best_index"])
    # refit on all samples with the winner == GridSearchCV(refit=True).best_estimator_
    for key, got in (("best_A", reg.A), ("best_B", reg.B), ("best_C", reg.C), ("best_W", reg.weights)):
        assert O.relerr(got, fx[key]) <= 1e-7, (key, O.relerr(got, fx[key]))


def test_cv_weights_and_score_against_oracle(engine):
    """One (kernel, fold): prediction weights of every gamma (batched factorisation) reproduce the oracle's predict."""
    rng = np.random.default_rng(3)
    n, d, p, m = 777, 7, 3, 131
    Xs, U, Y = O.synthetic(n, d, p, seed=9)
    X = np.hstack((Xs, U))
    Z = Y[rng.choice(n, m, replace=False)]
    ls = np.linspace(2.0, 4.0, d)
    gammas = [1e-1, 1e-3, 1e-4]
    s, e = 300, 455
    tr = np.r_[0:s, e:n]
    dev = lambda a: torch.as_tensor(np.ascontiguousarray(a), dtype=torch.float64).cuda()
    Zd, il = dev(Z), dev(1.0 / ls)
    G = engine.grams(dev(X[tr]), dev(Y[tr]), Zd, il, O.RBF, p)
    Kzz = engine.kzz(Zd, il, O.RBF)
    Wk, info = engine.cv_weights(G, Kzz, [g * len(tr) for g in gammas])
    assert info == [0, 0, 0]
    sse = engine.cv_score(Zd, il, O.RBF, Wk, dev(X[s:e]), dev(Y[s:e]), p).cpu().numpy()
    Kval = O.kernel_matrix(Z, Xs[s:e], O.RBF, ls)
    for b, g in enumerate(gammas):
        f = O.fit(X[tr], Y[tr], p, O.RBF, ls, g, Z=Z)
        want = O.predict(f["W"], Z, X[s:e], p, O.RBF, ls)
        got = (Wk[b].cpu().numpy() @ np.vstack((Kval, U[s:e].T))).T
        # 1e-9 where the regularised system is well conditioned (north-star bar); beyond that both solves (oracle and GPU,
        # different but equivalent factorisation orders) sit on the cond * eps floor of float64
        Kmm = f["Kzz"] + 1e-6 * np.eye(m)
        cond = np.linalg.cond(f["G"]["Gxx"] + g * len(tr) * Kmm)
        tol = max(1e-9, 1e-15 * cond)
        assert O.relerr(got, want) <= tol, (g, cond, O.relerr(got, want))
        assert np.abs(sse[b] - np.sum((want - Y[s:e]) ** 2, axis=0)).max() <= 10 * tol * np.sum((want - Y[s:e]) ** 2)


def test_cv_not_spd_candidate_scores_nan(engine):
    """A candidate whose regularised system is not positive definite gets a nan score (sklearn error_score=nan), the
    others are unaffected."""
    import regressors as R
    Xs, U, Y = O.synthetic(300, 3, 1, seed=2)
    X = np.hstack((Xs, U))
    np.random.seed(0)
    reg = R.KoopmanNystromRegressor(1, kernel=R.KernelWrapper([1.0] * 3), gamma=1e-3, m=20)
    res = reg.fit_cv(X, Y, [R.KernelWrapper([1.0] * 3)], [-10.0, 1e-3], n_splits=3, refit=False)
    assert np.isnan(res["mean_test_score"][0]) and np.isfinite(res["mean_test_score"][1])
    assert reg.best_index_ == 1
