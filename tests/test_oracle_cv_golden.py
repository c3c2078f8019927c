"""CPU: the oracle's hyper-parameter-search restatement against cv_results_ that scikit-learn's GridSearchCV produced
driving the unmodified reference estimator (tests/golden/make_golden_cv.py).  Scores are RMSEs (well conditioned):
tolerance 1e-8 relative on every split score; ranks and best index identical."""
import pathlib

import numpy as np
import pytest

from oracle import nk_oracle as O

CV_GOLDEN = sorted(pathlib.Path(__file__).parent.glob("golden/cv/*.npz"))


def test_kfold_bounds_match_sklearn():
    from sklearn.model_selection import KFold
    for n, k in ((1003, 5), (700, 5), (10, 3), (101, 4)):
        want = [(int(te[0]), int(te[-1]) + 1) for _, te in KFold(k).split(np.zeros((n, 1)))]
        assert O.kfold_bounds(n, k) == want


def test_neg_rmse_matches_sklearn_scorer():
    from sklearn.metrics import root_mean_squared_error
    rng = np.random.default_rng(0)
    a, b = rng.standard_normal((50, 4)), rng.standard_normal((50, 4))
    assert abs(O.neg_rmse(a, b) + root_mean_squared_error(a, b)) <= 1e-15


@pytest.mark.parametrize("path", CV_GOLDEN, ids=[p.stem for p in CV_GOLDEN])
def test_oracle_cv_reproduces_gridsearchcv(path):
    fx = np.load(path)
    kl = [(int(k), l) for k, l in zip(fx["kinds"], fx["ls"])]
    gammas = list(fx["gammas"])
    sc = O.cv_scores(fx["X"], fx["Y"], int(fx["n_inputs"]), kl, gammas, fx["Z"].T, int(fx["n_splits"]))
    got = np.array([sc[k, gammas.index(g)] for k, g in zip(fx["cand_kernel_index"], fx["cand_gamma"])])
    rel = np.abs(got - fx["split_test_score"]) / np.abs(fx["split_test_score"])
    assert rel.max() <= 1e-8, rel.max()
    mean = got.mean(axis=1)
    assert int(np.argmax(mean)) == int(fx["best_index"])
    assert np.array_equal(np.argsort(-mean, kind="stable"), np.argsort(fx["rank_test_score"], kind="stable"))
