"""Generates tests/golden/cv/*.npz: the reference's hyper-parameter search, run for real in this container.

`learn_hyperparams` of the three benchmark scripts (benchmark_lqr_cloth.py:39-66, _classic.py:44-64, _hjb.py:47-71) is
``GridSearchCV(KoopmanNystromRegressor(...), {'kernel': [...], 'gamma': [...]}, scoring='neg_root_mean_squared_error')``.
Here scikit-learn's own GridSearchCV drives the UNMODIFIED reference estimator (imported from /root/reference); the only
intervention is that every clone gets the same injected landmark set (the reference's supported way to fix landmarks:
the `None` checks at regressors.py:129,133), because a clone otherwise redraws them from the unseeded global RNG and the
search would not be reproducible.  The fixture stores inputs, landmarks, grid, and sklearn's cv_results_.

Run:  python tests/golden/make_golden_cv.py      (needs /root/reference; the fixtures are committed)
"""
import pathlib
import sys

import numpy as np

REF = pathlib.Path("/root/reference")
HERE = pathlib.Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parents[1]))
sys.path.insert(0, str(REF))
import regressors as ref                     # noqa: E402  the reference itself
from oracle import nk_oracle as O            # noqa: E402
from sklearn.model_selection import GridSearchCV   # noqa: E402
assert pathlib.Path(ref.__file__).parent == REF


class InjectedLandmarks(ref.KoopmanNystromRegressor):
    """The reference estimator with a fixed landmark set pushed in before every fit (clones included)."""
    Z = None

    def fit(self, X, Y):
        self.nystrom_centers_output = InjectedLandmarks.Z
        self.nystrom_centers_input = InjectedLandmarks.Z
        return super().fit(X, Y)


def run(name, X, Y, n_inputs, holders, kinds_ls, gammas, m, seed, n_splits=5):
    n = X.shape[0]
    np.random.seed(seed)
    idx = np.random.choice(np.arange(0, n), size=m, replace=False)
    InjectedLandmarks.Z = Y.T[:, idx].copy()
    est = InjectedLandmarks(n_inputs, kernel=holders[0], gamma=gammas[0], m=m)
    clf = GridSearchCV(est, {"kernel": holders, "gamma": list(gammas)}, scoring="neg_root_mean_squared_error", cv=n_splits, n_jobs=1, refit=True)
    clf.fit(X, Y)
    res = clf.cv_results_
    kidx = np.array([holders.index(pp["kernel"]) for pp in res["params"]])
    gam = np.array([pp["gamma"] for pp in res["params"]])
    split = np.stack([res[f"split{k}_test_score"] for k in range(n_splits)], axis=1)       # (candidates, folds)
    best = clf.best_estimator_
    out = dict(X=X, Y=Y, n_inputs=n_inputs, m=m, seed=seed, Z=InjectedLandmarks.Z, n_splits=n_splits,
               kinds=np.array([k for k, _ in kinds_ls]), ls=np.stack([np.asarray(l, dtype=float) for _, l in kinds_ls]), gammas=np.asarray(gammas, dtype=float),
               cand_kernel_index=kidx, cand_gamma=gam, split_test_score=split, mean_test_score=res["mean_test_score"],
               std_test_score=res["std_test_score"], rank_test_score=res["rank_test_score"], best_index=clf.best_index_,
               best_A=best.A, best_B=best.B, best_C=best.C, best_W=best.weights)
    (HERE / "cv").mkdir(exist_ok=True)
    np.savez_compressed(HERE / "cv" / f"{name}.npz", **out)
    print(name, "candidates", len(gam), "best", clf.best_index_, res["params"][clf.best_index_]["gamma"], "score", clf.best_score_)
    print("  mean scores", np.array2string(res["mean_test_score"], precision=6))


def main():
    Xs, U, Y = O.synthetic(1003, d=12, p=2, seed=11)        # 1003: folds of unequal size (201,201,201,200,200)
    X = np.hstack((Xs, U))
    trip = [(3.0, 4.0, 5.0), (6.0, 6.0, 6.0), (2.0, 8.0, 4.0)]
    holders = [ref.ThreeDimensionalKernel(a, b, c, 12) for a, b, c in trip]
    run("synthetic_rbf_cv", X, Y, 2, holders, [(O.RBF, np.resize(t, 12)) for t in trip], [1e-2, 1e-3, 1e-4, 1e-5], 48, 0)
    # Matern-5/2, d=2, p=1 (Duffing-like shapes, benchmark_lqr_classic.py:47-50 grid style)
    Xs, U, Y = O.synthetic(700, d=2, p=1, seed=5)
    X = np.hstack((Xs, U))
    lss = [[1.0, 1.0], [0.5, 2.0]]
    holders = [ref.KernelWrapper(l) for l in lss]
    run("synthetic_matern_cv", X, Y, 1, holders, [(O.MATERN52, np.asarray(l)) for l in lss], [1e-2, 1e-4, 1e-6], 15, 0)


if __name__ == "__main__":
    main()
