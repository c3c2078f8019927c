"""Timing probe of the batched hyper-parameter search (fit_cv) on synthetic data generated on the device."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import regressors as R


def main(n=200000, m=4096, d=192, p=6, nk=2, ng=16):
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(0)
    X = torch.randn(n, d + p, dtype=torch.float64, device=dev, generator=g)
    M = torch.randn(d, d, dtype=torch.float64, device=dev, generator=g) * (0.9 / d ** 0.5)
    Bu = 0.1 * torch.randn(d, p, dtype=torch.float64, device=dev, generator=g)
    Y = torch.tanh(X[:, :d] @ M.T) + X[:, d:] @ Bu.T
    np.random.seed(0)
    kernels = [R.ThreeDimensionalKernel(l, l, l, d) for l in np.logspace(0.9, 1.3, nk)]
    gammas = list(10.0 ** np.arange(-6.0, -6.0 + 0.25 * ng, 0.25))[:ng]
    reg = R.KoopmanNystromRegressor(p, kernel=kernels[0], gamma=gammas[0], m=m)
    reg.cv_profile = True
    for rep in range(2):
        torch.cuda.synchronize(); t0 = time.perf_counter()
        res = reg.fit_cv(X, Y, kernels, gammas, n_splits=5, refit=False)
        torch.cuda.synchronize(); dt = time.perf_counter() - t0
        print(json.dumps(dict(n=n, m=m, kernels=nk, gammas=ng, total_s=dt, profile=reg.cv_profile_, best=reg.best_index_,
                              best_score=reg.best_score_, nan=int(np.isnan(res["mean_test_score"]).sum()))), flush=True)


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
