"""Where a SMALL fit spends its time (the three script configurations: n = 3 000 ... 70 000, m = 10 ... 100).
    python tools/small_fit_probe.py            # on the GPU box
Prints, per shape: wall time of estimator.fit (host arrays in, numpy out) and of its phases with a device sync between them."""
import sys, time, pathlib
import numpy as np
import torch
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parents[1]))
import regressors as R
from nys_koop_lqr_b200.engine import Engine

eng = Engine.get(0)
sync = lambda: torch.cuda.synchronize()
SHAPES = [("duffing", 69900, 2, 1, 20, "matern", 1.0, 1e-6), ("duffing-m200", 69900, 2, 1, 200, "matern", 1.0, 1e-6),
          ("cloth", 3030, 192, 6, 100, "rbf", 10.0, 1e-7), ("hjb", 3980, 1, 1, 100, "matern", 1.0, 1e-3)]
for name, n, d, p, m, kern, ls, gamma in SHAPES:
    rng = np.random.default_rng(0)
    X = rng.standard_normal((n, d + p)); Y = np.tanh(X[:, :d]) + 0.1 * rng.standard_normal((n, d))
    holder = R.KernelWrapper([ls] * d) if kern == "matern" else R.ThreeDimensionalKernel(ls, ls, ls, d)
    Z = np.ascontiguousarray(Y[rng.choice(n, m, replace=False)].T)
    for chunk in (0, 128, 256, 1024, 2048):
        def fit():
            reg = R.KoopmanNystromRegressor(p, kernel=holder, gamma=gamma, m=m)
            reg.gram_chunk = chunk
            reg.nystrom_centers_output = Z.copy()
            reg.fit(X, Y)
            return reg
        for _ in range(3):
            fit()
        sync(); t0 = time.perf_counter()
        for _ in range(10):
            fit()
        sync(); wall = (time.perf_counter() - t0) / 10
        # phases
        reg = R.KoopmanNystromRegressor(p, kernel=holder, gamma=gamma, m=m); reg.gram_chunk = chunk; reg.nystrom_centers_output = Z.copy()
        t = [time.perf_counter()]
        dev = reg._device_state(d); sync(); t.append(time.perf_counter())
        Xd = torch.from_numpy(X).cuda(); Yd = torch.from_numpy(Y).cuda(); sync(); t.append(time.perf_counter())
        eng.gram_begin(dev["Z"], dev["inv_ls"], dev["kind"], p, chunk); sync(); t.append(time.perf_counter())
        eng.gram_update(Xd, Yd); sync(); t.append(time.perf_counter())
        G = eng.gram_finalize(); sync(); t.append(time.perf_counter())
        reg._solve(eng, dev, G, n, d); sync(); t.append(time.perf_counter())
        ph = [1e3 * (b - a) for a, b in zip(t[:-1], t[1:])]
        print(f"{name:13s} n={n} m={m} chunk={chunk or 512:5d}: fit {1e3*wall:7.2f} ms | landmark stage {ph[0]:.2f} upload {ph[1]:.2f} begin {ph[2]:.2f} "
              f"gram kernel {ph[3]:.2f} finalize {ph[4]:.2f} solve+download {ph[5]:.2f}", flush=True)
