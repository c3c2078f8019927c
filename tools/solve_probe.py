"""Host-side timing of the n-independent tail of a fit (solve, result download) -- what does not scale with GPUs."""
import sys, time, json
import numpy as np, torch
sys.path.insert(0, ".")
import regressors as R

def main(m=4096, d=192, p=6, n=60000):
    dev = torch.device("cuda")
    g = torch.Generator(device=dev); g.manual_seed(0)
    X = torch.randn(n, d + p, dtype=torch.float64, device=dev, generator=g)
    Y = torch.tanh(X[:, :d] * 0.5)
    np.random.seed(0)
    for rep in range(3):
        reg = R.KoopmanNystromRegressor(p, kernel=R.ThreeDimensionalKernel(10, 10, 10, d), gamma=1e-4, m=m)
        reg.nystrom_centers_output = np.ascontiguousarray(Y[:m].cpu().numpy().T)
        t0 = time.perf_counter(); dv = reg._device_state(d); torch.cuda.synchronize(); t1 = time.perf_counter()
        eng = dv["eng"]
        reg._accumulate_grams(eng, dv, X, Y); G = eng.gram_finalize(); torch.cuda.synchronize(); t2 = time.perf_counter()
        reg._solve(eng, dv, G, n, d); t3 = time.perf_counter()
        print(json.dumps(dict(landmark_stage_ms=(t1 - t0) * 1e3, gram_ms=(t2 - t1) * 1e3, solve_and_download_ms=(t3 - t2) * 1e3)), flush=True)

if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
