"""Timing probe of the n-independent dense stage and the rollout (CUDA events, synthetic SPD inputs)."""
import json
import sys

import torch

sys.path.insert(0, ".")
from nys_koop_lqr_b200.engine import Engine


def timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main(ms=(4096, 8192), nb=100000, steps=4):
    eng = Engine.get()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for m in ms:
        d, p = 192, 6
        Z = torch.randn(m, d, dtype=torch.float64, device="cuda", generator=g) * 0.5
        il = torch.full((d,), 0.1, dtype=torch.float64, device="cuda")
        A = torch.randn(m, m, dtype=torch.float64, device="cuda", generator=g)
        B = torch.randn(m, m, dtype=torch.float64, device="cuda", generator=g)
        out = torch.empty(m, m, dtype=torch.float64, device="cuda")
        t = timed(lambda: eng.gemm(A, B, transb=True, out=out))
        print(json.dumps(dict(op="gemm_nt", m=m, ms=t, tflops=2.0 * m ** 3 / t * 1e-9)), flush=True)
        t = timed(lambda: eng.gemm(A, B, out=out))
        print(json.dumps(dict(op="gemm_nn", m=m, ms=t, tflops=2.0 * m ** 3 / t * 1e-9)), flush=True)
        t = timed(lambda: torch.matmul(A, B.T, out=out))
        print(json.dumps(dict(op="cublas_dgemm_nt", m=m, ms=t, tflops=2.0 * m ** 3 / t * 1e-9)), flush=True)
        Kzz = eng.kzz(Z, il, 0)
        Kmm = Kzz.clone(); Kmm.diagonal().add_(1e-6)
        t = timed(lambda: eng.kzz(Z, il, 0))
        print(json.dumps(dict(op="kzz", m=m, ms=t)), flush=True)
        t = timed(lambda: eng.potrf(Kmm.clone()))
        print(json.dumps(dict(op="potrf(+clone)", m=m, ms=t, tflops=m ** 3 / 3.0 / t * 1e-9)), flush=True)
        t = timed(lambda: torch.linalg.cholesky(Kmm))
        print(json.dumps(dict(op="torch_cholesky", m=m, ms=t, tflops=m ** 3 / 3.0 / t * 1e-9)), flush=True)
        t = timed(lambda: eng.sym_sqrt(Kmm), reps=1)
        print(json.dumps(dict(op="sym_sqrt", m=m, ms=t, iters=eng.last_sqrt_iters)), flush=True)
        S, Sinv = eng.sym_sqrt(Kmm)
        # synthetic Grams: Phi = random features through the same kernel
        n = 4 * m
        X = torch.randn(n, d + p, dtype=torch.float64, device="cuda", generator=g) * 0.5
        Y = torch.randn(n, d, dtype=torch.float64, device="cuda", generator=g) * 0.5
        G = eng.grams(X, Y, Z, il, 0, p)
        t = timed(lambda: eng.solve_abc(G, Kzz, S, Sinv, 1e-4 * n), reps=1)
        print(json.dumps(dict(op="solve_abc", m=m, ms=t)), flush=True)
        Am, Bm, Cm, W = eng.solve_abc(G, Kzz, S, Sinv, 1e-4 * n)
        # rollout
        T = steps + 1
        Z0 = torch.randn(nb, m, dtype=torch.float64, device="cuda", generator=g) * 0.01
        U = torch.randn(T - 1, nb, p, dtype=torch.float64, device="cuda", generator=g)
        Yt = torch.randn(T, nb, d, dtype=torch.float64, device="cuda", generator=g)
        t = timed(lambda: eng.rollout(Am, Bm, Cm, Z0, U, Ytrue=Yt, return_traj=False), reps=1)
        fl = nb * ((T - 1) * (2.0 * m * m + 2.0 * m * p) + T * 2.0 * d * m)
        print(json.dumps(dict(op="rollout", m=m, nb=nb, T=T, ms=t, ms_per_step=t / (T - 1), tflops=fl / t * 1e-9)), flush=True)
        Xp = torch.randn(nb, d + p, dtype=torch.float64, device="cuda", generator=g) * 0.5
        t = timed(lambda: eng.predict(Z, il, 0, Sinv, W, Xp, p), reps=1)
        print(json.dumps(dict(op="predict", m=m, N=nb, ms=t, tflops=nb * (2.0 * m * (d + 2) + 2.0 * m * m + 2.0 * d * (m + p)) / t * 1e-9)), flush=True)
        del A, B, out, Z0, U, Yt, Xp, G
        torch.cuda.empty_cache()


if __name__ == "__main__":
    main(ms=tuple(int(a) for a in sys.argv[1:]) or (4096, 8192))
