"""Timing probe of the batched rollout (packed persistent GEMM, one launch per step) on random stable dynamics."""
import json
import sys

import torch

sys.path.insert(0, ".")
from nys_koop_lqr_b200.engine import Engine


def main(nb=100000, m=4096, steps=4, d=192, p=6):
    eng = Engine.get()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    A = torch.randn(m, m, dtype=torch.float64, device="cuda", generator=g) * (0.9 / m ** 0.5)
    B = torch.randn(m, p, dtype=torch.float64, device="cuda", generator=g)
    C = torch.randn(d, m, dtype=torch.float64, device="cuda", generator=g)
    Z0 = torch.randn(nb, m, dtype=torch.float64, device="cuda", generator=g)
    T = steps + 1
    U = torch.randn(T - 1, nb, p, dtype=torch.float64, device="cuda", generator=g)
    Yt = torch.randn(T, nb, d, dtype=torch.float64, device="cuda", generator=g)
    for rep in range(2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); eng.rollout(A, B, C, Z0, U, Ytrue=Yt, return_traj=False); e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        fl = nb * ((T - 1) * (2.0 * m * m + 2.0 * m * p) + T * 2.0 * d * m)
        print(json.dumps(dict(nb=nb, m=m, T=T, ms=ms, ms_per_step=ms / (T - 1), tflops=fl / ms * 1e-9)), flush=True)


if __name__ == "__main__":
    main(*[int(a) for a in sys.argv[1:]])
