"""Timing of the device LQR gain (nys_koop_lqr_b200/dare.py) at lifted dimensions where the host solver is impractical.
Synthetic model in the family of the fitted ones: spectral radius just above one, p = 6 inputs, Q = C'C with d = 192 outputs."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
from nys_koop_lqr_b200 import dare
from nys_koop_lqr_b200.engine import Engine


def main(ms=(1024, 4096)):
    eng = Engine.get()
    ops = dare.EngineOps(eng)
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    for m in ms:
        p, d = 6, 192
        A = torch.randn(m, m, dtype=torch.float64, device="cuda", generator=g) * (1.002 / np.sqrt(m))
        B = torch.randn(m, p, dtype=torch.float64, device="cuda", generator=g)
        C = torch.randn(d, m, dtype=torch.float64, device="cuda", generator=g) / np.sqrt(m)
        Q = ops.mm(C, C, ta=True)
        R = torch.eye(p, dtype=torch.float64, device="cuda")
        for rep in range(2):                      # first pass warms cuSOLVER and the GEMM descriptors
            torch.cuda.synchronize(); t0 = time.perf_counter()
            P, info = dare.solve_dare(A, B, Q, R, ops=ops)
            K = dare.gain_from_solution(A, B, R, P, ops)
            torch.cuda.synchronize(); t1 = time.perf_counter()
        flops = info["iterations"] * (5 * 2.0 * m ** 3 + (2.0 / 3 + 2 * 2.0) * m ** 3)
        print(json.dumps(dict(op="dlqr", m=m, p=p, seconds=t1 - t0, tflops=flops / (t1 - t0) * 1e-12, **info)), flush=True)


if __name__ == "__main__":
    main(ms=tuple(int(a) for a in sys.argv[1:]) or (1024, 4096))
