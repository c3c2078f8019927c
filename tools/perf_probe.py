"""Quick throughput probe of the fused lift+Gram kernel (synthetic, device-generated data)."""
import sys, time, json
import torch
sys.path.insert(0, ".")
from nys_koop_lqr_b200.engine import Engine

def main(n=400000, m=4096, d=192, p=6, chunk=512, reps=2):
    eng = Engine.get()
    g = torch.Generator(device="cuda"); g.manual_seed(0)
    Xa = torch.randn(n, d + p, dtype=torch.float64, device="cuda", generator=g)
    Mx = torch.randn(d, d, dtype=torch.float64, device="cuda", generator=g) * 0.9 / d ** 0.5
    Y = torch.tanh(Xa[:, :d] @ Mx.T)
    Z = Y[torch.randperm(n, device="cuda", generator=g)[:m]].contiguous()
    il = torch.full((d,), 0.1, dtype=torch.float64, device="cuda")
    for r in range(reps + 1):
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record()
        eng.gram_begin(Z, il, 0, p, chunk)
        eng.gram_update(Xa, Y)
        G = eng.gram_finalize()
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        F = 4.0 * m * m + 6.0 * m * d + 4.0 * m * p
        print(json.dumps(dict(n=n, m=m, d=d, chunk=chunk, ms=ms, samples_per_s=n / ms * 1e3, algo_tflops=F * n / ms * 1e-9,
                              exec_tflops=eng.gram_executed_flops() / ms * 1e-9, frac_of_37_1=F * n / ms * 1e-9 / 37.1)))
    # a perf number of a kernel that computes the wrong thing is worthless: check a small pass against an independent device path
    ns = 2048
    Gs = eng.grams(Xa[:ns], Y[:ns], Z, il, 0, p, chunk)
    Kx = eng.kernel_cross(Z, Xa[:ns, :d].contiguous(), il, 0)
    Ky = eng.kernel_cross(Z, Y[:ns].contiguous(), il, 0)
    U = Xa[:ns, d:]
    rel = lambda a, b: float((a - b).norm() / b.norm())
    errs = dict(Gxx=rel(Gs["Gxx"], Kx @ Kx.T), Gyx=rel(Gs["Gyx"], Ky @ Kx.T), Gyy=rel(Gs["Gyy"], Ky @ Ky.T), Gxu=rel(Gs["Gxu"], Kx @ U),
                Guu=rel(Gs["Guu"], U.T @ U), GYy=rel(Gs["GYy"], Y[:ns].T @ Ky.T))
    ok = max(errs.values()) <= 1e-11
    print(json.dumps(dict(check="fused Grams of the first %d samples vs kernel_cross + torch matmul" % ns, ok=ok, max_rel_err=max(errs.values()))))
    if not ok:
        raise SystemExit("perf_probe: WRONG RESULTS " + json.dumps(errs))
    return G

if __name__ == "__main__":
    args = [int(a) for a in sys.argv[1:]]
    main(*args)
