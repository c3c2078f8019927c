import sys, numpy as np, torch
sys.path.insert(0, ".")
from nys_koop_lqr_b200.engine import Engine
from oracle import nk_oracle as O
eng = Engine.get()
def run(n, d, p, m, kind, ls, chunk, seed=0):
    rng = np.random.default_rng(seed)
    Xs = rng.standard_normal((n, d)); U = rng.standard_normal((n, p)); Y = np.tanh(Xs) + 0.1
    Z = Y[rng.choice(n, m, replace=False)]
    lsv = np.full(d, ls)
    ref = O.grams(Xs, Y, U, Z, kind, lsv)
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a)).cuda()
    G = eng.grams(t(np.hstack((Xs, U))), t(Y), t(Z), t(1.0 / lsv), kind, p, chunk)
    torch.cuda.synchronize()
    print(f"--- n={n} d={d} p={p} m={m} kind={kind} chunk={chunk}")
    for k in ("Gxx", "Gyx", "Gyy", "Gxu", "Gyu", "Guu", "GYy"):
        if ref[k].size == 0: continue
        g = G[k].cpu().numpy(); e = np.abs(g - ref[k])
        bad = np.argwhere(e > 1e-9 * np.abs(ref[k]).max())
        print(k, "relerr %.3e" % O.relerr(g, ref[k]), "nbad", len(bad), "of", e.size,
              ("rows %d..%d cols %d..%d" % (bad[:,0].min(), bad[:,0].max(), bad[:,1].min(), bad[:,1].max())) if len(bad) else "")
        if len(bad) and k == "Gxx":
            r, c = bad[0]; print("   first bad", r, c, g[r, c], ref[k][r, c], "ratio", g[r,c]/ref[k][r,c])
            rows = np.unique(bad[:,0]); cols = np.unique(bad[:,1]); print("   bad rows", rows[:20], "bad cols", cols[:20])
run(128, 14, 2, 128, 0, 2.0, 128)
run(128, 14, 2, 100, 0, 2.0, 128)
run(100, 14, 2, 100, 0, 2.0, 128)
run(256, 14, 2, 100, 0, 2.0, 128)
run(256, 14, 2, 100, 0, 2.0, 256)
run(808, 192, 6, 100, 0, 10.0, 0)
