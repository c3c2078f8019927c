"""numpy prototype: symmetric sqrt via Cholesky + scaled Newton-Schulz polar (what the GPU dense stage runs)."""
import sys, numpy as np, scipy.linalg
sys.path.insert(0, "/root/repo")
from oracle import nk_oracle as O

def opt_cubic(l):
    # p(x) = a x - b x^3 equioscillating on [l,1]; returns a,b,new_l (after rescale to max 1)
    s = 1 + l + l*l
    xs = np.sqrt(s/3)
    # a = b*s ; p(xs) = (2a/3) xs ; p(1) = b (l + l^2); p(xs)+p(1)=2
    b = 2.0/((2*s/3)*xs + (l + l*l))
    a = b*s
    pmax = (2*a/3)*xs; pmin = a - b
    return a/pmax, b/pmax, pmin/pmax

def polar_sqrt(K, lam_min, iters_max=60, tol=1e-15):
    m = K.shape[0]
    L = np.linalg.cholesky(K); R = L.T
    nrm = np.sqrt(np.max(np.sum(np.abs(K), axis=1)))
    X = R/nrm
    l = np.sqrt(lam_min)/nrm*0.9
    its = 0
    while its < iters_max:
        M = X.T @ X
        err = np.linalg.norm(M - np.eye(m), 'fro')
        if l > 0.999 :
            a, b = 1.5, 0.5
        else:
            a, b, l = opt_cubic(l)
        X = X @ (a*np.eye(m) - b*M)
        its += 1
        if err < 1e-7 and l > 0.999:   # one more plain NS step squares the error
            M = X.T @ X; X = X @ (1.5*np.eye(m) - 0.5*M); its += 1
            break
    Q = X
    S = Q.T @ R; S = 0.5*(S+S.T)
    Sinv = scipy.linalg.solve_triangular(R, Q, lower=False)
    Sinv = 0.5*(Sinv+Sinv.T)
    return S, Sinv, its

def check(K, name):
    w, V = np.linalg.eigh(K)
    S0 = (V*np.sqrt(w))@V.T; Si0 = (V/np.sqrt(w))@V.T
    S, Sinv, its = polar_sqrt(K, 1e-6)
    Sq = scipy.linalg.sqrtm(K).real
    print(f"{name}: m={K.shape[0]} cond={w[-1]/w[0]:.2e} its={its} S:{O.relerr(S,S0):.2e} Sinv:{O.relerr(Sinv,Si0):.2e} "
          f"sqrtm-vs-eigh:{O.relerr(Sq,S0):.2e} resid:{O.relerr(S@S,K):.2e} SinvS-I:{np.linalg.norm(Sinv@S-np.eye(K.shape[0])):.2e}")

rng = np.random.default_rng(0)
Xs,U,Y = O.synthetic(20000)
np.random.seed(0)
for m in (128, 512, 1024):
    Z = Y[np.random.choice(20000, m, False)]
    K = O.kernel_matrix(Z,Z,O.RBF,10.0) + 1e-6*np.eye(m)
    check(K, "synthetic rbf l=10")
Z = Y[np.random.choice(20000, 512, False)]
K = O.kernel_matrix(Z,Z,O.RBF,30.0) + 1e-6*np.eye(512); check(K, "synthetic rbf l=30")
# cloth
import pathlib
p = pathlib.Path("/root/reference/8x8_cloth_swing_xyz")
trajs = [np.loadtxt(p/f"state_samples_cloth_swing_{i}.csv", delimiter=',') for i in range(10, 16)]
Yc = np.vstack([t[1:] for t in trajs])
np.random.seed(0)
for m in (100, 398):
    Z = Yc[np.random.choice(Yc.shape[0], m, False)]
    K = O.kernel_matrix(Z,Z,O.RBF,10.0) + 1e-6*np.eye(m); check(K, "cloth rbf l=10")
# duffing matern
Yd = np.loadtxt("/root/reference/duffing/duffing_y_forced.csv", delimiter=',').T
for m in (20, 200, 500):
    Z = Yd[np.random.choice(Yd.shape[0], m, False)]
    K = O.kernel_matrix(Z,Z,O.MATERN52,[1.0,1.0]) + 1e-6*np.eye(m); check(K, "duffing matern")
