"""CPU comparator estimators of the reference (NOT part of the B200 hot path, SURVEY 2.1 "out of scope"):

  * ``KoopmanKernelRegressor``  -- exact n x n kernel estimator (reference regressors.py:58-111), O(n^3);
  * ``KoopmanSplineRegressor``  -- thin-plate-spline EDMD of Korda & Mezic (reference regressors.py:181-234).

The reference's scripts instantiate them next to the Nystrom estimator (benchmark_lqr_classic.py:55,237,276;
benchmark_lqr_cloth.py:44,199,232; the hjb 'kernel' branch), so the drop-in ``regressors`` module has to export working
classes with the same constructor parameters, attributes and array layouts.  They are small numpy/scipy statements of the
published formulas -- plain host code, no GPU involvement, no effect on any number bench.py reports; parity with the
reference's classes is checked in tests/test_oracle_vs_reference.py.
"""
from __future__ import annotations

import numpy as np
import scipy.linalg


def _rows(a):
    return np.ascontiguousarray(np.asarray(a, dtype=np.float64))


class _KernelFitMixin:
    """fit / lift of the exact-kernel estimator (mixed into a KoopmanRegressor subclass by regressors.py)."""

    def _init_kernel_state(self, kernel):
        self.kernel = kernel
        self.training_inputs = None
        self.training_outputs = None
        self.jitter = 1e-6

    def fit(self, X, Y):
        Xc, Yc = _rows(X).T, _rows(Y).T                       # column-sample layout, as inside the reference
        n = Xc.shape[1]
        d = Xc.shape[0] - self.n_inputs
        reg = self.gamma * n
        if self.training_inputs is None:
            self.training_inputs = Xc
        if self.training_outputs is None:
            self.training_outputs = Yc
        tin, tout = self.training_inputs, self.training_outputs
        k = self.kernel.kernel
        states_in, ctrl = tin[:d].T, tin[d:]
        eye = np.eye(n)
        K_in = k(states_in, states_in) + ctrl.T @ ctrl + reg * eye      # state kernel + linear kernel on the controls
        K_out = k(tout.T, tout.T) + self.jitter * eye
        root = scipy.linalg.sqrtm(K_out).real
        root_pinv = scipy.linalg.pinv(root)
        self.Kout, self.Kout_sqrt_inv = K_out, root_pinv
        rhs = np.hstack(((root_pinv @ k(states_in, tout.T).T).T, ctrl.T))
        G = root @ scipy.linalg.solve(K_in, rhs, assume_a="her")
        self.A, self.B = G[:, :n], G[:, n:]
        Phi = scipy.linalg.solve(root, K_out).T
        self.C = tout @ scipy.linalg.solve(Phi @ Phi.T + reg * np.eye(Phi.shape[0]), Phi, assume_a="her")
        self.weights = self.C @ G

    def lift(self, X):
        return self.Kout_sqrt_inv @ self.kernel.kernel(self.training_outputs.T, np.asarray(X).T)


class _SplineFitMixin:
    """fit / lift of the thin-plate-spline EDMD comparator."""

    def _init_spline_state(self, state_bounds_params):
        self.state_bounds_params = state_bounds_params
        self.centers = None

    def compute_centers(self, X):
        if self.state_bounds_params is None:
            pick = np.random.choice(np.arange(0, X.shape[1]), size=self.m, replace=False)
            return X[:, pick]
        # polar sampling of a disc (two state dimensions), same two uniform draws and order as the reference
        radius = np.sqrt(np.random.uniform(0, self.state_bounds_params[0], size=(1, self.m)))
        theta = np.pi * np.random.uniform(0, self.state_bounds_params[1], size=(1, self.m))
        return np.vstack((radius * np.cos(theta), radius * np.sin(theta)))

    def lift(self, X):
        X = np.asarray(X, dtype=np.float64)
        if self.centers is None:
            self.centers = self.compute_centers(X)
        out = np.empty((self.m, X.shape[1]))
        for i in range(self.centers.shape[1]):                 # r^2 log r per centre, accumulated like the reference (axis-0 sum)
            r2 = np.sum(np.square(X - self.centers[:, i:i + 1]), axis=0)
            with np.errstate(divide="ignore", invalid="ignore"):
                out[i] = r2 * np.log(np.sqrt(r2))
        return np.nan_to_num(out, nan=0.0)

    def fit(self, X, Y):
        Xc, Yc = _rows(X).T, _rows(Y).T
        d = Xc.shape[0] - self.n_inputs
        reg = self.gamma * Xc.shape[1]
        feats_now = np.vstack((self.lift(Xc[:d]), Xc[d:]))        # [psi(x); u]
        feats_next = np.vstack((self.lift(Yc), Xc[:d]))           # [psi(x+); x]  (the last rows learn the reconstruction)
        cov = feats_now @ feats_now.T
        M = (feats_next @ feats_now.T) @ scipy.linalg.pinv(cov + reg * np.eye(cov.shape[0]))
        m = self.m
        self.A, self.B, self.C = M[:m, :m], M[:m, m:], M[m:, :m]
        self.weights = self.C @ M[:m, :]
