"""Simulators of the reference's closed-loop experiments  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The hjb / classic scripts close the loop on the TRUE system (benchmark_lqr_hjb.py:74-97, benchmark_lqr_classic.py:67-89):
per step one `lift` of the current state, u = K (phi_ref - phi), then the simulator's `update_SOM`.  The simulators
(reference dynamical_systems.py) are outside the hot path and are not part of the product; tests need them on the GPU box,
where /root/reference does not exist, so the two right-hand sides and the reference's Runge-Kutta step are restated here.

NOTE the reference's step is not the textbook RK4: its fourth stage re-uses k1 (`_k4(x,u) = f(x + k1*Ts, u)`,
dynamical_systems.py:42-43, 70-71, 101-102) -- restated as is, because the golden closed-loop trajectories were made with it.
tests/test_oracle_vs_reference.py checks these functions against the reference classes.
"""
from __future__ import annotations

import numpy as np


def duffing_rhs(x, u):
    """dynamical_systems.py:26-28 (DuffingOscillator._f_u): x (2, N)."""
    return -np.vstack((-x[1, :], 0.5 * x[1, :] + x[0, :] * (4 * x[0, :] ** 2 - 1) - 0.5 * u))


def hjb_rhs(x, u):
    """dynamical_systems.py:88-89 (HJB._f_u)."""
    return -x ** 3 + u


def reference_rk_step(rhs, x, u, Ts):
    """dynamical_systems.py:30-43: k2, k3 as in RK4; k4 evaluated at x + k1*Ts (the reference's own quirk)."""
    k1 = rhs(x, u)
    k2 = rhs(x + k1 * Ts / 2, u)
    k3 = rhs(x + k2 * Ts / 2, u)
    k4 = rhs(x + k1 * Ts, u)
    return x + (Ts / 6) * (k1 + 2 * k2 + 2 * k3 + k4)


def duffing_step(x, u, Ts=0.01):
    x = np.asarray(x, dtype=np.float64)
    if x.ndim == 1:
        x = x.reshape(-1, 1)
    return reference_rk_step(duffing_rhs, x, u, Ts)


def hjb_step(x, u, Ts=0.01):
    return reference_rk_step(hjb_rhs, np.asarray(x, dtype=np.float64), u, Ts)


def lqr_control_on_system(lift, C, K, step, initial_state, reference, num_steps):
    """The loop of benchmark_lqr_hjb.py:74-97 / benchmark_lqr_classic.py:67-89 for any `lift` callable: returns the TRUE states
    visited (d, num_steps+1), the reconstructions C phi (d, num_steps) and the controls (p, num_steps)."""
    x = np.asarray(initial_state, dtype=np.float64).reshape(-1, 1)
    phi_ref = lift(np.asarray(reference, dtype=np.float64).reshape(-1, 1))
    phi = lift(x)
    xs, recon, us = [x[:, 0].copy()], [], []
    for _ in range(num_steps):
        u = K @ (phi_ref - phi)
        us.append(u[:, 0].copy())
        recon.append((C @ phi)[:, 0])
        x = np.asarray(step(x, u), dtype=np.float64).reshape(-1, 1)
        xs.append(x[:, 0].copy())
        phi = lift(x)
    return np.array(xs).T, np.array(recon).T, np.array(us).T
