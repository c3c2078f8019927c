"""nys_koop_lqr_b200 -- B200-native (sm_100a) Nystrom-Koopman fit / lift / forecast hot path.

Drop-in for LCSL/nys-koop-lqr's ``regressors.KoopmanNystromRegressor`` (see ``regressors.py`` at the repo root
and ``nys_koop_lqr_b200/regressors.py``); the arithmetic lives in hand-written CUDA behind the C ABI of
``include/nk_b200.h`` (``libnkb200.so``, built in-tree by ``nys_koop_lqr_b200.build``).
"""
__all__ = ["build", "engine", "regressors"]
