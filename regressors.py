"""Drop-in module name used by the reference's scripts (``from regressors import *``, benchmark_lqr_hjb.py:20,
benchmark_lqr_classic.py:20, benchmark_lqr_cloth.py:15): re-exports the B200-backed estimator surface."""
from nys_koop_lqr_b200.regressors import *  # noqa: F401,F403
from nys_koop_lqr_b200.regressors import __all__  # noqa: F401
