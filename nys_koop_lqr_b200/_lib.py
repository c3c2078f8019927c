"""ctypes binding of libnkb200.so (the C ABI declared in include/nk_b200.h).

There is no fallback: if the shared library is missing or no sm_100 GPU is usable, calls raise.
"""
from __future__ import annotations

import ctypes as C
import pathlib

_HERE = pathlib.Path(__file__).resolve().parent
import os as _os
LIB_PATH = pathlib.Path(_os.environ.get("NK_LIB_PATH", _HERE / "libnkb200.so"))   # override: development A/B builds only

NK_KERNEL_RBF, NK_KERNEL_MATERN52 = 0, 1

_c_dp = C.c_void_p  # device pointers are passed as integers
_ll = C.c_longlong
_i = C.c_int
_d = C.c_double



class NkGrams(C.Structure):
    """struct nk_grams of include/nk_b200.h: the seven data-sample Grams, device pointers + leading dimensions."""
    _fields_ = [(f, t) for name in ("Gxx", "Gyx", "Gyy", "Gxu", "Gyu", "Guu", "GYy") for f, t in ((name, C.c_void_p), ("ld_" + name[0].lower() + name[1:], C.c_longlong))]


class NkLandmarks(C.Structure):
    """struct nk_landmarks: K_zz, S, S^-1 of the output landmarks; optional k(Z_in,Z_in), k(Z_in,Z_out) for distinct input landmarks."""
    _fields_ = [("Kzz", C.c_void_p), ("ld_kzz", C.c_longlong), ("S", C.c_void_p), ("ld_s", C.c_longlong), ("Sinv", C.c_void_p), ("ld_sinv", C.c_longlong),
                ("Kzz_in", C.c_void_p), ("ld_kzz_in", C.c_longlong), ("Kio", C.c_void_p), ("ld_kio", C.c_longlong)]


_pG, _pL = C.POINTER(NkGrams), C.POINTER(NkLandmarks)

_SIGNATURES = {
    "nk_version": ([], _i),
    "nk_create": ([C.POINTER(C.c_void_p), _i], _i),
    "nk_destroy": ([C.c_void_p], _i),
    "nk_last_error_string": ([C.c_void_p], C.c_char_p),
    "nk_device_sm_count": ([C.c_void_p], _i),
    "nk_release_scratch": ([C.c_void_p], _i),
    "nk_gram_begin": ([C.c_void_p, _c_dp, _ll, _i, _i, _i, _c_dp, _i, _i, C.c_void_p], _i),
    "nk_gram_begin_io": ([C.c_void_p, _c_dp, _ll, _c_dp, _ll, _i, _i, _i, _c_dp, _i, _i, C.c_void_p], _i),
    "nk_gram_status": ([C.c_void_p, C.c_void_p], _i),
    "nk_allreduce_grams": ([C.c_void_p, C.c_void_p, _c_dp, _ll, C.c_void_p], _i),
    "nk_gram_update": ([C.c_void_p, _c_dp, _ll, _c_dp, _ll, _ll, C.c_void_p], _i),
    "nk_gram_finalize": ([C.c_void_p] + [_c_dp, _ll] * 7 + [_i, C.c_void_p], _i),
    "nk_gram_plan": ([_i, _i, _i, _i, _i, C.POINTER(_i), C.POINTER(_i), _i], _i),
    "nk_gram_last_executed_flops": ([C.c_void_p], _d),
    "nk_launch_count": ([C.c_void_p], _ll),
    "nk_probe_dmma_tflops": ([C.c_void_p, _d, C.POINTER(_d)], _i),
    "nk_kernel_function": ([C.c_void_p, _i, _ll, _c_dp, _c_dp, C.c_void_p], _i),
    "nk_kzz": ([C.c_void_p, _c_dp, _ll, _i, _i, _c_dp, _i, _c_dp, _ll, C.c_void_p], _i),
    "nk_kernel_cross": ([C.c_void_p, _c_dp, _ll, _i, _i, _c_dp, _i, _c_dp, _ll, _ll, _c_dp, _ll, C.c_void_p], _i),
    "nk_gemm": ([C.c_void_p, _i, _i, _i, _i, _i, _d, _c_dp, _ll, _c_dp, _ll, _d, _c_dp, _ll, C.c_void_p], _i),
    "nk_potrf": ([C.c_void_p, _i, _c_dp, _ll, C.POINTER(_i), C.c_void_p], _i),
    "nk_trsm_lower": ([C.c_void_p, _i, _i, _i, _c_dp, _ll, _c_dp, _ll, C.c_void_p], _i),
    "nk_sym_sqrt": ([C.c_void_p, _i, _c_dp, _ll, _d, _c_dp, _ll, _c_dp, _ll, C.POINTER(_i), C.c_void_p], _i),
    "nk_solve_abc": ([C.c_void_p, _i, _i, _i, _d, _d, _pG, _pL] + [_c_dp, _ll] * 4 + [C.POINTER(_i), C.c_void_p], _i),
    "nk_solve_abc_part": ([C.c_void_p, _i, _i, _i, _d, _d, _pG, _pL, _i, _i, _c_dp, _ll, _i, _i, _c_dp, _ll, C.POINTER(_i), C.c_void_p], _i),
    "nk_solve_abc_finish": ([C.c_void_p, _i, _i, _i, _c_dp, _ll, _c_dp, _ll] + [_c_dp, _ll] * 4 + [C.c_void_p], _i),
    "nk_lift": ([C.c_void_p, _c_dp, _ll, _i, _i, _c_dp, _i, _c_dp, _ll, _c_dp, _ll, _ll, _c_dp, _ll, _c_dp, _ll, C.c_void_p], _i),
    "nk_predict": ([C.c_void_p, _c_dp, _ll, _i, _i, _i, _c_dp, _i, _c_dp, _ll, _c_dp, _ll, _c_dp, _ll, _ll, _c_dp, _ll, C.c_void_p], _i),
    "nk_rollout": ([C.c_void_p, _i, _i, _i, _i, _ll] + [_c_dp] * 10 + [C.c_void_p], _i),
    "nk_closed_loop": ([C.c_void_p, _i, _i, _i, _i, _ll] + [_c_dp] * 9 + [C.c_void_p], _i),
    "nk_cv_weights": ([C.c_void_p, _i, _i, _i, _i, C.POINTER(_d), _d, _pG, _c_dp, _ll, _c_dp, _ll, C.POINTER(_i), C.c_void_p], _i),
    "nk_cv_score": ([C.c_void_p, _c_dp, _ll, _i, _i, _i, _c_dp, _i, _c_dp, _i, _c_dp, _ll, _c_dp, _ll, _ll, _c_dp, C.c_void_p], _i),
    "nk_axpy": ([C.c_void_p, _ll, _d, _c_dp, _c_dp, C.c_void_p], _i),
}

EXPORTED_SYMBOLS = tuple(_SIGNATURES)

_lib = None


class NkError(RuntimeError):
    """A C-ABI call failed; `rc` is its status (NK_E_* of include/nk_b200.h: -1 invalid, -2 CUDA, -3 not SPD, -4 no memory, -5 state)."""

    def __init__(self, message, rc=None):
        super().__init__(message)
        self.rc = rc


NK_E_NOT_SPD = -3


def load() -> C.CDLL:
    """Load the in-tree shared library (built by nys_koop_lqr_b200.build / __graft_entry__.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise NkError(f"{LIB_PATH} is missing: run `python -m nys_koop_lqr_b200.build` (no CPU fallback exists)")
    lib = C.CDLL(str(LIB_PATH))
    for name, (argtypes, restype) in _SIGNATURES.items():
        fn = getattr(lib, name)   # AttributeError here means the .so does not export what include/nk_b200.h declares
        fn.argtypes = argtypes
        fn.restype = restype
    _lib = lib
    return lib


def check(handle, rc: int, what: str) -> None:
    if rc != 0:
        msg = load().nk_last_error_string(handle)
        raise NkError(f"{what} failed (rc={rc}): {msg.decode() if msg else ''}", rc=rc)
