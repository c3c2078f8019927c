"""CPU: the drop-in surface (names, params, clone, pickle), kernel introspection, and the C-ABI library contract."""
import ctypes
import pathlib
import pickle
import re

import numpy as np
import pytest

ROOT = pathlib.Path(__file__).resolve().parents[1]


def test_module_exports_reference_names():
    import regressors as R
    for name in ("np", "scipy", "ThreeDimensionalKernel", "KernelWrapper", "LinearKernelWrapper", "KoopmanRegressor",
                 "KoopmanKernelRegressor", "KoopmanNystromRegressor", "KoopmanSplineRegressor"):
        assert hasattr(R, name), name
    ns = {}
    exec("from regressors import *", ns)
    assert "KoopmanNystromRegressor" in ns and "np" in ns and "scipy" in ns


def test_constructor_params_and_clone():
    from sklearn.base import clone
    import regressors as R
    k = R.ThreeDimensionalKernel(1, 10, 100, 192)
    assert k.kernel.length_scale.shape == (1, 192)
    assert list(np.squeeze(k.kernel.length_scale)[:4]) == [1.0, 10.0, 100.0, 1.0]
    reg = R.KoopmanNystromRegressor(6, kernel=k, gamma=1e-7, m=100)
    assert reg.get_params() == {"n_inputs": 6, "kernel": k, "gamma": 1e-7, "m": 100}
    assert reg.A is None and reg.B is None and reg.C is None and reg.weights is None
    assert reg.nystrom_centers_input is None and reg.nystrom_centers_output is None and reg.jitter == 1e-6
    c = clone(reg)
    assert c.get_params()["m"] == 100 and c.nystrom_centers_output is None
    c.set_params(gamma=1e-3)
    assert c.gamma == 1e-3


def test_kernel_spec():
    import regressors as R
    from nys_koop_lqr_b200.regressors import kernel_spec
    kind, ls = kernel_spec(R.ThreeDimensionalKernel(1, 2, 3, 7), 7)
    assert kind == 0 and list(ls) == [1, 2, 3, 1, 2, 3, 1]
    kind, ls = kernel_spec(R.KernelWrapper([1, 1]), 2)
    assert kind == 1 and list(ls) == [1, 1]
    kind, ls = kernel_spec(R.KernelWrapper(0.5), 3)
    assert kind == 1 and list(ls) == [0.5] * 3
    with pytest.raises(NotImplementedError):
        kernel_spec(R.LinearKernelWrapper(1.0), 2)       # DotProduct: no GPU implementation, no CPU fallback
    from sklearn.gaussian_process.kernels import Matern
    with pytest.raises(NotImplementedError):
        kernel_spec(Matern(1.0, nu=1.5), 2)
    with pytest.raises(ValueError):
        kernel_spec(R.KernelWrapper([1, 1, 1]), 2)


def test_pickle_roundtrip_drops_device_state():
    import regressors as R
    reg = R.KoopmanNystromRegressor(1, kernel=R.KernelWrapper([1, 1]), gamma=1e-6, m=5)
    reg.nystrom_centers_output = np.arange(10.0).reshape(2, 5)
    reg.nystrom_centers_input = reg.nystrom_centers_output
    reg.A = np.eye(5)
    reg.__dict__["_dev"] = {"not": "picklable", "fn": lambda: None}
    reg2 = pickle.loads(pickle.dumps(reg))
    assert "_dev" not in reg2.__dict__
    assert np.array_equal(reg2.nystrom_centers_output, reg.nystrom_centers_output) and np.array_equal(reg2.A, reg.A)


def test_fit_argument_checks_do_not_need_gpu():
    import regressors as R
    reg = R.KoopmanNystromRegressor(1, kernel=R.KernelWrapper([1, 1]), gamma=1e-6, m=5)
    with pytest.raises(ValueError):
        reg.fit(np.zeros((10, 3)), np.zeros((9, 2)))
    reg2 = R.KoopmanNystromRegressor(1, kernel=None, gamma=None, m=5)
    with pytest.raises(ValueError):
        reg2.fit(np.zeros((10, 3)), np.zeros((10, 2)))


def test_more_landmarks_than_samples_raises_like_the_reference():
    """regressors.py:130 draws without replacement: m > n is numpy's own ValueError, raised before any device work."""
    import regressors as R
    reg = R.KoopmanNystromRegressor(1, kernel=R.KernelWrapper([1, 1]), gamma=1e-6, m=50)
    with pytest.raises(ValueError, match="larger sample than population"):
        reg.fit(np.zeros((10, 3)), np.zeros((10, 2)))
    assert reg.nystrom_centers_output is None


def test_no_cpu_fallback_without_gpu():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    import regressors as R
    from nys_koop_lqr_b200._lib import NkError
    np.random.seed(0)
    reg = R.KoopmanNystromRegressor(1, kernel=R.KernelWrapper([1, 1]), gamma=1e-6, m=5)
    with pytest.raises(NkError):
        reg.fit(np.random.rand(20, 3), np.random.rand(20, 2))


def _declared_symbols():
    text = (ROOT / "include" / "nk_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(nk_[a-z_0-9]+)\s*\(", text)))


def test_c_abi_library_loads_and_exports_every_declared_symbol():
    from nys_koop_lqr_b200 import _lib, build
    build.build()
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    declared = _declared_symbols()
    assert len(declared) >= 18
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in include/nk_b200.h but not exported by libnkb200.so"
    assert set(declared) == set(_lib.EXPORTED_SYMBOLS), "ctypes signatures and the header disagree"
    lib.nk_version.restype = ctypes.c_int
    assert lib.nk_version() >= 100


def test_c_abi_fails_loudly_without_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from nys_koop_lqr_b200 import _lib
    lib = _lib.load()
    h = ctypes.c_void_p()
    rc = lib.nk_create(ctypes.byref(h), 0)
    assert rc != 0 and not h.value
    assert b"no CUDA device" in lib.nk_last_error_string(None) or b"CUDA" in lib.nk_last_error_string(None)


def test_comparator_baselines_work_on_the_host():
    """The scripts' spline / exact-kernel branches (benchmark_lqr_classic.py:55,237,276; _cloth.py:44,199,232) instantiate these
    CPU comparators from the drop-in module: same ctor parameters (sklearn clone), fit -> A, B, C, weights, predict."""
    import regressors as R
    from sklearn.base import clone
    rng = np.random.default_rng(0)
    n, d, p = 90, 2, 1
    X, Y = rng.standard_normal((n, d + p)), rng.standard_normal((n, d))
    np.random.seed(1)
    sp = clone(R.KoopmanSplineRegressor(p, state_bounds_params=np.array([1.5, 2.0]), m=12, gamma=1e-4))
    sp.fit(X, Y)
    assert sp.A.shape == (12, 12) and sp.B.shape == (12, p) and sp.C.shape == (d, 12) and sp.predict(X[:5]).shape == (5, d)
    assert np.allclose(sp.predict(X[:5]), (sp.weights @ np.vstack((sp.lift(X[:5, :d].T), X[:5, d:].T))).T)
    ke = clone(R.KoopmanKernelRegressor(p, kernel=R.KernelWrapper([1.0, 1.0]), gamma=1e-3))
    ke.fit(X, Y)
    assert ke.A.shape == (n, n) and ke.B.shape == (n, p) and ke.C.shape == (d, n) and ke.lift(X[:3, :d].T).shape == (n, 3)
    assert np.isfinite(ke.predict(X)).all() and ke.predict(X).shape == (n, d)


def test_header_is_plain_c_and_links_from_c(tmp_path):
    """include/nk_b200.h is the drop-in boundary: it must compile as plain C99 (no C++ / torch types in the signatures) and a C
    program must link against libnkb200.so.  Without a GPU nk_create fails with NK_E_CUDA and says why (no CPU fallback)."""
    import shutil
    import subprocess
    from nys_koop_lqr_b200 import _lib, build
    gcc = shutil.which("gcc")
    if gcc is None:
        pytest.skip("no gcc")
    build.build()
    src = tmp_path / "abi.c"
    src.write_text(
        '#include <stdio.h>\n#include "nk_b200.h"\n'
        'int main(void) { nk_handle *h = 0; int rc = nk_create(&h, 0);\n'
        '  if (rc != NK_OK) { printf("%d %s\\n", rc, nk_last_error_string(0)); return 3; }\n'
        '  printf("ok %d SMs, ABI %d\\n", nk_device_sm_count(h), nk_version()); return nk_destroy(h); }\n')
    exe = tmp_path / "abi"
    libdir = _lib.LIB_PATH.parent
    subprocess.run([gcc, "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I", str(ROOT / "include"), str(src), "-o", str(exe),
                    "-L", str(libdir), "-lnkb200", f"-Wl,-rpath,{libdir}"], check=True, capture_output=True, text=True)
    r = subprocess.run([str(exe)], capture_output=True, text=True, timeout=120)
    import torch
    if torch.cuda.is_available():
        assert r.returncode == 0 and r.stdout.startswith("ok ")
    else:
        assert r.returncode == 3 and r.stdout.startswith("-2 ") and "no CPU fallback" in r.stdout


def test_result_buffer_pool_never_hands_out_a_viewed_buffer():
    """Result arrays are numpy VIEWS of a pooled page-locked buffer: the pool may reuse it only when every view is gone."""
    import gc
    import torch
    from nys_koop_lqr_b200.engine import Engine
    eng = Engine.__new__(Engine)                      # the pool logic needs no device; allocate pageable memory here
    alloc = lambda n: torch.empty(n, dtype=torch.float64)
    a, arr = eng.result_buffer(100, alloc)
    view = arr[10:30].reshape(4, 5)
    ptr = a.data_ptr()
    del a, arr
    b, arr_b = eng.result_buffer(50, alloc)
    assert b.data_ptr() != ptr and len(eng._result_pool) == 2      # `view` still looks at the first buffer
    view[:] = 7.0
    del view
    gc.collect()
    c, arr_c = eng.result_buffer(80, alloc)
    assert c.data_ptr() == ptr and len(eng._result_pool) == 2      # ... now it is free again
    d, _ = eng.result_buffer(1000, alloc)                           # too small buffers are not reused
    assert d.numel() >= 1000 and len(eng._result_pool) == 3
