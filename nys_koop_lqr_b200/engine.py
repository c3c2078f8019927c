"""Thin host layer over the C ABI (include/nk_b200.h): torch tensors are only device buffers + streams here.

Every method takes/returns float64 CUDA tensors (row-major, contiguous) and forwards raw pointers to
libnkb200.so.  Nothing in this module computes on the CPU; without the shared library or a B200 it raises.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib
from ._lib import NK_KERNEL_MATERN52, NK_KERNEL_RBF, NkError  # noqa: F401

JITTER = 1e-6  # regressors.py:120


def _ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)


def _f64(t, name):
    if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.is_contiguous()):
        raise TypeError(f"{name} must be a contiguous float64 CUDA tensor")
    return t


class Engine:
    """One handle per (process, device).  Not thread-safe (the C handle is not)."""

    _instances: dict = {}

    def __init__(self, device: int | None = None):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise NkError("no CUDA device: nys_koop_lqr_b200 has no CPU fallback")
        self.device = torch.cuda.current_device() if device is None else int(device)
        h = C.c_void_p()
        rc = self.lib.nk_create(C.byref(h), self.device)
        if rc != 0:
            raise NkError(f"nk_create failed (rc={rc}): {self.lib.nk_last_error_string(None).decode()}")
        self.h = h
        self.tdev = torch.device("cuda", self.device)
        self.gram_events = None   # set to [] to collect (start, end, n) CUDA-event triples per fused-kernel launch
        import os
        if not os.environ.get("NK_NO_WARM"):      # profiling runs skip the warm-up pass so that kernel filters see only their own launches
            self._warm()

    def _warm(self):
        """One tiny fit-shaped pass through every stage, so that the one-off costs of a process (lazy loading of the kernels
        into the context, opt-in shared-memory attributes, the cooperative-launch occupancy query, first workspace
        allocations) are paid when the engine is created and not inside the caller's first `fit` (it measured 3.1 s there)."""
        with torch.cuda.device(self.tdev):
            g = torch.Generator(device=self.tdev); g.manual_seed(0)
            n, d, p, m = 300, 3, 1, 24
            X = torch.randn(n, d + p, dtype=torch.float64, device=self.tdev, generator=g)
            Y = torch.randn(n, d, dtype=torch.float64, device=self.tdev, generator=g)
            Z = Y[:m].contiguous()
            il = torch.ones(d, dtype=torch.float64, device=self.tdev)
            G = self.grams(X, Y, Z, il, NK_KERNEL_RBF, p)
            Kzz = self.kzz(Z, il, NK_KERNEL_RBF)
            Kmm = Kzz.clone(); Kmm.diagonal().add_(JITTER)
            S, Sinv = self.sym_sqrt(Kmm)
            A, B, Cm, W = self.solve_abc(G, Kzz, S, Sinv, 1e-2 * n)
            self.predict(Z, il, NK_KERNEL_RBF, Sinv, W, X[:8], p)
            Z0 = self.lift(Z, il, NK_KERNEL_RBF, Sinv, X[:4, :d].contiguous(), transposed=True)
            self.rollout(A, B, Cm, Z0, torch.zeros(2, 4, p, dtype=torch.float64, device=self.tdev))
            torch.cuda.synchronize(self.tdev)

    @classmethod
    def get(cls, device: int | None = None) -> "Engine":
        _lib.load()
        if not torch.cuda.is_available():
            raise NkError("no CUDA device: nys_koop_lqr_b200 has no CPU fallback")
        dev = torch.cuda.current_device() if device is None else int(device)
        if dev not in cls._instances:
            cls._instances[dev] = Engine(dev)
        return cls._instances[dev]

    def close(self):
        if getattr(self, "h", None):
            self.lib.nk_destroy(self.h)
            self.h = None

    # ------------------------------------------------------------------ helpers
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.tdev).cuda_stream)

    def _ck(self, rc, what):
        _lib.check(self.h, rc, what)

    def empty(self, *shape):
        return torch.empty(*shape, dtype=torch.float64, device=self.tdev)

    def release_scratch(self):
        """Free the handle's grow-only device workspaces (and the pinned staging buffer)."""
        self._ck(self.lib.nk_release_scratch(self.h), "nk_release_scratch")
        self._pinned = None
        pool = self.__dict__.get("_result_pool", [])
        pool[:] = [e for e in pool if e[1]() is not None]          # keep only buffers that result arrays still view

    def pinned_staging(self, count):
        """Grow-only pinned host buffer of at least `count` doubles for device->host result copies."""
        buf = getattr(self, "_pinned", None)
        if buf is None or buf.numel() < count:
            buf = torch.empty(int(count), dtype=torch.float64, pin_memory=True)
            self._pinned = buf
        return buf

    def result_buffer(self, count, _alloc=None):
        """(tensor, ndarray view) of a page-locked host buffer of at least `count` doubles that nobody else uses: the result
        arrays of a fit are numpy views of `ndarray`, so the buffer must not be handed out again while any of them is alive.
        Page-locking costs ~1 ms per MB (183 ms for the 147 MB of an m=4096 fit), so buffers are pooled per engine and reused as
        soon as every array that viewed them is gone (a weak reference to the base ndarray tells)."""
        import weakref
        pool = self.__dict__.setdefault("_result_pool", [])
        for entry in pool:
            if entry[0].numel() >= count and entry[1]() is None:
                arr = entry[0].numpy()
                entry[1] = weakref.ref(arr)
                return entry[0], arr
        buf = (_alloc or (lambda n: torch.empty(n, dtype=torch.float64, pin_memory=True)))(int(count))
        arr = buf.numpy()
        pool.append([buf, weakref.ref(arr)])
        return buf, arr

    def side_stream(self):
        """A second stream of this device for transfers that may overlap with work queued on the current stream."""
        st = getattr(self, "_side", None)
        if st is None:
            st = self._side = torch.cuda.Stream(device=self.tdev)
        return st

    def launch_count(self) -> int:
        return int(self.lib.nk_launch_count(self.h))

    def probe_dmma_tflops(self, ms_target: float = 200.0) -> float:
        out = C.c_double(0.0)
        self._ck(self.lib.nk_probe_dmma_tflops(self.h, float(ms_target), C.byref(out)), "nk_probe_dmma_tflops")
        return out.value

    def sm_count(self) -> int:
        return int(self.lib.nk_device_sm_count(self.h))

    # ------------------------------------------------------------------ fused lift + Grams
    def gram_begin(self, Z, inv_ls, kind, p, chunk=0, Z_in=None):
        """Z: (m, d) output landmarks; Z_in: distinct input landmarks (m, d) or None (regressors.py:133-134)."""
        _f64(Z, "Z"); _f64(inv_ls, "inv_ls")
        m, d = Z.shape
        self._gram_shape = (m, d, int(p))
        if Z_in is None:
            self._ck(self.lib.nk_gram_begin(self.h, _ptr(Z), Z.stride(0), m, d, int(p), _ptr(inv_ls), int(kind), int(chunk),
                                            self._stream()), "nk_gram_begin")
        else:
            _f64(Z_in, "Z_in")
            if tuple(Z_in.shape) != (m, d):
                raise ValueError("input and output landmark sets must have the same shape")
            self._ck(self.lib.nk_gram_begin_io(self.h, _ptr(Z_in), Z_in.stride(0), _ptr(Z), Z.stride(0), m, d, int(p), _ptr(inv_ls),
                                               int(kind), int(chunk), self._stream()), "nk_gram_begin_io")

    def gram_status(self):
        """Synchronises the current stream; raises NkError if the fused kernel's dependence-wait watchdog fired."""
        self._ck(self.lib.nk_gram_status(self.h, self._stream()), "nk_gram_status")

    def gram_update(self, X_aug, Y):
        """X_aug (n, d+p) [state | controls], Y (n, d); row stride may exceed the width (views of wider buffers)."""
        m, d, p = self._gram_shape
        for t, nm, w in ((X_aug, "X_aug", d + p), (Y, "Y", d)):
            if not (t.is_cuda and t.dtype == torch.float64 and t.dim() == 2 and t.shape[1] == w and (t.stride(1) == 1 or w == 1)):
                raise TypeError(f"{nm} must be a float64 CUDA matrix with {w} unit-stride columns")
        n = X_aug.shape[0]
        if Y.shape[0] != n:
            raise ValueError("X_aug and Y must have the same number of rows")
        ev = None
        if self.gram_events is not None:     # CUDA events on the launching stream around the fused kernel (bench roofline)
            ev = (torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
            ev[0].record(torch.cuda.current_stream(self.tdev))
        self._ck(self.lib.nk_gram_update(self.h, _ptr(X_aug), X_aug.stride(0), _ptr(Y), Y.stride(0), n, self._stream()),
                 "nk_gram_update")
        if ev is not None:
            ev[1].record(torch.cuda.current_stream(self.tdev))
            self.gram_events.append((ev[0], ev[1], int(n)))

    def gram_finalize(self, out=None, accumulate=False):
        """Returns dict of the seven Grams (packed into one contiguous buffer `out['_flat']` for the allreduce)."""
        m, d, p = self._gram_shape
        sizes = dict(Gxx=(m, m), Gyx=(m, m), Gyy=(m, m), Gxu=(m, p), Gyu=(m, p), Guu=(p, p), GYy=(d, m))
        if out is None:
            total = sum(a * b for a, b in sizes.values())
            flat = torch.zeros(total, dtype=torch.float64, device=self.tdev)
            out = {"_flat": flat}
            o = 0
            for k, (a, b) in sizes.items():
                out[k] = flat[o:o + a * b].view(a, b)
                o += a * b
        args = []
        for k in ("Gxx", "Gyx", "Gyy", "Gxu", "Gyu", "Guu", "GYy"):
            t = out[k]
            args += [_ptr(t) if t.numel() else C.c_void_p(0), max(1, t.shape[1])]
        self._ck(self.lib.nk_gram_finalize(self.h, *args, int(bool(accumulate)), self._stream()), "nk_gram_finalize")
        return out

    def gram_views(self, flat, m, d, p):
        """Named views of a packed Gram buffer [Gxx|Gyx|Gyy|Gxu|Gyu|Guu|GYy] (the layout gram_finalize allocates)."""
        sizes = dict(Gxx=(m, m), Gyx=(m, m), Gyy=(m, m), Gxu=(m, p), Gyu=(m, p), Guu=(p, p), GYy=(d, m))
        out = {"_flat": flat}
        o = 0
        for k, (a, b) in sizes.items():
            out[k] = flat[o:o + a * b].view(a, b)
            o += a * b
        return out

    def grams(self, X_aug, Y, Z, inv_ls, kind, p, chunk=0, Z_in=None):
        self.gram_begin(Z, inv_ls, kind, p, chunk, Z_in=Z_in)
        self.gram_update(X_aug, Y)
        return self.gram_finalize()

    def gram_plan(self, m, d, p, chunk=0):
        """The work plan nk_gram_begin builds for these sizes on this GPU (host-only introspection, include/nk_b200.h)."""
        summ = (C.c_int * 12)()
        rc = self.lib.nk_gram_plan(int(m), int(d), int(p), int(chunk), self.sm_count(), summ, None, 0)
        if rc < 0:
            raise ValueError("nk_gram_plan: invalid sizes")
        keys = ("chunk", "MP", "KLS", "EP", "psi_rows", "nblk", "ntiles", "n_pack", "n_lift", "n_gram", "period_len", "nslots")
        return dict(zip(keys, [int(v) for v in summ]))

    def gram_executed_flops(self) -> float:
        return float(self.lib.nk_gram_last_executed_flops(self.h))

    # ------------------------------------------------------------------ dense stage
    def kernel_function(self, exponent, kind):
        """Elementwise kernel function of an array of exponents -r^2/2 (device implementation used by every lift)."""
        _f64(exponent, "exponent")
        out = torch.empty_like(exponent)
        self._ck(self.lib.nk_kernel_function(self.h, int(kind), exponent.numel(), _ptr(exponent), _ptr(out), self._stream()), "nk_kernel_function")
        return out

    def kzz(self, Z, inv_ls, kind):
        m, d = Z.shape
        K = self.empty(m, m)
        self._ck(self.lib.nk_kzz(self.h, _ptr(Z), Z.stride(0), m, d, _ptr(inv_ls), int(kind), _ptr(K), m, self._stream()), "nk_kzz")
        return K

    def kernel_cross(self, Z, X, inv_ls, kind):
        m, d = Z.shape
        N = X.shape[0]
        K = self.empty(m, N)
        self._ck(self.lib.nk_kernel_cross(self.h, _ptr(Z), Z.stride(0), m, d, _ptr(inv_ls), int(kind), _ptr(X), X.stride(0), N,
                                          _ptr(K), N, self._stream()), "nk_kernel_cross")
        return K

    def gemm(self, A, B, transa=False, transb=False, alpha=1.0, beta=0.0, out=None):
        _f64(A, "A"); _f64(B, "B")
        M, K = (A.shape[1], A.shape[0]) if transa else A.shape
        K2, N = (B.shape[1], B.shape[0]) if transb else B.shape
        if K != K2:
            raise ValueError("inner dimensions differ")
        if out is None:
            out = torch.zeros(M, N, dtype=torch.float64, device=self.tdev)
        self._ck(self.lib.nk_gemm(self.h, int(transa), int(transb), M, N, K, float(alpha), _ptr(A), A.stride(0), _ptr(B), B.stride(0),
                                  float(beta), _ptr(out), out.stride(0), self._stream()), "nk_gemm")
        return out

    def potrf(self, A):
        """In-place lower Cholesky of A (n,n). Raises NkError if not SPD."""
        _f64(A, "A")
        info = C.c_int(0)
        self._ck(self.lib.nk_potrf(self.h, A.shape[0], _ptr(A), A.stride(0), C.byref(info), self._stream()), "nk_potrf")
        return A

    def trsm_lower(self, L, B, trans=False):
        _f64(L, "L"); _f64(B, "B")
        self._ck(self.lib.nk_trsm_lower(self.h, int(trans), L.shape[0], B.shape[1], _ptr(L), L.stride(0), _ptr(B), B.stride(0),
                                        self._stream()), "nk_trsm_lower")
        return B

    def sym_sqrt(self, K, lambda_min_bound=JITTER):
        _f64(K, "K")
        n = K.shape[0]
        S, Sinv = self.empty(n, n), self.empty(n, n)
        iters = C.c_int(0)
        self._ck(self.lib.nk_sym_sqrt(self.h, n, _ptr(K), K.stride(0), float(lambda_min_bound), _ptr(S), n, _ptr(Sinv), n,
                                      C.byref(iters), self._stream()), "nk_sym_sqrt")
        self.last_sqrt_iters = iters.value
        return S, Sinv

    # -- struct marshalling (include/nk_b200.h: nk_grams, nk_landmarks) --
    @staticmethod
    def _mat(t, name, rows=None, cols=None):
        if t is None:
            return C.c_void_p(0), 0
        if not (isinstance(t, torch.Tensor) and t.is_cuda and t.dtype == torch.float64 and t.dim() == 2 and (t.stride(1) == 1 or t.shape[1] <= 1)):
            raise TypeError(f"{name} must be a float64 CUDA matrix with unit column stride")
        if t.numel() == 0:
            return C.c_void_p(0), max(1, t.shape[1])
        return C.c_void_p(t.data_ptr()), max(int(t.stride(0)), int(t.shape[1]), 1)

    def _grams_struct(self, G):
        g = _lib.NkGrams()
        for k in ("Gxx", "Gyx", "Gyy", "Gxu", "Gyu", "Guu", "GYy"):
            ptr, ld = self._mat(G[k], k)
            setattr(g, k, ptr)
            setattr(g, "ld_" + k[0].lower() + k[1:], ld)
        return g

    def _landmarks_struct(self, Kzz, S, Sinv, Kzz_in=None, Kio=None):
        lm = _lib.NkLandmarks()
        lm.Kzz, lm.ld_kzz = self._mat(Kzz, "Kzz")
        lm.S, lm.ld_s = self._mat(S, "S")
        lm.Sinv, lm.ld_sinv = self._mat(Sinv, "Sinv")
        lm.Kzz_in, lm.ld_kzz_in = self._mat(Kzz_in, "Kzz_in")
        lm.Kio, lm.ld_kio = self._mat(Kio, "Kio")
        return lm

    def solve_abc(self, G, Kzz, S, Sinv, gamma_n, jitter=JITTER, Kzz_in=None, Kio=None):
        """Grams -> A (m,m), B (m,p), C (d,m), W (d,m+p) (regressors.py:147-169).  Kzz_in / Kio: k(Z_in,Z_in), k(Z_in,Z_out) when the
        input landmarks differ from the output landmarks."""
        m = Kzz.shape[0]
        p = G["Guu"].shape[0]
        d = G["GYy"].shape[0]
        A, B, Cm, W = self.empty(m, m), self.empty(m, p), self.empty(d, m), self.empty(d, m + p)
        info = C.c_int(0)
        g, lm = self._grams_struct(G), self._landmarks_struct(Kzz, S, Sinv, Kzz_in, Kio)
        self._ck(self.lib.nk_solve_abc(self.h, m, p, d, float(gamma_n), float(jitter), C.byref(g), C.byref(lm), _ptr(A), m,
                                       _ptr(B) if p else C.c_void_p(0), max(p, 1), _ptr(Cm), m, _ptr(W), m + p, C.byref(info),
                                       self._stream()), "nk_solve_abc")
        return A, B, Cm, W

    def solve_abc_part(self, G, Kzz, S, Sinv, gamma_n, g_rows, c_rows, GT, CT, jitter=JITTER, Kzz_in=None, Kio=None):
        """A slice of the solve for sharding over devices: rows g_rows = (start, count) of G^T ((m+p, m): row c = column c of [A|B])
        are written to GT[:count], rows c_rows of C^T ((m, d)) to CT[:count].  Every device factors both systems and solves only
        its columns (include/nk_b200.h nk_solve_abc_part)."""
        m = Kzz.shape[0]
        p = G["Guu"].shape[0]
        d = G["GYy"].shape[0]
        info = C.c_int(0)
        g, lm = self._grams_struct(G), self._landmarks_struct(Kzz, S, Sinv, Kzz_in, Kio)
        self._ck(self.lib.nk_solve_abc_part(self.h, m, p, d, float(gamma_n), float(jitter), C.byref(g), C.byref(lm),
                                            int(g_rows[0]), int(g_rows[1]), _ptr(GT), GT.stride(0) if GT is not None else 0,
                                            int(c_rows[0]), int(c_rows[1]), _ptr(CT), CT.stride(0) if CT is not None else 0,
                                            C.byref(info), self._stream()), "nk_solve_abc_part")

    def solve_abc_finish(self, GT, CT, m, p, d):
        """Assembled G^T (m+p, m) and C^T (m, d) -> A, B, C, W = C G on this device."""
        A, B, Cm, W = self.empty(m, m), self.empty(m, p), self.empty(d, m), self.empty(d, m + p)
        self._ck(self.lib.nk_solve_abc_finish(self.h, m, p, d, _ptr(GT), GT.stride(0), _ptr(CT), CT.stride(0), _ptr(A), m,
                                              _ptr(B) if p else C.c_void_p(0), max(p, 1), _ptr(Cm), m, _ptr(W), m + p, self._stream()),
                 "nk_solve_abc_finish")
        return A, B, Cm, W

    def closed_loop(self, A, B, Cm, K, Z0, Zref, steps, return_final=False):
        """Batched lifted closed loop. Z0, Zref (nb, m); K (p, m). Returns Xs (steps, nb, d), Us (steps, nb, p)[, Zfinal]."""
        nb, m = Z0.shape
        d, p = Cm.shape[0], K.shape[0]
        Xs, Us = self.empty(steps, nb, d), self.empty(steps, nb, p)
        Zf = self.empty(nb, m) if return_final else None
        self._ck(self.lib.nk_closed_loop(self.h, m, p, d, int(steps), nb, _ptr(_f64(A, "A")), _ptr(_f64(B, "B")), _ptr(_f64(Cm, "C")),
                                         _ptr(_f64(K, "K")), _ptr(_f64(Z0, "Z0")), _ptr(_f64(Zref, "Zref")), _ptr(Xs), _ptr(Us), _ptr(Zf),
                                         self._stream()), "nk_closed_loop")
        return (Xs, Us, Zf) if return_final else (Xs, Us)

    # ------------------------------------------------------------------ cross-validation sweep
    def axpy(self, alpha, x, y):
        """y += alpha * x on the device (flat float64 buffers of equal length)."""
        _f64(x, "x"); _f64(y, "y")
        if x.numel() != y.numel():
            raise ValueError("axpy: length mismatch")
        self._ck(self.lib.nk_axpy(self.h, x.numel(), float(alpha), _ptr(x), _ptr(y), self._stream()), "nk_axpy")
        return y

    def cv_weights(self, G, Kzz, gamma_n, jitter=JITTER):
        """All regularisation values of one (kernel, training fold) as one batch.  gamma_n: sequence of gamma * n_train.
        Returns (Wk (nlam, d, m+p) prediction weights in kernel-matrix coordinates, info list: 0 ok / 1,2,3 not SPD)."""
        m = Kzz.shape[0]
        p = G["Guu"].shape[0]
        d = G["GYy"].shape[0]
        nlam = len(gamma_n)
        Wk = self.empty(nlam, d, m + p)
        gn = (C.c_double * nlam)(*[float(g) for g in gamma_n])
        info = (C.c_int * nlam)()
        g = self._grams_struct(G)
        kptr, kld = self._mat(Kzz, "Kzz")
        rc = self.lib.nk_cv_weights(self.h, m, p, d, nlam, gn, float(jitter), C.byref(g), kptr, kld, _ptr(Wk), m + p, info, self._stream())
        if rc not in (0, -3):      # NK_E_NOT_SPD is reported per value through info (sklearn's error_score=nan semantics)
            self._ck(rc, "nk_cv_weights")
        return Wk, list(info)

    def cv_score(self, Z, inv_ls, kind, Wk, X_aug, Y, p, sse=None):
        """sse (nlam, d) += per-output squared error of Yhat = Wk [k(Z,x); u] over the held-out rows."""
        nlam, d, _ = Wk.shape
        m = Z.shape[0]
        if sse is None:
            sse = torch.zeros(nlam, d, dtype=torch.float64, device=self.tdev)
        N = X_aug.shape[0]
        self._ck(self.lib.nk_cv_score(self.h, _ptr(Z), Z.stride(0), m, d, int(p), _ptr(inv_ls), int(kind), _ptr(_f64(Wk, "Wk")), nlam * d,
                                      _ptr(X_aug), X_aug.stride(0), _ptr(Y), Y.stride(0), N, _ptr(sse), self._stream()), "nk_cv_score")
        return sse

    # ------------------------------------------------------------------ lift / predict / rollout
    def lift(self, Z, inv_ls, kind, Sinv, X_rows, transposed=False):
        """X_rows (N,d) -> phi (m,N) (or (N,m) if transposed)."""
        m, d = Z.shape
        N = X_rows.shape[0]
        if transposed:
            out = self.empty(N, m)
            args = (C.c_void_p(0), 0, _ptr(out), m)
        else:
            out = self.empty(m, N)
            args = (_ptr(out), N, C.c_void_p(0), 0)
        self._ck(self.lib.nk_lift(self.h, _ptr(Z), Z.stride(0), m, d, _ptr(inv_ls), int(kind), _ptr(Sinv), Sinv.stride(0),
                                  _ptr(X_rows), X_rows.stride(0), N, *args, self._stream()), "nk_lift")
        return out

    def predict(self, Z, inv_ls, kind, Sinv, W, X_aug, p):
        m, d = Z.shape
        N = X_aug.shape[0]
        out = self.empty(N, d)
        self._ck(self.lib.nk_predict(self.h, _ptr(Z), Z.stride(0), m, d, int(p), _ptr(inv_ls), int(kind), _ptr(Sinv), Sinv.stride(0),
                                     _ptr(W), W.stride(0), _ptr(X_aug), X_aug.stride(0), N, _ptr(out), d, self._stream()), "nk_predict")
        return out

    def rollout(self, A, B, Cm, Z0, U, Ytrue=None, return_traj=True, return_final=False):
        """Z0 (nb,m); U (T-1,nb,p); Ytrue (T,nb,d) optional.  Returns dict(Yhat (T,nb,d), sq_err (nb), sq_sim (nb), Zfinal)."""
        nb, m = Z0.shape
        d = Cm.shape[0]
        p = B.shape[1] if B is not None and B.numel() else 0
        T = (U.shape[0] + 1) if U is not None else (Ytrue.shape[0] if Ytrue is not None else 1)
        res = {}
        Yhat = self.empty(T, nb, d) if return_traj else None
        se = ss = None
        if Ytrue is not None:
            _f64(Ytrue, "Ytrue")
            se, ss = self.empty(nb), self.empty(nb)
        Zf = self.empty(nb, m) if return_final else None
        self._ck(self.lib.nk_rollout(self.h, m, p, d, T, nb, _ptr(_f64(A, "A")), _ptr(B) if p else C.c_void_p(0), _ptr(_f64(Cm, "C")),
                                     _ptr(_f64(Z0, "Z0")), _ptr(U) if (p and T > 1) else C.c_void_p(0), _ptr(Yhat), _ptr(Ytrue),
                                     _ptr(se), _ptr(ss), _ptr(Zf), self._stream()), "nk_rollout")
        res.update(Yhat=Yhat, sq_err=se, sq_sim=ss, Zfinal=Zf)
        return res
