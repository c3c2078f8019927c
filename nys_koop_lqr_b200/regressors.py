"""Drop-in ``regressors`` module: the reference's estimator surface, backed by the sm_100a kernels.

Mirrors LCSL/nys-koop-lqr ``regressors.py`` for the Nystrom-Koopman path (reference lines in each docstring):
same class names, constructor parameters, attribute names and array layouts, so ``benchmark_lqr_hjb.py``,
``benchmark_lqr_classic.py`` and ``benchmark_lqr_cloth.py`` (``from regressors import *``) run against it.
All arithmetic of ``KoopmanNystromRegressor.fit / lift / predict`` and of the batched ``forecast`` runs on the
GPU through the C ABI (``include/nk_b200.h``); there is no CPU path -- without the library or a B200 the
methods raise.  The landmark draw stays the reference's own ``np.random.choice`` call on the host
(regressors.py:129-132) so results are comparable seed for seed.
"""
from __future__ import annotations

import os
import zlib

import numpy as np
import scipy
import scipy.linalg
import scipy.signal
from sklearn.base import BaseEstimator
from sklearn.gaussian_process.kernels import RBF, DotProduct, Matern

__all__ = [
    "np", "scipy", "RBF", "Matern", "DotProduct", "BaseEstimator",
    "ThreeDimensionalKernel", "KernelWrapper", "LinearKernelWrapper",
    "KoopmanRegressor", "KoopmanNystromRegressor", "KoopmanKernelRegressor", "KoopmanSplineRegressor",
]


# ----------------------------------------------------------------------------------------------
# kernel holder types (regressors.py:15-30): plain containers of a scikit-learn kernel in `.kernel`
# ----------------------------------------------------------------------------------------------
class ThreeDimensionalKernel:
    """Anisotropic RBF whose length scale cycles (lx, ly, lz) over the state, shape (1, n_states) (regressors.py:15-22)."""

    def __init__(self, lx, ly, lz, n_states):
        scales = np.resize(np.array([lx, ly, lz], dtype=float), n_states).reshape(1, n_states)
        self.kernel = RBF(scales)


class KernelWrapper:
    """Matern nu=5/2 with the given length scale(s) (regressors.py:24-26)."""

    def __init__(self, ls):
        self.kernel = Matern(ls, nu=2.5)


class LinearKernelWrapper:
    """DotProduct holder (regressors.py:28-30); only referenced from a comment upstream. Not supported by the GPU fit."""

    def __init__(self, sigma):
        self.kernel = DotProduct(sigma_0=sigma)


def kernel_spec(kernel_holder, n_states):
    """(kind, length_scale[n_states]) from a holder's scikit-learn kernel; raises for kernels the hot path lacks."""
    from ._lib import NK_KERNEL_MATERN52, NK_KERNEL_RBF
    k = getattr(kernel_holder, "kernel", kernel_holder)
    if isinstance(k, Matern):   # Matern subclasses RBF: test it first
        if float(k.nu) != 2.5:
            raise NotImplementedError(f"Matern nu={k.nu}: only nu=2.5 is implemented on the GPU path (no CPU fallback)")
        kind = NK_KERNEL_MATERN52
    elif isinstance(k, RBF):
        kind = NK_KERNEL_RBF
    else:
        raise NotImplementedError(f"kernel {type(k).__name__} is not implemented on the GPU path (no CPU fallback)")
    ls = np.asarray(k.length_scale, dtype=np.float64).reshape(-1)
    if ls.size == 1:
        ls = np.full(n_states, float(ls[0]))
    if ls.size != n_states:
        raise ValueError(f"length_scale has {ls.size} entries for {n_states} state dimensions")
    if not np.all(ls > 0):
        raise ValueError("length_scale must be positive")
    return kind, ls


# ----------------------------------------------------------------------------------------------
# estimator base (regressors.py:32-55)
# ----------------------------------------------------------------------------------------------
class KoopmanRegressor(BaseEstimator):
    def __init__(self, n_inputs, gamma, m=None):
        self.gamma = gamma
        self.m = m
        self.A = None
        self.B = None
        self.C = None
        self.weights = None
        self.n_inputs = n_inputs

    def lift(self, X):
        raise NotImplementedError

    def fit(self, X, Y):
        raise NotImplementedError

    def predict(self, X_aug):
        """Rows of X_aug are [state | input]; returns (N, n_states) one-step predictions (regressors.py:48-55)."""
        X_aug = np.asarray(X_aug)
        n_states = X_aug.shape[1] - self.n_inputs
        lifted = self.lift(X_aug[:, :n_states].T)
        return (self.weights @ np.vstack((lifted, X_aug[:, n_states:].T))).T


def _engine():
    from .engine import Engine
    return Engine.get()


def _as_device_rows(eng, a, width=None):
    """numpy / torch(CPU or CUDA) 2-D -> float64 CUDA tensor with unit column stride (no copy if already there)."""
    import torch
    if isinstance(a, torch.Tensor):
        t = a
        if t.dtype != torch.float64:
            t = t.double()
        if not t.is_cuda:
            t = t.to(eng.tdev, non_blocking=True)
        if t.stride(-1) != 1:
            t = t.contiguous()
        return t
    arr = np.ascontiguousarray(np.asarray(a, dtype=np.float64))
    return torch.from_numpy(arr).to(eng.tdev)


def _host_copy(src, threads=None):
    """Fresh numpy array with the contents of `src` (a view of the engine's pinned staging buffer).  Large results are copied by
    a few threads: a first-touch copy into newly mapped pages runs at ~5 GB/s on one core (27 ms for A at m=4096)."""
    dst = np.empty_like(src)
    if src.nbytes < (16 << 20) or src.ndim != 2:
        np.copyto(dst, src)
        return dst
    from concurrent.futures import ThreadPoolExecutor
    threads = threads or max(1, min(8, (os.cpu_count() or 1)))
    step = -(-src.shape[0] // threads)
    with ThreadPoolExecutor(threads) as ex:                          # numpy releases the GIL inside the copy loop
        list(ex.map(lambda r: np.copyto(dst[r:r + step], src[r:r + step]), range(0, src.shape[0], step)))
    return dst


def _fingerprint(a):
    """Cheap content key of a (small) numpy array: the device-side cache must notice IN-PLACE edits (sklearn kernels are
    mutable, callers assign into the centre arrays), which object identity does not."""
    if a is None:
        return None
    a = np.asarray(a)
    if a.nbytes <= (64 << 20):
        return (a.shape, a.dtype.str, zlib.crc32(np.ascontiguousarray(a).view(np.uint8).reshape(-1)))
    flat = a.reshape(-1)
    step = max(1, flat.size // (1 << 20))
    return (a.shape, a.dtype.str, zlib.crc32(np.ascontiguousarray(flat[::step]).view(np.uint8)), float(flat.sum()))


def _lazy_result(name):
    """A / B / C / weights: plain numpy attributes for callers (as upstream), backed by a pending DEVICE result after a
    sharded fit -- the download happens on first access, on the ranks that look (results once per node)."""
    key = "_" + name

    def getter(self):
        v = self.__dict__.get(key)
        if v is None and name in (self.__dict__.get("_pending") or {}):
            self._materialize()
            v = self.__dict__.get(key)
        return v

    def setter(self, v):
        self.__dict__[key] = v
        pend = self.__dict__.get("_pending")
        if pend:
            pend.pop(name, None)
        if name == "weights" and self.__dict__.get("_dev"):
            self.__dict__["_dev"]["W"] = None          # assigned / loaded weights: the device copy used by predict is stale

    return property(getter, setter)


class KoopmanNystromRegressor(KoopmanRegressor):
    """Nystrom-Koopman estimator (regressors.py:114-178) on the B200 kernels.

    fit(X (n, d+p), Y (n, d)) -> None; lift(X (d, N)) -> (m, N); predict(X_aug (N, d+p)) -> (N, d).
    After fit: A (m,m), B (m,p), C (d,m), weights (d, m+p) are numpy float64; centres are (d, m) like upstream.
    Additive API: ``forecast`` (batched open-loop rollout), ``fit_cv`` (batched grid search), ``fit_distributed`` (sample-sharded fit, one NCCL
    allreduce of the Grams), ``closed_loop`` / ``lqr_closed_loop`` (the scripts' ``lqr_control`` loops), ``lqr_gain`` (``control.dlqr`` on the
    device), ``stream_block`` (rows per host->device block of the streaming fit).
    """

    stream_block = 262144     # samples per host->device block when inputs live in host memory
    device_block = 1 << 24    # samples per fused-kernel launch when inputs are already on the device
    gram_chunk = 0            # samples per on-chip feature chunk (0: library default 512)
    pinned_results = True     # A / B / C / weights are numpy views of one page-locked buffer per fit (False: pageable copies)

    def __init__(self, n_inputs, kernel=None, gamma=None, m=None):
        super().__init__(n_inputs, gamma, m)
        self.kernel = kernel
        self.nystrom_centers_input = None
        self.nystrom_centers_output = None
        self.jitter = 1e-6

    A = _lazy_result("A")
    B = _lazy_result("B")
    C = _lazy_result("C")
    weights = _lazy_result("weights")

    # -- device-side cache (never pickled) ------------------------------------------------------
    def __getstate__(self):
        self._materialize()
        state = dict(self.__dict__)
        state.pop("_dev", None)
        state.pop("_pending", None)
        for k in ("A", "B", "C", "weights"):          # plain attribute names in the pickle, as the reference class writes them
            state[k] = state.pop("_" + k, None)
        return state

    def __setstate__(self, state):
        state = dict(state)
        for k in ("A", "B", "C", "weights"):          # also accepts pickles written by the reference class itself
            if k in state:
                state["_" + k] = state.pop(k)
        self.__dict__.update(state)

    def _distinct_input_centers(self):
        zi, zo = self.nystrom_centers_input, self.nystrom_centers_output
        return zi is not None and zi is not zo and not np.array_equal(zi, zo)

    def _cache_key(self, n_states):
        kind, ls = kernel_spec(self.kernel, n_states)
        zi = self.nystrom_centers_input if self._distinct_input_centers() else None
        return (kind, tuple(ls.tolist()), float(self.jitter), _fingerprint(self.nystrom_centers_output), _fingerprint(zi))

    def _device_state(self, n_states, landmark_stage=True):
        """Landmarks, 1/l, K_zz, S, S^-1 on the device; rebuilt lazily (e.g. after unpickling) and whenever the kernel's
        hyper-parameters, the jitter or the CONTENTS of the centre arrays change (the key is by value, not identity: sklearn
        kernels and the centre arrays are mutable).  With ``landmark_stage=False`` only the landmarks and 1/l are set up (K_zz,
        S, S^-1 are filled in later: the sample-sharded fit computes them on one rank and broadcasts them)."""
        import torch
        key = self._cache_key(n_states)
        dev = self.__dict__.get("_dev")
        if dev is not None and dev["key"] == key and (dev.get("S") is not None or not landmark_stage):
            return dev
        eng = _engine()
        kind, ls = kernel_spec(self.kernel, n_states)
        rows = lambda c: torch.from_numpy(np.ascontiguousarray(np.asarray(c, dtype=np.float64).T)).to(eng.tdev)   # (m, d) rows
        Z = rows(self.nystrom_centers_output)
        if Z.shape[1] != n_states:
            raise ValueError(f"landmarks have {Z.shape[1]} state dimensions, the data {n_states}")
        Z_in = rows(self.nystrom_centers_input) if self._distinct_input_centers() else None
        if Z_in is not None and tuple(Z_in.shape) != tuple(Z.shape):
            raise ValueError("nystrom_centers_input and nystrom_centers_output must have the same shape (regressors.py:143 adds jitter * eye(m))")
        inv_ls = torch.from_numpy(1.0 / ls).to(eng.tdev)
        dev = dict(key=key, eng=eng, kind=kind, ls=ls, Z=Z, Z_in=Z_in, inv_ls=inv_ls, Kzz=None, S=None, Sinv=None, Kzz_in=None, Kio=None)
        if landmark_stage:
            self._landmark_stage(dev)
        self.__dict__["_dev"] = dev
        return dev

    def _landmark_stage(self, dev):
        """K_zz (regressors.py:144), K_mm = K_zz + jitter I (:139,143), S = K_mm^(1/2) and S^-1 (:140,152-153,163): the part
        of the fit that depends on the landmarks only.  With distinct input landmarks also k(Z_in, Z_in) (:143) and
        k(Z_in, Z_out) (:144)."""
        eng = dev["eng"]
        Kzz = eng.kzz(dev["Z"], dev["inv_ls"], dev["kind"])
        Kmm = Kzz.clone()
        Kmm.diagonal().add_(self.jitter)
        S, Sinv = eng.sym_sqrt(Kmm, lambda_min_bound=self.jitter)
        dev.update(Kzz=Kzz, S=S, Sinv=Sinv)
        if dev["Z_in"] is not None:
            dev["Kzz_in"] = eng.kzz(dev["Z_in"], dev["inv_ls"], dev["kind"])
            dev["Kio"] = eng.kernel_cross(dev["Z_in"], dev["Z"], dev["inv_ls"], dev["kind"])     # (m, m): rows = input landmarks

    # -- landmarks --------------------------------------------------------------------------------
    def _ensure_centers(self, Y_rows_getter, n):
        """regressors.py:129-134: one np.random.choice draw on the global legacy RNG, landmarks are next-state
        samples, input centres alias the output centres unless the caller injected its own; persisted so a refit reuses them."""
        if self.nystrom_centers_output is None:
            idx = np.random.choice(np.arange(0, n), size=self.m, replace=False)
            self.nystrom_centers_output = Y_rows_getter(idx)          # (d, m)
        if self.nystrom_centers_input is None:
            self.nystrom_centers_input = self.nystrom_centers_output
        self.m = int(np.asarray(self.nystrom_centers_output).shape[1]) if self.m is None else self.m

    @staticmethod
    def _rows_to_centers(Y, idx):
        import torch
        if isinstance(Y, torch.Tensor):
            sel = Y[torch.as_tensor(idx, device=Y.device, dtype=torch.long)]
            return np.ascontiguousarray(sel.double().cpu().numpy().T)
        return np.ascontiguousarray(np.asarray(Y, dtype=np.float64)[idx].T)

    # -- Grams ------------------------------------------------------------------------------------
    def _accumulate_grams(self, eng, dev, X, Y):
        """Streams (X, Y) through the fused lift+Gram kernel. Host inputs go up in double-buffered blocks."""
        import torch
        n = X.shape[0]
        eng.gram_begin(dev["Z"], dev["inv_ls"], dev["kind"], self.n_inputs, self.gram_chunk, Z_in=dev["Z_in"])
        on_device = isinstance(X, torch.Tensor) and X.is_cuda and isinstance(Y, torch.Tensor) and Y.is_cuda
        if on_device:
            Xd, Yd = _as_device_rows(eng, X), _as_device_rows(eng, Y)
            for s in range(0, n, self.device_block):       # one launch for n <= 2^24; nk_gram_update bounds the chunks per call
                eng.gram_update(Xd[s:s + self.device_block], Yd[s:s + self.device_block])
            return
        blk = int(self.stream_block)
        if n <= blk:
            eng.gram_update(_as_device_rows(eng, X), _as_device_rows(eng, Y))
            return
        # double-buffered upload on a side stream, overlapped with the Gram kernel of the previous block
        Xt = X if isinstance(X, torch.Tensor) else torch.from_numpy(np.asarray(X, dtype=np.float64))
        Yt = Y if isinstance(Y, torch.Tensor) else torch.from_numpy(np.asarray(Y, dtype=np.float64))
        wx, wy = Xt.shape[1], Yt.shape[1]
        bufs = [(torch.empty(blk, wx, dtype=torch.float64, device=eng.tdev), torch.empty(blk, wy, dtype=torch.float64, device=eng.tdev))
                for _ in range(2)]
        copy_stream = torch.cuda.Stream(device=eng.tdev)
        main = torch.cuda.current_stream(eng.tdev)
        ready = [torch.cuda.Event() for _ in range(2)]
        freed = [torch.cuda.Event() for _ in range(2)]
        for b in range(2):
            freed[b].record(main)
        nblk = (n + blk - 1) // blk
        for i in range(nblk):
            s, e = i * blk, min(n, (i + 1) * blk)
            b = i & 1
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(freed[b])
                bufs[b][0][: e - s].copy_(Xt[s:e], non_blocking=True)
                bufs[b][1][: e - s].copy_(Yt[s:e], non_blocking=True)
                ready[b].record(copy_stream)
            main.wait_event(ready[b])
            eng.gram_update(bufs[b][0][: e - s], bufs[b][1][: e - s])
            freed[b].record(main)
        self._h2d_bytes = int(n) * (wx + wy) * 8

    # -- Grams -> A, B, C, weights ------------------------------------------------------------------
    @staticmethod
    def _shift_grams(G, scale=64.0):
        """Tikhonov retry for a numerically indefinite regularised system: add delta = scale * eps * max diagonal entry to the
        diagonals of G_xx, G_uu, G_yy (i.e. to inner_term and inner_term_rec).  The reference's lstsq (LAPACK gelsd, rcond = eps)
        discards the directions below eps * sigma_max instead; both regularise the same near-null space.  Stated deviation."""
        import torch
        delta = scale * float(np.finfo(np.float64).eps) * float(torch.maximum(G["Gxx"].diagonal().max(), G["Gyy"].diagonal().max()))
        for k in ("Gxx", "Gyy", "Guu"):
            if G[k].numel():
                G[k].diagonal().add_(delta)
        return delta

    def _solve_guarded(self, fn, G):
        """Runs a solve; if a regularised system is not positive definite in float64 (cond >~ 1e16: the tiniest gammas of the
        scripts' grids), retries once with a diagonal shift (see _shift_grams) and records it in ``spd_shift_``."""
        from ._lib import NK_E_NOT_SPD, NkError
        self.spd_shift_ = 0.0
        try:
            return fn()
        except NkError as exc:
            if exc.rc != NK_E_NOT_SPD:
                raise
            import warnings
            self.spd_shift_ = self._shift_grams(G)
            warnings.warn(f"KoopmanNystromRegressor: regularised system not positive definite in float64; retrying with a diagonal "
                          f"shift of {self.spd_shift_:.3e} (the reference's lstsq truncates those directions instead)", RuntimeWarning)
            return fn()

    def _set_device_results(self, dev, A, B, C, W, eager):
        dev["W"], dev["W_src"] = W, None
        self.__dict__["_pending"] = dict(A=A, B=B, C=C, weights=W)
        for k in ("A", "B", "C", "weights"):
            self.__dict__["_" + k] = None
        if eager:
            self._materialize()

    def _materialize(self):
        """Pending device results -> numpy attributes: one pinned staging buffer (engine-owned, grow-only), A and B on a side
        stream, first-touch host copies of the large matrices by a few threads."""
        pend = self.__dict__.get("_pending")
        if not pend:
            return
        import torch
        eng = self.__dict__["_dev"]["eng"] if self.__dict__.get("_dev") else _engine()
        names = [k for k in ("A", "B", "C", "weights") if k in pend]
        tensors = [pend[k] for k in names]
        total = sum(t.numel() for t in tensors)
        # One page-locked buffer per fit from the engine's pool (reused once the arrays of an earlier fit are gone); the numpy
        # attributes are VIEWS of it -- no second, first-touch copy on the host (that copy was 10 of the 12 ms of a 147 MB
        # result download).  The buffer stays out of the pool as long as any of the arrays lives.
        if self.pinned_results:
            host, arr = eng.result_buffer(total)
        else:
            host = eng.pinned_staging(total)
            arr = host.numpy()
        main = torch.cuda.current_stream(eng.tdev)
        o, slots = 0, []
        for t in tensors:
            slots.append((o, tuple(t.shape)))
            host[o:o + t.numel()].copy_(t.reshape(-1), non_blocking=True)
            o += t.numel()
        main.synchronize()
        for k, (off, shape) in zip(names, slots):
            view = arr[off:off + int(np.prod(shape))].reshape(shape)
            self.__dict__["_" + k] = view if self.pinned_results else _host_copy(view)
        self._d2h_bytes = int(total) * 8
        self.__dict__["_pending"] = {}

    def _solve(self, eng, dev, G, n_total, d):
        """Grams -> A, B, C, weights on the device (nk_solve_abc), then numpy copies on the host."""
        gamma_n = float(self.gamma) * float(n_total)                 # regressors.py:127
        A, B, C, W = self._solve_guarded(lambda: eng.solve_abc(G, dev["Kzz"], dev["S"], dev["Sinv"], gamma_n, self.jitter,
                                                               Kzz_in=dev["Kzz_in"], Kio=dev["Kio"]), G)
        self._set_device_results(dev, A, B, C, W, eager=True)

    # -- public API -------------------------------------------------------------------------------
    def fit(self, X, Y):
        """regressors.py:122-169.  X (n, d+p) rows [x_t | u_t], Y (n, d) rows x_{t+1}; numpy, torch CPU (pinned
        preferred) or torch CUDA.  Returns None like the reference."""
        n = int(X.shape[0])
        d = int(X.shape[1]) - self.n_inputs
        if int(Y.shape[0]) != n or int(Y.shape[1]) != d:
            raise ValueError("X must be (n, n_states + n_inputs) and Y (n, n_states)")
        if self.gamma is None or self.kernel is None:
            raise ValueError("kernel and gamma must be set before fit")
        self._ensure_centers(lambda idx: self._rows_to_centers(Y, idx), n)
        dev = self._device_state(d)
        eng = dev["eng"]
        self._accumulate_grams(eng, dev, X, Y)
        G = eng.gram_finalize()
        self._solve(eng, dev, G, n, d)

    # -- helpers of the multi-process paths ---------------------------------------------------------
    @staticmethod
    def _agree(group, device, fn, what):
        """Runs a rank-local stage and makes its outcome collective: a failure on ANY rank (NkError from the library, a shape
        error, ...) raises on EVERY rank before the next data collective, instead of leaving the others blocked in it."""
        import torch
        import torch.distributed as dist
        err, out = None, None
        try:
            out = fn()
        except Exception as exc:  # noqa: BLE001
            err = exc
        flag = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=device)
        dist.all_reduce(flag, op=dist.ReduceOp.MAX, group=group)
        if int(flag.item()):
            if err is not None:
                raise err
            from ._lib import NkError
            raise NkError(f"{what}: failed on another rank")
        return out

    def _draw_shared_landmarks(self, n_total, off, n_local, Y_local, d, group, device):
        """regressors.py:129-132 over the GLOBAL sample index.  Every rank makes the reference's draw (so that each process's
        legacy RNG advances exactly as in a single-process run), but rank 0's indices are the ones used: they are broadcast, so
        ranks whose numpy RNG states differ still assemble ONE landmark set."""
        import torch
        import torch.distributed as dist
        from . import sharding
        idx = np.random.choice(np.arange(0, n_total), size=self.m, replace=False)
        t = torch.from_numpy(np.ascontiguousarray(idx, dtype=np.int64)).to(device)
        dist.broadcast(t, src=dist.get_global_rank(group, 0) if group is not None else 0, group=group)
        idx = t.cpu().numpy()
        rows_of = lambda loc: torch.from_numpy(np.ascontiguousarray(self._rows_to_centers(Y_local, loc).T))
        Z = sharding.assemble_landmarks(idx, off, n_local, rows_of, d, group, device)
        self.nystrom_centers_output = np.ascontiguousarray(Z.cpu().numpy().T)

    def fit_distributed(self, X_local, Y_local, group=None, head_rank=0, materialize="lazy"):
        """Sample-sharded fit: every rank holds a contiguous block of samples; landmarks are drawn with the
        reference's call over the GLOBAL sample index (rank 0's draw, broadcast), the local Grams are summed with ONE allreduce
        -- the only collective on the n-proportional path.

        What does not shrink with the rank count is kept off the critical path:
          * the landmark-only stage (K_zz, the symmetric square root S and S^-1) is computed by ``head_rank`` alone, after its
            own Gram pass, and broadcast -- the other ranks are still streaming samples then if the head's shard is smaller by
            ``sharding.head_samples(m, d, p)`` (``sharding.balanced_bounds``); ``head_rank=None``: every rank computes it;
          * the two regularised solves are SHARDED by right-hand-side column: every rank factors both systems, solves (m+p)/G
            columns of [A|B] and m/G columns of C (``nk_solve_abc_part``), and two all-gathers assemble G^T and C^T;
          * the results stay on the device until somebody reads them: ``materialize="lazy"`` (default) downloads A / B / C /
            weights on first attribute access on the ranks that look, ``"all"`` downloads on every rank before returning."""
        import time
        import torch
        import torch.distributed as dist
        from . import sharding
        eng = _engine()
        n_local = int(X_local.shape[0])
        d = int(X_local.shape[1]) - self.n_inputs
        p = self.n_inputs
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        prof = {} if os.environ.get("NK_PROFILE") else None      # per-phase wall times (device-synchronised) in self.profile_

        def tick(name, t0):
            if prof is None:
                return 0.0
            torch.cuda.synchronize(eng.tdev)
            now = time.perf_counter()
            if name:
                prof[name] = prof.get(name, 0.0) + (now - t0) * 1e3
            return now
        t = tick(None, 0.0)
        n_total, off = sharding.global_layout(n_local, group, eng.tdev)
        if self.nystrom_centers_output is None:
            self._draw_shared_landmarks(n_total, off, n_local, Y_local, d, group, eng.tdev)
        self._ensure_centers(None, n_total)
        split = head_rank is not None and world > 1
        t = tick("layout_ms", t)

        def local_stage():
            nonlocal t
            dev = self._device_state(d, landmark_stage=not split)
            self._accumulate_grams(eng, dev, X_local, Y_local)
            G = eng.gram_finalize()
            t = tick("gram_pass_ms", t)
            lm = None
            if split and dev.get("S") is None:
                m = dev["Z"].shape[0]
                nmat = 5 if dev["Z_in"] is not None else 3
                lm = torch.empty(nmat, m, m, dtype=torch.float64, device=eng.tdev)
                if rank == head_rank:
                    self._landmark_stage(dev)
                    for i, k in enumerate(("Kzz", "S", "Sinv", "Kzz_in", "Kio")[:nmat]):
                        lm[i].copy_(dev[k])
                t = tick("landmark_stage_ms", t)
            return dev, G, lm
        dev, G, lm = self._agree(group, eng.tdev, local_stage, "fit_distributed: Gram pass / landmark stage")
        t = tick("wait_for_slowest_rank_ms", t)
        sharding.allreduce_grams(G["_flat"], group)                                   # the only data-path collective
        t = tick("allreduce_ms", t)
        if lm is not None:
            dist.broadcast(lm, src=dist.get_global_rank(group, head_rank) if group is not None else head_rank, group=group)
            dev.update(Kzz=lm[0], S=lm[1], Sinv=lm[2])
            if lm.shape[0] == 5:
                dev.update(Kzz_in=lm[3], Kio=lm[4])
            t = tick("broadcast_ms", t)
        # ---- column-sharded solve ----
        m = dev["Z"].shape[0]
        N1 = m + p
        g_per, g0, gc = sharding.solve_row_ranges(N1, world, rank)
        c_per, c0, cc = sharding.solve_row_ranges(m, world, rank)
        g_rows, c_rows = (g0, gc), (c0, cc)
        GT_all = torch.zeros(world * g_per, m, dtype=torch.float64, device=eng.tdev)
        CT_all = torch.zeros(world * c_per, d, dtype=torch.float64, device=eng.tdev)
        GT_mine, CT_mine = GT_all[rank * g_per:(rank + 1) * g_per], CT_all[rank * c_per:(rank + 1) * c_per]
        gamma_n = float(self.gamma) * float(n_total)                                  # regressors.py:127
        part = lambda: eng.solve_abc_part(G, dev["Kzz"], dev["S"], dev["Sinv"], gamma_n, g_rows, c_rows, GT_mine, CT_mine, self.jitter,
                                          Kzz_in=dev["Kzz_in"], Kio=dev["Kio"])
        self._agree(group, eng.tdev, lambda: self._solve_guarded(part, G), "fit_distributed: regularised solves")
        t = tick("sharded_solve_ms", t)
        sharding.gather_rows(GT_all, GT_mine, group)
        sharding.gather_rows(CT_all, CT_mine, group)
        A, B, C, W = eng.solve_abc_finish(GT_all, CT_all, m, p, d)
        t = tick("gather_finish_ms", t)
        self._set_device_results(dev, A, B, C, W, eager=(materialize == "all"))
        t = tick("materialize_ms", t)
        self.profile_ = prof

    def fit_cv(self, X, Y, kernels, gammas, n_splits=5, refit=True):
        """Batched hyper-parameter search: the reference's ``learn_hyperparams`` (benchmark_lqr_cloth.py:39-66,
        _classic.py:44-64, _hjb.py:47-71), i.e. ``GridSearchCV(estimator, {'kernel': kernels, 'gamma': gammas},
        scoring='neg_root_mean_squared_error')`` with sklearn's default unshuffled ``KFold(n_splits)``, on the GPU.

        Per kernel the samples stream through the fused lift+Gram kernel ONCE (one pass per fold block); a training
        fold's Grams are the sum of the other folds' Grams; all gammas of a (kernel, fold) are factored as one batch
        (``nk_cv_weights``) and scored together on the held-out block (``nk_cv_score``).  gamma_n = gamma * n_train
        as in regressors.py:127.

        Deviation from GridSearchCV, stated: every clone there redraws its landmarks from its own training fold with
        the unseeded global RNG; here ONE landmark set (this estimator's centres, drawn from Y with the reference's
        call if unset) is shared by all candidates and folds -- what sharing Grams across folds requires.
        Returns a ``cv_results_``-style dict (same keys and candidate order as sklearn: gamma outer, kernel inner) and
        sets ``cv_results_``, ``best_index_``, ``best_params_``, ``best_score_``; with ``refit`` the estimator is then
        fitted on all samples with the best (kernel, gamma)."""
        import torch
        from scipy.stats import rankdata
        kernels, gammas = list(kernels), [float(g) for g in gammas]
        n = int(X.shape[0])
        d = int(X.shape[1]) - self.n_inputs
        p = self.n_inputs
        if int(Y.shape[0]) != n or int(Y.shape[1]) != d:
            raise ValueError("X must be (n, n_states + n_inputs) and Y (n, n_states)")
        self._ensure_centers(lambda idx: self._rows_to_centers(Y, idx), n)
        if self._distinct_input_centers():
            raise NotImplementedError("fit_cv shares one landmark set across candidates and folds; distinct input centres are only supported by fit")
        eng = _engine()
        Xd, Yd = _as_device_rows(eng, X), _as_device_rows(eng, Y)
        sizes = np.full(n_splits, n // n_splits, dtype=int)
        sizes[: n % n_splits] += 1
        stops = np.cumsum(sizes)
        folds = [(int(e - s), int(e)) for s, e in zip(sizes, stops)]
        Z = torch.from_numpy(np.ascontiguousarray(np.asarray(self.nystrom_centers_output, dtype=np.float64).T)).to(eng.tdev)
        m = Z.shape[0]
        scores = np.full((len(kernels), len(gammas), n_splits), np.nan)
        prof = {"gram_s": 0.0, "weights_s": 0.0, "score_s": 0.0} if getattr(self, "cv_profile", False) else None

        def tick():
            if prof is None:
                return 0.0
            import time
            torch.cuda.synchronize(eng.tdev)
            return time.perf_counter()
        for ki, holder in enumerate(kernels):
            kind, ls = kernel_spec(holder, d)
            inv_ls = torch.from_numpy(1.0 / ls).to(eng.tdev)
            Kzz = eng.kzz(Z, inv_ls, kind)
            fold_grams = []
            t0 = tick()
            for s, e in folds:                                   # one pass over the samples per kernel
                eng.gram_begin(Z, inv_ls, kind, p, self.gram_chunk)
                eng.gram_update(Xd[s:e], Yd[s:e])
                fold_grams.append(eng.gram_finalize()["_flat"])
            train = torch.empty_like(fold_grams[0])
            if prof is not None:
                prof["gram_s"] += tick() - t0
            for fi, (s, e) in enumerate(folds):
                t1 = tick()
                train.zero_()
                for fj in range(n_splits):
                    if fj != fi:
                        eng.axpy(1.0, fold_grams[fj], train)
                n_train = n - (e - s)
                Wk, info = eng.cv_weights(eng.gram_views(train, m, d, p), Kzz, [g * n_train for g in gammas], self.jitter)
                t2 = tick()
                sse = eng.cv_score(Z, inv_ls, kind, Wk, Xd[s:e], Yd[s:e], p).cpu().numpy()
                if prof is not None:
                    prof["weights_s"] += t2 - t1
                    prof["score_s"] += tick() - t2
                sc = -np.mean(np.sqrt(sse / (e - s)), axis=1)
                sc[np.asarray(info) != 0] = np.nan               # sklearn error_score=nan for a candidate whose fit fails
                scores[ki, :, fi] = sc
            del fold_grams, train
        self.cv_profile_ = prof
        return self._finish_cv(scores, kernels, gammas, n_splits, refit, lambda: self.fit(Xd, Yd))

    def fit_cv_distributed(self, X_local, Y_local, kernels, gammas, n_splits=5, refit=True, group=None):
        """`fit_cv` over several GPUs (one process per GPU, `torch.distributed`).  Every rank holds a contiguous block of the
        samples; folds are sklearn's unshuffled KFold over the GLOBAL index.  Per kernel: each rank streams its part of
        every fold through the fused kernel, ONE all_reduce sums the stacked per-fold Grams, the (fold, gamma-slice) solves
        are dealt out round-robin (`sharding.cv_tasks`) and their prediction weights exchanged with one all_reduce of a
        zero-filled buffer, every rank scores the held-out samples it owns, and one small all_reduce sums the squared
        errors.  All ranks end with identical `cv_results_`; with `refit` the winner is fitted by `fit_distributed`."""
        import torch
        import torch.distributed as dist
        from scipy.stats import rankdata
        from . import sharding
        kernels, gammas = list(kernels), [float(g) for g in gammas]
        eng = _engine()
        world, rank = dist.get_world_size(group), dist.get_rank(group)
        n_local = int(X_local.shape[0])
        d = int(X_local.shape[1]) - self.n_inputs
        p = self.n_inputs
        n_total, off = sharding.global_layout(n_local, group, eng.tdev)
        if self.nystrom_centers_output is None:
            self._draw_shared_landmarks(n_total, off, n_local, Y_local, d, group, eng.tdev)
        self._ensure_centers(None, n_total)
        if self._distinct_input_centers():
            raise NotImplementedError("fit_cv shares one landmark set across candidates and folds; distinct input centres are only supported by fit")
        Xd, Yd = _as_device_rows(eng, X_local), _as_device_rows(eng, Y_local)
        folds = sharding.kfold_bounds(n_total, n_splits)
        local = sharding.fold_local_ranges(n_total, n_splits, off, n_local)
        tasks = sharding.cv_tasks(n_splits, len(gammas), world)
        Z = torch.from_numpy(np.ascontiguousarray(np.asarray(self.nystrom_centers_output, dtype=np.float64).T)).to(eng.tdev)
        m, nlam = Z.shape[0], len(gammas)
        scores = np.full((len(kernels), nlam, n_splits), np.nan)
        prof = {"gram_s": 0.0, "weights_s": 0.0, "score_s": 0.0} if getattr(self, "cv_profile", False) else None

        def tick():
            if prof is None:
                return 0.0
            import time
            torch.cuda.synchronize(eng.tdev)
            return time.perf_counter()
        for ki, holder in enumerate(kernels):
            kind, ls = kernel_spec(holder, d)
            inv_ls = torch.from_numpy(1.0 / ls).to(eng.tdev)
            Kzz = eng.kzz(Z, inv_ls, kind)
            t0 = tick()

            def gram_pass():
                stacked = None
                for fi, (lo, hi) in enumerate(local):
                    eng.gram_begin(Z, inv_ls, kind, p, self.gram_chunk)
                    if hi > lo:
                        eng.gram_update(Xd[lo:hi], Yd[lo:hi])
                    flat = eng.gram_finalize()["_flat"]
                    if stacked is None:
                        stacked = torch.empty(n_splits, flat.numel(), dtype=torch.float64, device=eng.tdev)
                    stacked[fi].copy_(flat)
                return stacked
            stacked = self._agree(group, eng.tdev, gram_pass, "fit_cv_distributed: Gram pass")   # a failure on one rank raises on all
            sharding.allreduce_sum(stacked, group)                                  # per-fold Grams of ALL samples
            t1 = tick()
            Wall = torch.zeros(n_splits, nlam, d, m + p, dtype=torch.float64, device=eng.tdev)
            bad = torch.zeros(n_splits, nlam, dtype=torch.int32, device=eng.tdev)
            train = torch.empty_like(stacked[0])

            def my_tasks():
                last_fold = -1
                for trank, fi, g0, g1 in tasks:
                    if trank != rank:
                        continue
                    if fi != last_fold:
                        train.zero_()
                        for fj in range(n_splits):
                            if fj != fi:
                                eng.axpy(1.0, stacked[fj], train)
                        last_fold = fi
                    n_train = n_total - (folds[fi][1] - folds[fi][0])
                    Wk, info = eng.cv_weights(eng.gram_views(train, m, d, p), Kzz, [g * n_train for g in gammas[g0:g1]], self.jitter)
                    Wall[fi, g0:g1].copy_(Wk)
                    bad[fi, g0:g1] = torch.as_tensor(info, dtype=torch.int32, device=eng.tdev)
            self._agree(group, eng.tdev, my_tasks, "fit_cv_distributed: batched solves")
            sharding.allreduce_sum(Wall, group)                                     # every slice was written by exactly one rank
            sharding.allreduce_sum(bad, group)
            t2 = tick()
            sse = torch.zeros(n_splits, nlam, d, dtype=torch.float64, device=eng.tdev)
            for fi, (lo, hi) in enumerate(local):
                if hi > lo:
                    eng.cv_score(Z, inv_ls, kind, Wall[fi], Xd[lo:hi], Yd[lo:hi], p, sse=sse[fi])
            sharding.allreduce_sum(sse, group)
            sse_h, bad_h = sse.cpu().numpy(), bad.cpu().numpy()
            if prof is not None:
                t3 = tick()
                prof["gram_s"] += t1 - t0; prof["weights_s"] += t2 - t1; prof["score_s"] += t3 - t2
            for fi, (s, e) in enumerate(folds):
                sc = -np.mean(np.sqrt(sse_h[fi] / (e - s)), axis=1)
                sc[bad_h[fi] != 0] = np.nan
                scores[ki, :, fi] = sc
            del stacked, Wall, train
        self.cv_profile_ = prof
        return self._finish_cv(scores, kernels, gammas, n_splits, refit,
                               lambda: self.fit_distributed(Xd, Yd, group=group))

    def _finish_cv(self, scores, kernels, gammas, n_splits, refit, refit_fn):
        """cv_results_ in sklearn's candidate order (ParameterGrid: keys sorted, 'gamma' outer, 'kernel' inner)."""
        from scipy.stats import rankdata
        params, split = [], []
        for gi, g in enumerate(gammas):
            for ki, holder in enumerate(kernels):
                params.append({"gamma": g, "kernel": holder})
                split.append(scores[ki, gi])
        split = np.array(split)
        mean, std = split.mean(axis=1), split.std(axis=1)
        res = {"params": params, "param_gamma": np.array([pp["gamma"] for pp in params]), "param_kernel": [pp["kernel"] for pp in params],
               "mean_test_score": mean, "std_test_score": std,
               "rank_test_score": rankdata(-np.where(np.isnan(mean), -np.inf, mean), method="min").astype(np.int32)}
        for k in range(n_splits):
            res[f"split{k}_test_score"] = split[:, k]
        self.cv_results_ = res
        self.best_index_ = int(np.argmin(res["rank_test_score"]))
        self.best_params_ = params[self.best_index_]
        self.best_score_ = float(mean[self.best_index_])
        if refit:
            self.kernel = self.best_params_["kernel"]
            self.gamma = self.best_params_["gamma"]
            refit_fn()
        return res

    def lift(self, X):
        """regressors.py:171-178: X (d, N) column-samples -> phi = S^-1 k(Z, X), (m, N).  S^-1 is cached on the
        device (the reference recomputes sqrtm(K_mm) on every call)."""
        import torch
        X = np.asarray(X, dtype=np.float64)
        if X.ndim == 1:
            X = X.reshape(-1, 1)
        d, N = X.shape
        dev = self._device_state(d)
        eng = dev["eng"]
        m = int(dev["Z"].shape[0])                       # not self.m: centres may have been injected without a fit
        out = np.empty((m, N))
        step = max(1, int(2 ** 27 // max(m, 1)))
        for s in range(0, N, step):
            e = min(N, s + step)
            rows = torch.from_numpy(np.ascontiguousarray(X[:, s:e].T)).to(eng.tdev)
            out[:, s:e] = eng.lift(dev["Z"], dev["inv_ls"], dev["kind"], dev["Sinv"], rows).cpu().numpy()
        return out

    def predict(self, X_aug):
        """regressors.py:48-55 on the device: (W [phi(x); u])^T for X_aug (N, d+p) -> (N, d)."""
        import torch
        X_aug = np.ascontiguousarray(np.asarray(X_aug, dtype=np.float64))
        N = X_aug.shape[0]
        d = X_aug.shape[1] - self.n_inputs
        dev = self._device_state(d)
        eng = dev["eng"]
        W = dev.get("W")                                  # set by fit; dropped when weights are assigned or the state is rebuilt
        if W is None:
            W = torch.from_numpy(np.ascontiguousarray(np.asarray(self.weights, dtype=np.float64))).to(eng.tdev)
            dev["W"] = W
        out = np.empty((N, d))
        step = max(1, int(2 ** 27 // max(int(dev["Z"].shape[0]), 1)))
        for s in range(0, N, step):
            e = min(N, s + step)
            rows = torch.from_numpy(X_aug[s:e]).to(eng.tdev)
            out[s:e] = eng.predict(dev["Z"], dev["inv_ls"], dev["kind"], dev["Sinv"], W, rows, self.n_inputs).cpu().numpy()
        return out

    def forecast(self, x0, controls, true_trajectories=None):
        """Batched open-loop rollout (benchmark_lqr_cloth.py:18-36 for many trajectories at once).

        x0 (d, nb) initial states, controls (p, T-1, nb) (or (p, T-1) for one trajectory).
        Returns simulated states (d, T, nb) (or (d, T)); with ``true_trajectories`` of that shape also returns
        (rmse, rmse_percent) per trajectory: sqrt(mean((true-sim)^2)) and ||true-sim||_F/||sim||_F*100."""
        import torch
        x0 = np.asarray(x0, dtype=np.float64)
        controls = np.asarray(controls, dtype=np.float64)
        single = controls.ndim == 2
        if single:
            controls = controls[:, :, None]
            x0 = x0.reshape(-1, 1)
            if true_trajectories is not None:
                true_trajectories = np.asarray(true_trajectories)[:, :, None]
        d, nb = x0.shape
        p, Tm1, _ = controls.shape
        dev = self._device_state(d)
        eng = dev["eng"]
        to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(eng.tdev)
        Z0 = eng.lift(dev["Z"], dev["inv_ls"], dev["kind"], dev["Sinv"], to(x0.T), transposed=True)        # (nb, m)
        U = to(np.transpose(controls, (1, 2, 0)))                                                        # (T-1, nb, p)
        Yt = to(np.transpose(np.asarray(true_trajectories, dtype=np.float64), (1, 2, 0))) if true_trajectories is not None else None
        res = eng.rollout(to(self.A), to(self.B), to(self.C), Z0, U, Ytrue=Yt)
        sim = np.transpose(res["Yhat"].cpu().numpy(), (2, 0, 1))                                          # (d, T, nb)
        if single:
            sim = sim[:, :, 0]
        if true_trajectories is None:
            return sim
        se, ss = res["sq_err"].cpu().numpy(), res["sq_sim"].cpu().numpy()
        rmse = np.sqrt(se / (d * (Tm1 + 1)))
        pct = np.sqrt(se) / np.sqrt(ss) * 100
        if single:
            rmse, pct = float(rmse[0]), float(pct[0])
        return sim, rmse, pct


# (closed_loop is attached to KoopmanNystromRegressor below)
def _closed_loop(self, K, initial_states, references, num_steps):
    """Batched lifted closed loop: `lqr_control` of benchmark_lqr_cloth.py:69-104 (loop body :80-84) for many
    (initial state, reference) pairs at once.  K (p, m) is the LQR gain (``control.dlqr`` on the host, as upstream);
    initial_states, references: (d, nb) (or (d,) / (d,1) for one).  Returns (states (d, num_steps, nb), controls
    (p, num_steps, nb)): states[:, i] = C z_i, controls[:, i] = K (phi_ref - z_i), z_{i+1} = A z_i + B u_i."""
    import torch
    x0 = np.asarray(initial_states, dtype=np.float64)
    xr = np.asarray(references, dtype=np.float64)
    single = x0.ndim == 1 or x0.shape[1] == 1
    x0, xr = x0.reshape(x0.shape[0], -1), xr.reshape(xr.shape[0], -1)
    d = x0.shape[0]
    dev = self._device_state(d)
    eng = dev["eng"]
    to = lambda a: torch.from_numpy(np.ascontiguousarray(np.asarray(a, dtype=np.float64))).to(eng.tdev)
    Z0 = eng.lift(dev["Z"], dev["inv_ls"], dev["kind"], dev["Sinv"], to(x0.T), transposed=True)
    Zr = eng.lift(dev["Z"], dev["inv_ls"], dev["kind"], dev["Sinv"], to(xr.T), transposed=True)
    if Zr.shape[0] == 1 and Z0.shape[0] > 1:
        Zr = Zr.expand(Z0.shape[0], -1).contiguous()
    Xs, Us = eng.closed_loop(to(self.A), to(self.B), to(self.C), to(K), Z0, Zr, int(num_steps))
    states = np.transpose(Xs.cpu().numpy(), (2, 0, 1))
    controls = np.transpose(Us.cpu().numpy(), (2, 0, 1))
    if single:
        return states[:, :, 0], controls[:, :, 0]
    return states, controls


KoopmanNystromRegressor.closed_loop = _closed_loop


def _lqr_closed_loop(self, K, initial_state, reference, num_steps, step):
    """Closed loop on the TRUE system with one `lift` per step: `lqr_control` of benchmark_lqr_hjb.py:74-97 and
    benchmark_lqr_classic.py:67-89 (u_i = K (phi_ref - phi(x_i)); x_{i+1} = step(x_i, u_i); phi recomputed from x_{i+1}).
    `step(x (d,1), u (p,1)) -> x_next` is the caller's simulator (host code, e.g. DuffingOscillator.update_SOM).  The lift runs
    on the device against the CACHED S^-1 (the reference re-runs scipy sqrtm inside every lift, 2000 times per seed); the state
    goes up and the lifted state comes down through pinned buffers, a few tens of microseconds per step.
    Returns (true states (d, num_steps+1), reconstructions C phi (d, num_steps), controls (p, num_steps))."""
    import torch
    x = np.asarray(initial_state, dtype=np.float64).reshape(-1, 1)
    d = x.shape[0]
    dev = self._device_state(d)
    eng = dev["eng"]
    K = np.asarray(K, dtype=np.float64)
    C = np.asarray(self.C, dtype=np.float64)
    m = dev["Z"].shape[0]
    x_pin = torch.empty(1, d, dtype=torch.float64).pin_memory()
    phi_pin = torch.empty(m, dtype=torch.float64).pin_memory()
    x_dev = torch.empty(1, d, dtype=torch.float64, device=eng.tdev)
    stream = torch.cuda.current_stream(eng.tdev)

    def lift_point(col):
        x_pin[0].copy_(torch.from_numpy(np.ascontiguousarray(col[:, 0])))
        x_dev.copy_(x_pin, non_blocking=True)
        phi = eng.lift(dev["Z"], dev["inv_ls"], dev["kind"], dev["Sinv"], x_dev, transposed=True)     # (1, m)
        phi_pin.copy_(phi[0], non_blocking=True)
        stream.synchronize()
        return phi_pin.numpy().reshape(-1, 1).copy()
    phi_ref = lift_point(np.asarray(reference, dtype=np.float64).reshape(-1, 1))
    phi = lift_point(x)
    xs = np.empty((d, num_steps + 1)); recon = np.empty((d, num_steps)); us = np.empty((K.shape[0], num_steps))
    xs[:, 0] = x[:, 0]
    for i in range(int(num_steps)):
        u = K @ (phi_ref - phi)
        us[:, i] = u[:, 0]
        recon[:, i] = (C @ phi)[:, 0]
        x = np.asarray(step(x, u), dtype=np.float64).reshape(-1, 1)
        xs[:, i + 1] = x[:, 0]
        phi = lift_point(x)
    return xs, recon, us


KoopmanNystromRegressor.lqr_closed_loop = _lqr_closed_loop


def _lqr_gain(self, Q=None, R=None, q_scale=1.0, return_all=False):
    """LQR gain of the fitted lifted model, on the device: ``control.dlqr(A, B, Q, R)[0]`` of the scripts with their defaults
    Q = q_scale * C'C symmetrised (benchmark_lqr_cloth.py:239-240 uses 0.0075, _classic.py:284 and _hjb.py:289 use 1) and
    R = I (p x p).  The Riccati equation is solved by doubling (``nys_koop_lqr_b200/dare.py``): m x m products on the FP64
    tensor-core GEMM, so a gain at m = 4096-8192 costs about a second where scipy's QZ-based solver needs hours
    (SURVEY 8f row 4).  Returns K (p, m) numpy; with ``return_all`` also the Riccati solution P (m, m) and the
    iteration record dict(iterations, delta, residual)."""
    import torch
    from . import dare
    eng = self.__dict__["_dev"]["eng"] if self.__dict__.get("_dev") else _engine()
    pend = self.__dict__.get("_pending") or {}
    dev_of = lambda name: pend[name] if name in pend else torch.from_numpy(
        np.ascontiguousarray(np.asarray(getattr(self, name), dtype=np.float64))).to(eng.tdev)
    A, B, C = dev_of("A"), dev_of("B"), dev_of("C")
    ops = dare.EngineOps(eng)
    m, p = A.shape[0], B.shape[1]
    if Q is None:
        Qd = ops.mm(C, C, ta=True) * float(q_scale)
    else:
        Qd = torch.from_numpy(np.ascontiguousarray(np.asarray(Q, dtype=np.float64))).to(eng.tdev)
    Rd = (torch.eye(p, dtype=torch.float64, device=eng.tdev) if R is None
          else torch.from_numpy(np.ascontiguousarray(np.atleast_2d(np.asarray(R, dtype=np.float64)))).to(eng.tdev))
    P, info = dare.solve_dare(A, B.contiguous(), Qd, Rd, ops=ops)
    K = dare.gain_from_solution(A, B.contiguous(), Rd, P, ops).cpu().numpy()
    self.lqr_info_ = info
    return (K, P.cpu().numpy(), info) if return_all else K


KoopmanNystromRegressor.lqr_gain = _lqr_gain


# ----------------------------------------------------------------------------------------------
# comparator baselines (regressors.py:58-111 exact-kernel, :181-234 thin-plate splines): CPU classes, outside the B200 hot
# path -- the scripts' 'splines' / exact-kernel branches instantiate them, so the drop-in module provides them
# ----------------------------------------------------------------------------------------------
from .baselines import _KernelFitMixin, _SplineFitMixin  # noqa: E402


class KoopmanKernelRegressor(_KernelFitMixin, KoopmanRegressor):
    """Exact (n x n) kernel estimator, reference regressors.py:58-111 (CPU comparator)."""

    def __init__(self, n_inputs, kernel=None, gamma=None):
        KoopmanRegressor.__init__(self, n_inputs, gamma)
        self._init_kernel_state(kernel)


class KoopmanSplineRegressor(_SplineFitMixin, KoopmanRegressor):
    """Thin-plate-spline EDMD (Korda & Mezic), reference regressors.py:181-234 (CPU comparator)."""

    def __init__(self, n_inputs, state_bounds_params=None, m=None, gamma=None):
        KoopmanRegressor.__init__(self, n_inputs, gamma, m)
        self._init_spline_state(state_bounds_params)
