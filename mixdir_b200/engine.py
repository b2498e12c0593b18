"""Thin object wrapper over the C ABI (include/mixdir_b200.h).  numpy in, numpy out."""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import _capi
from ._capi import FitParams, Timing

# the stop() text of /root/reference/R/mixdir_vi.R:78-80
UNDERFLOW_MESSAGE = (
    "There was an underflow in the calculation of zeta. Cannot continue.\n"
    "The problem probably came from the large number of columns in the input "
    "data ({ncol}). Is it possible that you want to work on t(X)?")
# the warning() texts of R/mixdir_vi.R:106-107 and R/mixdir_vi_dp.R:160-161
WARN_ELBO_TEXT = ("The ELBO decreased. This should not happen, it might be due to numerical "
                  "instabilities or a bug in the code. Please contact the maintainer to report this.\n")
WARN_DP_TEXT = ("The model put considerable weight on the last component.\n"
                "            Consider re-running the model with an increased number of latent categories.")


class MixdirError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(message)
        self.status = status


class ZetaUnderflow(MixdirError):
    pass


def _check(status, n_features=None):
    if status == _capi.OK:
        return
    msg = _capi.last_error()
    if status == _capi.UNDERFLOW:
        raise ZetaUnderflow(status, UNDERFLOW_MESSAGE.format(ncol=n_features))
    raise MixdirError(status, f"mixdir_b200 status {status}: {msg}")


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


class Comm:
    """NCCL communicator for the one-process-per-GPU flavour (mixdir_b200_comm_init);
    collective over all ranks, shared by any number of Engine handles."""

    def __init__(self, device, rank, world, unique_id=None):
        self._c = C.c_void_p()
        uid = (C.c_char * 128).from_buffer_copy(unique_id) if unique_id is not None else None
        _check(_capi.lib().mixdir_b200_comm_init(int(device), int(rank), int(world), uid,
                                                 C.byref(self._c)))
        self.device, self.rank, self.world = device, rank, world

    def close(self):
        if getattr(self, "_c", None) is not None and self._c:
            _capi.lib().mixdir_b200_comm_destroy(self._c)
            self._c = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class Engine:
    """One uploaded code matrix (the `X` of mixdir_vi, /root/reference/R/mixdir_vi.R:2)."""

    def __init__(self, codes, ncat, devices=None, dist=None):
        """codes: [N, J] uint8/uint16, 0 = NA.  ncat: C_j per feature.
        devices: list of CUDA device ids for the single-process multi-GPU flavour.
        dist: a Comm for the one-process-per-GPU flavour (codes = this rank's row shard)."""
        L = _capi.lib()
        codes = np.ascontiguousarray(codes)
        if codes.dtype not in (np.uint8, np.uint16):
            raise TypeError("codes must be uint8 or uint16")
        if codes.ndim != 2:
            raise ValueError("codes must be 2-D [rows, features]")
        self.n_rows, self.n_features = codes.shape
        self.ncat = np.ascontiguousarray(ncat, dtype=np.int32)
        if self.ncat.size != self.n_features:
            raise ValueError("len(ncat) != number of features")
        self._h = C.c_void_p()
        # the upload is asynchronous (include/mixdir_b200.h): the source must stay alive and untouched
        # until the first call that resolves it; keep it for the life of the handle
        self._codes = codes
        ip = self.ncat.ctypes.data_as(C.POINTER(C.c_int))
        if dist is not None:
            self._comm = dist  # keep the communicator alive as long as the handle
            st = L.mixdir_b200_create_dist(_ptr(codes), codes.dtype.itemsize, self.n_rows,
                                           self.n_features, ip, dist._c, C.byref(self._h))
        else:
            dev = None if devices is None else (C.c_int * len(devices))(*devices)
            st = L.mixdir_b200_create(_ptr(codes), codes.dtype.itemsize, self.n_rows,
                                      self.n_features, ip, 0 if devices is None else len(devices),
                                      dev, C.byref(self._h))
        _check(st)

    @classmethod
    def from_r_matrix(cls, X, ncat, devices=None, dist=None):
        """X: [N, J] float64, values 1..C_j, NaN = NA (the matrix mixdir() builds, R/mixdir.R:117-118).
        dist: a Comm for the one-process-per-GPU flavour (X = this rank's rows)."""
        self = cls.__new__(cls)
        L = _capi.lib()
        X = np.asfortranarray(X, dtype=np.float64)
        self._codes = None  # the library has packed and copied X when create returns
        self.n_rows, self.n_features = X.shape
        self.ncat = np.ascontiguousarray(ncat, dtype=np.int32)
        self._h = C.c_void_p()
        ip = self.ncat.ctypes.data_as(C.POINTER(C.c_int))
        if dist is not None:
            self._comm = dist
            st = L.mixdir_b200_create_dist_from_r_matrix(_ptr(X), self.n_rows, self.n_features, ip,
                                                         dist._c, C.byref(self._h))
        else:
            dev = None if devices is None else (C.c_int * len(devices))(*devices)
            st = L.mixdir_b200_create_from_r_matrix(_ptr(X), self.n_rows, self.n_features, ip,
                                                    0 if devices is None else len(devices), dev,
                                                    C.byref(self._h))
        _check(st)
        return self

    @staticmethod
    def nccl_unique_id() -> bytes:
        buf = (C.c_char * 128)()
        _check(_capi.lib().mixdir_b200_nccl_unique_id(buf))
        return bytes(buf)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h:
            _capi.lib().mixdir_b200_destroy(self._h)
            self._h = C.c_void_p()
            self._codes = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    @property
    def max_code(self):
        """largest code present = the reference's n_cat (waits for the asynchronous upload)"""
        v = _capi.lib().mixdir_b200_max_code(self._h)
        if v < 0:
            raise MixdirError(_capi.BAD_ARGUMENT, f"mixdir_b200: {_capi.last_error()}")
        return v

    @property
    def n_missing(self):
        v = _capi.lib().mixdir_b200_n_missing(self._h)
        if v < 0:
            raise MixdirError(_capi.BAD_ARGUMENT, f"mixdir_b200: {_capi.last_error()}")
        return v

    def set_engine(self, engine: int):
        _check(_capi.lib().mixdir_b200_set_engine(self._h, engine))

    def set_pairing(self, on: bool):
        """feature pairing of the fused kernel (default on)"""
        _check(_capi.lib().mixdir_b200_set_pairing(self._h, 1 if on else 0))

    def ragged_size(self, K):
        return int(K * self.ncat.sum())

    def timing(self):
        t = Timing()
        _check(_capi.lib().mixdir_b200_last_timing(self._h, C.byref(t)))
        return {f: getattr(t, f) for f, _ in Timing._fields_}

    def fit(self, n_latent, colsum_init, phi_init, dp=False, alpha=None, beta=None, max_iter=100,
            epsilon=1e-3, n_cat_bound=0, want_class_prob=True, want_pred_class=True):
        """-> dict with the fields of the reference's result list (R/mixdir_vi.R:143-155)."""
        K = int(n_latent)
        a1, a2 = 1.0, 1.0  # defaults R/mixdir_vi.R:10-17, R/mixdir_vi_dp.R:10-20
        if alpha is not None:
            al = np.atleast_1d(np.asarray(alpha, dtype=float))
            if al.size >= 1:
                a1 = float(al[0])
                a2 = float(al[1]) if (dp and al.size == 2) else float(al[0])
        b = 0.1 if beta is None else float(np.atleast_1d(beta)[0])  # R/mixdir_vi.R:19-26
        p = FitParams(K, 1 if dp else 0, a1, a2, b, int(max_iter), float(epsilon), int(n_cat_bound))
        cs = np.ascontiguousarray(colsum_init, dtype=np.float64)
        ph = np.ascontiguousarray(phi_init, dtype=np.float64)
        nrag = self.ragged_size(K)
        if cs.size != K:
            raise ValueError("colsum_init must have n_latent entries")
        if ph.size != nrag:
            # stop() of R/mixdir_vi.R:41-47
            raise ValueError("phi_init has the wrong number of elements")
        trace = np.full(max(int(max_iter), 1), np.nan)
        n_iter, conv, warn = C.c_int(0), C.c_int(0), C.c_int(0)
        elbo = C.c_double(float("nan"))
        lam = np.zeros(K)
        prior = np.zeros(2 * K)
        phi = np.zeros(nrag)
        U = np.zeros(nrag)
        cp = np.zeros((self.n_rows, K), order="F") if want_class_prob else None
        pc = np.zeros(self.n_rows, dtype=np.int32) if want_pred_class else None
        st = _capi.lib().mixdir_b200_fit(self._h, C.byref(p), _ptr(cs), _ptr(ph), _ptr(trace),
                                         C.byref(n_iter), C.byref(conv), C.byref(elbo), _ptr(lam),
                                         _ptr(prior), _ptr(phi), _ptr(U), _ptr(cp), _ptr(pc),
                                         C.byref(warn))
        _check(st, self.n_features)
        n = n_iter.value
        out = dict(converged=bool(conv.value), convergence=trace[:n].copy(), ELBO=elbo.value,
                   n_iter=n, lambda_=lam, pred_class=pc, class_prob=cp, category_prob=U, phi=phi,
                   warn_flags=warn.value)
        if dp:
            out["kappa1"], out["kappa2"] = prior[:K].copy(), prior[K:].copy()
        else:
            out["omega"] = prior[:K].copy()
        return out

    def fit_batch(self, n_latent, colsum_inits, phi_inits, dp=False, alpha=None, beta=None, max_iter=100,
                  epsilon=1e-3, n_cat_bound=0, want_class_prob=True, want_pred_class=True):
        """The `repetitions` loop of mixdir() (R/mixdir.R:120-134) in one call: one fit per row of
        colsum_inits [R, K] / phi_inits [R, ragged], small fits run concurrently on the device.
        -> (list of per-fit dicts, index of the kept fit); class_prob / pred_class are computed for
        the kept fit only."""
        K = int(n_latent)
        a1, a2 = 1.0, 1.0
        if alpha is not None:
            al = np.atleast_1d(np.asarray(alpha, dtype=float))
            if al.size >= 1:
                a1 = float(al[0])
                a2 = float(al[1]) if (dp and al.size == 2) else float(al[0])
        b = 0.1 if beta is None else float(np.atleast_1d(beta)[0])
        p = FitParams(K, 1 if dp else 0, a1, a2, b, int(max_iter), float(epsilon), int(n_cat_bound))
        cs = np.ascontiguousarray(colsum_inits, dtype=np.float64)
        ph = np.ascontiguousarray(phi_inits, dtype=np.float64)
        nrag = self.ragged_size(K)
        if cs.ndim != 2 or cs.shape[1] != K:
            raise ValueError("colsum_inits must be [repetitions, n_latent]")
        R = cs.shape[0]
        if ph.shape != (R, nrag):
            raise ValueError("phi_inits must be [repetitions, ragged size]")
        mi = max(int(max_iter), 1)
        trace = np.full((R, mi), np.nan)
        n_iter = np.zeros(R, dtype=np.int32)
        conv = np.zeros(R, dtype=np.int32)
        warn = np.zeros(R, dtype=np.int32)
        status = np.zeros(R, dtype=np.int32)
        elbo = np.full(R, np.nan)
        lam = np.zeros((R, K))
        prior = np.zeros((R, (2 if dp else 1) * K))
        phi = np.zeros((R, nrag))
        U = np.zeros((R, nrag))
        best = C.c_int(-1)
        cp = np.zeros((self.n_rows, K), order="F") if want_class_prob else None
        pc = np.zeros(self.n_rows, dtype=np.int32) if want_pred_class else None
        st = _capi.lib().mixdir_b200_fit_batch(
            self._h, C.byref(p), R, _ptr(cs), _ptr(ph), _ptr(trace), _ptr(n_iter), _ptr(conv), _ptr(elbo),
            _ptr(lam), _ptr(prior), _ptr(phi), _ptr(U), _ptr(warn), _ptr(status), C.byref(best), _ptr(cp),
            _ptr(pc))
        _check(st, self.n_features)
        outs = []
        for i in range(R):
            n = int(n_iter[i])
            o = dict(converged=bool(conv[i]), convergence=trace[i, :n].copy(), ELBO=float(elbo[i]), n_iter=n,
                     lambda_=lam[i].copy(), pred_class=None, class_prob=None, category_prob=U[i].copy(),
                     phi=phi[i].copy(), warn_flags=int(warn[i]))
            if dp:
                o["kappa1"], o["kappa2"] = prior[i, :K].copy(), prior[i, K:].copy()
            else:
                o["omega"] = prior[i].copy()
            outs.append(o)
        outs[best.value]["pred_class"] = pc
        outs[best.value]["class_prob"] = cp
        return outs, int(best.value)

    def predict(self, n_latent, lambda_, category_prob, want_class_prob=True, want_pred_class=True):
        K = int(n_latent)
        lam = np.ascontiguousarray(lambda_, dtype=np.float64)
        U = np.ascontiguousarray(category_prob, dtype=np.float64)
        if lam.size != K or U.size != self.ragged_size(K):
            raise ValueError("lambda / category_prob have the wrong size")
        cp = np.zeros((self.n_rows, K), order="F") if want_class_prob else None
        pc = np.zeros(self.n_rows, dtype=np.int32) if want_pred_class else None
        _check(_capi.lib().mixdir_b200_predict(self._h, K, _ptr(lam), _ptr(U), _ptr(cp), _ptr(pc)))
        return cp, pc

    def feature_divergence(self, n_latent, lambda_, category_prob, active, measure="JS", rows=None):
        """Quality loss per still-active feature if it were removed as well -- the inner loop of
        find_defining_features (R/predict.R:299-310) -- and, last entry, of the active set itself."""
        K = int(n_latent)
        lam = np.ascontiguousarray(lambda_, dtype=np.float64)
        U = np.ascontiguousarray(category_prob, dtype=np.float64)
        if lam.size != K or U.size != self.ragged_size(K):
            raise ValueError("lambda / category_prob have the wrong size")
        act = np.ascontiguousarray(np.asarray(active).astype(bool), dtype=np.uint8)
        if act.size != self.n_features:
            raise ValueError("active must have one flag per feature")
        m = {"JS": 0, "ARI": 1}[measure]
        out = np.zeros(self.n_features + 1)
        r = None if rows is None else np.ascontiguousarray(rows, dtype=np.int64)
        _check(_capi.lib().mixdir_b200_feature_divergence(
            self._h, K, _ptr(lam), _ptr(U), _ptr(act), m, _ptr(r), 0 if r is None else r.size, _ptr(out)))
        return out

    def estep(self, n_latent, log_prior, table, engine=0, want_zeta=False):
        """Unit-test door: one E+M pass with caller tables (see include/mixdir_b200.h)."""
        K = int(n_latent)
        lp = np.ascontiguousarray(log_prior, dtype=np.float64)
        T = np.ascontiguousarray(table, dtype=np.float64)
        S = np.zeros(self.ragged_size(K))
        cs = np.zeros(K)
        H = C.c_double(0.0)
        uf = C.c_int(0)
        zeta = np.zeros((self.n_rows, K), order="F") if want_zeta else None
        _check(_capi.lib().mixdir_b200_estep(self._h, K, _ptr(lp), _ptr(T), int(engine), _ptr(S),
                                             _ptr(cs), C.byref(H), C.byref(uf), _ptr(zeta)))
        return dict(S=S, cs=cs, H=H.value, underflow=uf.value, zeta=zeta)
