"""TEST INFRASTRUCTURE ONLY: ctypes doors into the CPU oracle.

  liboracle  = oracle/libmixdir_oracle.so        C restatement (oracle/mixdir_oracle.c)
  libref     = oracle/_ref/libmixdir_ref.so      the reference's own src/mixdir_impl.cpp,
                                                 compiled verbatim against oracle/shim/Rcpp.h

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
reference legs may import this module; the product package never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_dp = C.POINTER(C.c_double)
_ip = C.POINTER(C.c_int)


def build(quiet: bool = True) -> None:
    """Compile the oracle (and oracle/_ref when /root/reference is present)."""
    subprocess.run(["make", "-C", _HERE], check=True,
                   stdout=subprocess.DEVNULL if quiet else None)


def _load(path: str):
    if not os.path.exists(path):
        build()
    return C.CDLL(path)


_lib = None
_ref = None


def lib():
    global _lib
    if _lib is None:
        _lib = _load(os.path.join(_HERE, "libmixdir_oracle.so"))
        _lib.oracle_digamma.restype = C.c_double
        _lib.oracle_digamma.argtypes = [C.c_double]
        _lib.oracle_expec_log_x.restype = C.c_double
    return _lib


def ref():
    """The reference TU; None if it was never built (e.g. no /root/reference and no prebuilt)."""
    global _ref
    if _ref is None:
        p = os.path.join(_HERE, "_ref", "libmixdir_ref.so")
        if not os.path.exists(p):
            try:
                build()
            except Exception:
                return None
        if not os.path.exists(p):
            return None
        _ref = C.CDLL(p)
        _ref.ref_expec_log_x.restype = C.c_double
    return _ref


def _d(a):
    return a.ctypes.data_as(_dp)


def _f64(a, order="F"):
    return np.require(np.asarray(a, dtype=np.float64), requirements=["ALIGNED", "WRITEABLE"]).copy(order=order)


def digamma(x):
    f = lib().oracle_digamma
    return np.vectorize(lambda v: f(float(v)), otypes=[float])(x)


def phi_to_array(phi_rag, ncat, K):
    """R/mixdir_vi.R:196-207 conv_phi_to_array: ragged (j,k,r) -> [J,K,Cmax] col-major, NaN padded."""
    J = len(ncat)
    cdim = int(max(ncat))
    phia = np.full((J, K, cdim), np.nan, order="F")
    o = 0
    for j in range(J):
        c = int(ncat[j])
        phia[j, :, :c] = np.asarray(phi_rag[o:o + K * c]).reshape(K, c)
        o += K * c
    return phia


def update_zeta(X, phia, omega, n_cat, which="oracle"):
    """un-normalised zeta (src/mixdir_impl.cpp:57-87).  which = 'oracle' | 'ref'."""
    X = _f64(X)
    N, J = X.shape
    phia = _f64(phia)
    K, cdim = phia.shape[1], phia.shape[2]
    omega = _f64(omega)
    zeta = np.zeros((N, K), order="F")
    if which == "ref":
        ref().ref_update_zeta(_d(zeta), _d(X), _d(phia), cdim, _d(omega), N, J, K, int(n_cat))
    else:
        lib().oracle_update_zeta(_d(zeta), _d(X), _d(phia), cdim, _d(omega), N, J, K, int(n_cat))
    return zeta


def update_zeta_dp(X, phia, kappa1, kappa2, n_cat, which="oracle"):
    X = _f64(X)
    N, J = X.shape
    phia = _f64(phia)
    K, cdim = phia.shape[1], phia.shape[2]
    k1, k2 = _f64(kappa1), _f64(kappa2)
    zeta = np.zeros((N, K), order="F")
    if which == "ref":
        ref().ref_update_zeta_dp(_d(zeta), _d(X), _d(phia), cdim, _d(k1), _d(k2), N, J, K, int(n_cat))
    else:
        lib().oracle_update_zeta_dp(_d(zeta), _d(X), _d(phia), cdim, _d(k1), _d(k2), N, J, K, int(n_cat))
    return zeta


def expec_log_x(X, phia, zeta, n_cat, which="oracle"):
    X = _f64(X)
    N, J = X.shape
    phia = _f64(phia)
    K, cdim = phia.shape[1], phia.shape[2]
    zeta = _f64(zeta)
    if which == "ref":
        return ref().ref_expec_log_x(_d(X), N, J, _d(phia), cdim, _d(zeta), K, int(n_cat))
    return lib().oracle_expec_log_x(_d(X), N, _d(phia), cdim, _d(zeta), J, K, int(n_cat))


def codes_to_r_matrix(codes):
    """compact codes (0 = NA) -> the reference's double matrix (NaN = NA), column-major."""
    X = np.asarray(codes).astype(np.float64)
    X[X == 0] = np.nan
    return np.asfortranarray(X)


def fit(X, ncat, K, dp, alpha1, alpha2, beta, max_iter, epsilon, zeta_init, phi_init,
        which="oracle", final_pass=True):
    """Literal restatement of mixdir_vi / mixdir_vi_dp; returns a dict like the R result list.
    which='ref': same loop, but the three native steps run the reference's own
    src/mixdir_impl.cpp code (oracle/_ref)."""
    L = lib()
    if which == "ref":
        L = ref()
        L.ref_install_hooks(1)
    X = _f64(X)
    N, J = X.shape
    ncat = np.ascontiguousarray(ncat, dtype=np.int32)
    nphi = int(K * ncat.sum())
    zeta_init = _f64(zeta_init)
    phi_init = np.ascontiguousarray(phi_init, dtype=np.float64)
    assert zeta_init.shape == (N, K) and phi_init.size == nphi
    hist = np.full(max(max_iter, 1), np.nan)
    n_iter, conv, warn = C.c_int(0), C.c_int(0), C.c_int(0)
    lam = np.zeros(K)
    cp = np.zeros((N, K), order="F") if final_pass else None
    pc = np.zeros(N, dtype=np.int32)
    U = np.zeros(nphi)
    prior = np.zeros(2 * K)
    phi = np.zeros(nphi)
    st = L.oracle_fit(_d(X), N, J, ncat.ctypes.data_as(_ip), K, int(dp), C.c_double(alpha1),
                          C.c_double(alpha2), C.c_double(beta), int(max_iter), C.c_double(epsilon),
                          _d(zeta_init), _d(phi_init), _d(hist), C.byref(n_iter), C.byref(conv),
                          _d(lam), _d(cp) if final_pass else None, pc.ctypes.data_as(_ip), _d(U),
                   _d(prior), _d(phi), C.byref(warn))
    n = n_iter.value
    out = dict(status=st, converged=bool(conv.value), convergence=hist[:n].copy(),
               ELBO=float(hist[n - 1]) if n else float("nan"), n_iter=n, lambda_=lam,
               pred_class=pc, class_prob=cp, category_prob=U, phi=phi, warn_flags=warn.value)
    if dp:
        out["kappa1"], out["kappa2"] = prior[:K].copy(), prior[K:].copy()
    else:
        out["omega"] = prior[:K].copy()
    return out


def predict(X, ncat, K, lambda_, U):
    X = _f64(X)
    N, J = X.shape
    ncat = np.ascontiguousarray(ncat, dtype=np.int32)
    lam = np.ascontiguousarray(lambda_, dtype=np.float64)
    U = np.ascontiguousarray(U, dtype=np.float64)
    cp = np.zeros((N, K), order="F")
    pc = np.zeros(N, dtype=np.int32)
    lib().oracle_predict(_d(X), N, J, ncat.ctypes.data_as(_ip), K, _d(lam), _d(U), _d(cp),
                         pc.ctypes.data_as(_ip))
    return cp, pc


def estep_stats(codes, ncat, K, logprior, T, row0=0, row1=None, want_zeta=False):
    """Fused-state pass (oracle_estep_stats).  logprior INCLUDES the reference's -1."""
    codes = np.ascontiguousarray(codes)
    assert codes.dtype in (np.uint8, np.uint16)
    N, J = codes.shape
    row1 = N if row1 is None else row1
    ncat = np.ascontiguousarray(ncat, dtype=np.int32)
    nphi = int(K * ncat.sum())
    lp = np.ascontiguousarray(logprior, dtype=np.float64)
    T = np.ascontiguousarray(T, dtype=np.float64)
    S = np.zeros(nphi)
    cs = np.zeros(K)
    H = C.c_double(0.0)
    uf = C.c_int(0)
    zeta = np.zeros((row1 - row0, K), order="F") if want_zeta else None
    lib().oracle_estep_stats(codes.ctypes.data_as(C.c_void_p), codes.dtype.itemsize,
                             C.c_int64(row0), C.c_int64(row1), J, ncat.ctypes.data_as(_ip), K,
                             _d(lp), _d(T), _d(S), _d(cs), C.byref(H), C.byref(uf),
                             _d(zeta) if want_zeta else None)
    return dict(S=S, cs=cs, H=H.value, underflow=uf.value, zeta=zeta)


def adjusted_rand(c1, c2):
    """mcclust::arandi(cl1, cl2, adjust=TRUE) (package not vendored with the reference; the
    published Hubert-Arabie formula it implements)."""
    c1 = np.asarray(c1)
    c2 = np.asarray(c2)
    _, i1 = np.unique(c1, return_inverse=True)
    _, i2 = np.unique(c2, return_inverse=True)
    tab = np.zeros((i1.max() + 1, i2.max() + 1))
    np.add.at(tab, (i1, i2), 1)

    def ch2(x):
        return x * (x - 1) / 2.0
    correc = ch2(tab.sum(1)).sum() * ch2(tab.sum(0)).sum() / ch2(float(c1.size))
    with np.errstate(invalid="ignore", divide="ignore"):
        return (ch2(tab).sum() - correc) / (0.5 * ch2(tab.sum(1)).sum() + 0.5 * ch2(tab.sum(0)).sum() - correc)


def r_order(values):
    """R's order(): ascending, stable, NA/NaN last."""
    v = np.asarray(values, dtype=float)
    idx = np.arange(v.size)
    nan = np.isnan(v)
    return np.concatenate([idx[~nan][np.argsort(v[~nan], kind="stable")], idx[nan]])


def find_defining_features(X, ncat, K, lambda_, U, subsample, n_features=np.inf, measure="JS",
                           step_size=np.inf, exponential_decay=True):
    """Literal restatement of find_defining_features (/root/reference/R/predict.R:266-339) on the
    numeric code matrix X (NaN = NA): every candidate is a full predict_class call with the dropped
    columns all-NA (R/predict.R:66-67: a missing column is a column of NAs).  Features are column
    indices; `subsample` replaces R's sample().  Returns (features, quality, rounds) where rounds
    is the list of per-round divergence vectors (for the tests)."""
    X = np.asarray(X, dtype=np.float64)
    J = X.shape[1]
    exponential_step = 2.0 if exponential_decay is True else (
        float(exponential_decay) if not isinstance(exponential_decay, bool) else np.nan)
    early_stop = 0 if np.isinf(n_features) else int(n_features)
    Xs = X[np.asarray(subsample)]

    def pred(cols):
        Z = np.full_like(Xs, np.nan)
        Z[:, cols] = Xs[:, cols]
        return predict(Z, ncat, K, lambda_, U)[0]

    def quality_of(q, p):
        if measure == "JS":  # R/predict.R:304
            with np.errstate(invalid="ignore", divide="ignore"):
                return float(np.sum(q * (np.log(q) - np.log(p)) + p * (np.log(p) - np.log(q))) / 2)
        return float(adjusted_rand(np.argmax(q, axis=1), np.argmax(p, axis=1)))

    p = pred(list(range(J)))
    remaining = list(range(J))
    removed, quality, rounds = [], [], []
    diverg = np.zeros(0)
    while len(remaining) > early_stop:
        diverg = np.zeros(len(remaining))
        for n, i in enumerate(remaining):  # R/predict.R:299-310
            v = quality_of(pred([c for c in remaining if c != i]), p)
            diverg[n] = v if measure == "JS" else 1 - v
        rounds.append((list(remaining), diverg.copy()))
        exp_step = np.floor(len(remaining) * (1 - 1 / exponential_step)) if exponential_decay else np.inf
        last_rm_idx = int(max(min(step_size, exp_step, len(remaining) - early_stop), 1))
        order = r_order(diverg)[:last_rm_idx]
        removed += [remaining[i] for i in order]
        quality += [float(diverg[i]) for i in order]
        remaining = [j for j in range(J) if j not in removed]
    if measure == "ARI":
        quality = [1 - q for q in quality]
    if not np.isinf(n_features) and n_features != 0:
        q = quality_of(pred(remaining), p)
        head = -np.asarray(diverg[:len(remaining)], dtype=float)
        return [remaining[i] for i in r_order(head)], q, rounds
    return (removed + remaining)[::-1], quality[::-1], rounds
