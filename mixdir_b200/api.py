"""Python mirror of the reference's user entry points, on top of the CUDA engine.

  mixdir()                  <->  /root/reference/R/mixdir.R:79-138
  predict()                 <->  /root/reference/R/predict.R:29-105
  find_defining_features()  <->  /root/reference/R/predict.R:266-339

R is not available in this image, so this module plays the role of the R adapter
(r_adapter/ holds the Rcpp/R sources a maintainer would add):
same argument names, defaults, result fields, errors and warnings.  All numeric
work happens in the C-ABI library; nothing here falls back to the CPU.
"""
from __future__ import annotations

import warnings

import numpy as np

from . import _capi
from .coding import code_columns, recode_for_predict
from .engine import WARN_DP_TEXT, WARN_ELBO_TEXT, Engine


def _as_columns(X):
    """data.frame-like input -> (names, list of columns)."""
    if hasattr(X, "columns") and hasattr(X, "__getitem__") and not isinstance(X, np.ndarray):
        names = [str(c) for c in X.columns]
        return names, [list(X[c]) for c in X.columns]
    if isinstance(X, dict):
        names = list(X.keys())
        cols = [list(v) if isinstance(v, (list, tuple, np.ndarray)) else [v] for v in X.values()]
        return names, cols
    A = np.asarray(X, dtype=object)
    if A.ndim == 1:
        A = A.reshape(1, -1)
    names = [f"V{j + 1}" for j in range(A.shape[1])]
    return names, [list(A[:, j]) for j in range(A.shape[1])]


def default_inits(n_rows, ncat, n_latent, rng):
    """The reference's random initial values (R/mixdir_vi.R:28-39): zeta_init ~ Dirichlet(1)
    rows, phi_init in {1,2,3}.  R's RNG stream is not reproducible outside R; parity tests
    pass explicit inits instead."""
    zeta = rng.dirichlet(np.ones(n_latent), size=n_rows)
    phi = rng.integers(1, 4, size=int(n_latent * np.sum(ncat))).astype(np.float64)
    return zeta, phi


def mixdir(X, n_latent=3, alpha=None, beta=None, select_latent=False, max_iter=100, epsilon=1e-3,
           na_handle="ignore", repetitions=1, zeta_init=None, phi_init=None, seed=None,
           devices=None, verbose=False):
    """Cluster categorical data; returns a dict with the reference's result fields
    (converged, convergence, ELBO, lambda, pred_class, class_prob, category_prob,
    specific_params, na_handle) -- R/mixdir.R:79-138, R/mixdir_vi.R:143-155."""
    names, cols = _as_columns(X)
    codes, categories = code_columns(cols, na_handle=na_handle)
    ncat = np.array([len(c) for c in categories], dtype=np.int32)
    n_rows = codes.shape[0]
    rng = np.random.default_rng(seed)
    K = int(n_latent)
    R = int(repetitions)
    inits = []
    for _rep in range(R):  # one set of initial values per repetition (R/mixdir_vi.R:28-39)
        z0, p0 = default_inits(n_rows, ncat, K, rng)
        if zeta_init is not None:
            z0 = np.asarray(zeta_init, dtype=float)
        if phi_init is not None:
            p0 = np.asarray(phi_init, dtype=float)
        inits.append((z0.sum(axis=0), p0))
    best = None
    with Engine(codes, ncat, devices=devices) as eng:
        # the repetitions loop (R/mixdir.R:120-134) is one library call: the restarts share the
        # uploaded code matrix and, when small, the GPU
        if R == 1:
            fits, keep = [eng.fit(K, inits[0][0], inits[0][1], dp=bool(select_latent), alpha=alpha, beta=beta,
                                  max_iter=max_iter, epsilon=epsilon)], 0
        else:
            fits, keep = eng.fit_batch(K, np.stack([c for c, _ in inits]), np.stack([p for _, p in inits]),
                                       dp=bool(select_latent), alpha=alpha, beta=beta, max_iter=max_iter,
                                       epsilon=epsilon)
        for res in fits:
            if res["warn_flags"] & _capi.WARN_ELBO_DECREASED:
                warnings.warn(WARN_ELBO_TEXT)
            if res["warn_flags"] & _capi.WARN_DP_LAST_COMPONENT:
                warnings.warn(WARN_DP_TEXT)
            if verbose:
                for i, e in enumerate(res["convergence"], 1):
                    if i % 10 == 0:  # R/mixdir_vi.R:111
                        print(f"Iter: {i} ELBO: {e:.4g}")
        best = fits[keep]
    return _as_result(best, names, categories, K, bool(select_latent), na_handle)


def _ragged_to_lists(flat, categories, K):
    """ragged (j,k,r) -> list[J] of list[K] of dict(level -> value) (named vectors in R)."""
    out, o = [], 0
    for cats in categories:
        c = len(cats)
        out.append([dict(zip(cats, flat[o + k * c:o + (k + 1) * c])) for k in range(K)])
        o += K * c
    return out


def _as_result(res, names, categories, K, dp, na_handle):
    cat_prob = dict(zip(names, _ragged_to_lists(res["category_prob"], categories, K)))
    phi = _ragged_to_lists(res["phi"], categories, K)
    spec = dict(kappa1=res["kappa1"], kappa2=res["kappa2"], phi=phi) if dp else \
        dict(omega=res["omega"], phi=phi)
    return {
        "converged": res["converged"], "convergence": res["convergence"], "ELBO": res["ELBO"],
        "lambda": res["lambda_"], "pred_class": res["pred_class"], "class_prob": res["class_prob"],
        "category_prob": cat_prob, "specific_params": spec, "na_handle": na_handle,
        "_names": names, "_categories": categories, "_U_flat": res["category_prob"],
    }


def predict(obj, newdata=None):
    """predict.mixdir (R/predict.R:29-36): class_prob of the training data, or the posterior
    class probabilities of new rows under (lambda, category_prob)."""
    if newdata is None:
        return obj["class_prob"]
    categories = obj["_categories"]
    codes = _recode_new(obj, newdata)
    ncat = np.array([len(c) for c in categories], dtype=np.int32)
    K = len(obj["lambda"])
    with Engine(codes, ncat) as eng:
        cp, _ = eng.predict(K, obj["lambda"], obj["_U_flat"], want_pred_class=False)
    return cp


def _recode_new(obj, newdata):
    """name matching, unseen-level warning and NA handling of predict_class (R/predict.R:64-84)."""
    names, categories = obj["_names"], obj["_categories"]
    nd_names, nd_cols = _as_columns(newdata)
    # columns the model does not know are ignored: predict_class rebuilds X from names(categories)
    # only (R/predict.R:64-86), which makes its later stop() for unknown columns (:95-96) unreachable
    by_name = dict(zip(nd_names, nd_cols))
    codes, unseen = recode_for_predict(by_name, names, categories, obj.get("na_handle") or "ignore")
    for col, vals in unseen.items():  # warning() of R/predict.R:75-80
        lst = ",".join(f'"{v}"' for v in sorted(vals))
        warnings.warn(f'The new data has a response ({lst}) for category "{col}" which is not in '
                      "category_prob. Handling X[j, i] it as if it was missing.\n")
    return codes


def _r_order(values):
    """R's order(): ascending, stable, NA/NaN last."""
    v = np.asarray(values, dtype=float)
    idx = np.arange(v.size)
    nan = np.isnan(v)
    return np.concatenate([idx[~nan][np.argsort(v[~nan], kind="stable")], idx[nan]])


def find_defining_features(obj, X, n_features=np.inf, measure="JS", subsample_size=np.inf,
                           step_size=np.inf, exponential_decay=True, verbose=False, subsample=None,
                           seed=None, devices=None):
    """find_defining_features (R/predict.R:266-339): greedy feature elimination by the loss of
    clustering quality.  Returns dict(features=..., quality=...).  The control flow below is the
    reference's; each round's |remaining| predict() calls are one device pass
    (Engine.feature_divergence).  `subsample` (0-based row indices) replaces R's sample(), whose
    stream is not reproducible outside R."""
    if measure not in ("JS", "ARI"):  # match.arg
        raise ValueError("'arg' should be one of \"JS\", \"ARI\"")
    if exponential_decay is True:
        exponential_step = 2.0
    elif isinstance(exponential_decay, (int, float)) and not isinstance(exponential_decay, bool):
        exponential_step = float(exponential_decay)
    else:
        exponential_step = np.nan
    early_stop = 0 if np.isinf(n_features) else int(n_features)
    codes = _recode_new(obj, X)
    names = obj["_names"]
    J, n = len(names), codes.shape[0]
    if subsample is None:
        rng = np.random.default_rng(seed)
        size = int(min(n, subsample_size))
        subsample = rng.permutation(n)[:size]  # sample(seq_len(nrow(X)), size=..., replace=FALSE)
    subsample = np.asarray(subsample, dtype=np.int64)
    ncat = np.array([len(c) for c in obj["_categories"]], dtype=np.int32)
    K = len(obj["lambda"])
    remaining = list(range(J))
    removed, quality = [], []
    diverg = np.zeros(0)
    with Engine(codes, ncat, devices=devices) as eng:
        def divergences(cols):
            active = np.zeros(J, dtype=np.uint8)
            active[cols] = 1
            return eng.feature_divergence(K, obj["lambda"], obj["_U_flat"], active, measure,
                                          rows=subsample)
        while len(remaining) > early_stop:
            diverg = divergences(remaining)[remaining]  # R/predict.R:299-310
            exp_step = (np.floor(len(remaining) * (1 - 1 / exponential_step))
                        if exponential_decay else np.inf)
            last_rm_idx = int(max(min(step_size, exp_step, len(remaining) - early_stop), 1))
            order = _r_order(diverg)[:last_rm_idx]
            removed += [remaining[i] for i in order]
            quality += [float(diverg[i]) for i in order]
            remaining = [j for j in range(J) if j not in removed]  # setdiff(colnames(X), removed_vars)
            if verbose:
                print(f"Remaining variables: {len(remaining)}")
        if measure == "ARI":
            quality = [1 - q for q in quality]
        if not np.isinf(n_features) and n_features != 0:
            if diverg.size == 0:  # n_features >= ncol(X): the reference stops here as well
                raise ValueError("object 'diverg' not found")
            last = divergences(remaining)[J]  # predict with exactly the remaining variables
            q = float(last) if measure == "JS" else float(1 - last)
            # "Not perfect order but better than random" (R/predict.R:333)
            head = -np.asarray(diverg[:len(remaining)], dtype=float)
            feats = [names[remaining[i]] for i in _r_order(head)]
            return {"features": feats, "quality": q}
    feats = [names[j] for j in (removed + remaining)][::-1]
    return {"features": feats, "quality": quality[::-1]}
