"""Generate the committed golden fixtures under tests/golden/.  Run in the BUILD container
(needs /root/reference for the mushroom data and for oracle/_ref):

    python tests/golden/make_golden.py

What is pinned, and by what:
  native_vectors.npz   inputs + outputs of the REFERENCE'S OWN native functions
                       (/root/reference/src/mixdir_impl.cpp compiled verbatim, oracle/_ref):
                       update_zeta_cpp, update_zeta_dp_cpp, expec_log_x_cpp -- true reference
                       outputs (with the oracle's digamma standing in for R::digamma).
  mushroom.npz         the reference's data/mushroom.rda decoded (oracle/rda.py) and factor-coded
                       like mixdir() does (R/mixdir.R:97-118); cross-checked here against
                       data-raw/agaricus-lepiota.data.txt.
  fit_*.npz            full fits by the oracle's literal restatement of mixdir_vi / mixdir_vi_dp
                       (oracle/mixdir_oracle.c), whose native steps are bit-identical to
                       oracle/_ref (asserted below).  The R-level loop itself cannot be run
                       here (no R): these fixtures pin the restatement, not R.
Initial values come from mixdir_b200.synth (pure hash of a seed), so only the seed is stored.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from mixdir_b200 import synth  # noqa: E402
from mixdir_b200.coding import code_columns  # noqa: E402
from oracle import pyoracle as po  # noqa: E402
from oracle.rda import read_mushroom  # noqa: E402

OUT = os.path.dirname(os.path.abspath(__file__))


def save(name, **kw):
    np.savez_compressed(os.path.join(OUT, name), **kw)
    print("wrote", name, {k: getattr(v, "shape", None) for k, v in kw.items()})


def fit_fixture(name, codes, ncat, K, dp, seed, max_iter=100, epsilon=1e-3, alpha=(1.0, 1.0),
                beta=0.1, keep_rows=None, store_codes=False):
    X = po.codes_to_r_matrix(codes)
    N = codes.shape[0]
    z0 = synth.zeta_init(seed, N, K)
    p0 = synth.phi_init(seed, ncat, K)
    r = po.fit(X, ncat, K, dp, alpha[0], alpha[1], beta, max_iter, epsilon, z0, p0)
    rows = np.arange(N) if keep_rows is None else np.asarray(keep_rows)
    kw = dict(ncat=np.asarray(ncat, dtype=np.int32), K=K, dp=dp, seed=seed, max_iter=max_iter,
              epsilon=epsilon, alpha=np.asarray(alpha), beta=beta, status=r["status"],
              converged=r["converged"], convergence=r["convergence"], ELBO=r["ELBO"],
              lambda_=r["lambda_"], pred_class=r["pred_class"], rows=rows,
              class_prob_rows=r["class_prob"][rows], category_prob=r["category_prob"],
              phi=r["phi"], warn_flags=r["warn_flags"],
              prior=np.concatenate([r["kappa1"], r["kappa2"]]) if dp else r["omega"])
    if store_codes:
        kw["codes"] = codes
    save(name, **kw)
    print("   ", name, "iters", r["n_iter"], "converged", r["converged"], "ELBO", r["ELBO"],
          "warn", r["warn_flags"])
    return r


def main():
    assert po.ref() is not None, "oracle/_ref is needed (build container)"

    # ---- mushroom ---------------------------------------------------------
    names, cols = read_mushroom()
    codes, cats = code_columns(cols)
    # cross-check against the raw UCI file the reference built the .rda from
    raw = open("/root/reference/data-raw/agaricus-lepiota.data.txt").read().strip().split("\n")
    hdr = raw[0].split(",")
    rawrows = [l.split(",") for l in raw[1:]]
    assert len(rawrows) == codes.shape[0] == 8124
    for j, nm in enumerate(names):
        col = [r[hdr.index(nm)] for r in rawrows]
        # same partition of the rows: letter code <-> level must be a bijection
        pairs = {(c, cols[j][i]) for i, c in enumerate(col)}
        assert len({p[0] for p in pairs}) == len(pairs), nm
    assert sum(v is None for v in cols[names.index("stalk-root")]) == 2480
    save("mushroom.npz", codes=codes, names=np.array(names),
         categories=np.array(json.dumps(cats)))

    # ---- the reference's native functions: true reference outputs ----------
    rng_rows = (synth.hash_u64(11, 9, np.arange(8124)) % np.uint64(8124)).astype(int)[:300]
    sub = [[c[i] for i in rng_rows] for c in cols]
    scodes, scats = code_columns(sub)
    sncat = np.array([len(c) for c in scats], dtype=np.int32)
    X = po.codes_to_r_matrix(scodes)
    n_cat = int(np.nanmax(X))
    for K, tag in ((3, "k3"), (10, "k10")):
        phi = synth.phi_init(21, sncat, K) + 0.1 * synth.hash_u01(22, 5, np.arange(int(K * sncat.sum())))
        phia = po.phi_to_array(phi, sncat, K)
        omega = 1.0 + 300.0 * synth.hash_u01(23, 6, np.arange(K))
        kappa2 = 1.0 + 300.0 * synth.hash_u01(24, 7, np.arange(K))
        zu = po.update_zeta(X, phia, omega, n_cat, "ref")
        zd = po.update_zeta_dp(X, phia, omega, kappa2, n_cat, "ref")
        assert np.array_equal(zu, po.update_zeta(X, phia, omega, n_cat, "oracle"))
        assert np.array_equal(zd, po.update_zeta_dp(X, phia, omega, kappa2, n_cat, "oracle"))
        zn = zu / zu.sum(axis=1, keepdims=True)
        ex = po.expec_log_x(X, phia, zn, n_cat, "ref")
        assert ex == po.expec_log_x(X, phia, zn, n_cat, "oracle")
        if tag == "k3":
            nat = dict(codes=scodes, ncat=sncat)
        nat.update({f"phi_{tag}": phi, f"omega_{tag}": omega, f"kappa2_{tag}": kappa2,
                    f"zeta_{tag}": zu, f"zeta_dp_{tag}": zd, f"zeta_norm_{tag}": zn,
                    f"expec_log_x_{tag}": ex})
    save("native_vectors.npz", **nat)

    # ---- M1: mushroom[1:1000, 1:5], n_latent=3 (README example) --------------
    m1cols = [c[:1000] for c in cols[:5]]
    m1codes, m1cats = code_columns(m1cols)
    m1ncat = np.array([len(c) for c in m1cats], dtype=np.int32)
    assert list(m1ncat) == [2, 5, 4, 3, 2]
    for seed in (1, 2, 3, 4):
        fit_fixture(f"fit_m1_seed{seed}.npz", m1codes, m1ncat, 3, 0, seed)

    # ---- M2: full mushroom, Dirichlet-process variant ----------------------
    ncat = np.array([len(c) for c in cats], dtype=np.int32)
    keep = np.arange(0, 8124, 9)
    fit_fixture("fit_m2_dp_k10.npz", codes, ncat, 10, 1, 1, keep_rows=keep)
    fit_fixture("fit_m2_dp_k20.npz", codes, ncat, 20, 1, 2, keep_rows=keep)
    fit_fixture("fit_m2_k5.npz", codes, ncat, 5, 0, 4, keep_rows=keep)  # test-mixdir.R:42-46

    # ---- toy data of the reference tests (helper.R:2-48) ---------------------
    tcols = [["ABC"[v - 1] for v in synth.TOY_DATA[:, j]] for j in range(3)]
    tcodes, tcats = code_columns(tcols)
    tncat = np.array([len(c) for c in tcats], dtype=np.int32)
    fit_fixture("fit_toy_k3.npz", tcodes, tncat, 3, 0, 1, store_codes=True)
    fit_fixture("fit_toy_dp_k10.npz", tcodes, tncat, 10, 1, 1, store_codes=True)
    tna = tcodes.copy()
    tna[1, 0] = tna[4, 2] = tna[8, 1] = 0  # three NA cells (test-mixdir.R:21-34)
    fit_fixture("fit_toy_na_k3.npz", tna, tncat, 3, 0, 1, store_codes=True)
    fit_fixture("fit_toy_na_dp_k10.npz", tna, tncat, 10, 1, 1, store_codes=True)

    # ---- S3-like synthetic (scaled down): 3000 x 60, C=10, K=10, 10% NA ------
    sn = np.full(60, 10, dtype=np.int32)
    scod, _ = synth.make_codes(1, 3000, sn, 10, p_na=0.10)
    fit_fixture("fit_s3_small.npz", scod, sn, 10, 0, 1, max_iter=30, epsilon=-np.inf,
                keep_rows=np.arange(0, 3000, 5), store_codes=True)
    fit_fixture("fit_s3_small_dp.npz", scod, sn, 10, 1, 1, max_iter=30, epsilon=-np.inf,
                keep_rows=np.arange(0, 3000, 5), store_codes=True)
    # ---- S4-like shape (scaled down): ragged C_j, K=32 -------------------------
    pat = np.array([16, 12, 8, 4, 2, 10] * 10, dtype=np.int32)
    c4, _ = synth.make_codes(2, 2000, pat, 32, p_na=0.0)
    fit_fixture("fit_s4_small.npz", c4, pat, 32, 0, 2, max_iter=15, epsilon=-np.inf,
                keep_rows=np.arange(0, 2000, 5), store_codes=True)
    # ---- trailing unused level: n_cat = max(X) < dim(phia)[3] quirk ----
    qn = np.array([3, 4, 6], dtype=np.int32)  # feature 2 declares 6 levels, only 1..4 observed
    qc, _ = synth.make_codes(3, 400, np.array([3, 4, 4]), 3)
    fit_fixture("fit_unused_level.npz", qc, qn, 3, 0, 3, store_codes=True)


if __name__ == "__main__":
    main()
