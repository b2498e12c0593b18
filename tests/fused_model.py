"""TEST INFRASTRUCTURE: numpy model of the engine's fused-state iteration.

Mirrors, step by step, what the CUDA kernels in mixdir_b200/csrc/kernels.cuh compute
(k_prior -> E+M pass -> [all-reduce] -> k_param -> k_finalize -> k_results), on top of the
oracle's row pass (oracle_estep_stats).  Used on the CPU to check that the fused-state algebra
(state = class masses + phi; D-term = sum S*T + K*nNA; DP re-ordering applied to the reduced
statistics) reproduces the literal restatement of the reference loop, and to emulate several
ranks by splitting rows and summing the partial statistics buffers.
"""
import math

import numpy as np

from oracle import pyoracle as po


def _rag_views(flat, ncat, K):
    out, o = [], 0
    for c in ncat:
        out.append(flat[o:o + K * c].reshape(K, c))
        o += K * c
    return out


def table_from_phi(phi, ncat, K, n_cat_bound):
    T = np.zeros_like(phi)
    for pj, tj, c in zip(_rag_views(phi, ncat, K), _rag_views(T, ncat, K), ncat):
        for k in range(K):
            sb = 0.0
            for r in range(min(c, n_cat_bound)):
                sb += pj[k, r]
            tj[k, :] = po.digamma(pj[k, :]) - po.digamma(sb)
    return T


def prior_step(cs, K, dp, a1, a2):
    lg = math.lgamma
    if not dp:
        omega = a1 + cs
        so = 0.0
        for v in omega:
            so += v
        e1 = po.digamma(omega) - po.digamma(so)
        A = lg(K * a1) - K * lg(a1) + float(np.sum((a1 - 1) * e1))
        E = -lg(so) + sum(lg(v) for v in omega) - float(np.sum((omega - 1) * e1))
        return dict(omega=omega, e1=e1, e2=None, logprior=e1.copy(), A=A, E=E)
    k1 = a1 + cs
    tail = np.concatenate([np.cumsum(cs[::-1])[::-1][1:], [0.0]])
    k2 = a2 + tail
    d1, d2, d12 = po.digamma(k1), po.digamma(k2), po.digamma(k1 + k2)
    e1, e2 = d1 - d12, d2 - d12
    logprior = e1 + np.concatenate([[0.0], np.cumsum(e2)[:-1]])
    A = float(np.sum((a1 - 1) * e1 + (a2 - 1) * e2))
    E = sum(lg(a) + lg(b) - lg(a + b) for a, b in zip(k1, k2)) - float(
        np.sum((k1 - 1) * d1 + (k2 - 1) * d2 - (k1 + k2 - 2) * d12))
    return dict(omega=k1, kappa2=k2, e1=e1, e2=e2, logprior=logprior, A=A, E=E)


def fit(codes, ncat, K, dp, a1, a2, beta, max_iter, epsilon, colsum_init, phi_init,
        n_shards=1, n_cat_bound=None):
    codes = np.ascontiguousarray(codes)
    N, J = codes.shape
    ncat = np.asarray(ncat, dtype=np.int32)
    if n_cat_bound is None:
        n_cat_bound = int(codes.max())
    n_missing = int((codes == 0).sum())
    lg = math.lgamma
    cs = np.asarray(colsum_init, dtype=float).copy()
    phi = np.asarray(phi_init, dtype=float).copy()
    T = table_from_phi(phi, ncat, K, n_cat_bound)
    trace, conv, warn = [], False, 0
    bounds = [N * d // n_shards for d in range(n_shards + 1)]
    for it in range(max_iter):
        pr = prior_step(cs, K, dp, a1, a2)
        S = np.zeros_like(phi)
        csn = np.zeros(K)
        H = 0.0
        under = 0
        for d in range(n_shards):  # one partial buffer per rank, then the "all-reduce"
            st = po.estep_stats(codes, ncat, K, pr["logprior"] - 1.0, T, bounds[d], bounds[d + 1])
            S += st["S"]
            csn += st["cs"]
            H += st["H"]
            under |= st["underflow"]
        if under:
            return dict(status=1, convergence=np.array(trace), n_iter=it + 1)
        perm = np.argsort(-csn, kind="stable") if dp else np.arange(K)
        csp = csn[perm]
        tC = tD = tG = 0.0
        o = 0
        new_phi = np.zeros_like(phi)
        for j, c in enumerate(ncat):
            Sj = S[o:o + K * c].reshape(K, c)[perm]
            pj = Sj + beta
            new_phi[o:o + K * c] = pj.ravel()
            for k in range(K):
                sa = sb = 0.0
                for r in range(c):
                    sa += pj[k, r]
                    if r < n_cat_bound:
                        sb += pj[k, r]
                dg = po.digamma(pj[k])
                dsa, dsb = float(po.digamma(sa)), float(po.digamma(sb))
                tC += lg(c * beta) - c * lg(beta) + float(np.sum((beta - 1) * (dg - dsa)))
                tD += float(np.sum(Sj[k] * (dg - dsb)))
                tG += -lg(sa) + sum(lg(v) for v in pj[k]) - float(np.sum((pj[k] - 1) * (dg - dsa)))
            o += K * c
        phi = new_phi
        T = table_from_phi(phi, ncat, K, n_cat_bound)
        tD += K * n_missing
        if not dp:
            tB = float(np.sum(csp * pr["e1"]))
        else:
            tail = np.concatenate([np.cumsum(csp[::-1])[::-1][1:], [0.0]])
            tB = float(np.sum(tail * pr["e2"] + csp * pr["e1"]))
        elbo = pr["A"] + tB + tC + tD + pr["E"] + H + tG
        if it != 0 and not math.isinf(elbo):
            if elbo - trace[-1] < -epsilon:
                warn |= 1
            if elbo - trace[-1] < epsilon:
                conv = True
        trace.append(elbo)
        cs = csp
        if conv:
            break
    # results
    U = np.zeros_like(phi)
    for pj, uj in zip(_rag_views(phi, ncat, K), _rag_views(U, ncat, K)):
        uj[:] = pj / pj.sum(axis=1, keepdims=True)
    if not dp:
        omega = a1 + cs
        lam = omega / omega.sum()
        prior = omega
    else:
        k1 = a1 + cs
        k2 = a2 + np.concatenate([np.cumsum(cs[::-1])[::-1][1:], [0.0]])
        nu = k1 / (k1 + k2)
        lam = nu * np.concatenate([[1.0], np.cumprod(1 - nu)[:-1]])
        if lam[-1] > 0.01:
            warn |= 2
        prior = np.concatenate([k1, k2])
    return dict(status=0, convergence=np.array(trace), n_iter=len(trace), converged=conv,
                ELBO=trace[-1] if trace else float("nan"), lambda_=lam, phi=phi, category_prob=U,
                prior=prior, warn_flags=warn)
