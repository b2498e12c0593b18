"""Deterministic synthetic inputs for tests and benchmarks (host-side utility, no GPU math).

Everything is a pure function of (seed, index) through a 64-bit counter hash, so the same
rows come out whatever the shard layout, GPU count, numpy version or device.

Data model: the reference's own generator, tests/testthat/helper.R:52-82
(`generate_categorical_dataset`): per (class, feature) Dirichlet weights
rbetabinom(C, 100, 0.05, 10) + 1, labels ~ Categorical(lambda_true), cells ~
DirMult(1, weights) == Categorical(weights / sum).  rmutil's rbetabinom(size=100, m=.05, s=10)
is Binomial(100, p) with p ~ Beta(s*m, s*(1-m)) = Beta(0.5, 9.5).

Initial values mirror R/mixdir_vi.R:28-39: zeta_init rows ~ Dirichlet(1,...,1) (normalised
Exp(1) draws), phi_init in {1,2,3}.
"""
from __future__ import annotations

import numpy as np

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(x):
    """splitmix64 finaliser on uint64 arrays."""
    x = np.asarray(x, dtype=np.uint64)
    with np.errstate(over="ignore"):
        x = (x + np.uint64(0x9E3779B97F4A7C15)) & _M64
        x = ((x ^ (x >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
        x = ((x ^ (x >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
        x = x ^ (x >> np.uint64(31))
    return x


def hash_u64(seed, stream, idx):
    idx = np.asarray(idx, dtype=np.uint64)
    with np.errstate(over="ignore"):
        key = _mix(np.uint64(seed) * np.uint64(0x100000001B3) + np.uint64(stream))
        return _mix(idx ^ key)


def hash_u01(seed, stream, idx):
    """uniform in (0,1), 53 bits."""
    return ((hash_u64(seed, stream, idx) >> np.uint64(11)).astype(np.float64) + 0.5) / float(1 << 53)


def zeta_init(seed, n_rows, K, row0=0):
    """Dirichlet(1,...,1) rows (R/mixdir_vi.R:32), rows [row0, row0+n_rows)."""
    idx = (np.arange(row0, row0 + n_rows, dtype=np.uint64)[:, None] * np.uint64(K)
           + np.arange(K, dtype=np.uint64)[None, :])
    e = -np.log(hash_u01(seed, 1, idx))
    return e / e.sum(axis=1, keepdims=True)


def phi_init(seed, ncat, K):
    """ragged (j,k,r) values in {1,2,3} (R/mixdir_vi.R:35-39: sample(1:3, C_j, replace=TRUE))."""
    n = int(K * np.sum(ncat))
    return 1.0 + (hash_u64(seed, 2, np.arange(n, dtype=np.uint64)) % np.uint64(3)).astype(np.float64)


def class_weights(seed, ncat, K_true):
    """W[j] : [K_true, C_j] integer weights  Binomial(100, Beta(0.5, 9.5)) + 1  (helper.R:63)."""
    rng = np.random.Generator(np.random.Philox(key=int(seed) + 7919))
    out = []
    for c in ncat:
        p = rng.beta(0.5, 9.5, size=(K_true, int(c)))
        out.append(rng.binomial(100, p).astype(np.float64) + 1.0)
    return out


def make_codes(seed, n_rows, ncat, K_true, p_na=0.0, row0=0, dtype=None):
    """Synthetic code matrix rows [row0, row0+n_rows): uint8/uint16 [n_rows, J], 0 = NA.
    Returns (codes, true_latent)."""
    ncat = np.asarray(ncat, dtype=np.int64)
    J = ncat.size
    W = class_weights(seed, ncat, K_true)
    rows = np.arange(row0, row0 + n_rows, dtype=np.uint64)
    z = (hash_u64(seed, 3, rows) % np.uint64(K_true)).astype(np.int64)  # sample(1:K_true) (helper.R:70)
    if dtype is None:
        dtype = np.uint8 if ncat.max() <= 255 else np.uint16
    codes = np.zeros((n_rows, J), dtype=dtype)
    for j in range(J):
        cdf = np.cumsum(W[j], axis=1)
        cdf /= cdf[:, -1:]
        u = hash_u01(seed, 100 + j, rows)
        x = (u[:, None] > cdf[z]).sum(axis=1) + 1  # Categorical(W/sum W)  (helper.R:76)
        codes[:, j] = np.minimum(x, ncat[j])
        if p_na > 0:
            codes[hash_u01(seed, 100000 + j, rows) < p_na, j] = 0  # MCAR
    return codes, z


def make_codes_torch(seed, n_rows, ncat, K_true, p_na=0.0, row0=0, device="cuda", chunk=1 << 20):
    """Same model as make_codes, generated on `device` with torch (bench-size inputs).  The random
    stream differs from make_codes (torch Philox), the class weights are identical."""
    import torch

    ncat = np.asarray(ncat, dtype=np.int64)
    J = ncat.size
    W = class_weights(seed, ncat, K_true)
    dt = torch.uint8 if ncat.max() <= 255 else torch.int16
    out = torch.empty((n_rows, J), dtype=dt, device=device)
    cdfs = []
    for j in range(J):
        cdf = np.cumsum(W[j], axis=1)
        cdf /= cdf[:, -1:]
        cdfs.append(torch.tensor(cdf, device=device, dtype=torch.float64))
    g = torch.Generator(device=device)
    for r0 in range(0, n_rows, chunk):
        n = min(chunk, n_rows - r0)
        g.manual_seed(int(seed) * 1000003 + row0 + r0)
        z = torch.randint(0, K_true, (n,), device=device, generator=g)
        for j in range(J):
            u = torch.rand((n,), device=device, dtype=torch.float64, generator=g)
            x = (u[:, None] > cdfs[j][z]).sum(dim=1) + 1
            x = torch.clamp(x, max=int(ncat[j]))
            if p_na > 0:
                x = torch.where(torch.rand((n,), device=device, generator=g) < p_na,
                                torch.zeros_like(x), x)
            out[r0:r0 + n, j] = x.to(dt)
    return out


# the 10 x 3 toy matrix of the reference's tests (tests/testthat/helper.R:2-48):
# codes A/B/C = 1/2/3; rows 2,7 = CCC; rows 5,6,8,9,10 = ABB
TOY_DATA = np.array([
    [1, 2, 1], [3, 3, 3], [1, 2, 3], [3, 2, 1], [1, 2, 2],
    [1, 2, 2], [3, 3, 3], [1, 2, 2], [1, 2, 2], [1, 2, 2]], dtype=np.uint8)
