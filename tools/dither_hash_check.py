"""Host check of k_em_seg's rounding dither (mixdir_b200/csrc/seg.cuh: dither_base / dither_mix): the
32-bit mix of (global row, class) that makes the 32-bit rounding of the responsibilities unbiased.
Prints uniformity (chi-square over 256 bins), correlations between neighbouring rows / classes and
the variance of block sums against n/12 (what decides how the rounding noise of a statistics cell
grows with its rows).  tests/test_host_logic.py runs `check()` with loose bounds."""
import numpy as np

M32 = np.uint64(0xFFFFFFFF)


def _mul(a, b):
    return (a * np.uint64(b)) & M32


def dither(rows, classes):
    """x(row, class) exactly as the kernel computes it (rows < 2^32)."""
    rows = np.asarray(rows, dtype=np.uint64)[:, None]
    ks = np.asarray(classes, dtype=np.uint64)[None, :]
    h = (_mul(rows, 0x9E3779B1) + _mul(ks, 0xC2B2AE3D)) & M32
    h ^= h >> np.uint64(15)
    h = _mul(h, 0x2C1B3C6D)
    h ^= h >> np.uint64(13)
    h = _mul(h, 0x297A2D39)
    return h


def check(n_rows=200_000, n_classes=32):
    u = dither(np.arange(n_rows), np.arange(n_classes)).astype(np.float64) / 2.0**32
    hist = np.histogram(u, bins=256, range=(0, 1))[0]
    e = u.size / 256
    corr = lambda a, b: float(np.corrcoef(a.ravel(), b.ravel())[0, 1])
    out = {
        "mean": float(u.mean()),
        "var_times_12": float(u.var() * 12),
        "chi2_per_dof": float(((hist - e) ** 2 / e).sum() / 255),
        "corr_next_row": corr(u[:-1], u[1:]),
        "corr_row_plus_8": corr(u[:-8], u[8:]),
        "corr_row_plus_240": corr(u[:-240], u[240:]),
        "corr_next_class": corr(u[:, :-1], u[:, 1:]),
    }
    for n in (16, 240, 4096):
        s = (u[: n_rows // n * n] - 0.5).reshape(-1, n, n_classes).sum(axis=1)
        out[f"block_{n}_var_ratio"] = float(s.var() / (n / 12))
    return out


if __name__ == "__main__":
    for k, v in check().items():
        print(f"{k:24s} {v:.6f}")
