"""`repetitions` (/root/reference/R/mixdir.R:120-134): mixdir_b200_fit_batch runs the
restarts of one mixdir() call in one library call -- concurrently when the fits are small -- and keeps
the first fit with the largest ELBO.  Every fit of a batch must be bit-identical to the same fit made
alone; the kept one must be what the reference's loop keeps (checked against the CPU oracle run
repetition by repetition)."""
import time

import numpy as np
import pytest

from conftest import subset_recode
from mixdir_b200 import Engine, ZetaUnderflow, synth
from mixdir_b200 import api
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _inits(seeds, N, ncat, K):
    cs = np.stack([synth.zeta_init(s, N, K).sum(axis=0) for s in seeds])
    ph = np.stack([synth.phi_init(s, ncat, K) for s in seeds])
    return cs, ph


@pytest.mark.parametrize("dp", [False, True])
def test_batch_equals_sequential_fits_and_oracle_selection(dp, mushroom):
    codes, cats = subset_recode(mushroom["codes"], mushroom["categories"], range(1000), range(5))
    ncat = np.array([len(c) for c in cats], dtype=np.int32)
    N, K = codes.shape[0], (10 if dp else 3)
    seeds = [1, 2, 3, 4, 5, 6]
    cs, ph = _inits(seeds, N, ncat, K)
    with Engine(codes, ncat) as eng:
        seq = [eng.fit(K, cs[i], ph[i], dp=dp) for i in range(len(seeds))]
        fits, keep = eng.fit_batch(K, cs, ph, dp=dp)
        fits2, keep2 = eng.fit_batch(K, cs, ph, dp=dp)  # the views are reused
    assert keep == keep2
    for a, b, c in zip(seq, fits, fits2):
        for o in (b, c):
            assert a["n_iter"] == o["n_iter"] and a["converged"] == o["converged"]
            assert np.array_equal(a["convergence"], o["convergence"])  # bit-identical traces
            assert a["ELBO"] == o["ELBO"] and a["warn_flags"] == o["warn_flags"]
            for f in ("lambda_", "phi", "category_prob"):
                assert np.array_equal(a[f], o[f])
    elbos = [s["ELBO"] for s in seq]
    assert keep == int(np.argmax(elbos))  # np.argmax = first maximum = R's strict `<` update
    assert np.array_equal(fits[keep]["pred_class"], seq[keep]["pred_class"])
    assert np.max(np.abs(fits[keep]["class_prob"] - seq[keep]["class_prob"])) < 1e-12
    assert all(f["pred_class"] is None for i, f in enumerate(fits) if i != keep)
    # the reference's loop on the CPU: same repetition kept, same result
    X = po.codes_to_r_matrix(codes)
    refs = [po.fit(X, ncat, K, int(dp), 1.0, 1.0, 0.1, 100, 1e-3, synth.zeta_init(s, N, K),
                   synth.phi_init(s, ncat, K)) for s in seeds]
    kept = 0
    for i in range(1, len(refs)):
        if refs[kept]["ELBO"] < refs[i]["ELBO"]:
            kept = i
    assert kept == keep
    r, g = refs[kept], fits[keep]
    assert r["n_iter"] == g["n_iter"]
    assert np.max(np.abs(g["convergence"] - r["convergence"]) / np.abs(r["convergence"])) < 1e-9
    assert np.array_equal(g["pred_class"], r["pred_class"])
    assert np.max(np.abs(g["class_prob"] - r["class_prob"])) < 1e-6


def test_batch_is_cheaper_than_sequential_fits(mushroom):
    """BASELINE configs[0] with repetitions = 8: the restarts overlap on the device."""
    codes, cats = subset_recode(mushroom["codes"], mushroom["categories"], range(1000), range(5))
    ncat = np.array([len(c) for c in cats], dtype=np.int32)
    N, K, R = codes.shape[0], 3, 8
    cs, ph = _inits(range(1, R + 1), N, ncat, K)
    with Engine(codes, ncat) as eng:
        kw = dict(max_iter=40, epsilon=-np.inf, want_class_prob=False, want_pred_class=False)
        eng.fit_batch(K, cs, ph, **kw)  # warm-up: views, workspaces, module load
        eng.fit(K, cs[0], ph[0], **kw)
        t1, tb = [], []
        for _ in range(5):
            t = time.perf_counter()
            eng.fit(K, cs[0], ph[0], **kw)
            t1.append(time.perf_counter() - t)
            t = time.perf_counter()
            eng.fit_batch(K, cs, ph, **kw)
            tb.append(time.perf_counter() - t)
    one, batch = min(t1), min(tb)
    print(f"one fit {one * 1e3:.2f} ms, batch of {R} {batch * 1e3:.2f} ms ({batch / one:.2f}x)")
    assert batch < 0.5 * R * one  # at least twice as fast as the sequential loop


def test_batch_error_and_large_or_sharded_handles():
    # an underflowing repetition ends the call with the reference's stop() (R/mixdir_vi.R:77-81)
    ncat = np.full(1000, 2, dtype=np.int32)  # test-mixdir.R:109-114: 50 x 1000 columns
    codes, _ = synth.make_codes(1, 50, ncat, 2)
    cs, ph = _inits([1, 2], 50, ncat, 2)
    with Engine(codes, ncat) as eng:
        with pytest.raises(ZetaUnderflow, match="underflow in the calculation of zeta"):
            eng.fit_batch(2, cs, ph, max_iter=5)
    # bad shapes
    with Engine(synth.TOY_DATA, [3, 3, 3]) as eng:
        with pytest.raises(ValueError):
            eng.fit_batch(2, np.ones((2, 3)), np.ones((2, 18)))
        with pytest.raises(ValueError):
            eng.fit_batch(2, np.ones((2, 2)), np.ones((3, 18)))


def test_mixdir_repetitions_keeps_the_best_restart(mushroom):
    codes, cats = subset_recode(mushroom["codes"], mushroom["categories"], range(1000), range(5))
    cols = {f"c{j}": [cats[j][v - 1] if v else None for v in codes[:, j]] for j in range(codes.shape[1])}
    one = [api.mixdir(cols, n_latent=3, seed=s)["ELBO"] for s in (11,)]
    many = api.mixdir(cols, n_latent=3, seed=11, repetitions=5)
    assert many["ELBO"] >= one[0]  # the first repetition draws the same initial values
    assert many["class_prob"].shape == (1000, 3) and many["pred_class"].shape == (1000,)
    assert abs(many["class_prob"].sum(axis=1) - 1).max() < 1e-12
