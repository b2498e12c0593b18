"""Multi-GPU parity (skipped on boxes with one GPU): the single-process flavour
(mixdir_b200_create with n_devices > 1: contiguous row shards, ncclCommInitAll, one grouped
all-reduce of the statistics buffer per iteration) must reproduce the one-GPU fit."""
import numpy as np
import pytest

from mixdir_b200 import Engine, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _ngpu():
    import torch
    return torch.cuda.device_count()


@pytest.mark.parametrize("engine,K", [(0, 20), (0, 32), (4, 20)], ids=["auto_seg_k20", "auto_seg_k32", "units_fp64"])
@pytest.mark.parametrize("dp", [False, True])
def test_single_process_two_devices_match_one(dp, engine, K):
    """engine 0 picks the sorted-segment kernel here (>= 16,384 rows): its statistics AND class masses
    are exact integer sums, so the whole trajectory (phi, class_prob) must come out BIT-identical on
    one and on two devices; only the ELBO (fp64 entropy sum) may differ in the last bits.  The fp64
    kernel (engine 4) re-associates its sums across shards.  (K = 32 runs two CTAs of 16 classes with two
    rows per warp instruction; the second shard starts at row 25,000 = 8 mod 16, so its rows sit in the
    other half-warp than on one device: the logit sums must not depend on that.)"""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    pat = np.array([16, 12, 8, 4, 2, 10] * 5, dtype=np.int32)
    N = 50_001
    codes, _ = synth.make_codes(5, N, pat, K, p_na=0.05)
    z0 = synth.zeta_init(5, N, K)
    p0 = synth.phi_init(5, pat, K)
    outs = []
    for devs in (None, [0, 1]):
        with Engine(codes, pat, devices=devs) as eng:
            assert eng.n_missing == int((codes == 0).sum())
            eng.set_engine(engine)
            outs.append(eng.fit(K, z0.sum(axis=0), p0, dp=dp, max_iter=8, epsilon=-np.inf))
            assert eng.timing()["engine"] == (5 if engine == 0 else 4)
    a, b = outs
    assert np.max(np.abs(a["convergence"] - b["convergence"]) / np.abs(a["convergence"])) < 1e-12
    assert np.array_equal(a["pred_class"], b["pred_class"])
    if engine == 0:
        assert np.array_equal(a["phi"], b["phi"])
        assert np.array_equal(a["class_prob"], b["class_prob"])
    else:
        assert np.max(np.abs(a["class_prob"] - b["class_prob"])) < 1e-10
        np.testing.assert_allclose(a["phi"], b["phi"], rtol=1e-10, atol=1e-10)
    # and against the oracle on a prefix small enough for the CPU
    sub = slice(0, 2000)
    ref = po.fit(po.codes_to_r_matrix(codes[sub]), pat, K, int(dp), 1.0, 1.0, 0.1, 5, -np.inf,
                 synth.zeta_init(5, 2000, K), p0)
    with Engine(codes[sub], pat, devices=[0, 1]) as eng:
        r = eng.fit(K, synth.zeta_init(5, 2000, K).sum(axis=0), p0, dp=dp, max_iter=5, epsilon=-np.inf)
        cp, pc = eng.predict(K, r["lambda_"], r["category_prob"])
    assert np.max(np.abs(r["convergence"] - ref["convergence"]) / np.abs(ref["convergence"])) < 1e-9
    assert np.max(np.abs(r["class_prob"] - ref["class_prob"])) < 1e-6
    assert np.array_equal(r["pred_class"], ref["pred_class"])
    assert np.max(np.abs(cp - r["class_prob"])) < 1e-12 and np.array_equal(pc, r["pred_class"])


@pytest.mark.parametrize("measure", ["JS", "ARI"])
def test_feature_divergence_two_devices_match_one(measure):
    """mixdir_b200_feature_divergence on row shards (per-shard sums / contingency tables, one grouped
    all-reduce) == the one-GPU pass, with and without a row subsample."""
    if _ngpu() < 2:
        pytest.skip("needs 2 GPUs")
    pat = np.array([4, 2, 6, 3, 5, 2, 3, 8], dtype=np.int32)
    N, K = 20_001, 5
    codes, _ = synth.make_codes(9, N, pat, K, p_na=0.1)
    with Engine(codes, pat) as eng:
        r = eng.fit(K, synth.zeta_init(9, N, K).sum(axis=0), synth.phi_init(9, pat, K), max_iter=6,
                    epsilon=-np.inf, want_class_prob=False, want_pred_class=False)
    active = np.array([1, 1, 0, 1, 1, 1, 0, 1], dtype=np.uint8)
    sub = np.random.default_rng(3).permutation(N)[:7000]
    outs = []
    for devs in (None, [0, 1]):
        with Engine(codes, pat, devices=devs) as eng:
            outs.append((eng.feature_divergence(K, r["lambda_"], r["category_prob"], active, measure),
                         eng.feature_divergence(K, r["lambda_"], r["category_prob"], active, measure, rows=sub)))
    for a, b in zip(outs[0], outs[1]):
        np.testing.assert_allclose(a, b, rtol=1e-11, atol=1e-12, equal_nan=True)
