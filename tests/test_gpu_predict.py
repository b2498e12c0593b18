"""predict / final pass (/root/reference/R/predict.R:88-103 == R/mixdir_vi.R:136-148)
through the C ABI against the oracle's predict_class restatement, over the shapes that select the
different code paths of k_predict_units: 8 / 16 / 32 classes per CTA, clusters of 2..8 CTAs for
K > 32, odd row counts (scalar column stores), tails of the last tile, NA cells, features without
NAs (no NA row in the pair tables), all-NA rows, and rows whose terms all underflow (NaN -> 0)."""
import os

import numpy as np
import pytest

from mixdir_b200 import Engine, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu


def _model(seed, ncat, K):
    """a plausible (lambda, category_prob): Dirichlet-ish rows from the seeded hash generator"""
    rng = np.random.default_rng(seed)
    lam = rng.dirichlet(np.ones(K))
    U = []
    for c in ncat:
        U.append(rng.dirichlet(np.full(int(c), 0.3), size=K).ravel())  # (k, r) order inside a feature
    return lam, np.concatenate(U)


@pytest.mark.parametrize("K,J,N,p_na", [(3, 5, 1001, 0.1), (8, 23, 2049, 0.0), (9, 12, 777, 0.2),
                                        (16, 30, 4096, 0.05), (17, 60, 3001, 0.0), (32, 60, 5000, 0.1),
                                        (33, 20, 2000, 0.1), (64, 100, 3000, 0.0), (100, 10, 1025, 0.3),
                                        (256, 6, 513, 0.1)])
def test_predict_matches_oracle(K, J, N, p_na):
    ncat = np.array(([16, 12, 8, 4, 2, 10, 3, 7] * 13)[:J], dtype=np.int32)
    codes, _ = synth.make_codes(K + J, N, ncat, min(K, 16), p_na=p_na)
    if p_na > 0:
        codes[:, 0] = np.maximum(codes[:, 0], 1)  # one feature without any NA next to features with NAs
    else:
        codes[N // 2] = 0  # an all-NA row (the only NAs of the data): class_prob = lambda / sum(lambda)
    lam, U = _model(K * 1000 + J, ncat, K)
    ocp, opc = po.predict(po.codes_to_r_matrix(codes), ncat, K, lam, U)
    with Engine(codes, ncat) as eng:
        cp, pc = eng.predict(K, lam, U)
        tm = eng.timing()
        _, pc_only = eng.predict(K, lam, U, want_class_prob=False)
        cp_only, _ = eng.predict(K, lam, U, want_pred_class=False)
    assert tm["final_pass_engine"] == 6  # the unit kernel
    assert np.max(np.abs(cp - ocp)) < 1e-12
    assert np.array_equal(pc, opc) and np.array_equal(pc_only, opc)
    assert np.array_equal(cp_only, cp)
    if p_na == 0:
        np.testing.assert_allclose(cp[N // 2], lam / lam.sum(), rtol=1e-12)
    assert abs(cp.sum(axis=1) - 1).max() < 1e-12


def test_predict_underflowing_rows_are_nan_like_the_reference():
    """No max shift in predict_class: when every lambda_k * exp(...) underflows, rowSums is 0 and the row
    is NaN; which.max of an all-NaN row is integer(0) -> 0 here."""
    J, K, N = 400, 4, 300
    ncat = np.full(J, 2, dtype=np.int32)
    codes = np.ones((N, J), dtype=np.uint8)
    codes[::3] = 2
    codes[1] = 0  # all NA: finite
    lam = np.full(K, 0.25)
    U = np.tile(np.array([1e-3, 1 - 1e-3]), J * K)  # code 1 everywhere: (1e-3)^400 underflows
    ocp, opc = po.predict(po.codes_to_r_matrix(codes), ncat, K, lam, U)
    with Engine(codes, ncat) as eng:
        cp, pc = eng.predict(K, lam, U)
    assert np.isnan(ocp[2]).all() and opc[2] == 0 and not np.isnan(ocp[0]).any() and not np.isnan(ocp[1]).any()
    assert np.array_equal(np.isnan(cp), np.isnan(ocp))
    assert np.array_equal(pc, opc)
    np.testing.assert_allclose(cp[~np.isnan(cp)], ocp[~np.isnan(ocp)], rtol=1e-12)


def test_unit_kernel_equals_the_feature_table_kernels():
    """k_predict_units against the round-1 kernels (feature table in shared memory / in L2), which stay
    in the library for 2-byte codes: same pred_class, class_prob to the last few ulps (the pair table adds
    log U_a + log U_b before the row sum: a re-association of R/predict.R:91-99)."""
    ncat = np.array([16, 12, 8, 4, 2, 10] * 10, dtype=np.int32)
    N, K = 30_001, 32
    codes, _ = synth.make_codes(3, N, ncat, K, p_na=0.05)
    lam, U = _model(7, ncat, K)
    with Engine(codes, ncat) as eng:
        cp, pc = eng.predict(K, lam, U)
        new = eng.timing()["final_pass_engine"]
        os.environ["MIXDIR_B200_OLD_PREDICT"] = "1"
        try:
            cp0, pc0 = eng.predict(K, lam, U)
            old = eng.timing()["final_pass_engine"]
        finally:
            del os.environ["MIXDIR_B200_OLD_PREDICT"]
    assert (new, old) == (6, 7)
    assert np.array_equal(pc, pc0)
    assert np.max(np.abs(cp - cp0)) < 1e-13
    # 2-byte codes take the feature-table kernel by themselves
    with Engine(codes.astype(np.uint16), ncat) as eng:
        cp2, pc2 = eng.predict(K, lam, U)
        assert eng.timing()["final_pass_engine"] in (7, 8)
    assert np.array_equal(pc2, pc) and np.max(np.abs(cp2 - cp)) < 1e-13
