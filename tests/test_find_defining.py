"""find_defining_features (/root/reference/R/predict.R:266-339): the device pass
(mixdir_b200_feature_divergence) and the Python mirror of the R control flow against the literal
CPU restatement (oracle/pyoracle.py), plus the reference's own test of this function
(tests/testthat/test-mixdir.R:192-204) re-expressed."""
import numpy as np
import pytest

from conftest import subset_recode
from mixdir_b200 import synth
from oracle import pyoracle as po


def small_problem(seed=3, N=400, J=7, K=4, p_na=0.1):
    ncat = np.array(([4, 2, 6, 3, 5, 2, 3] * 3)[:J], dtype=np.int32)
    codes, _ = synth.make_codes(seed, N, ncat, K, p_na=p_na)
    z0 = synth.zeta_init(seed, N, K)
    p0 = synth.phi_init(seed, ncat, K)
    fit = po.fit(po.codes_to_r_matrix(codes), ncat, K, 0, 1.0, 1.0, 0.1, 30, 1e-3, z0, p0)
    return codes, ncat, K, fit


def test_oracle_matches_reference_test_invariants():
    """test-mixdir.R:192-204: lengths, and quality(n_features=2) == quality(all)[3] -- the same greedy
    path is followed and the final quality of the 2 survivors is the 'cumulative' entry.  (Holds
    when the halving schedule passes through 3 -> 2 remaining features, as for mushroom's 23
    columns: 23 -> 12 -> 6 -> 3 -> 2; here 6 -> 3 -> 2.)"""
    codes, ncat, K, fit = small_problem(J=6)
    X = po.codes_to_r_matrix(codes)
    sub = np.arange(codes.shape[0])
    f_all, q_all, _ = po.find_defining_features(X, ncat, K, fit["lambda_"], fit["category_prob"], sub)
    assert len(f_all) == len(q_all) == ncat.size and sorted(f_all) == list(range(ncat.size))
    f2, q2, _ = po.find_defining_features(X, ncat, K, fit["lambda_"], fit["category_prob"], sub,
                                          n_features=2)
    assert len(f2) == 2 and np.isscalar(q2)
    assert q2 == pytest.approx(q_all[2], rel=1e-9)


def test_adjusted_rand_known_answers():
    a = np.array([0, 0, 1, 1, 2, 2])
    assert po.adjusted_rand(a, a) == pytest.approx(1.0)
    assert po.adjusted_rand(a, (a + 1) % 3) == pytest.approx(1.0)  # label permutation
    # hand-computed: tab = [[1,1],[1,1]] -> (0 - 4/6) / (2 - 4/6)
    assert po.adjusted_rand([0, 0, 1, 1], [0, 1, 0, 1]) == pytest.approx(-0.5)


@pytest.mark.gpu
@pytest.mark.parametrize("measure", ["JS", "ARI"])
def test_feature_divergence_matches_oracle_rounds(measure):
    from mixdir_b200 import Engine
    codes, ncat, K, fit = small_problem()
    X = po.codes_to_r_matrix(codes)
    N, J = codes.shape
    sub = np.random.default_rng(1).permutation(N)[:250]
    _, _, rounds = po.find_defining_features(X, ncat, K, fit["lambda_"], fit["category_prob"], sub,
                                             measure=measure, step_size=1, exponential_decay=False)
    with Engine(codes, ncat) as eng:
        for cols, diverg in rounds:
            active = np.zeros(J, dtype=np.uint8)
            active[cols] = 1
            out = eng.feature_divergence(K, fit["lambda_"], fit["category_prob"], active, measure, rows=sub)
            np.testing.assert_allclose(out[cols], diverg, rtol=1e-9, atol=1e-11)
            assert np.all(np.isnan(out[:J][active == 0]))
        # every row, no subsample; last entry = nothing removed = 0 divergence against itself
        full = eng.feature_divergence(K, fit["lambda_"], fit["category_prob"], np.ones(J), measure)
        assert abs(full[J]) < 1e-12


@pytest.mark.gpu
@pytest.mark.parametrize("kw", [dict(), dict(n_features=2), dict(n_features=3, step_size=1),
                                dict(exponential_decay=False), dict(measure="ARI"),
                                dict(measure="ARI", n_features=2), dict(exponential_decay=3, step_size=2)])
def test_find_defining_features_mirror_matches_oracle(kw):
    from mixdir_b200 import api
    codes, ncat, K, fit = small_problem(seed=5, N=300, J=9, K=3)
    J = ncat.size
    names = [f"f{j}" for j in range(J)]
    cats = [[str(c) for c in range(1, n + 1)] for n in ncat]
    obj = {"lambda": fit["lambda_"], "_U_flat": fit["category_prob"], "_names": names,
           "_categories": cats, "na_handle": "ignore"}
    Xd = {names[j]: [None if v == 0 else str(v) for v in codes[:, j]] for j in range(J)}
    sub = np.random.default_rng(2).permutation(codes.shape[0])[:200]
    got = api.find_defining_features(obj, Xd, subsample=sub, **kw)
    feats, quality, _ = po.find_defining_features(po.codes_to_r_matrix(codes), ncat, K, fit["lambda_"],
                                                  fit["category_prob"], sub, **kw)
    assert got["features"] == [names[j] for j in feats]
    np.testing.assert_allclose(got["quality"], quality, rtol=1e-8, atol=1e-10)


@pytest.mark.gpu
def test_find_defining_features_on_mushroom(mushroom):
    """the reference's own test (test-mixdir.R:192-204) on mushroom[1:30, ] through the product API"""
    from mixdir_b200 import api
    codes, cats = subset_recode(mushroom["codes"], mushroom["categories"], range(30), range(23))
    cols = {n: [None if v == 0 else cats[j][v - 1] for v in codes[:, j]]
            for j, n in enumerate(mushroom["names"])}
    res = api.mixdir(cols, beta=1, seed=4)
    allf = api.find_defining_features(res, cols, subsample=np.arange(30))
    assert len(allf["quality"]) == 23 and len(allf["features"]) == 23
    two = api.find_defining_features(res, cols, n_features=2, subsample=np.arange(30))
    assert len(two["features"]) == 2 and np.isscalar(two["quality"])
    assert two["quality"] == pytest.approx(allf["quality"][2], rel=1e-8, abs=1e-12)


def test_host_mirror_control_flow_on_cpu(monkeypatch):
    """The Python twin of find_defining_features (api.py) follows R/predict.R:266-339 step by step;
    here WITHOUT a GPU: the device pass is replaced by an oracle-backed stand-in, so what is tested
    is the host logic (removal schedule, ordering, quality bookkeeping, final re-ordering quirk)."""
    from mixdir_b200 import api

    codes, ncat, K, fit = small_problem(seed=7, N=200, J=9, K=3)
    J = ncat.size
    X = po.codes_to_r_matrix(codes)

    class FakeEngine:
        def __init__(self, c, n, devices=None):
            assert np.array_equal(c, codes)

        def __enter__(self):
            return self

        def __exit__(self, *a):
            pass

        def feature_divergence(self, K_, lam, U, active, measure, rows=None):
            Xs = X[rows]
            cols = [j for j in range(J) if active[j]]

            def pred(cs):
                Z = np.full_like(Xs, np.nan)
                Z[:, cs] = Xs[:, cs]
                return po.predict(Z, ncat, K_, lam, U)[0]

            def q_of(q, p):
                if measure == "JS":
                    return float(np.sum(q * (np.log(q) - np.log(p)) + p * (np.log(p) - np.log(q))) / 2)
                return 1 - float(po.adjusted_rand(np.argmax(q, axis=1), np.argmax(p, axis=1)))
            p = pred(list(range(J)))
            out = np.full(J + 1, np.nan)
            for i in cols:
                out[i] = q_of(pred([c for c in cols if c != i]), p)
            out[J] = q_of(pred(cols), p)
            return out

    monkeypatch.setattr(api, "Engine", FakeEngine)
    names = [f"f{j}" for j in range(J)]
    cats = [[str(c) for c in range(1, n + 1)] for n in ncat]
    obj = {"lambda": fit["lambda_"], "_U_flat": fit["category_prob"], "_names": names,
           "_categories": cats, "na_handle": "ignore"}
    Xd = {names[j]: [None if v == 0 else str(v) for v in codes[:, j]] for j in range(J)}
    sub = np.random.default_rng(4).permutation(codes.shape[0])[:120]
    for kw in (dict(), dict(n_features=2), dict(n_features=4, step_size=1), dict(exponential_decay=False),
               dict(measure="ARI"), dict(measure="ARI", n_features=3), dict(exponential_decay=3, step_size=2)):
        got = api.find_defining_features(obj, Xd, subsample=sub, **kw)
        feats, quality, _ = po.find_defining_features(X, ncat, K, fit["lambda_"], fit["category_prob"],
                                                      sub, **kw)
        assert got["features"] == [names[j] for j in feats], kw
        np.testing.assert_allclose(got["quality"], quality, rtol=1e-12, atol=1e-14)
    with pytest.raises(ValueError):
        api.find_defining_features(obj, Xd, measure="bogus")
    with pytest.raises(ValueError, match="diverg"):
        api.find_defining_features(obj, Xd, n_features=J, subsample=sub)
