"""GPU parity tests: the CUDA engine, called through the C ABI (ctypes), against the CPU oracle
and the committed golden fixtures.  Tolerances are north_star's: pred_class identical,
class_prob within 1e-6 absolute, ELBO within 1e-9 relative (BASELINE.json)."""
import numpy as np
import pytest

from conftest import load_golden, subset_recode
from mixdir_b200 import Engine, ZetaUnderflow, MixdirError, synth
from oracle import pyoracle as po

pytestmark = pytest.mark.gpu

CP_ATOL = 1e-6    # class_prob, absolute   (north_star)
ELBO_RTOL = 1e-9  # ELBO, relative         (north_star)

# generic (global tables + atomics), fused (row-major codes), units (fp64 histograms in shared memory),
# seg (default: sorted-segment M-step)
ENGINES = [1, 2, 4, 5]
ENGINE_IDS = ["generic", "fused", "units", "seg"]
# Engine 5 feeds the M-step with responsibilities in 32-bit fixed point, rounded without bias
# (seg.cuh; exact integer sums).  The north_star tolerances (ELBO 1e-9 relative, class_prob 1e-6,
# identical pred_class) apply unchanged on every fixture; the statistics themselves carry zero-mean
# noise of ~2^-32 * sqrt(rows of the cell) (test_seg_kernel_drift_is_bounded looks at long traces).
SEG_S_ATOL = 1e-6


def elbo_tol(engine, n_rows):
    """north_star's 1e-9, except for engine 5 FORCED onto fewer than 16384 rows (the automatic choice
    takes it only from there on): rounding the responsibilities to 32 bits perturbs the ELBO by
    ~1e-10 / sqrt(rows) of its natural scale rows x features, and the first, far-from-converged
    iterations of a random start amplify the perturbation (seg.cuh)."""
    return 1e-6 if (engine == 5 and n_rows < 16384) else ELBO_RTOL

FIT_FIXTURES = [
    "fit_toy_k3.npz", "fit_toy_dp_k10.npz", "fit_toy_na_k3.npz", "fit_toy_na_dp_k10.npz",
    "fit_m1_seed1.npz", "fit_m1_seed2.npz", "fit_m1_seed3.npz", "fit_m1_seed4.npz",
    "fit_m2_dp_k10.npz", "fit_m2_dp_k20.npz", "fit_m2_k5.npz",
    "fit_s3_small.npz", "fit_s3_small_dp.npz", "fit_s4_small.npz", "fit_unused_level.npz",
]


def fixture_codes(name, g, mushroom):
    if "codes" in g:
        return g["codes"]
    if name.startswith("fit_m1"):
        return subset_recode(mushroom["codes"], mushroom["categories"], range(1000), range(5))[0]
    return mushroom["codes"]


def rel(a, b):
    return np.max(np.abs(np.asarray(a) - np.asarray(b)) / np.maximum(np.abs(np.asarray(b)), 1e-300))


@pytest.mark.parametrize("engine", ENGINES, ids=ENGINE_IDS)
@pytest.mark.parametrize("name", FIT_FIXTURES)
def test_fit_matches_golden(name, engine, mushroom):
    g = load_golden(name)
    codes = fixture_codes(name, g, mushroom)
    ncat, K, dp, seed = g["ncat"], int(g["K"]), int(g["dp"]), int(g["seed"])
    N = codes.shape[0]
    z0 = synth.zeta_init(seed, N, K)
    p0 = synth.phi_init(seed, ncat, K)
    with Engine(codes, ncat) as eng:
        eng.set_engine(engine)
        try:
            r = eng.fit(K, z0.sum(axis=0), p0, dp=bool(dp), alpha=g["alpha"], beta=float(g["beta"]),
                        max_iter=int(g["max_iter"]), epsilon=float(g["epsilon"]))
        except MixdirError as e:
            if engine >= 2 and "does not fit" in str(e):
                pytest.skip("fused plan does not fit this shape")
            raise
        t = eng.timing()
    assert t["engine"] == engine
    ref_trace = g["convergence"]
    assert r["n_iter"] == len(ref_trace), (r["n_iter"], len(ref_trace))
    assert rel(r["convergence"], ref_trace) < elbo_tol(engine, N)
    assert r["converged"] == bool(g["converged"])
    assert abs(r["ELBO"] - float(g["ELBO"])) <= elbo_tol(engine, N) * abs(float(g["ELBO"]))
    small_seg = engine == 5 and N < 16384  # see elbo_tol
    np.testing.assert_allclose(r["lambda_"], g["lambda_"], rtol=1e-7 if small_seg else 1e-9, atol=1e-12)
    # (engine 5: the ascent amplifies the 32-bit rounding noise of the responsibilities while classes
    # merge; the north_star quantities below keep their tolerances)
    np.testing.assert_allclose(r["phi"], g["phi"], rtol=1e-5 if engine == 5 else 1e-8,
                               atol=1e-5 if engine == 5 else 1e-9)
    np.testing.assert_allclose(r["category_prob"], g["category_prob"], rtol=1e-8,
                               atol=1e-7 if engine == 5 else 1e-10)
    prior = np.concatenate([r["kappa1"], r["kappa2"]]) if dp else r["omega"]
    np.testing.assert_allclose(prior, g["prior"], rtol=1e-7 if small_seg else 1e-9, atol=1e-9)
    assert np.max(np.abs(r["class_prob"][g["rows"]] - g["class_prob_rows"])) < CP_ATOL
    assert np.array_equal(r["pred_class"], g["pred_class"])  # identical class assignments
    assert r["warn_flags"] == int(g["warn_flags"])


@pytest.mark.parametrize("engine", ENGINES, ids=ENGINE_IDS)
@pytest.mark.parametrize("tag", ["k3", "k10"])
def test_estep_matches_reference_native_functions(tag, engine):
    """kernel zeta / statistics vs the reference's own update_zeta_cpp / update_zeta_dp_cpp /
    expec_log_x_cpp outputs (tests/golden/native_vectors.npz, produced by oracle/_ref);
    mirrors tests/testthat/test_cpp.R:15-122."""
    g = load_golden("native_vectors.npz")
    codes, ncat = g["codes"], g["ncat"]
    K = 3 if tag == "k3" else 10
    phi, omega, kappa2 = g[f"phi_{tag}"], g[f"omega_{tag}"], g[f"kappa2_{tag}"]
    n_cat = int(codes.max())
    import sys, os
    sys.path.insert(0, os.path.dirname(__file__))
    import fused_model as fm
    T = fm.table_from_phi(phi, ncat, K, n_cat)
    with Engine(codes, ncat) as eng:
        for dp, zref in ((0, g[f"zeta_{tag}"]), (1, g[f"zeta_dp_{tag}"])):  # noqa: E501
            pr = fm.prior_step(omega - 1.0, K, 0, 1.0, 1.0) if not dp else None
            if dp:  # prior term of update_zeta_dp_cpp from explicit kappa1/kappa2
                d12 = po.digamma(omega + kappa2)
                e1, e2 = po.digamma(omega) - d12, po.digamma(kappa2) - d12
                lp = e1 + np.concatenate([[0.0], np.cumsum(e2)[:-1]])
            else:
                lp = pr["logprior"]
            out = eng.estep(K, lp, T, engine=engine, want_zeta=True)
            zn = zref / zref.sum(axis=1, keepdims=True)
            assert out["underflow"] == 0
            assert np.max(np.abs(out["zeta"] - zn)) < 1e-12
            np.testing.assert_allclose(out["cs"], zn.sum(axis=0), rtol=1e-9 if engine == 5 else 1e-12)
            # M-step statistics and the collapsed expec_log_x
            st = po.estep_stats(codes, ncat, K, lp - 1.0, T)
            np.testing.assert_allclose(out["S"], st["S"], rtol=1e-11, atol=SEG_S_ATOL if engine == 5 else 1e-12)
            assert abs(out["H"] - st["H"]) < 1e-10 * max(1.0, abs(st["H"]))
            if not dp:
                ex = float(np.sum(out["S"] * T)) + K * int((codes == 0).sum())
                assert abs(ex - float(g[f"expec_log_x_{tag}"])) < (1e-9 if engine == 5 else 1e-10) * abs(float(g[f"expec_log_x_{tag}"]))


def test_fit_live_oracle_random_shapes():
    """seeded random shapes, CUDA path vs the literal oracle run live."""
    rng = np.random.default_rng(7)
    for case in range(6):
        J = int(rng.integers(1, 40))
        ncat = rng.integers(1, 9, size=J).astype(np.int32)
        K = int(rng.integers(1, 12))
        dp = case % 2
        N = int(rng.integers(5, 600))
        codes, _ = synth.make_codes(100 + case, N, ncat, max(K, 2), p_na=0.15 * (case % 3))
        z0 = synth.zeta_init(case, N, K)
        p0 = synth.phi_init(case, ncat, K)
        ref = po.fit(po.codes_to_r_matrix(codes), ncat, K, dp, 1.0, 1.0, 0.1, 25, 1e-3, z0, p0)
        for engine in ENGINES:
            with Engine(codes, ncat) as eng:
                eng.set_engine(engine)
                r = eng.fit(K, z0.sum(axis=0), p0, dp=bool(dp), max_iter=25, epsilon=1e-3,
                            n_cat_bound=int(np.nanmax(po.codes_to_r_matrix(codes))))
            assert r["n_iter"] == ref["n_iter"]
            assert rel(r["convergence"], ref["convergence"]) < elbo_tol(engine, N)
            assert np.max(np.abs(r["class_prob"] - ref["class_prob"])) < CP_ATOL
            assert np.array_equal(r["pred_class"], ref["pred_class"])


def test_predict_consistency(mushroom):
    """test-mixdir.R:147-152,168-172: predict(res, training rows) == class_prob;
    predict(all-NA) == lambda (DP: lambda / sum(lambda))."""
    codes, cats = subset_recode(mushroom["codes"], mushroom["categories"], range(30), range(23))
    ncat = np.array([len(c) for c in cats], dtype=np.int32)
    for dp, K in ((False, 3), (True, 10)):
        z0 = synth.zeta_init(5, 30, K)
        p0 = synth.phi_init(5, ncat, K)
        with Engine(codes, ncat) as eng:
            r = eng.fit(K, z0.sum(axis=0), p0, dp=dp)
            cp, pc = eng.predict(K, r["lambda_"], r["category_prob"])
        assert np.max(np.abs(cp - r["class_prob"])) < 1e-12
        assert np.array_equal(pc, r["pred_class"])
        with Engine(np.zeros((1, 23), dtype=np.uint8), ncat) as e2:
            cp0, _ = e2.predict(K, r["lambda_"], r["category_prob"])
        np.testing.assert_allclose(cp0[0] * r["lambda_"].sum(), r["lambda_"], rtol=1e-12)
        # and against the oracle's predict
        ocp, opc = po.predict(po.codes_to_r_matrix(codes), ncat, K, r["lambda_"], r["category_prob"])
        assert np.max(np.abs(cp - ocp)) < 1e-12
        assert np.array_equal(pc, opc)


@pytest.mark.parametrize("dp", [False, True])
def test_underflow_is_an_error(dp):
    """test-mixdir.R:109-114: 50 x 1000 columns -> the zeta-underflow stop()."""
    ncat = np.full(1000, 2, dtype=np.int32)
    codes, _ = synth.make_codes(1, 50, ncat, 2)
    z0 = synth.zeta_init(1, 50, 2)
    p0 = synth.phi_init(1, ncat, 2)
    ref = po.fit(po.codes_to_r_matrix(codes), ncat, 2, int(dp), 1.0, 1.0, 0.1, 5, 1e-3, z0, p0)
    assert ref["status"] == 1
    for engine in (1, 2):
        with Engine(codes, ncat) as eng:
            # J=1000 features do not fit the fused kernel's shared-memory plan: engine 0 picks
            # the generic kernel by itself, engine 2 (forced) must say so
            eng.set_engine(0 if engine == 2 else 1)
            with pytest.raises(ZetaUnderflow, match="underflow in the calculation of zeta"):
                eng.fit(2, z0.sum(axis=0), p0, dp=dp, max_iter=5)
    # a shape that does fit the fused kernel: few features, many categories -> tiny T values
    ncat2 = np.full(40, 2, dtype=np.int32)
    codes2, _ = synth.make_codes(1, 64, ncat2, 2)
    p_big = synth.phi_init(1, ncat2, 2).copy()
    p_big[::2] = 1e-300  # psi(1e-300) ~ -1e300: every logit underflows
    for engine in (2, 4, 5):
        with Engine(codes2, ncat2) as eng:
            eng.set_engine(engine)
            with pytest.raises(ZetaUnderflow):
                eng.fit(2, [32.0, 32.0], p_big, dp=dp, max_iter=3)


@pytest.mark.parametrize("engine", ENGINES, ids=ENGINE_IDS)
def test_subnormal_band_is_renormalised_not_quantised(engine):
    """A stated deviation (precision): the reference exponentiates the UNSHIFTED logit
    (src/mixdir_impl.cpp:83) and normalises afterwards (R/mixdir_vi.R:82), so a row whose largest
    logit lies in about [-745, -708] has subnormal, i.e. coarsely quantised, weights -- logits
    (-740, -746) give zeta = (1, 0) in R.  The engine shifts by the row maximum and returns the
    exact soft-max (0.99753, 0.00247); only the all-zero row (largest logit below -745.13) is the
    reference's underflow error.  This test pins both sides of that statement."""
    ncat = np.array([2], dtype=np.int32)
    codes = np.array([[1], [2], [1]], dtype=np.uint8)
    lp = np.array([0.0, 0.0])
    #           class 0: r=1, r=2      class 1: r=1, r=2        (ragged (j,k,r) order)
    T = np.array([-739.0, -2.0, -745.0, -3.0])  # logits row 1/3: (-740, -746); row 2: (-3, -4)
    with Engine(codes, ncat) as eng:
        out = eng.estep(2, lp, T, engine=engine, want_zeta=True)
    assert out["underflow"] == 0
    exact = 1.0 / (1.0 + np.exp(-6.0))
    np.testing.assert_allclose(out["zeta"][0], [exact, 1.0 - exact], rtol=1e-12)
    np.testing.assert_allclose(out["zeta"][1], [1.0 / (1.0 + np.exp(-1.0)), 1.0 - 1.0 / (1.0 + np.exp(-1.0))], rtol=1e-12)
    # what the reference's arithmetic gives for row 1: exp(-740) is a subnormal, exp(-746) is 0
    ref = np.exp(np.array([-740.0, -746.0]))
    assert ref[0] > 0 and ref[1] == 0 and np.array_equal(ref / ref.sum(), [1.0, 0.0])
    # one step further down the reference stops with its underflow error, and so does the engine
    T2 = T.copy()
    T2[0], T2[2] = -745.0, -751.0  # logits (-746, -752): rowSums(zeta) == 0
    with Engine(codes, ncat) as eng:
        assert eng.estep(2, lp, T2, engine=engine, want_zeta=True)["underflow"] == 1


def test_seg_kernel_drift_is_bounded():
    """The price of the sorted-segment kernel's 32-bit responsibilities: over 25
    iterations from a random start at K = 32 the ELBO trace stays within ~1e-10 of the fp64-statistics
    kernel on most iterations; while classes merge the ascent amplifies the rounding noise for a few
    iterations (survey over 36 starts, profiles/r02_seg_precision.txt: median of the per-fit maxima
    4e-10, worst 2e-8), and the traces come together again (final difference <= 4e-10); pred_class
    is identical.  With round-to-nearest instead of unbiased rounding the same runs sat at 5e-9 .. 9e-9
    on EVERY late iteration -- that is what this test guards against."""
    ncat = np.array([16, 12, 8, 4, 2, 10] * 4, dtype=np.int32)
    N, K = 65_536, 32
    meds, finals = [], []
    for seed in (2, 3, 4):
        codes, _ = synth.make_codes(seed, N, ncat, K, p_na=0.05)
        z0 = synth.zeta_init(seed, N, K).sum(axis=0)
        p0 = synth.phi_init(seed, ncat, K)
        out = {}
        for engine in (5, 4):
            with Engine(codes, ncat) as eng:
                eng.set_engine(engine)
                out[engine] = eng.fit(K, z0, p0, max_iter=25, epsilon=-np.inf, want_class_prob=False)
        d = np.abs(out[5]["convergence"] - out[4]["convergence"]) / np.abs(out[4]["convergence"])
        assert np.array_equal(out[5]["pred_class"], out[4]["pred_class"])
        assert d.max() < 1e-7            # the amplified transients
        assert np.median(d) < ELBO_RTOL  # the typical iteration
        assert d[-1] < ELBO_RTOL         # and the end of the fit
        meds.append(np.median(d))
        finals.append(d[-1])
    assert np.median(meds) < 3e-10


def test_seg_global_statistics_variant(monkeypatch):
    """MIXDIR_B200_SEG_GLOBAL_STATS=1 (A/B switch): k_em_seg without a statistics
    table in shared memory -- K = 32 then runs in ONE CTA per SM, the S5 shape K = 64, J = 100 in
    clusters of 4 x 16 classes.  Checked against the default plan (K = 32; another pair plan adds the
    logit terms in another order, so the rounded responsibilities differ in their last unit here and
    there) and against the fp64 unit kernel where the default plan does not fit (K = 64)."""
    for K, J in ((32, 60), (64, 100)):
        pat = np.array(([16, 12, 8, 4, 2, 10] * 20)[:J], dtype=np.int32)
        N = 20_000
        codes, _ = synth.make_codes(31, N, pat, 16, p_na=0.03)
        z0 = synth.zeta_init(31, N, K).sum(axis=0)
        p0 = synth.phi_init(31, pat, K)
        out = {}
        for tag, env, engine in (("default", None, 5 if K == 32 else 4), ("global", "1", 5)):
            if env:
                monkeypatch.setenv("MIXDIR_B200_SEG_GLOBAL_STATS", env)
            else:
                monkeypatch.delenv("MIXDIR_B200_SEG_GLOBAL_STATS", raising=False)
            with Engine(codes, pat) as eng:
                eng.set_engine(engine)
                out[tag] = eng.fit(K, z0, p0, max_iter=6, epsilon=-np.inf)
                t = eng.timing()
            assert t["engine"] == engine
            if tag == "global":
                assert (t["cluster_size"], t["classes_per_cta"]) == ((1, 32) if K == 32 else (4, 16)), t
        monkeypatch.delenv("MIXDIR_B200_SEG_GLOBAL_STATS", raising=False)
        a, b = out["default"], out["global"]
        assert np.array_equal(a["pred_class"], b["pred_class"])
        assert rel(b["convergence"], a["convergence"]) < ELBO_RTOL
        np.testing.assert_allclose(a["phi"], b["phi"], rtol=1e-5, atol=1e-5)
        # K = 64 on 20,000 rows generated from 16 classes: four fitted classes per true one merge during
        # these iterations and the ascent amplifies the 32-bit rounding noise -- a
        # regime the automatic engine choice never puts on this kernel
        assert np.max(np.abs(a["class_prob"] - b["class_prob"])) < (CP_ATOL if K == 32 else 2e-5)


def test_warp_specialised_seg_kernel_is_bit_identical(monkeypatch):
    """MIXDIR_B200_SEG_WS=1 (A/B switch): k_em_seg_ws runs the E-step of tile t+1 on
    15 warps with 96 registers next to the M-step of tile t on 16 warps with 32 (setmaxnreg), handing
    the zeta tile over through cluster-wide mbarriers.  Same plan, same arithmetic, exact integer
    statistics: every result must have the bits of the default kernel's; ragged row counts, NAs."""
    pat = np.array([16, 12, 8, 4, 2, 10] * 10, dtype=np.int32)
    K = 32
    for N in (20_000, 100_003):
        codes, _ = synth.make_codes(41, N, pat, K, p_na=0.04)
        z0 = synth.zeta_init(41, min(N, 65536), K).sum(axis=0) * (N / min(N, 65536))
        p0 = synth.phi_init(41, pat, K)
        out = []
        for ws in (False, True):
            if ws:
                monkeypatch.setenv("MIXDIR_B200_SEG_WS", "1")
            else:
                monkeypatch.delenv("MIXDIR_B200_SEG_WS", raising=False)
            with Engine(codes, pat) as eng:
                eng.set_engine(5)
                out.append(eng.fit(K, z0, p0, max_iter=5, epsilon=-np.inf))
        monkeypatch.delenv("MIXDIR_B200_SEG_WS", raising=False)
        a, b = out
        assert np.array_equal(a["phi"], b["phi"]) and np.array_equal(a["class_prob"], b["class_prob"])
        assert np.array_equal(a["pred_class"], b["pred_class"])
        assert rel(b["convergence"], a["convergence"]) < 1e-14  # (the entropy is an fp64 sum per CTA)


def test_create_from_r_matrix_and_u16():
    ncat = np.array([300, 4, 7], dtype=np.int32)  # > 255 levels -> 2-byte codes
    codes, _ = synth.make_codes(9, 500, ncat, 4, p_na=0.1)
    assert codes.dtype == np.uint16
    K = 4
    z0 = synth.zeta_init(2, 500, K)
    p0 = synth.phi_init(2, ncat, K)
    ref = po.fit(po.codes_to_r_matrix(codes), ncat, K, 0, 1.0, 1.0, 0.1, 10, 1e-3, z0, p0)
    for make in (lambda: Engine(codes, ncat),
                 lambda: Engine.from_r_matrix(po.codes_to_r_matrix(codes), ncat)):
        for engine in (1, 2):
            with make() as eng:
                assert eng.n_missing == int((codes == 0).sum())
                assert eng.max_code == int(codes.max())
                eng.set_engine(engine)
                try:
                    r = eng.fit(K, z0.sum(axis=0), p0, max_iter=10)
                except MixdirError as e:
                    if engine >= 2 and "does not fit" in str(e):
                        continue
                    raise
            assert rel(r["convergence"], ref["convergence"]) < ELBO_RTOL
            assert np.array_equal(r["pred_class"], ref["pred_class"])


def test_bad_arguments():
    codes = np.array([[1, 2], [3, 1]], dtype=np.uint8)
    # code 3 > C_0 = 2: the upload is asynchronous, the first call that needs the data reports it
    with Engine(codes, [2, 2]) as eng:
        with pytest.raises(MixdirError, match="exceeds"):
            eng.max_code
    with Engine(codes, [2, 2]) as eng:
        with pytest.raises(MixdirError, match="exceeds"):
            eng.fit(2, [1.0, 1.0], np.ones(8), max_iter=3, n_cat_bound=2)
    with Engine.from_r_matrix(np.array([[1.0, 2.5], [np.nan, 1.0]]), [2, 3]) as eng:
        with pytest.raises(MixdirError, match="not a positive integer"):
            eng.n_missing
    with Engine(codes, [3, 2]) as eng:
        with pytest.raises(ValueError):
            eng.fit(2, [1.0, 1.0], np.ones(3))  # wrong phi_init length (R/mixdir_vi.R:41-47)
        with pytest.raises(MixdirError):
            eng.fit(1000, np.ones(1000), np.ones(5000))


def test_full_size_properties():
    """BASELINE S4 shape at reduced N (fits the test budget): size-independent invariants +
    fused vs generic kernel agreement + a row-subsample checked against the oracle."""
    pat = np.array([16, 12, 8, 4, 2, 10] * 10, dtype=np.int32)
    N, K = 300_000, 32
    codes, _ = synth.make_codes(2, N, pat, K)
    p0 = synth.phi_init(2, pat, K)
    cs0 = np.full(K, N / K)
    outs = {}
    for engine in ENGINES:
        with Engine(codes, pat) as eng:
            eng.set_engine(engine)
            outs[engine] = eng.fit(K, cs0, p0, max_iter=4, epsilon=-np.inf, want_class_prob=False)
            outs[engine]["timing"] = eng.timing()
    assert outs[4]["timing"]["engine"] == 4 and outs[4]["timing"]["unit_pairs"] > 0
    assert outs[5]["timing"]["engine"] == 5 and outs[5]["timing"]["unit_pairs"] > 0
    assert rel(outs[5]["convergence"], outs[2]["convergence"]) < ELBO_RTOL
    assert np.array_equal(outs[5]["pred_class"], outs[2]["pred_class"])
    np.testing.assert_allclose(outs[5]["phi"], outs[2]["phi"], rtol=1e-8, atol=1e-4)
    assert rel(outs[4]["convergence"], outs[2]["convergence"]) < 1e-11
    assert np.array_equal(outs[4]["pred_class"], outs[2]["pred_class"])
    np.testing.assert_allclose(outs[4]["phi"], outs[2]["phi"], rtol=1e-9, atol=1e-9)
    a, b = outs[1], outs[2]
    assert rel(a["convergence"], b["convergence"]) < 1e-11
    np.testing.assert_allclose(a["phi"], b["phi"], rtol=1e-9, atol=1e-9)
    assert np.array_equal(a["pred_class"], b["pred_class"])
    # sum_k colSums(zeta) == N ; sum_r S_jkr summed over k == number of non-NA cells in column j
    np.testing.assert_allclose(b["omega"].sum(), N + K * 1.0, rtol=1e-12)
    o = 0
    for j, c in enumerate(pat):
        tot = (b["phi"][o:o + K * c] - 0.1).sum()
        assert abs(tot - N) < 1e-6 * N
        o += K * c
    # fused kernel really ran as planned: K=32 needs two CTAs of 16 classes
    assert b["timing"]["engine"] == 2 and b["timing"]["cluster_size"] >= 1
    # oracle on the same inputs, literal loop, same init
    sub = slice(0, 3000)
    ref = po.fit(po.codes_to_r_matrix(codes[sub]), pat, K, 0, 1.0, 1.0, 0.1, 4, -np.inf,
                 synth.zeta_init(1, 3000, K), p0)
    with Engine(codes[sub], pat) as eng:
        r = eng.fit(K, synth.zeta_init(1, 3000, K).sum(axis=0), p0, max_iter=4, epsilon=-np.inf)
    assert rel(r["convergence"], ref["convergence"]) < ELBO_RTOL
    assert np.max(np.abs(r["class_prob"] - ref["class_prob"])) < CP_ATOL


def test_full_size_s3_fit_to_convergence():
    """BASELINE.json configs[2] at FULL size: synthetic 70k x 60, 10 categories per feature, K=10,
    10 % missing values, the reference's defaults (max_iter=100, epsilon=1e-3), fitted to convergence
    on the GPU and by the reference's CPU path (oracle/_ref native functions when present + the
    restated R loop; ~2 s per iteration on one host core): same stopping iteration, ELBO trace
    within 1e-9 relative, class_prob within 1e-6, identical pred_class (R/mixdir_vi.R:58-155)."""
    ncat = np.full(60, 10, dtype=np.int32)
    N, K = 70_000, 10
    codes, _ = synth.make_codes(1, N, ncat, K, p_na=0.10)
    z0 = synth.zeta_init(1, N, K)
    p0 = synth.phi_init(1, ncat, K)
    with Engine(codes, ncat) as eng:
        r = eng.fit(K, z0.sum(axis=0), p0, max_iter=100, epsilon=1e-3)
        tm = eng.timing()
    assert tm["engine"] == 5  # the default kernel at this size
    ref = po.fit(po.codes_to_r_matrix(codes), ncat, K, 0, 1.0, 1.0, 0.1, 100, 1e-3, z0, p0,
                 which="ref" if po.ref() is not None else "oracle")
    assert r["n_iter"] == ref["n_iter"] and r["converged"] == ref["converged"]
    assert rel(r["convergence"], ref["convergence"]) < ELBO_RTOL
    assert np.max(np.abs(r["class_prob"] - ref["class_prob"])) < CP_ATOL
    assert np.array_equal(r["pred_class"], ref["pred_class"])
    np.testing.assert_allclose(r["lambda_"], ref["lambda_"], rtol=1e-8)
    np.testing.assert_allclose(r["category_prob"], ref["category_prob"], rtol=1e-6, atol=1e-9)
    assert int((codes == 0).sum()) > 0.09 * codes.size


@pytest.mark.parametrize("K,J,dp", [(64, 100, 0), (64, 100, 1), (40, 30, 0), (20, 23, 1), (5, 3, 0),
                                    (1, 4, 0), (33, 7, 1), (128, 12, 0), (256, 5, 0)])
def test_cluster_shapes_against_oracle(K, J, dp):
    """every cluster plan of the fused kernel (KC in {8,16,32}, P = 1..8, incl. the S5 shape
    K=64, J=100 -> 8 CTAs x 8 classes) and large K against the literal oracle."""
    pat = np.array(([16, 12, 8, 4, 2, 10] * 20)[:J], dtype=np.int32)
    N = 1500
    codes, _ = synth.make_codes(11, N, pat, min(K, 16), p_na=0.05)
    z0 = synth.zeta_init(3, N, K)
    p0 = synth.phi_init(3, pat, K)
    ref = po.fit(po.codes_to_r_matrix(codes), pat, K, dp, 1.0, 1.0, 0.1, 4, -np.inf, z0, p0)
    for engine in (0, 1, 2, 4, 5):
        with Engine(codes, pat) as eng:
            eng.set_engine(engine)
            try:
                r = eng.fit(K, z0.sum(axis=0), p0, dp=bool(dp), max_iter=4, epsilon=-np.inf)
            except MixdirError as e:
                if engine in (2, 4, 5) and "does not fit" in str(e):
                    continue
                raise
            t = eng.timing()
        if engine == 0 and K <= 64:
            assert t["engine"] >= 2, t  # these shapes must run on a fused kernel
        assert rel(r["convergence"], ref["convergence"]) < elbo_tol(engine, N)
        assert np.max(np.abs(r["class_prob"] - ref["class_prob"])) < CP_ATOL
        assert np.array_equal(r["pred_class"], ref["pred_class"])


def test_recycled_buffers_do_not_leak_stale_codes():
    """Device blocks are recycled across handles (block cache).  A handle whose rows carry padding
    bytes (J*code_bytes not a multiple of 4) must not see the previous owner's bytes there."""
    n = 1000
    big = np.full((n, 8), 250, dtype=np.uint8)  # fills every byte of an (n+512) x 8 block
    with Engine(big, np.full(8, 250, dtype=np.int32)) as eng:
        assert eng.max_code == 250
    ncat = np.array([2, 5, 4, 3, 2], dtype=np.int32)  # J=5 -> 8-byte rows, 3 padding bytes
    codes, _ = synth.make_codes(3, n, ncat, 3)
    z0 = synth.zeta_init(3, n, 3)
    p0 = synth.phi_init(3, ncat, 3)
    ref = po.fit(po.codes_to_r_matrix(codes), ncat, 3, 0, 1.0, 1.0, 0.1, 6, -np.inf, z0, p0)
    for engine in (2, 1, 4, 5):
        with Engine(codes, ncat) as eng:
            eng.set_engine(engine)
            r = eng.fit(3, z0.sum(axis=0), p0, max_iter=6, epsilon=-np.inf)
        assert rel(r["convergence"], ref["convergence"]) < elbo_tol(engine, n)
        assert np.array_equal(r["pred_class"], ref["pred_class"])


@pytest.mark.parametrize("K,J,p_na", [(32, 60, 0.0), (32, 60, 0.1), (10, 23, 0.05), (64, 100, 0.0),
                                      (3, 5, 0.0), (7, 2, 0.2), (12, 61, 0.0)])
def test_feature_pair_units_against_oracle(K, J, p_na):
    """Feature pairing (two features per table gather / histogram update, kernels.cuh "Units"):
    the same fit with pairing on and off, both against the literal oracle; the pair plan must
    really be in use for the BASELINE S4 shape."""
    pat = np.array(([16, 12, 8, 4, 2, 10] * 20)[:J], dtype=np.int32)
    N = 2500
    codes, _ = synth.make_codes(21, N, pat, min(K, 16), p_na=p_na)
    z0 = synth.zeta_init(5, N, K)
    p0 = synth.phi_init(5, pat, K)
    ref = po.fit(po.codes_to_r_matrix(codes), pat, K, 0, 1.0, 1.0, 0.1, 5, -np.inf, z0, p0)
    pairs = {}
    for engine, on in ((2, True), (2, False), (4, True), (4, False), (5, True), (5, False)):
        with Engine(codes, pat) as eng:
            eng.set_engine(engine)
            eng.set_pairing(on)
            try:
                r = eng.fit(K, z0.sum(axis=0), p0, max_iter=5, epsilon=-np.inf)
            except MixdirError as e:
                if engine == 5 and "does not fit" in str(e):  # K=64, J=100: tables + statistics too large
                    pairs[(engine, on)] = 0 if not on else 1
                    continue
                raise
            t = eng.timing()
            # a second fit on the same handle reuses the packed unit codes
            r2 = eng.fit(K, z0.sum(axis=0), p0, max_iter=5, epsilon=-np.inf)
        assert t["engine"] == engine
        pairs[(engine, on)] = t["unit_pairs"]
        assert np.array_equal(r["convergence"], r2["convergence"])
        assert rel(r["convergence"], ref["convergence"]) < elbo_tol(engine, N)
        assert np.max(np.abs(r["class_prob"] - ref["class_prob"])) < CP_ATOL
        assert np.array_equal(r["pred_class"], ref["pred_class"])
        np.testing.assert_allclose(r["phi"], ref["phi"], rtol=1e-10, atol=1e-4 if engine == 5 else 1e-10)
    assert pairs[(2, False)] == 0 and pairs[(4, False)] == 0 and pairs[(5, False)] == 0
    if (J + 3) // 4 > (J - J // 2 + 3) // 4:  # pairing is used when it saves code words per row
        assert pairs[(2, True)] > 0 and pairs[(4, True)] > 0, pairs


@pytest.mark.parametrize("engine", [2, 4, 5])
def test_chunked_first_iteration_equals_resident_fit(engine):
    """With n_cat_bound given, fit() starts while the upload is still in flight: the first pass runs
    chunk by chunk (scan -> pack unit codes -> E+M on the rows that have landed, partial statistics
    accumulate over the launches).  Same results as the fit on resident data, and as the oracle on a
    prefix."""
    pat = np.array([16, 12, 8, 4, 2, 10] * 6, dtype=np.int32)
    N, K = 150_000, 24  # 8 upload chunks of 20480 rows; tiles straddle chunk ends for some plans
    codes, _ = synth.make_codes(31, N, pat, K, p_na=0.03)
    cs0 = synth.zeta_init(31, 4000, K).sum(axis=0) * (N / 4000.0)
    p0 = synth.phi_init(31, pat, K)
    ncb = int(codes.max())
    with Engine(codes, pat) as eng:
        eng.set_engine(engine)
        a = eng.fit(K, cs0, p0, max_iter=3, epsilon=-np.inf, n_cat_bound=ncb)  # upload in flight
        ta = eng.timing()
        b = eng.fit(K, cs0, p0, max_iter=3, epsilon=-np.inf)                   # resident
    assert ta["engine"] == engine
    assert rel(a["convergence"], b["convergence"]) < 1e-12
    assert np.array_equal(a["pred_class"], b["pred_class"])
    np.testing.assert_allclose(a["phi"], b["phi"], rtol=1e-11, atol=1e-11)
    sub = slice(0, 2500)
    z0 = synth.zeta_init(31, 2500, K)
    ref = po.fit(po.codes_to_r_matrix(codes[sub]), pat, K, 0, 1.0, 1.0, 0.1, 3, -np.inf, z0, p0)
    with Engine(codes[sub], pat) as eng:
        eng.set_engine(engine)
        r = eng.fit(K, z0.sum(axis=0), p0, max_iter=3, epsilon=-np.inf, n_cat_bound=ncb)
    assert rel(r["convergence"], ref["convergence"]) < elbo_tol(engine, 2500)
    assert np.max(np.abs(r["class_prob"] - ref["class_prob"])) < CP_ATOL

