"""CPU tests of the host-side mirror of mixdir()'s input handling (R/mixdir.R:91-118,
R/predict.R:64-84) and of the deterministic synthetic generators."""
import os

import numpy as np
import pytest

from mixdir_b200 import synth
from mixdir_b200.coding import MISSING_LEVEL, code_columns, recode_for_predict


def test_factor_coding_sorted_levels_and_na():
    cols = [["b", "a", None, "c", "a"], ["x", "x", "x", float("nan"), "x"]]
    codes, cats = code_columns(cols)
    assert cats == [["a", "b", "c"], ["x"]]
    assert codes.dtype == np.uint8
    assert codes.tolist() == [[2, 1], [1, 1], [0, 1], [3, 0], [1, 1]]


def test_na_handle_category_adds_level_to_every_column():
    """R/mixdir.R:103-106: "(Missing)" is appended to EVERY column."""
    cols = [["b", "a", None], ["x", "y", "x"]]
    codes, cats = code_columns(cols, na_handle="category")
    assert cats == [["a", "b", MISSING_LEVEL], ["x", "y", MISSING_LEVEL]]
    assert codes.tolist() == [[2, 1], [1, 2], [3, 1]]
    assert (codes != 0).all()


def test_empty_column_is_an_error():
    with pytest.raises(ValueError, match="is empty"):  # R/mixdir.R:112-115
        code_columns([[None, None]])
    with pytest.raises(ValueError):
        code_columns([["a"]], na_handle="bogus")  # match.arg


def test_numeric_input_is_coded_by_its_string_form():
    # create_data() passes a numeric matrix; mixdir() does as.factor(as.character(col))
    codes, cats = code_columns([[1, 3, 1, 3], [2, 2, 3, 2]])
    assert cats == [["1", "3"], ["2", "3"]] and codes.tolist() == [[1, 1], [2, 1], [1, 2], [2, 1]]


def test_wide_features_use_two_byte_codes():
    col = [str(i).zfill(4) for i in range(300)]
    codes, cats = code_columns([col])
    assert codes.dtype == np.uint16 and codes.max() == 300


def test_recode_for_predict_missing_column_and_unseen_level():
    names, cats = ["f1", "f2"], [["a", "b"], ["x", "y"]]
    codes, unseen = recode_for_predict({"f2": ["y", "zzz", None]}, names, cats)
    assert codes.tolist() == [[0, 2], [0, 0], [0, 0]]  # missing column -> NA, unseen -> NA
    assert unseen == {"f2": {"zzz"}}
    codes, unseen = recode_for_predict({"f1": [None]}, names, [c + [MISSING_LEVEL] for c in cats],
                                       na_handle="category")
    assert codes.tolist() == [[3, 0]] and not unseen


def test_synth_is_a_pure_function_of_seed_and_row():
    ncat = np.array([4, 2, 9], dtype=np.int32)
    full, z = synth.make_codes(3, 1000, ncat, 5, p_na=0.2)
    part, z2 = synth.make_codes(3, 300, ncat, 5, p_na=0.2, row0=500)
    assert np.array_equal(full[500:800], part) and np.array_equal(z[500:800], z2)
    assert full.max() <= 9 and (full[:, 1] <= 2).all()
    assert 0.15 < (full == 0).mean() < 0.25
    zi = synth.zeta_init(1, 50, 7)
    np.testing.assert_allclose(zi.sum(axis=1), 1.0)
    assert np.array_equal(synth.zeta_init(1, 20, 7, row0=30), zi[30:])
    p = synth.phi_init(1, ncat, 7)
    assert set(np.unique(p)) <= {1.0, 2.0, 3.0} and p.size == 7 * 15


def test_toy_data_matches_helper_r():
    t = synth.TOY_DATA
    assert t.shape == (10, 3)
    assert (t[1] == 3).all() and (t[6] == 3).all()          # rows 2, 7 = CCC
    for i in (4, 5, 7, 8, 9):
        assert t[i].tolist() == [1, 2, 2]                   # rows 5,6,8,9,10 = ABB


def test_fast_exp_accuracy(tmp_path):
    """mixdir_b200/csrc/fastexp.cuh (the branch-free exp of the soft-max weights) compiled for the HOST
    against libm/long double: < 1 ulp over the normal range, 0 below -708, for -inf and for NaN."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    src = tmp_path / "t.cpp"
    src.write_text(r'''
#include "fastexp.cuh"
#include <cstdio>
#include <random>
int main() {
  std::mt19937_64 g(7);
  std::uniform_real_distribution<double> wide(-700.0, 0.01), near(-40.0, 0.001);
  double worst = 0;
  for (long i = 0; i < 4000000; i++) {
    const double x = (i & 1) ? wide(g) : near(g);
    const double a = mxd::exp_nonpos(x);
    const long double ref = expl((long double)x);
    const double ulp = std::ldexp(1.0, std::ilogb((double)ref) - 52);
    const double err = std::fabs((double)((long double)a - ref)) / ulp;
    if (err > worst) worst = err;
  }
  const bool edges = mxd::exp_nonpos(0.0) == 1.0 && mxd::exp_nonpos(-709.0) == 0.0 &&
                     mxd::exp_nonpos(-INFINITY) == 0.0 && mxd::exp_nonpos(NAN) == 0.0 &&
                     mxd::exp_nonpos(INFINITY) == 0.0 && mxd::exp_nonpos(-1e300) == 0.0;
  std::printf("%.4f %d\n", worst, edges ? 1 : 0);
}
''')
    exe = tmp_path / "t"
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    subprocess.run([gxx, "-O2", "-std=c++17", "-I", os.path.join(root, "mixdir_b200", "csrc"), "-o", str(exe), str(src)],
                   check=True)
    worst, edges = subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.split()
    assert float(worst) < 1.0 and edges == "1"


def test_seg_kernel_rounding_dither_is_uniform_and_uncorrelated():
    """k_em_seg rounds the responsibilities to 32 bits with a dither u = mix(row, class) / 2^32
    (seg.cuh); the rounding is unbiased only if u is uniform and the noise of a statistics cell grows
    like sqrt(rows) only if neighbouring rows / classes are uncorrelated.  tools/dither_hash_check.py
    restates the mix; an independent uniform sample would give the same numbers."""
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "dither_hash_check", os.path.join(os.path.dirname(__file__), "..", "tools", "dither_hash_check.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    r = mod.check(n_rows=100_000)
    assert abs(r["mean"] - 0.5) < 1e-3 and abs(r["var_times_12"] - 1) < 5e-3
    assert r["chi2_per_dof"] < 1.3
    for k in ("corr_next_row", "corr_row_plus_8", "corr_row_plus_240", "corr_next_class"):
        assert abs(r[k]) < 5e-3, k
    assert abs(r["block_16_var_ratio"] - 1) < 0.05 and abs(r["block_240_var_ratio"] - 1) < 0.1


@pytest.mark.parametrize("J", [1, 3, 4, 7, 8, 12, 23, 60, 61])
def test_host_packer_fp64_to_bytes(J):
    """mixdir_b200_create_from_r_matrix's host packer (csrc/hostpack.cpp; host code, runs without a GPU):
    the reference's fp64 column-major X with NA (R/mixdir.R:117-118) -> 1-byte row-major codes, 0 = NA;
    every column-group path (8, 4, rest), ragged row counts, a row offset, padded rows; cells that are not
    NA or an integer in 1..limit raise the flag and come out as 0."""
    import ctypes as C
    from mixdir_b200 import _capi
    lib = _capi.lib()
    if not lib.mxd_have_avx2():
        pytest.skip("no AVX2 on this host")
    fn = lib.mxd_pack_u8_avx2
    fn.restype = C.c_int
    fn.argtypes = [C.c_void_p, C.c_int64, C.c_int64, C.c_int64, C.c_int, C.c_int, C.c_void_p, C.c_int]
    rng = np.random.default_rng(J)
    for N, r0, r1 in ((5000, 0, 5000), (4099, 13, 4099), (7, 0, 7), (2100, 256, 2049)):
        X = np.asfortranarray(rng.integers(1, 200, size=(N, J)).astype(np.float64))
        X[rng.random((N, J)) < 0.1] = np.nan
        row_bytes = (J + 3) // 4 * 4
        out = np.full((r1 - r0, row_bytes), 0xEE, dtype=np.uint8)
        bad = fn(X.ctypes.data, N, r0, r1, J, row_bytes, out.ctypes.data, 255)
        want = np.nan_to_num(X[r0:r1], nan=0.0).astype(np.uint8)
        assert bad == 0
        assert np.array_equal(out[:, :J], want)
        assert (out[:, J:] == 0xEE).all()  # padding bytes are the caller's
        # invalid cells: a fraction, a zero, a negative, a value above the limit
        for v in (2.5, 0.0, -3.0, 256.0, np.inf):
            Y = X.copy(order="F")
            i, j = int(rng.integers(r0, r1)), int(rng.integers(0, J))
            Y[i, j] = v
            out2 = np.zeros_like(out)
            assert fn(Y.ctypes.data, N, r0, r1, J, row_bytes, out2.ctypes.data, 255) == 1, v
            want2 = want.copy()
            want2[i - r0, j] = 0
            assert np.array_equal(out2[:, :J], want2)
