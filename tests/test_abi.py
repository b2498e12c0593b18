"""CPU tests of the C-ABI boundary: the library loads, exports every symbol the header declares,
and fails loudly (status codes, no CPU fallback) without a GPU.  No compute calls."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from mixdir_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def header_symbols():
    src = open(os.path.join(ROOT, "include", "mixdir_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(mixdir_b200_[a-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    L = _capi.lib()
    names = header_symbols()
    assert len(names) >= 14
    for n in names:
        assert hasattr(L, n), n
    assert sorted(_capi.SYMBOLS) == names
    assert L.mixdir_b200_abi_version() == 2


def test_struct_layouts_match_header():
    src = open(os.path.join(ROOT, "include", "mixdir_b200.h")).read()
    body = re.search(r"typedef struct mixdir_b200_fit_params \{(.*?)\}", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int|double)\s+(\w+)\s*;", body)
    assert fields == [f for f, _ in _capi.FitParams._fields_]
    body = re.search(r"typedef struct mixdir_b200_timing \{(.*?)\}", src, re.S).group(1)
    body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
    fields = re.findall(r"\b(?:int|double)\s+(\w+)\s*;", body)
    assert fields == [f for f, _ in _capi.Timing._fields_]


def test_ragged_size():
    L = _capi.lib()
    ncat = (C.c_int * 3)(2, 5, 4)
    assert L.mixdir_b200_ragged_size(3, ncat, 7) == 7 * 11


def test_argument_checks_do_not_need_a_gpu():
    L = _capi.lib()
    h = C.c_void_p()
    ncat = (C.c_int * 2)(2, 2)
    codes = np.ones((3, 2), dtype=np.uint8)
    p = codes.ctypes.data_as(C.c_void_p)
    assert L.mixdir_b200_create(p, 3, 3, 2, ncat, 0, None, C.byref(h)) == _capi.BAD_ARGUMENT
    assert b"code_bytes" in L.mixdir_b200_last_error()
    bad = (C.c_int * 2)(2, 0)  # a feature without categories: mixdir() stops (R/mixdir.R:112-115)
    assert L.mixdir_b200_create(p, 1, 3, 2, bad, 0, None, C.byref(h)) == _capi.BAD_ARGUMENT
    big = (C.c_int * 2)(2, 300)
    assert L.mixdir_b200_create(p, 1, 3, 2, big, 0, None, C.byref(h)) == _capi.BAD_ARGUMENT
    c = C.c_void_p()
    assert L.mixdir_b200_comm_init(0, 5, 2, None, C.byref(c)) == _capi.BAD_ARGUMENT  # rank >= world
    assert L.mixdir_b200_create_dist(p, 1, 3, 2, ncat, None, C.byref(h)) == _capi.BAD_ARGUMENT
    assert not h


def test_no_cpu_fallback_without_a_device():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from mixdir_b200 import Engine, MixdirError
    with pytest.raises(MixdirError, match="no CUDA device|CUDA"):
        Engine(np.ones((4, 3), dtype=np.uint8), [2, 2, 2])


def test_product_never_imports_the_oracle():
    """only tests/, __graft_entry__.smoke() and bench.py's cpu legs may touch oracle/."""
    pkg = os.path.join(ROOT, "mixdir_b200")
    for dirpath, _d, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "pyoracle" not in txt and "libmixdir_oracle" not in txt and \
                    "import oracle" not in txt and "from oracle" not in txt, os.path.join(dirpath, f)


def test_r_adapter_parses():
    """r_adapter/mixdir_b200_rcpp.cpp (the Rcpp glue a maintainer adds to the reference,
    INTEGRATION.md) type-checks against include/mixdir_b200.h and a stand-in for <Rcpp.h> that declares
    the Rcpp signatures it uses (the image has no R): every C-ABI call in it matches the header."""
    import shutil
    import subprocess
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    src = os.path.join(ROOT, "r_adapter", "mixdir_b200_rcpp.cpp")
    p = subprocess.run([gxx, "-std=c++17", "-fsyntax-only", "-Wall", "-Werror",
                        "-I", os.path.join(ROOT, "tests", "rcpp_stub"), "-I", os.path.join(ROOT, "include"), src],
                       capture_output=True, text=True)
    assert p.returncode == 0, p.stderr
    text = open(src).read()
    for fn in ("mixdir_b200_create_from_r_matrix", "mixdir_b200_fit_batch", "mixdir_b200_predict",
               "mixdir_b200_feature_divergence", "mixdir_b200_destroy"):
        assert fn in text
    assert text.count("[[Rcpp::export]]") == 5
