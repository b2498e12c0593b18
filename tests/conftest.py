import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on a B200)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.fixture(scope="session")
def mushroom():
    g = load_golden("mushroom.npz")
    return dict(codes=g["codes"], names=[str(s) for s in g["names"]],
                categories=json.loads(str(g["categories"])))


def subset_recode(codes, categories, rows, cols):
    """mixdir() codes the rows/columns it is given: levels = observed strings of the subset
    (R/mixdir.R:99-102).  Re-level a slice of the globally coded matrix accordingly."""
    sub = codes[np.asarray(rows)][:, np.asarray(cols)]
    out = np.zeros_like(sub)
    cats = []
    for jj, j in enumerate(cols):
        present = sorted(set(int(v) for v in sub[:, jj] if v != 0))  # global levels are sorted
        remap = np.zeros(256, dtype=sub.dtype)
        for new, old in enumerate(present, 1):
            remap[old] = new
        out[:, jj] = remap[sub[:, jj]]
        cats.append([categories[j][o - 1] for o in present])
    return out, cats
