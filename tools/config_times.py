"""Per-iteration and whole-fit times of the small BASELINE configs (M1, M2, S3) on the GPU next to the
reference's CPU path (oracle/_ref + restated R loop, 1 thread).
    python tools/config_times.py [--no-cpu]        (--no-cpu: skip the CPU leg, 40 s)"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden, subset_recode  # noqa: E402
from mixdir_b200 import Engine, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (CPU baseline leg)

import json  # noqa: E402

g = load_golden("mushroom.npz")
mush = dict(codes=g["codes"], categories=json.loads(str(g["categories"])))


NO_CPU = "--no-cpu" in sys.argv


def run(name, codes, ncat, K, dp, max_iter=100, eps=1e-3, cpu=True):
    cpu = cpu and not NO_CPU
    N = codes.shape[0]
    z0 = synth.zeta_init(1, N, K)
    p0 = synth.phi_init(1, ncat, K)
    with Engine(codes, ncat) as e:  # warm-up: module load, plan, caches
        e.fit(K, z0.sum(0), p0, dp=dp, max_iter=2, epsilon=-np.inf)
    t = time.perf_counter()
    with Engine(codes, ncat) as e:
        r = e.fit(K, z0.sum(0), p0, dp=dp, max_iter=max_iter, epsilon=eps)
        tm = e.timing()
    wall = time.perf_counter() - t
    line = (f"{name:34s} GPU: {r['n_iter']:3d} iterations, {tm['fit_loop_ms'] / r['n_iter'] * 1e3:8.1f} us/iter on the "
            f"device, whole mixdir_vi call (create+fit+results+destroy) {wall * 1e3:8.2f} ms, engine {tm['engine']}"
            f" (KC {tm['classes_per_cta']}, P {tm['cluster_size']}, {tm['grid_ctas']} CTAs"
            + (f", E+M pass {tm['em_kernel_ms'] / r['n_iter'] * 1e3:.1f} us" if tm['em_kernel_ms'] > 0 else "") + ")")
    if cpu:
        t = time.perf_counter()
        ref = po.fit(po.codes_to_r_matrix(codes), ncat, K, int(dp), 1.0, 1.0, 0.1, max_iter, eps, z0, p0,
                     which="ref" if po.ref() is not None else "oracle")
        cw = time.perf_counter() - t
        same = np.array_equal(ref["pred_class"], r["pred_class"]) and ref["n_iter"] == r["n_iter"]
        line += f" | CPU reference path: {cw:8.3f} s ({cw / ref['n_iter'] * 1e3:9.2f} ms/iter), same result: {same}"
    print(line, flush=True)


c1, cats1 = subset_recode(mush["codes"], mush["categories"], range(1000), range(5))
run("M1 mushroom[1:1000,1:5] K=3", c1, np.array([len(c) for c in cats1], dtype=np.int32), 3, False)
c2, cats2 = subset_recode(mush["codes"], mush["categories"], range(8124), range(23))
run("M2 mushroom 8124x23 DP K=10", c2, np.array([len(c) for c in cats2], dtype=np.int32), 10, True)
ncat3 = np.full(60, 10, dtype=np.int32)
c3, _ = synth.make_codes(1, 70_000, ncat3, 10, p_na=0.10)
run("S3 70k x 60, C=10, K=10, 10% NA", c3, ncat3, 10, False, max_iter=20, eps=-np.inf)
