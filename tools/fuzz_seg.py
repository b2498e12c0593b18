"""Hardening aid (GPU): random shapes at sizes the sorted-segment kernel (engine 5) serves, every plan it
can take (16 / 32 classes per CTA, one CTA or a cluster, feature split), against the fp64-statistics unit
kernel (engine 4) on the same data: gross errors only -- the two differ by the 32-bit rounding noise of
DESIGN.md section 5, so the bounds are loose.    python tools/fuzz_seg.py [seed]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mixdir_b200 import Engine, synth, MixdirError

rng = np.random.default_rng(int(sys.argv[1]) if len(sys.argv) > 1 else 0)
bad = n = 0
plans = {}
for ci in range(28):
    N = int(rng.integers(16384, 150000))
    J = int(rng.integers(2, 90))
    K = int(rng.choice([2, 3, 7, 10, 16, 17, 24, 32, 33, 48, 64]))
    cmax = int(rng.choice([2, 4, 10, 16, 30]))
    ncat = rng.integers(1, cmax + 1, size=J).astype(np.int32)
    p_na = float(rng.choice([0.0, 0.0, 0.05, 0.3]))
    dp = bool(rng.integers(0, 2))
    codes, _ = synth.make_codes(2000 + ci, N, ncat, max(2, min(K, 8)), p_na=p_na)
    z0 = synth.zeta_init(ci, min(N, 65536), K).sum(axis=0) * (N / min(N, 65536))
    p0 = synth.phi_init(ci, ncat, K)
    out = {}
    try:
        for e in (5, 4):
            with Engine(codes, ncat) as eng:
                eng.set_engine(e)
                out[e] = eng.fit(K, z0, p0, dp=dp, max_iter=6, epsilon=-np.inf)
                if e == 5:
                    t = eng.timing()
                    key = (t["classes_per_cta"], t["cluster_size"])
                    plans[key] = plans.get(key, 0) + 1
    except MixdirError as err:
        if "does not fit" in str(err):
            continue
        print("CASE", N, J, K, cmax, p_na, dp, "ERROR", err); bad += 1; continue
    n += 1
    a, b = out[5], out[4]
    rel = np.max(np.abs(a["convergence"] - b["convergence"]) / np.abs(b["convergence"]))
    dcp = float(np.nanmax(np.abs(a["class_prob"] - b["class_prob"])))
    same = float(np.mean(a["pred_class"] == b["pred_class"]))
    ok = rel < 1e-7 and (dcp < 1e-4 or dp) and (same > 0.9999 or dp)
    if not ok:
        print("CASE", dict(N=N, J=J, K=K, cmax=cmax, p_na=p_na, dp=dp), "plan", key, "rel", rel, "dcp", dcp, "same", same); bad += 1
print("cases", n, "bad", bad, "plans (classes per CTA, cluster size):", plans)
