"""A handful of iterations of the small BASELINE configs (M1, M2, S3) -- meant to run under
`ncu --metrics gpu__time_duration.sum` for a per-kernel launch list of the launch-bound regime.
    python tools/small_fit_profile.py [m1|m2|s3] [iterations]"""
import json
import os
import sys

os.environ.setdefault("MIXDIR_B200_TIME_EM", "1")  # small fits are not timed per pass by default

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from conftest import load_golden, subset_recode  # noqa: E402
from mixdir_b200 import Engine, synth  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "m1"
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 6
engine = int(sys.argv[3]) if len(sys.argv) > 3 else 0
g = load_golden("mushroom.npz")
cats = json.loads(str(g["categories"]))
if which == "m1":
    codes, c = subset_recode(g["codes"], cats, range(1000), range(5))
    ncat, K, dp = np.array([len(x) for x in c], dtype=np.int32), 3, False
elif which == "m2":
    codes, c = subset_recode(g["codes"], cats, range(8124), range(23))
    ncat, K, dp = np.array([len(x) for x in c], dtype=np.int32), 10, True
else:
    ncat, K, dp = np.full(60, 10, dtype=np.int32), 10, False
    codes, _ = synth.make_codes(1, 70_000, ncat, 10, p_na=0.10)
N = codes.shape[0]
z0 = synth.zeta_init(1, N, K)
p0 = synth.phi_init(1, ncat, K)
with Engine(codes, ncat) as e:
    e.set_engine(engine)
    for rep in range(2):
        r = e.fit(K, z0.sum(0), p0, dp=dp, max_iter=iters, epsilon=-np.inf, want_class_prob=False)
        tm = e.timing()
        print(which, "engine", tm["engine"], "KC", tm["classes_per_cta"], "P", tm["cluster_size"], "ctas",
              tm["grid_ctas"], "iters", r["n_iter"], "us/iter %.1f" % (tm["fit_loop_ms"] / r["n_iter"] * 1e3),
              "em us/iter %.1f" % (tm["em_kernel_ms"] / r["n_iter"] * 1e3), "launches", tm["total_launches"])
