"""How far the ELBO trace of the sorted-segment kernel (engine 5: responsibilities rounded to 32-bit
fixed point before they enter the phi statistics, exact integer sums) drifts from the fp64-statistics
unit kernel (engine 4) on the same data and start: relative difference per iteration, over sizes, K
and seeds; and the largest class_prob difference when both fits stop after 6, 12 and 25 iterations
(north_star: class_prob within 1e-6).   python tools/seg_precision.py [quick]"""
import os
import sys

sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402

from mixdir_b200 import Engine, synth  # noqa: E402

quick = len(sys.argv) > 1
ncat = np.array([16, 12, 8, 4, 2, 10] * 4, dtype=np.int32)
cases = [(32, 16384, range(2, 4)), (10, 70000, range(2, 4))] if quick else \
    [(32, 16384, range(2, 10)), (32, 65536, range(2, 18)), (32, 300000, range(2, 6)), (32, 2000000, range(2, 4)),
     (10, 70000, range(2, 6)), (64, 100000, range(2, 4))]
# (K = 64 runs on the fp64 unit kernel by default; MIXDIR_B200_SEG_GLOBAL_STATS=1 puts it on k_em_seg)
for K, N, seeds in cases:
    mx, fin, med, cpmax = [], [], [], []
    for seed in seeds:
        codes, _ = synth.make_codes(seed, N, ncat, K, p_na=0.05)
        z0 = synth.zeta_init(seed, min(N, 65536), K).sum(0) * (N / min(N, 65536))
        p0 = synth.phi_init(seed, ncat, K)
        out = {}
        for eng_id in (5, 4):
            with Engine(codes, ncat) as e:
                e.set_engine(eng_id)
                out[eng_id] = e.fit(K, z0, p0, max_iter=25, epsilon=-np.inf, want_class_prob=False)
        a, b = out[5]["convergence"], out[4]["convergence"]
        rel = np.abs(a - b) / np.abs(b)
        same = np.array_equal(out[5]["pred_class"], out[4]["pred_class"])
        mx.append(rel.max()); fin.append(rel[-1]); med.append(np.median(rel))
        cps = []
        if N <= 300000 and seed < 6:
            for stop in (6, 12, 25):
                cp = {}
                for eng_id in (5, 4):
                    with Engine(codes, ncat) as e:
                        e.set_engine(eng_id)
                        cp[eng_id] = e.fit(K, z0, p0, max_iter=stop, epsilon=-np.inf, want_pred_class=False)["class_prob"]
                cps.append(f"{np.max(np.abs(cp[5] - cp[4])):.1e}")
                cpmax.append(float(np.max(np.abs(cp[5] - cp[4]))))
        print(f"K={K} N={N} seed={seed}: max {rel.max():.2e} (iteration {rel.argmax() + 1}), median {np.median(rel):.2e}, "
              f"after 25 iterations {rel[-1]:.2e}, pred_class identical {same}"
              + (f", class_prob max abs diff after 6/12/25 iterations {'/'.join(cps)}" if cps else ""), flush=True)
    print(f"== K={K} N={N}: over {len(mx)} starts: worst max {max(mx):.2e}, median of the maxima {np.median(mx):.2e}, "
          f"worst final {max(fin):.2e}, median of the per-fit medians {np.median(med):.2e}"
          + (f", class_prob: worst {max(cpmax):.1e}, median {np.median(cpmax):.1e}" if cpmax else ""), flush=True)
