"""Timing of one find_defining_features round (GPU pass vs the reference's |remaining| predict()
calls restated on the CPU).   python tools/feature_div_timing.py [--rows N]"""
import argparse
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixdir_b200 import Engine, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402  (CPU baseline leg only)

C_PATTERN = [16, 12, 8, 4, 2, 10]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=2_000_000)
    ap.add_argument("--features", type=int, default=60)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--cpu-rows", type=int, default=2000)
    a = ap.parse_args()
    J, K = a.features, a.k
    ncat = np.array((C_PATTERN * ((J + 5) // 6))[:J], dtype=np.int32)
    codes = synth.make_codes_torch(2, a.rows, ncat, K, device="cuda:0").cpu().numpy()
    torch.cuda.empty_cache()
    with Engine(codes, ncat) as eng:
        r = eng.fit(K, np.full(K, a.rows / K), synth.phi_init(2, ncat, K), max_iter=5, epsilon=-np.inf,
                    want_class_prob=False, want_pred_class=False)
        act = np.ones(J, dtype=np.uint8)
        for measure in ("JS", "ARI"):
            eng.feature_divergence(K, r["lambda_"], r["category_prob"], act, measure)
            t = time.perf_counter()
            out = eng.feature_divergence(K, r["lambda_"], r["category_prob"], act, measure)
            dt = time.perf_counter() - t
            print(f"GPU  {measure}: {a.rows} rows x {J} features, K={K}, all {J} candidates in one pass: "
                  f"{dt * 1e3:.1f} ms  ({a.rows * J * J / dt:.3g} row-feature-candidates/s); "
                  f"min/max loss {np.nanmin(out[:J]):.4g} / {np.nanmax(out[:J]):.4g}")
    sub = np.arange(a.cpu_rows)
    X = po.codes_to_r_matrix(codes[:a.cpu_rows])
    t = time.perf_counter()
    po.find_defining_features(X, ncat, K, r["lambda_"], r["category_prob"], sub, n_features=J - 1,
                              step_size=1, exponential_decay=False)
    dt = time.perf_counter() - t
    print(f"CPU  JS: oracle (literal: {J} predict_class calls), {a.cpu_rows} rows, 1 thread: {dt:.2f} s per "
          f"round ({a.cpu_rows * J * J / dt:.3g} row-feature-candidates/s)")


if __name__ == "__main__":
    main()
