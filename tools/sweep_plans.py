"""Tuning aid (GPU): time one VI iteration of the unit kernel under forced plans
(MIXDIR_B200_FORCE_PLAN="KC,RG,pairs"; write x for a free field) on the bench workload.
    python tools/sweep_plans.py [--rows N] [--features J] [--k K] plan [plan ...]
"""
import argparse
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixdir_b200 import Engine, synth  # noqa: E402

C_PATTERN = [16, 12, 8, 4, 2, 10]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--rows", type=int, default=10_000_000)
    ap.add_argument("--features", type=int, default=60)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--engine", type=int, default=4)
    ap.add_argument("plans", nargs="*", default=["x,x,x"])
    a = ap.parse_args()
    ncat = np.array((C_PATTERN * ((a.features + 5) // 6))[:a.features], dtype=np.int32)
    dev = synth.make_codes_torch(2, a.rows, ncat, a.k, device="cuda:0")
    host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True)
    host.copy_(dev)
    del dev
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    codes = host.numpy()
    cs0 = np.full(a.k, a.rows / a.k)
    p0 = synth.phi_init(2, ncat, a.k)
    with Engine(codes, ncat) as eng:
        for plan in a.plans:
            if plan == "old":
                os.environ.pop("MIXDIR_B200_FORCE_PLAN", None)
                eng.set_engine(2)
            else:
                os.environ["MIXDIR_B200_FORCE_PLAN"] = plan.replace("x", "-1")
                eng.set_engine(a.engine)
            try:
                eng.fit(a.k, cs0, p0, max_iter=2, epsilon=-np.inf, want_class_prob=False, want_pred_class=False)
                r = eng.fit(a.k, cs0, p0, max_iter=a.steps, epsilon=-np.inf, want_class_prob=False,
                            want_pred_class=False)
            except Exception as e:  # noqa: BLE001
                print(f"plan {plan}: {e}")
                continue
            t = eng.timing()
            print(f"plan {plan:>10}: {t['fit_loop_ms'] / a.steps:8.3f} ms/iter  em {t['em_kernel_ms'] / a.steps:8.3f}  "
                  f"engine {t['engine']} KC {t['classes_per_cta']} P {t['cluster_size']} ctas {t['grid_ctas']} "
                  f"smem {t['smem_bytes']} U {t['unit_elements']} pairs {t['unit_pairs']} rows2 {t['unit_table_rows']} "
                  f"elbo {r['convergence'][-1]:.6f}", flush=True)


if __name__ == "__main__":
    main()
