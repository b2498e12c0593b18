#!/usr/bin/env python
"""Benchmark of the mixdir VI hot path (BASELINE.json metric: observation-features/sec per VI
iteration, 10M x 60, K=32, on 1/2/4/8 B200; % of HBM peak).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one coordinate-ascent iteration (E-step + M-step + ELBO + convergence test,
/root/reference/R/mixdir_vi.R:58-115) over the whole synthetic code matrix.

ours:       value   K iterations timed on the device (CUDA events on the engine's stream,
                    max over ranks) with the codes already resident in HBM;
            e2e     the same metric through the C-ABI with HOST buffers: every step uploads the
                    code matrix from pinned host memory (mixdir_b200_create*), runs one
                    iteration (mixdir_b200_fit) and reads the results back;
            roofline  the fused E+M kernel: algorithmic bytes = N*J*code_bytes per launch over
                    its CUDA-event duration, against the measured HBM copy bandwidth;
            cpu_baseline  the reference's CPU path (oracle/_ref: src/mixdir_impl.cpp compiled
                    verbatim + the restated R loop) on a bounded row sample, 1 thread.
reference:  the reference's CPU implementation of the path on this box's host cores (same
            oracle/_ref code), bounded sample per step; rank 0 only.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = "observation-features/sec per VI iteration"
UNIT = "obs-features/s"
C_PATTERN = [16, 12, 8, 4, 2, 10]  # S4 categories per feature, repeating


def workload(args):
    if args.cats > 0:  # uniform number of categories (e.g. S3: 10 per feature)
        return np.full(args.features, args.cats, dtype=np.int32)
    ncat = np.array((C_PATTERN * ((args.features + 5) // 6))[:args.features], dtype=np.int32)
    return ncat


def hbm_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """dram bytes per row of the hot kernel from the committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", "em_traffic.json")) as f:
            return json.load(f)
    except Exception:
        return None


def shared_memory_roofline(traffic, tm, N, J, K, kernel_ms, clocks):
    """Secondary roofline: shared-memory wavefronts of the hot kernel -- counted by ncu
    on the committed capture, scaled to this launch's rows -- over the kernel's duration, against the
    pipe's peak of one wavefront per SM and cycle at the SM clock sampled during the run."""
    if not (traffic and tm["engine"] == traffic.get("engine") and J == 60 and K == 32
            and traffic.get("smem_wavefronts_per_row")):
        return None
    mhz = (clocks or {}).get("sm_mhz") or 1965.0
    sms = 148
    wf = traffic["smem_wavefronts_per_row"] * N
    peak = sms * mhz * 1e6
    ach = wf / (kernel_ms * 1e-3)
    return {"bound": "shared-memory wavefronts", "achieved": ach, "peak": peak, "unit": "wavefronts/s",
            "frac": ach / peak, "wavefronts_per_row": traffic["smem_wavefronts_per_row"],
            "source": traffic.get("source")}


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.index)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), line.strip()))

    def stop(self, t0, t1):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        sm, mx, reasons = [], None, set()
        for t, line in self.rows:
            if t < t0 or t > t1:
                continue
            f = [x.strip() for x in line.split(",")]
            try:
                sm.append(float(f[0]))
                mx = float(f[1])
            except Exception:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:  # region shorter than the sampling period: take what there is
            for t, line in self.rows[-3:]:
                f = [x.strip() for x in line.split(",")]
                try:
                    sm.append(float(f[0]))
                    mx = float(f[1])
                except Exception:
                    pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": mx,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_iteration(ncat, K, rows, iters, seed=2):
    """Time the reference's CPU path on `rows` synthetic rows: oracle/_ref when present
    (kind 'reference'), else the oracle port.  Returns (obs_features_per_s, kind, seconds_per_iter)."""
    from mixdir_b200 import synth
    from oracle import pyoracle as po

    J = len(ncat)
    codes, _ = synth.make_codes(seed, rows, ncat, K)
    X = po.codes_to_r_matrix(codes)
    z0 = synth.zeta_init(seed, rows, K)
    p0 = synth.phi_init(seed, ncat, K)
    which = "ref" if po.ref() is not None else "oracle"
    t = time.perf_counter()
    po.fit(X, ncat, K, 0, 1.0, 1.0, 0.1, iters, -np.inf, z0, p0, which=which, final_pass=False)
    dt = (time.perf_counter() - t) / iters
    return rows * J / dt, ("reference" if which == "ref" else "port"), dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    ncat = workload(args)
    K, J = args.k, args.features
    # calibrate, then size the per-step sample so the whole run stays within ~2 minutes
    _, kind, dt = cpu_reference_iteration(ncat, K, 2000, 1)
    per_row = dt / 2000
    budget = 120.0 / max(args.steps + args.warmup, 1)
    rows = int(max(1000, min(args.rows, budget / per_row, 50000)))
    from mixdir_b200 import synth
    from oracle import pyoracle as po
    codes, _ = synth.make_codes(2, rows, ncat, K)
    X = po.codes_to_r_matrix(codes)
    z0 = synth.zeta_init(2, rows, K)
    p0 = synth.phi_init(2, ncat, K)
    which = "ref" if kind == "reference" else "oracle"
    for _ in range(args.warmup):
        po.fit(X, ncat, K, 0, 1.0, 1.0, 0.1, 1, -np.inf, z0, p0, which=which, final_pass=False)
    t = time.perf_counter()
    po.fit(X, ncat, K, 0, 1.0, 1.0, 0.1, args.steps, -np.inf, z0, p0, which=which, final_pass=False)
    sec = time.perf_counter() - t
    ms = sec / args.steps * 1e3
    value = rows * J / (ms * 1e-3)
    sample = (f"{rows} rows x {J} features, K={K} (prefix-sized sample of the {args.rows}-row "
              f"workload; cost is linear in rows), {args.steps} iterations, 1 thread: the reference "
              "has no threading (SURVEY.md 2)")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, ncat),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "host": {"nproc": os.cpu_count()},
    }
    print(json.dumps(line))
    return 0


def config_dict(args, ncat):
    return {"workload": f"synthetic {args.rows} x {args.features} per GPU, C_j pattern "
                        f"{C_PATTERN if args.cats <= 0 else args.cats} (sum C_j = {int(ncat.sum())}), K={args.k}, fixed-K Dirichlet "
                        "variant, alpha=1, beta=0.1, p_NA=%g" % args.p_na,
            "rows_per_gpu": args.rows, "features": args.features, "n_latent": args.k,
            "code_bytes": 1 if ncat.max() <= 255 else 2,
            "parallelism": f"rows sharded over {args.gpus} GPU(s), one NCCL all-reduce of the "
                           "statistics buffer per iteration" if args.gpus > 1 else "single GPU",
            "l2_policy": "inputs larger than L2 (code matrix %.0f MB per GPU vs 126 MB L2)"
                         % (args.rows * args.features / 1e6)}


def run_ours(args):
    import torch
    import torch.distributed as dist

    from mixdir_b200 import Comm, Engine, synth

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the engine has no CPU fallback")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    ncat = workload(args)
    K, J, N = args.k, args.features, args.rows

    # ---- synthetic shard, generated on the device, kept in pinned host memory ----
    dev_codes = synth.make_codes_torch(2, N, ncat, K, p_na=args.p_na, row0=rank * N,
                                       device=f"cuda:{local}")
    host_codes = torch.empty(dev_codes.shape, dtype=dev_codes.dtype, pin_memory=True)
    host_codes.copy_(dev_codes)
    del dev_codes
    torch.cuda.synchronize()
    torch.cuda.empty_cache()
    codes_np = host_codes.numpy()
    if codes_np.dtype == np.int16:
        codes_np = codes_np.view(np.uint16)

    uid = None
    if world > 1:
        t = torch.zeros(128, dtype=torch.uint8, device=f"cuda:{local}")
        if rank == 0:
            t.copy_(torch.frombuffer(bytearray(Engine.nccl_unique_id()), dtype=torch.uint8))
        dist.broadcast(t, 0)
        uid = bytes(t.cpu().numpy().tobytes())

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def maxr(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=f"cuda:{local}")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    cs0 = np.full(K, N * world / K)
    p0 = synth.phi_init(2, ncat, K)
    ncb = int(ncat.max())  # n_cat = max(X) (R/mixdir_vi.R:8): every category occurs in this data
    comm = Comm(local, rank, world, uid)  # one NCCL communicator, shared by every handle below
    mk = lambda: Engine(codes_np, ncat, dist=comm)  # noqa: E731

    # ---- value: device-resident, K iterations, CUDA events inside the engine ----
    eng = mk()
    if args.engine:
        eng.set_engine(args.engine)
    eng.fit(K, cs0, p0, max_iter=args.warmup, epsilon=-np.inf, want_class_prob=False,
            want_pred_class=False)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)
    barrier()
    t0 = time.perf_counter()
    res = eng.fit(K, cs0, p0, max_iter=args.steps, epsilon=-np.inf, want_class_prob=False,
                  want_pred_class=False)
    barrier()
    t1 = time.perf_counter()
    tm = eng.timing()
    clocks = sampler.stop(t0, t1) if rank == 0 else None
    loop_ms = maxr(tm["fit_loop_ms"])
    em_ms = maxr(tm["em_kernel_ms"])
    ms_step = loop_ms / args.steps
    value = N * world * J / (ms_step * 1e-3)
    assert res["n_iter"] == args.steps

    # ---- final pass (reported separately) ----
    _ = eng.predict(K, res["lambda_"], res["category_prob"], want_class_prob=False)  # loads the kernel
    barrier()
    tp = time.perf_counter()
    _ = eng.predict(K, res["lambda_"], res["category_prob"], want_class_prob=False)
    barrier()
    predict_ms = (time.perf_counter() - tp) * 1e3
    eng.close()

    # ---- e2e: host buffers through the C ABI, upload + 1 iteration + read-back per step ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(1):
        with mk() as e:
            e.fit(K, cs0, p0, max_iter=1, epsilon=-np.inf, want_class_prob=False, want_pred_class=False,
                  n_cat_bound=ncb)
    barrier()
    d2h = 0
    step_ms = []
    for _ in range(e2e_steps):
        barrier()
        ts = time.perf_counter()
        with mk() as e:
            tc = time.perf_counter()
            r = e.fit(K, cs0, p0, max_iter=1, epsilon=-np.inf, want_class_prob=False,
                      want_pred_class=False, n_cat_bound=ncb)
            tf = time.perf_counter()
        step_ms.append((time.perf_counter() - ts) * 1e3)
        if os.environ.get("BENCH_DEBUG") and rank == 0:
            print("e2e step: create %.2f fit %.2f destroy %.2f ms" % (
                (tc - ts) * 1e3, (tf - tc) * 1e3, (time.perf_counter() - tf) * 1e3), file=sys.stderr)
        d2h = r["convergence"].nbytes + r["lambda_"].nbytes + r["phi"].nbytes + \
            r["category_prob"].nbytes + r["omega"].nbytes + 32
    barrier()
    # the host side of this box is shared: report the median step (mean and min beside it)
    e2e_ms = maxr(statistics.median(step_ms))
    e2e_mean, e2e_min = maxr(sum(step_ms) / len(step_ms)), maxr(min(step_ms))
    e2e_value = N * world * J / (e2e_ms * 1e-3)
    h2d = codes_np.nbytes + cs0.nbytes + p0.nbytes

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return 0

    peak, peak_src = hbm_peak()
    cb = codes_np.dtype.itemsize
    alg_bytes = N * J * cb  # per launch, per GPU: each code read once
    em_ms_launch = em_ms / max(tm["em_launches"], 1)
    achieved = alg_bytes / (em_ms_launch * 1e-3) / 1e9
    traffic = ncu_traffic()
    cpu_val, cpu_kind, cpu_dt = cpu_reference_iteration(ncat, K, args.cpu_rows, 2)
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, ncat),
        "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": e2e_ms, "steps": e2e_steps,
                "ms_per_step_mean": e2e_mean, "ms_per_step_min": e2e_min,
                "what": "mixdir_b200_create_dist(host codes, pinned; asynchronous chunked upload) + "
                        "mixdir_b200_fit(max_iter=1; consumes the chunks as they land) + result "
                        "read-back + mixdir_b200_destroy, per step"},
        "gpu_launches": int(tm["total_launches"]),
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                     "frac": achieved / peak,
                     "traffic": (int(traffic["dram_bytes_per_row"] * N) if traffic and
                                 tm["engine"] == traffic.get("engine", 2) and J == 60 and K == 32
                                 else None),
                     "traffic_source": (traffic or {}).get("source"),
                     "kernel": {2: "k_em_fused", 4: "k_em_units", 5: "k_em_seg"}.get(tm["engine"],
                                                                                       "k_em_generic"),
                     "algorithmic_bytes_per_launch": int(alg_bytes),
                     "kernel_ms_per_launch": em_ms_launch, "peak_source": peak_src,
                     "kernel_share_of_step": em_ms / loop_ms,
                     "shared_memory": shared_memory_roofline(traffic, tm, N, J, K, em_ms_launch, clocks),
                     "note": "shared-memory bound, not HBM bound: see DESIGN.md 4.1 (per code, K fp64 "
                             "table gathers + K fp64 histogram read-modify-writes; the shared-memory "
                             "pipe runs at ~81 % of its wavefront peak, profiles/README.md)"},
        "cpu_baseline": {"value": cpu_val, "unit": UNIT, "cores": 1, "kind": cpu_kind,
                         "sample": f"{args.cpu_rows} rows x {J} features, K={K}, 2 iterations "
                                   f"({cpu_dt:.2f} s/iter), same generator as the GPU workload; "
                                   "1 thread (the reference is single-threaded)",
                         "host_nproc": os.cpu_count()},
        "clocks": clocks,
        "engine": {"engine": tm["engine"], "classes_per_cta": tm["classes_per_cta"],
                   "cluster_size": tm["cluster_size"], "grid_ctas": tm["grid_ctas"],
                   "smem_bytes": tm["smem_bytes"], "unit_elements": tm["unit_elements"],
                   "feature_pairs": tm["unit_pairs"], "unit_table_rows": tm["unit_table_rows"]},
        "final_pass_ms": predict_ms,
        "elbo_last": float(res["convergence"][-1]),
        "wall_ms_per_step": (t1 - t0) * 1e3 / args.steps,
    }
    print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=10_000_000, help="rows per GPU")
    ap.add_argument("--features", type=int, default=60)
    ap.add_argument("--k", type=int, default=32)
    ap.add_argument("--p-na", dest="p_na", type=float, default=0.0)
    ap.add_argument("--cats", type=int, default=0, help="uniform categories per feature (0: S4 pattern)")
    ap.add_argument("--engine", type=int, default=0)
    ap.add_argument("--e2e-steps", dest="e2e_steps", type=int, default=5)
    ap.add_argument("--cpu-rows", dest="cpu_rows", type=int, default=20000)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    if args.impl == "reference":
        return run_reference(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
