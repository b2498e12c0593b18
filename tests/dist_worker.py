"""Worker of tests/test_gpu_dist.py: one process per GPU (launched with torch.distributed.run), the
flavour bench.py --gpus N uses: mixdir_b200_comm_init + mixdir_b200_create_dist[_from_r_matrix].
Every rank fits its contiguous row shard; rank 0 compares the sharded result with
  (a) the same library on ONE GPU with the whole matrix,
  (b) the CPU oracle (R loop restated, oracle/mixdir_oracle.c) on the whole matrix,
and writes a JSON report.  torch.distributed (gloo) only carries the NCCL id and the test's own
result gathering; the data path's all-reduce is the library's NCCL communicator."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mixdir_b200 import Comm, Engine, synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b) / np.abs(b))) if a.shape == b.shape and a.size else float("inf")


def gather_rows(x, rank, world):
    """variable-length gather of per-shard row results onto every rank (gloo, CPU tensors)"""
    out = [None] * world
    dist.all_gather_object(out, np.ascontiguousarray(x))
    return np.concatenate(out, axis=0)


def main():
    out_path = sys.argv[1]
    rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
    local = int(os.environ.get("LOCAL_RANK", rank))
    dist.init_process_group("gloo")
    torch.cuda.set_device(local)
    uid = [Engine.nccl_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(uid, 0)
    comm = Comm(local, rank, world, uid[0])
    report = {"world": world, "cases": []}

    cases = [
        # name, N, ncat, K, dp, p_na, iterations, engine (0 = automatic), from_r_matrix
        ("seg_auto_k20", 40_003, [16, 12, 8, 4, 2, 10] * 4, 20, False, 0.05, 6, 0, False),
        ("seg_auto_dp_k32", 40_003, [16, 12, 8, 4, 2, 10] * 4, 32, True, 0.0, 6, 0, False),
        ("units_k12", 9_001, [5, 3, 7, 2, 4, 6, 3, 3], 12, False, 0.1, 8, 4, False),
        ("fused_k12_dp", 9_001, [5, 3, 7, 2, 4, 6, 3, 3], 12, True, 0.1, 8, 2, False),
        ("generic_k5", 3_001, [4, 3, 5, 2], 5, False, 0.2, 8, 1, False),
        ("r_matrix_k10", 20_001, [10] * 12, 10, False, 0.1, 6, 0, True),
        # rank 0's shard has no NA at all, the other ranks do: the NA-row rule needs the all-reduced scan
        ("na_only_on_last_ranks", 30_000, [6, 4, 9, 3, 5, 2], 8, False, -1.0, 6, 0, False),
    ]
    for name, N, ncat, K, dp, p_na, iters, engine, from_r in cases:
        ncat = np.asarray(ncat, dtype=np.int32)
        if p_na < 0:
            codes, _ = synth.make_codes(11, N, ncat, K, p_na=0.0)
            tail, _ = synth.make_codes(12, N, ncat, K, p_na=0.15)
            first = N // world
            codes[first:] = tail[first:]
        else:
            codes, _ = synth.make_codes(11, N, ncat, K, p_na=p_na)
        z0 = synth.zeta_init(11, N, K)
        cs0 = z0.sum(axis=0)
        p0 = synth.phi_init(11, ncat, K)
        lo, hi = N * rank // world, N * (rank + 1) // world
        shard = np.ascontiguousarray(codes[lo:hi])
        if from_r:
            eng = Engine.from_r_matrix(po.codes_to_r_matrix(shard), ncat, dist=comm)
        else:
            eng = Engine(shard, ncat, dist=comm)
        with eng:
            if engine:
                eng.set_engine(engine)
            n_missing = eng.n_missing
            r = eng.fit(K, cs0, p0, dp=dp, max_iter=iters, epsilon=-np.inf)
            used = eng.timing()["engine"]
            cp2, pc2 = eng.predict(K, r["lambda_"], r["category_prob"])
        pc_all = gather_rows(r["pred_class"], rank, world)
        cp_all = gather_rows(r["class_prob"], rank, world)
        pred_same = bool(np.array_equal(pc2, r["pred_class"]) and np.max(np.abs(cp2 - r["class_prob"])) < 1e-12)
        flags = [None] * world
        dist.all_gather_object(flags, (pred_same, float(r["convergence"][-1]), r["phi"].tobytes()))
        if rank != 0:
            continue
        with Engine(codes, ncat) as e1:
            if engine:
                e1.set_engine(engine)
            s = e1.fit(K, cs0, p0, dp=dp, max_iter=iters, epsilon=-np.inf)
        ref = po.fit(po.codes_to_r_matrix(codes), ncat, K, int(dp), 1.0, 1.0, 0.1, iters, -np.inf, z0, p0)
        case = {
            "name": name, "engine": used, "n_missing_ok": bool(n_missing == int((codes == 0).sum())),
            "ranks_agree": bool(all(f[1] == flags[0][1] and f[2] == flags[0][2] for f in flags)),
            "predict_equals_fit_final_pass": bool(all(f[0] for f in flags)),
            "elbo_rel_vs_single": rel(r["convergence"], s["convergence"]),
            # k_em_seg: statistics and class masses are exact integer sums -> same bits on 1 and N GPUs
            "phi_bit_identical_vs_single": bool(np.array_equal(r["phi"], s["phi"])),
            "elbo_rel_vs_oracle": rel(r["convergence"], ref["convergence"]),
            "phi_rel_vs_oracle": rel(r["phi"], ref["phi"]),
            "class_prob_abs_vs_oracle": float(np.max(np.abs(cp_all - ref["class_prob"]))),
            "pred_class_identical_vs_single": bool(np.array_equal(pc_all, s["pred_class"])),
            "pred_class_identical_vs_oracle": bool(np.array_equal(pc_all, ref["pred_class"])),
        }
        case["ok"] = bool(case["n_missing_ok"] and case["ranks_agree"] and case["predict_equals_fit_final_pass"]
                          and case["elbo_rel_vs_single"] < 1e-11 and case["elbo_rel_vs_oracle"] < 1e-9
                          and case["phi_rel_vs_oracle"] < 1e-7 and case["class_prob_abs_vs_oracle"] < 1e-6
                          and case["pred_class_identical_vs_single"] and case["pred_class_identical_vs_oracle"]
                          and (case["phi_bit_identical_vs_single"] or used != 5))
        report["cases"].append(case)
    if rank == 0:
        report["ok"] = bool(all(c["ok"] for c in report["cases"]))
        with open(out_path, "w") as f:
            json.dump(report, f, indent=1)
    dist.barrier()
    comm.close()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
