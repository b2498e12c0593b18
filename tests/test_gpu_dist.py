"""Multi-PROCESS multi-GPU parity (one process per GPU, the flavour bench.py --gpus N measures):
mixdir_b200_comm_init + mixdir_b200_create_dist[_from_r_matrix] on 2 ranks (4 when the box has
them) against one GPU and against the CPU oracle, every kernel family, DP and fixed K, uneven
shards, NAs present on some ranks only.  Skipped on a one-GPU box; bench.py repeats the core of it
before every N > 1 measurement (`parity_vs_single`)."""
import json
import os
import socket
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


@pytest.mark.parametrize("world", [2, 4])
def test_one_process_per_gpu_matches_single_gpu_and_oracle(world, tmp_path):
    if _ngpu() < world:
        pytest.skip(f"needs {world} GPUs")
    out = tmp_path / "report.json"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}",
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()),
           os.path.join(HERE, "dist_worker.py"), str(out)]
    p = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert p.returncode == 0, p.stdout[-3000:] + p.stderr[-3000:]
    rep = json.loads(out.read_text())
    bad = [c for c in rep["cases"] if not c["ok"]]
    assert rep["ok"] and not bad, json.dumps(bad, indent=1)
    assert len(rep["cases"]) == 7
    # the automatic choice took the sorted-segment kernel where the shard is large enough
    assert {c["name"]: c["engine"] for c in rep["cases"]}["seg_auto_k20"] == 5
