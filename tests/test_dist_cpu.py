"""world_size-2 CPU test (gloo): the multi-rank protocol of the engine -- contiguous row shards,
per-rank partial statistics buffers, ONE all-reduce(sum) of [S | cs | H | flag] per iteration,
parameter step replicated on every rank -- reproduces the single-rank fit.  The row pass is the
oracle's (no GPU here); the sharding and the buffer layout are the engine's."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

sys.path.insert(0, os.path.dirname(__file__))


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank),
                      WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import fused_model as fm
    from mixdir_b200 import synth
    from oracle import pyoracle as po

    ncat = np.array([4, 3, 5, 2, 6, 3], dtype=np.int32)
    N, K = 401, 5
    codes, _ = synth.make_codes(4, N, ncat, K, p_na=0.1)
    lo, hi = N * rank // world, N * (rank + 1) // world  # engine's shard rule (engine.cu make_shards)
    local = codes[lo:hi]
    cs = synth.zeta_init(4, N, K).sum(axis=0)
    phi = synth.phi_init(4, ncat, K)
    n_cat = int(codes.max())
    nmiss = torch.tensor([float((local == 0).sum())], dtype=torch.float64)
    dist.all_reduce(nmiss)
    trace = []
    for _it in range(6):
        pr = fm.prior_step(cs, K, 1, 1.0, 1.0)
        T = fm.table_from_phi(phi, ncat, K, n_cat)
        st = po.estep_stats(local, ncat, K, pr["logprior"] - 1.0, T)
        buf = torch.from_numpy(np.concatenate([st["S"], st["cs"], [st["H"]], [float(st["underflow"])]]))
        dist.all_reduce(buf)  # the one collective per iteration
        b = buf.numpy()
        S, csn, H = b[:phi.size], b[phi.size:phi.size + K], b[phi.size + K]
        perm = np.argsort(-csn, kind="stable")
        new_phi = np.zeros_like(phi)
        o = 0
        for c in ncat:
            new_phi[o:o + K * c] = (S[o:o + K * c].reshape(K, c)[perm] + 0.1).ravel()
            o += K * c
        phi, cs = new_phi, csn[perm]
        trace.append(float(H))
    if rank == 0:
        q.put((phi, cs, np.array(trace), float(nmiss.item())))
    dist.destroy_process_group()


def test_two_rank_gloo_matches_single_rank():
    import fused_model as fm
    from mixdir_b200 import synth

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    phi2, cs2, h2, nmiss = q.get(timeout=300)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    ncat = np.array([4, 3, 5, 2, 6, 3], dtype=np.int32)
    N, K = 401, 5
    codes, _ = synth.make_codes(4, N, ncat, K, p_na=0.1)
    assert nmiss == float((codes == 0).sum())
    r = fm.fit(codes, ncat, K, 1, 1.0, 1.0, 0.1, 6, -np.inf, synth.zeta_init(4, N, K).sum(axis=0),
               synth.phi_init(4, ncat, K))
    np.testing.assert_allclose(phi2, r["phi"], rtol=1e-11, atol=1e-12)
    np.testing.assert_allclose(np.sort(cs2)[::-1], cs2)  # DP ordering applied to the reduced masses
    assert abs(cs2.sum() - N) < 1e-9
