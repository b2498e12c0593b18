"""Where does an e2e step go?  create (upload + scan) / fit(1 iteration) / destroy, S4 shape."""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixdir_b200 import Engine, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
ncat = np.array([16, 12, 8, 4, 2, 10] * 10, dtype=np.int32); K = 32
dev = synth.make_codes_torch(2, N, ncat, K)
host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True); host.copy_(dev); del dev
torch.cuda.synchronize()
codes = host.numpy()
p0 = synth.phi_init(2, ncat, K); cs0 = np.full(K, N / K)
# raw pinned H2D bandwidth for reference
d = torch.empty_like(host, device="cuda")
torch.cuda.synchronize(); t = time.perf_counter(); d.copy_(host, non_blocking=True); torch.cuda.synchronize()
print("torch pinned H2D: %.2f ms  (%.1f GB/s)" % ((time.perf_counter() - t) * 1e3, host.numel() / (time.perf_counter() - t) / 1e9))
del d
for rep in range(4):
    t0 = time.perf_counter(); e = Engine(codes, ncat); t1 = time.perf_counter()
    r = e.fit(K, cs0, p0, max_iter=1, epsilon=-np.inf, want_class_prob=False, want_pred_class=False, n_cat_bound=16)
    t2 = time.perf_counter(); tm = e.timing(); e.close(); t3 = time.perf_counter()
    print("create %.2f ms  fit %.2f ms (loop %.2f, em %.2f)  destroy %.2f ms  total %.2f" % (
        (t1 - t0) * 1e3, (t2 - t1) * 1e3, tm["fit_loop_ms"], tm["em_kernel_ms"], (t3 - t2) * 1e3, (t3 - t0) * 1e3))
