"""Final pass / predict timings at a bench shape: kernel (CUDA events inside the library) and wall
clock incl. the read-back, pred_class only and with class_prob, new unit kernel vs the older ones.
    python tools/predict_timing.py [rows] [K] [J]"""
import os, sys, time
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mixdir_b200 import Engine, synth
N = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
K = int(sys.argv[2]) if len(sys.argv) > 2 else 32
J = int(sys.argv[3]) if len(sys.argv) > 3 else 60
ncat = np.array(([16, 12, 8, 4, 2, 10] * 17)[:J], dtype=np.int32)
dev = synth.make_codes_torch(2, N, ncat, K)
host = torch.empty(dev.shape, dtype=dev.dtype, pin_memory=True); host.copy_(dev); del dev
torch.cuda.synchronize()
codes = host.numpy()
p0 = synth.phi_init(2, ncat, K); cs0 = np.full(K, N / K)
e = Engine(codes, ncat)
r = e.fit(K, cs0, p0, max_iter=3, epsilon=-np.inf, want_class_prob=False, want_pred_class=False, n_cat_bound=16)
alg_pc = N * J + N * 4
alg_cp = N * J + N * K * 8 + N * 4
for rep in range(3):
    t = time.perf_counter(); cp, pc = e.predict(K, r["lambda_"], r["category_prob"], want_class_prob=False)
    w = (time.perf_counter() - t) * 1e3; tm = e.timing()
    print("pred_class only: kernel %.3f ms (engine %d, %.0f GB/s algorithmic), wall incl. read-back %.2f ms"
          % (tm["final_pass_kernel_ms"], tm["final_pass_engine"], alg_pc / tm["final_pass_kernel_ms"] / 1e6, w))
cp = None
for rep in range(2):
    t = time.perf_counter(); cp, pc2 = e.predict(K, r["lambda_"], r["category_prob"])
    w = (time.perf_counter() - t) * 1e3; tm = e.timing()
    print("class_prob + pred_class: kernel %.3f ms (engine %d, %.0f GB/s algorithmic), wall incl. %.2f GB read-back %.1f ms"
          % (tm["final_pass_kernel_ms"], tm["final_pass_engine"], alg_cp / tm["final_pass_kernel_ms"] / 1e6, N * K * 8 / 1e9, w))
    assert np.array_equal(pc, pc2)
os.environ["MIXDIR_B200_OLD_PREDICT"] = "1"
t = time.perf_counter(); cp0, pc0 = e.predict(K, r["lambda_"], r["category_prob"], want_class_prob=N <= 4_000_000)
w = (time.perf_counter() - t) * 1e3; tm = e.timing()
print("older kernel: kernel %.3f ms (engine %d), wall %.1f ms; pred_class identical %s%s"
      % (tm["final_pass_kernel_ms"], tm["final_pass_engine"], w, np.array_equal(pc, pc0),
         "" if cp0 is None else ", class_prob max abs diff %.2e" % np.max(np.abs(cp - cp0))))
