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
e = Engine(codes, ncat)
r = e.fit(K, cs0, p0, max_iter=3, epsilon=-np.inf, want_class_prob=False, want_pred_class=False, n_cat_bound=16)
for rep in range(3):
    t = time.perf_counter(); cp, pc = e.predict(K, r["lambda_"], r["category_prob"], want_class_prob=False)
    print("predict (pred_class only): %.2f ms" % ((time.perf_counter() - t) * 1e3))
t = time.perf_counter(); r2 = e.fit(K, cs0, p0, max_iter=1, epsilon=-np.inf, want_class_prob=False, want_pred_class=True, n_cat_bound=16)
print("fit(1 iter)+pred_class: %.2f ms" % ((time.perf_counter() - t) * 1e3))
if N <= 2_000_000:
    t = time.perf_counter(); cp, pc = e.predict(K, r["lambda_"], r["category_prob"])
    print("predict with class_prob: %.2f ms" % ((time.perf_counter() - t) * 1e3))
