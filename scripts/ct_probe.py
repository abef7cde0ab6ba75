"""ComputeTest at a 2048-component world + 5 clients (launch-list probe)."""
import sys, time
import numpy as np
sys.path.insert(0, ".")
from lia_ral_b200 import capi, synth
capi.init(0)
C, D, T = 2048, 60, 131072
w, mean, cov = synth.make_ubm(C, D, seed=1)
X = synth.make_frames(w, mean, cov, T, seed=2)
g = capi.GMM(w, mean, cov)
clients = [capi.GMM(*synth.perturb_ubm(w, mean, cov, seed=30 + i, frac=0.3, scale=0.3)) for i in range(5)]
capi.compute_test(g, clients, X, K=10)
t0 = time.perf_counter(); capi.compute_test(g, clients, X, K=10); print("s", time.perf_counter() - t0)
