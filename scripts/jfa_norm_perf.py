"""Throughput of lr_jfa_normalize_features (JFAAcc::normalizeFeatures) at 2048c/60d: frames/s through the C ABI with
host frames (H2D, likelihood pass with S dumped, k_jfa_compensate, D2H inside the call)."""
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lia_ral_b200 import capi, synth  # noqa: E402

capi.init(0)
C, D, T = 2048, 60, 200_000
w, mean, cov = synth.make_ubm(C, D, seed=1)
X = synth.make_frames(w, mean, cov * 3.0, T, seed=2)
ux = 0.3 * np.sqrt(cov) * np.random.default_rng(3).standard_normal((C, D))
g = capi.GMM(w, mean + ux, cov * 3.0)
g.jfa_normalize_features(ux, X[:4096], [(0, 4096, 0)])
capi.profile(True)
t0 = time.perf_counter()
Y = g.jfa_normalize_features(ux, X, [(0, T, 0)])
dt = time.perf_counter() - t0
lse_ms, n = capi.profile_read(0)
capi.profile(False)
print(json.dumps({"frames": T, "seconds": dt, "frames_per_s": T / dt, "likelihood_pass_ms": lse_ms, "launches": n,
                  "moved": float(np.abs(Y - X).max())}))
