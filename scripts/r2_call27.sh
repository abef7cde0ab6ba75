#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r2c27
timeout -k 10 900 python -m pytest tests/test_gmm_gpu.py tests/test_cli_gpu.py -x -q -m gpu > $O.pytest.log 2>&1; echo "rc=$?" >> $O.pytest.log
tail -n 12 $O.pytest.log | cut -c1-200
python - <<'PY'
import time, numpy as np, sys
sys.path.insert(0, ".")
from lia_ral_b200 import capi, synth
capi.init(0)
C, D, T = 2048, 60, 200000
w, mean, cov = synth.make_ubm(C, D, seed=1)
X = synth.make_frames(w, mean, cov, T, seed=2)
g = capi.GMM(w, mean, cov)
clients = [capi.GMM(*synth.perturb_ubm(w, mean, cov, seed=30 + i, frac=0.3, scale=0.3)) for i in range(5)]
for kern in (1, 0):
    capi.set_gmm_kernel(kern)
    capi.compute_test(g, clients, X[:20000], K=10)
    t0 = time.perf_counter(); mw, mc = capi.compute_test(g, clients, X, K=10); dt = time.perf_counter() - t0
    print("kernel", kern, "ComputeTest 2048c world, 5 clients:", T / dt / 1e6, "M frames/s", mw, mc[:, 0] - mw)
PY
