"""Timing probe of the digit GEMM at the L = N TETt shape (1024 x 80200 x 2048, 6 planes)."""
import sys, time, ctypes as ct
import numpy as np
sys.path.insert(0, ".")
from lia_ral_b200 import capi
capi.init(0)
rng = np.random.default_rng(0)
M, N, K = 1024, 20096, 2048
A = rng.random((M, K)); B = rng.standard_normal((N, K))
for planes in (6,):
    capi.gemm_digits(A, B, planes=planes)
    import torch
    torch.cuda.synchronize()
    # time only the device part via the launch profile: use cuda events around the call is polluted by copies;
    # so report wall of 3 calls minus copy estimate -- simpler: nvtx-free approach = ncu. Here: wall.
    t0 = time.perf_counter(); capi.gemm_digits(A, B, planes=planes); dt = time.perf_counter() - t0
    print("planes", planes, "wall ms (incl. H2D/D2H)", dt * 1e3)
