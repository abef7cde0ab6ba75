"""Profiling experiment: time the tcgen05 kernels with parts disabled (flags) / cluster off (32)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lia_ral_b200 import capi, synth
capi.init(0)
C, D, T = 2048, 60, 1_000_000
w, mean, cov = synth.make_ubm(C, D, seed=1)
X = torch.randn(T, D, device="cuda") * 2.0
torch.cuda.synchronize()
feats = capi.Feats(device_ptr=X.data_ptr(), T=T, ldx=D, D=D)
g = capi.GMM(w, mean, cov)
capi.set_gmm_kernel(2)
stats = torch.zeros(g.em_stats_len(), dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
flag_sets = [int(a) for a in sys.argv[1:]] or [0, 32, 2, 8, 15, 15 | 32]
for flags in flag_sets:
    capi.lib().lr_debug_flags(flags)  # needs a `make DEBUG=1` build
    for rep in range(2):
        capi.profile(True)
        g.em_accumulate_dev(feats, 0, T, 1.0, stats.data_ptr())
        a = capi.profile_read(0); b = capi.profile_read(1)
        capi.profile(False)
    print(f"flags {flags:2d}: lse {a[0]:.3f} ms ({T / a[0] / 1e3:.0f} M/s)  acc {b[0]:.3f} ms ({T / b[0] / 1e3:.0f} M/s)", flush=True)
capi.lib().lr_debug_flags(0)
