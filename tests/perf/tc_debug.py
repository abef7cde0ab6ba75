"""Bring-up check of the tcgen05 path against the SIMT path and the oracle (run on the GPU box)."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lia_ral_b200 import capi, synth
from oracle.ffi import Oracle

capi.init(0)
orc = Oracle()
C, D, T = int(sys.argv[1]) if len(sys.argv) > 1 else 2048, 60, int(sys.argv[2]) if len(sys.argv) > 2 else 1000
w, mean, cov = synth.make_ubm(C, D, seed=1)
X = synth.make_frames(w, mean, cov, T, seed=2)
w, mean, cov = synth.perturb_ubm(w, mean, cov * 4.0, seed=3, frac=0.5, scale=1.0)
g = capi.GMM(w, mean, cov)
o = orc.gmm(w, mean, cov)
ref = orc.llk_all(o, X, -1e9, 1e9)
capi.set_gmm_kernel(1)
simt = g.llk(X, -1e9, 1e9)
print("simt llk max err", np.abs(simt - ref).max(), flush=True)
capi.set_gmm_kernel(2)
t0 = time.time()
tc = g.llk(X, -1e9, 1e9)
print("tc   llk max err", np.abs(tc - ref).max(), "time", time.time() - t0, flush=True)
print("ref[:4]", ref[:4], "tc[:4]", tc[:4])
bad = np.argsort(-np.abs(tc - ref))[:5]
print("worst frames", bad, (tc - ref)[bad])
U = 3
segs = [(0, T // 2, 0), (T // 2, T // 4, 1), (3 * T // 4, T - 3 * T // 4, 2)]
f2r = np.zeros(T, dtype=np.int32); f2r[T // 2:] = 1; f2r[3 * T // 4:] = 2
N_ref, F_ref = orc.bwstats(o, X, f2r, U)
N, F = g.bwstats(X, segs, U)
print("tc bw  N relmax", np.abs(N - N_ref).max() / np.abs(N_ref).max(), "F relmax", np.abs(F - F_ref).max() / np.abs(F_ref).max(), flush=True)
Fc = F.reshape(U, C, D) - N[:, :, None] * mean[None]
Fc_ref = F_ref.reshape(U, C, D) - N_ref[:, :, None] * mean[None]
print("tc bw  centred F relmax", np.abs(Fc - Fc_ref).max() / np.abs(Fc_ref).max())
llk_r, n_r, occ_r, m1_r, m2_r = orc.em_accumulate(o, X)
llk, n, occ, m1, m2 = g.em_accumulate(X)
print("tc em llk", llk, llk_r, "occ relmax", np.abs(occ - occ_r).max() / occ_r.max(), "m1", np.abs(m1 - m1_r).max() / np.abs(m1_r).max(), "m2", np.abs(m2 - m2_r).max() / np.abs(m2_r).max(), flush=True)
wd, mud, cvd = orc.em_get(o, occ, m1, m2)
wo, muo, cvo = orc.em_get(o, occ_r, m1_r, m2_r)
heavy = occ_r > 1.0
print("tc em var relmax (heavy comps)", np.abs(cvd[heavy] - cvo[heavy]).max() / np.abs(cvo[heavy]).max())
capi.set_gmm_kernel(0)
