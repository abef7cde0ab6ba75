import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lia_ral_b200 import capi, synth
from oracle.ffi import Oracle
capi.init(0); oracle = Oracle()
C, D, U, per = 2048, 60, 10, 20000
w, mean, cov = synth.make_ubm(C, D, seed=1)
X = synth.make_frames(w, mean, cov, U * per, seed=21)
w2, m2, c2 = w, mean * 0.15, cov * 5.0
o = oracle.gmm(w2, m2, c2)
f2r = (np.arange(U * per) // per).astype(np.int32)
N_ref, F_ref = oracle.bwstats(o, X, f2r, U, threads=os.cpu_count() or 1)
for k in (1, 2, 3):
    capi.set_gmm_kernel(k)
    g = capi.GMM(w2, m2, c2)
    N, F = g.bwstats(X, [(u * per, per, u) for u in range(U)], U)
    rel = np.abs(N - N_ref) / np.maximum(N_ref, 1e-300)
    big = N_ref >= 100.0; mid = (N_ref >= 1.0) & ~big
    F3, F3r = F.reshape(U, C, D), F_ref.reshape(U, C, D)
    relF = np.linalg.norm(F3 - F3r, axis=2) / np.maximum(np.linalg.norm(F3r, axis=2), 1e-300)
    print("kernel", k, "n_big", big.sum(), "n_mid", mid.sum(), "rowsum err", np.abs(N.sum(1) - per).max(),
          "relN big %.2e mid*sqrt %.2e | relF big %.2e mid*sqrt %.2e | max-norm N %.2e F %.2e" % (
          rel[big].max() if big.any() else 0, (rel[mid] * np.sqrt(N_ref[mid])).max(), relF[big].max() if big.any() else 0,
          (relF[mid] * np.sqrt(N_ref[mid])).max(), np.abs(N - N_ref).max() / N_ref.max(), np.abs(F - F_ref).max() / np.abs(F_ref).max()))
    # rms of the scaled error
    print("   rms rel*sqrt(N) over mid: %.2e ; over big: %.2e" % (np.sqrt(np.mean((rel[mid] * np.sqrt(N_ref[mid]))**2)), np.sqrt(np.mean((rel[big]*np.sqrt(N_ref[big]))**2)) if big.any() else 0))

# i-vector level
R = 40
invvar = (1.0 / c2).reshape(-1)
Tm = synth.make_T(R, C, D, invvar, seed=23, scale=0.05)
tett = oracle.tv_tett(Tm, invvar, C, D, threads=os.cpu_count() or 1)
W_ref = oracle.tv_ivectors(N_ref, oracle.tv_subtract_m(N_ref, F_ref, m2.reshape(-1)), Tm, invvar, tett)
for k in (1, 2, 3):
    capi.set_gmm_kernel(k)
    g = capi.GMM(w2, m2, c2)
    N, F = g.bwstats(X, [(u * per, per, u) for u in range(U)], U)
    W = oracle.tv_ivectors(N, oracle.tv_subtract_m(N, F, m2.reshape(-1)), Tm, invvar, tett)
    llk = g.llk(X[:50000], -1e9, 1e9); ref = oracle.llk_all(o, X[:50000], -1e9, 1e9)
    print("kernel", k, "ivec rel err %.2e (max|W| %.2f)" % (np.abs(W - W_ref).max() / np.abs(W_ref).max(), np.abs(W_ref).max()), "llk max abs err %.2e rel %.2e" % (np.abs(llk-ref).max(), np.abs(llk-ref).max()/np.abs(ref).max()))
