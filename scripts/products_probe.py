"""Accuracy of the one-pass kernel per product level (lr_set_gmm_products) against the fp64 oracle:
the full-size blurred-model BW statistics case of tests/test_gmm_gpu.py + one EM accumulation on the
generator's own model (occupations, second moments, mean log-likelihood)."""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from lia_ral_b200 import capi, synth  # noqa: E402
from oracle.ffi import Oracle  # noqa: E402  (checker only)

capi.init(0)
orc = Oracle()
thr = os.cpu_count() or 1
C, D, U, per = 2048, 60, 10, 20000
w, mean, cov = synth.make_ubm(C, D, seed=1)
X = synth.make_frames(w, mean, cov, U * per, seed=21)
w2, m2, c2 = w, mean * 0.15, cov * 5.0
o = orc.gmm(w2, m2, c2)
f2r = (np.arange(U * per) // per).astype(np.int32)
N_ref, F_ref = orc.bwstats(o, X, f2r, U, threads=thr)
llk_r, n_r, occ_r, m1_r, m2_r = orc.em_accumulate(o, X, threads=thr)
R = 40
invvar = (1.0 / c2).reshape(-1)
Tm = synth.make_T(R, C, D, invvar, seed=23, scale=0.05)
tett = orc.tv_tett(Tm, invvar, C, D, threads=thr)
W_ref = orc.tv_ivectors(N_ref, orc.tv_subtract_m(N_ref, F_ref, m2.reshape(-1)), Tm, invvar, tett)
capi.set_gmm_kernel(2)
for level in (0, 1, 2):
    capi.set_gmm_products(level)
    g = capi.GMM(w2, m2, c2)
    N, F = g.bwstats(X, [(u * per, per, u) for u in range(U)], U)
    rel = np.abs(N - N_ref) / np.maximum(N_ref, 1e-300)
    big = N_ref >= 100.0
    mid = (N_ref >= 1.0) & ~big
    F3, F3r = F.reshape(U, C, D), F_ref.reshape(U, C, D)
    relF = np.linalg.norm(F3 - F3r, axis=2) / np.maximum(np.linalg.norm(F3r, axis=2), 1e-300)
    W = orc.tv_ivectors(N, orc.tv_subtract_m(N, F, m2.reshape(-1)), Tm, invvar, tett)
    llk, n, occ, m1, m2_ = g.em_accumulate(X)
    # EM quantities on the components that matter (occupation >= 100)
    keep = occ_r >= 100.0
    mean_g, mean_r = m1[keep] / occ[keep, None], m1_r[keep] / occ_r[keep, None]
    var_g = m2_[keep] / occ[keep, None] - mean_g ** 2
    var_r = m2_r[keep] / occ_r[keep, None] - mean_r ** 2
    print(json.dumps({
        "level": level,
        "N_rel_big": float(rel[big].max()), "N_rel_mid_sqrtocc": float((rel[mid] * np.sqrt(N_ref[mid])).max()),
        "F_rel_big": float(relF[big].max()), "F_rel_mid_sqrtocc": float((relF[mid] * np.sqrt(N_ref[mid])).max()),
        "ivec_rel": float(np.abs(W - W_ref).max() / np.abs(W_ref).max()),
        "llk_mean_rel": float(abs(llk / n - llk_r / n_r) / abs(llk_r / n_r)),
        "occ_rel": float((np.abs(occ - occ_r)[keep] / occ_r[keep]).max()),
        "mean_err_in_sigma": float((np.abs(mean_g - mean_r) / np.sqrt(var_r)).max()),
        "var_rel": float((np.abs(var_g - var_r) / var_r).max()),
        "components_checked": int(keep.sum())}), flush=True)
capi.set_gmm_products(0)
capi.set_gmm_kernel(0)
