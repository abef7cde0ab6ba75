"""Measurements of the widened rows through the C ABI on one B200, each with
the CPU restatement of the reference loop timed beside it on a bounded sample (oracle -O3 -ffast-math):
approximate i-vector extraction (ubmWeight / eigenDecomposition), the PldaDev statistics and
normalisation matrices, cosine / Mahalanobis / two-covariance scoring, one PLDA EM iteration.
Writes one JSON object per line; summarised in profiles/r01_extra.md."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lia_ral_b200 import capi, synth
from oracle.ffi import Oracle

capi.init(0)
orc = Oracle(fast=True)
CORES = os.cpu_count() or 1


def timed(fn, reps=1):
    capi.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        r = fn()
    capi.synchronize()
    return (time.perf_counter() - t0) / reps, r


def emit(**kv):
    print(json.dumps(kv), flush=True)


# ---- f1: approximate i-vector extraction, 2048c/60d, R = 400
C, D, R, U = 2048, 60, 400, 1024
w, mean, cov = synth.make_ubm(C, D, seed=1)
invvar = (1.0 / cov).reshape(-1)
N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=3000, active=64, seed=5)
T0 = synth.make_T(R, C, D, invvar, seed=4, scale=0.02)
tv = capi.TV(C, D, R, U, mean.reshape(-1), invvar)
tv.set_stats(N, F)
tv.set_T(T0)
tv.norm_t()
t_wc, Wcov = timed(lambda: tv.weighted_cov(w))
t_wc, Wcov = timed(lambda: tv.weighted_cov(w))
tv.norm_statistics()
timed(lambda: tv.estimate_w_ubm_weight(Wcov))
t_ubm, _ = timed(lambda: tv.estimate_w_ubm_weight(Wcov), reps=3)
t_eig, (Q, lam) = timed(lambda: capi.eigen_problem(Wcov))
t_eig, (Q, lam) = timed(lambda: capi.eigen_problem(Wcov))
t_tc, Dm = timed(lambda: tv.approximate_tctc(Q))
t_tc, Dm = timed(lambda: tv.approximate_tctc(Q))
timed(lambda: tv.estimate_w_eigen_decomposition(Dm, Q))
t_ed, _ = timed(lambda: tv.estimate_w_eigen_decomposition(Dm, Q), reps=3)
us = 8
Tn = tv.get_T()
Nn, Fn = tv.get_stats()
t0 = time.perf_counter(); orc.tv_ivectors_ubm_weight(Nn[:us], Fn[:us], Tn, Wcov); c_ubm = time.perf_counter() - t0
t0 = time.perf_counter(); orc.tv_ivectors_eigen(Nn[:us], Fn[:us], Tn, Dm, Q); c_ed = time.perf_counter() - t0
emit(config="f1 approximate i-vector extraction 2048c/60d R=400", utterances=U,
     weighted_cov_s=t_wc, eigen_problem_s=t_eig, approximate_tctc_s=t_tc,
     ubm_weight_ivectors_per_s=U / t_ubm, eigen_decomposition_ivectors_per_s=U / t_ed,
     cpu_baseline={"kind": "port", "cores": 1, "sample": f"{us} utterances (the reference has no threaded variant of these estimators)",
                   "ubm_weight_ivectors_per_s": us / c_ubm, "eigen_decomposition_ivectors_per_s": us / c_ed})
del tv, N, F, Nn, Fn, Tn

# ---- f3: development-set statistics, d = 400, 5 000 speakers x 10 sessions
d, n_spk, per = 400, 5000, 10
rng = np.random.default_rng(7)
cls = np.repeat(np.arange(n_spk), per).astype(np.int32)
n = len(cls)
data = (rng.standard_normal((d, n_spk)) * 1.2)[:, cls] + rng.standard_normal((d, n))
capi.iv_cov_mat(data[:, :1000], cls[:1000], 100)
t_cov, (mu, sm, S, W, B) = timed(lambda: capi.iv_cov_mat(data, cls, n_spk))
capi.iv_wccn_chol(data[:, :5000], cls[:5000], 500)
t_wccn, _ = timed(lambda: capi.iv_wccn_chol(data, cls, n_spk))
capi.iv_efr_matrix(S)
t_efr, E = timed(lambda: capi.iv_efr_matrix(S))
capi.iv_lda(W, B, 200)
t_lda, L = timed(lambda: capi.iv_lda(W, B, 200))
t_norm, _ = timed(lambda: capi.iv_normalize(data, mu=mu, M=E, length_norm=True))
ns = 2000
t0 = time.perf_counter(); orc.iv_cov_mat(data[:, :ns], cls[:ns], ns // per); c_cov = time.perf_counter() - t0
emit(config="f3 PldaDev statistics d=400, 50 000 sessions / 5 000 speakers", cov_mat_s=t_cov, wccn_chol_s=t_wccn,
     efr_matrix_s=t_efr, lda_rank200_s=t_lda, efr_apply_s=t_norm, sessions_per_s_cov_mat=n / t_cov,
     cpu_baseline={"kind": "port", "cores": 1, "sample": f"{ns} sessions (computeCovMatUnThreaded loop)",
                   "sessions_per_s_cov_mat": ns / c_cov})

# ---- f3: scorings, d = 400, 20 000 models x 10 000 tests
nm, nt = int(os.environ.get("NM", 20000)), 10000
models, segments = rng.standard_normal((d, nm)), rng.standard_normal((d, nt))
Mah = np.linalg.inv(W)
capi.iv_cosine_scoring(models[:, :64], segments[:, :64])
t_cos, _ = timed(lambda: capi.iv_cosine_scoring(models, segments))
capi.iv_mahalanobis_scoring(models[:, :64], segments[:, :64], Mah)
t_mah, _ = timed(lambda: capi.iv_mahalanobis_scoring(models, segments, Mah))
capi.iv_two_cov_scoring(models[:, :64], segments[:, :64], W, B)
t_2c, _ = timed(lambda: capi.iv_two_cov_scoring(models, segments, W, B))
cm, ct_ = 32, 128
t0 = time.perf_counter(); orc.iv_cosine(models[:, :cm], segments[:, :ct_]); c_cos = time.perf_counter() - t0
t0 = time.perf_counter(); orc.iv_mahalanobis(models[:, :cm], segments[:, :ct_], Mah); c_mah = time.perf_counter() - t0
t0 = time.perf_counter(); orc.iv_two_cov(models[:, :cm], segments[:, :ct_], W, B); c_2c = time.perf_counter() - t0
emit(config="f3 IvTest scorings d=400", models=nm, tests=nt,
     cosine_trials_per_s=nm * nt / t_cos, mahalanobis_trials_per_s=nm * nt / t_mah, two_cov_trials_per_s=nm * nt / t_2c,
     note="host-buffer calls: H2D of the vectors and D2H of the fp64 score matrix included",
     cpu_baseline={"kind": "port", "cores": 1, "sample": f"{cm} models x {ct_} tests",
                   "cosine_trials_per_s": cm * ct_ / c_cos, "mahalanobis_trials_per_s": cm * ct_ / c_mah,
                   "two_cov_trials_per_s": cm * ct_ / c_2c, "two_cov_note": "includes the four d x d inversions of the model"})
del models, segments

# ---- f3: one PLDA EM iteration, d = 400, rankF = 200, rankG = 0
rF = 200
F0 = rng.standard_normal((d, rF))
X0 = data - mu[:, None]
capi.plda_em_iteration(X0[:, :2000], cls[:2000], 200, F0, None, S, np.zeros(d))
t_em, _ = timed(lambda: capi.plda_em_iteration(X0, cls, n_spk, F0, None, S, np.zeros(d)))
ns = 1000
t0 = time.perf_counter(); orc.plda_em_iteration(X0[:, :ns], cls[:ns], ns // per, F0, None, S, np.zeros(d)); c_em = time.perf_counter() - t0
emit(config="f3 PLDA EM iteration d=400 rankF=200 rankG=0, 50 000 sessions / 5 000 speakers", seconds=t_em,
     sessions_per_s=n / t_em,
     cpu_baseline={"kind": "port", "cores": 1, "sample": f"{ns} sessions / {ns // per} speakers (getExpectedValuesUnThreaded + mStep)",
                   "sessions_per_s": ns / c_em})
