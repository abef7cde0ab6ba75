"""Secondary measurements through the C ABI on one B200:
i-vectors/s, TotalVariability EM iteration time, PLDA trials/s, ComputeTest LLR throughput,
each with the CPU restatement of the reference loop (oracle -O3 -ffast-math build, all host
threads where the reference has a threaded variant) timed beside it on a bounded sample.
Writes one JSON object per line; results are summarised under profiles/."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from lia_ral_b200 import capi, synth

from oracle.ffi import Oracle

capi.init(0)
out = []
orc = Oracle(fast=True)
CORES = os.cpu_count() or 1
CPU = os.environ.get("NO_CPU", "") == ""


def cpu_timed(fn):
    t0 = time.perf_counter()
    r = fn()
    return time.perf_counter() - t0, r


def timed(fn, reps=1):
    capi.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    capi.synchronize()
    return (time.perf_counter() - t0) / reps


C, D = 2048, 60
w, mean, cov = synth.make_ubm(C, D, seed=1)
invvar = (1.0 / cov).reshape(-1)

# ---- cfg3: BW statistics (device-resident frames) + i-vector solve, R = 400
U, FPU, R = int(os.environ.get("U3", 1024)), 3000, 400
X = synth.make_frames(w, mean, cov, U * FPU, seed=2)
g = capi.GMM(w, mean, cov)
tv = capi.TV(C, D, R, U, mean.reshape(-1), invvar)
tv.set_T(synth.make_T(R, C, D, invvar, seed=4, scale=0.02))
feats = capi.Feats(X)
segs = [(u * FPU, FPU, u) for u in range(U)]
t_bw = timed(lambda: g.bwstats_dev(feats, segs, U, tv.dev_N(), tv.dev_F()))
t_bw = timed(lambda: g.bwstats_dev(feats, segs, U, tv.dev_N(), tv.dev_F()))
tv.set_stats(*[a / 3.0 for a in tv.get_stats()])   # three accumulations above
t_sub = timed(tv.subtract_m)
t_tett = timed(tv.estimate_tett)
t_tett = timed(tv.estimate_tett)   # steady state: the first call of every cuBLAS kernel loads its module
timed(tv.estimate_w)
t_w = timed(tv.estimate_w, reps=3)
flop_iv = U * (C * R * (R + 1) + 2 * C * D * R + R ** 3 / 3 + 2 * R * R)
cpu3 = None
if CPU:
    Nh, Fh = tv.get_stats()
    Th = tv.get_T()
    us = 32
    t_ct, tett_h = cpu_timed(lambda: orc.tv_tett(Th, invvar, C, D, threads=CORES))
    t_cw, _ = cpu_timed(lambda: orc.tv_ivectors(Nh[:us], Fh[:us], Th, invvar, tett_h, threads=CORES))
    cpu3 = {"kind": "port", "cores": CORES, "sample": f"{us} utterances (estimateW), TETt once",
            "ivectors_per_s": us / t_cw, "tett_s": t_ct}
    del Nh, Fh, Th, tett_h
out.append({"config": "cfg3 IvExtractor 2048c/60d R=400", "utterances": U, "frames_per_utt": FPU, "cpu_baseline": cpu3,
            "bwstats_frames_per_s": U * FPU / t_bw, "tett_s": t_tett, "substractM_s": t_sub,
            "ivector_solve_s": t_w, "ivectors_per_s_solve_only": U / t_w,
            "ivectors_per_s_incl_bwstats": U / (t_bw + t_sub + t_w),
            "solve_algorithmic_tflops": flop_iv / t_w / 1e12})
print(json.dumps(out[-1]), flush=True)
del feats, X, tv

# ---- cfg4: T-matrix EM iteration, R = 600, synthesised statistics
U, R = int(os.environ.get("U4", 1024)), 600
N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=3000, active=64, seed=5)
tv = capi.TV(C, D, R, U, mean.reshape(-1), invvar)
tv.set_stats(N, F)
tv.set_T(synth.make_T(R, C, D, invvar, seed=4, scale=0.02))
tv.reset_tmp_acc()
tv.subtract_m()
t_tett = timed(tv.estimate_tett)
t_tett = timed(tv.estimate_tett)
timed(tv.estimate_a_and_c)         # warm-up (Cmx accumulates across calls; timing only)
t_e = timed(tv.estimate_a_and_c)
timed(tv.update_t)
tv.estimate_tett()                 # update_t factors A inside the TETt buffer
t_m = timed(tv.update_t)
t_md = timed(lambda: tv.min_divergence(float(U)))
flop_e = U * (2 * C * R * (R + 1) + 4 * R * C * D + R ** 3)
cpu4 = None
if CPU:
    us = 16
    Th = synth.make_T(R, C, D, invvar, seed=4, scale=0.02)
    Fc = orc.tv_subtract_m(N[:us], F[:us], mean.reshape(-1))
    tett_h = orc.tv_tett(Th, invvar, C, D, threads=CORES)
    t_ce, _ = cpu_timed(lambda: orc.tv_estep(N[:us], Fc, Th, invvar, tett_h, threads=CORES))
    cpu4 = {"kind": "port", "cores": CORES, "sample": f"{us} utterances (estimateAandC)", "utterances_per_s_estep": us / t_ce}
    del Th, Fc, tett_h
out.append({"config": "cfg4 TotalVariability EM 2048c/60d R=600", "utterances": U, "tett_s": t_tett, "cpu_baseline": cpu4,
            "estep_s": t_e, "mstep_s": t_m, "mindiv_s": t_md, "utterances_per_s_estep": U / t_e,
            "estep_algorithmic_tflops": flop_e / t_e / 1e12})
print(json.dumps(out[-1]), flush=True)
del tv, N, F

# ---- cfg5: PLDA native scoring, d = 400, rank 200
nm, nt = int(os.environ.get("NM5", 20000)), 10000
Fm, G, Sigma, models, model_of, segments = synth.make_plda(d=400, rF=200, rG=0, n_models=nm, n_test=nt, seed=6)
capi.plda_native_scoring(Fm, G, Sigma, models[:, :256], model_of[:256], segments[:, :256])   # warm-up
t_p = timed(lambda: capi.plda_native_scoring(Fm, G, Sigma, models, model_of, segments))
cpu5 = None
if CPU:
    ms_, ts_ = 64, 256
    t_cp, _ = cpu_timed(lambda: orc.plda_native_scoring(Fm, G, Sigma, models[:, :ms_], model_of[:ms_], segments[:, :ts_]))
    cpu5 = {"kind": "port", "cores": 1, "sample": f"{ms_} models x {ts_} tests (pldaScoringUnThreaded loop)",
            "trials_per_s": ms_ * ts_ / t_cp}
out.append({"config": "cfg5 IvTest PLDA d=400 r=200", "models": nm, "tests": nt, "seconds": t_p, "cpu_baseline": cpu5,
            "trials_per_s": nm * nt / t_p, "note": "host-buffer call: includes H2D of i-vectors and D2H of the fp64 score matrix"})
print(json.dumps(out[-1]), flush=True)

# ---- cfg1: ComputeTest LLR, 64c/60d UBM, 100 utterances x 3000 frames, 5 clients, top-10
C1 = 64
w1, m1, c1 = synth.make_ubm(C1, D, seed=1)
world = capi.GMM(w1, m1, c1)
clients = [capi.GMM(*synth.perturb_ubm(w1, m1, c1, seed=10 + i, frac=0.1, scale=0.3)) for i in range(5)]
Xs = [synth.make_frames(w1, m1, c1, 3000, seed=100 + u) for u in range(100)]
def run():
    for Xu in Xs:
        capi.compute_test(world, clients, Xu, K=10, complete=True)
run()
t_c = timed(run)
cpu1 = None
if CPU:
    ow = orc.gmm(w1, m1, c1)
    ocl = [orc.gmm(*synth.perturb_ubm(w1, m1, c1, seed=10 + i, frac=0.1, scale=0.3)) for i in range(5)]
    def cpu_run():
        for Xu in Xs[:20]:
            _, idx, _, rest, _ = orc.llk_determine_top(ow, Xu, 10)
            for oc in ocl:
                orc.llk_use_top(oc, Xu, idx, rest)
    t_cc, _ = cpu_timed(cpu_run)
    cpu1 = {"kind": "port", "cores": 1, "sample": "20 utterances x 3000 frames x 5 clients (the reference ComputeTest frame loop is single-threaded)",
            "frames_per_s": 20 * 3000 / t_cc, "llr_per_s": 100 / t_cc}
out.append({"config": "cfg1 ComputeTest 64c/60d 100 utt x 3000 frames x 5 clients top-10", "seconds": t_c, "cpu_baseline": cpu1,
            "frames_per_s": 100 * 3000 / t_c, "llr_per_s": 500 / t_c})
print(json.dumps(out[-1]), flush=True)
