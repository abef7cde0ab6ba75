"""Per-kernel breakdown helper for the TV rows: runs estimate_w (R=400) and estimate_a_and_c +
update_t (R=600) on synthesised statistics; meant to be run under
`ncu --metrics gpu__time_duration.sum --clock-control none` (launch list) or bare (wall times)."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from lia_ral_b200 import capi, synth

capi.init(0)
C, D = 2048, 60
w, mean, cov = synth.make_ubm(C, D, seed=1)
invvar = (1.0 / cov).reshape(-1)


def timed(fn):
    capi.synchronize()
    t0 = time.perf_counter()
    fn()
    capi.synchronize()
    return time.perf_counter() - t0


for R, U, what in ((400, int(os.environ.get("U3", 512)), "w"), (600, int(os.environ.get("U4", 256)), "e")):
    N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=3000, active=64, seed=5)
    tv = capi.TV(C, D, R, U, mean.reshape(-1), invvar)
    tv.set_stats(N, F)
    tv.set_T(synth.make_T(R, C, D, invvar, seed=4, scale=0.02))
    tv.reset_tmp_acc()
    tv.subtract_m()
    t_tett = timed(tv.estimate_tett)
    if what == "w":
        timed(tv.estimate_w)
        t = timed(tv.estimate_w)
        print(json.dumps({"R": R, "U": U, "tett_s": t_tett, "estimate_w_s": t, "ivectors_per_s": U / t}), flush=True)
    else:
        timed(tv.estimate_a_and_c)
        tv.reset_tmp_acc()
        t = timed(tv.estimate_a_and_c)
        t_m0 = timed(tv.update_t)
        t_m = timed(tv.update_t)
        print(json.dumps({"R": R, "U": U, "tett_s": t_tett, "estep_s": t, "utt_per_s": U / t,
                          "mstep_first_call_s": t_m0, "mstep_s": t_m}), flush=True)
    del tv
