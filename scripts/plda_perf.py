"""configs[4] per-GPU shard of the PLDA trial matrix: (1 M / 8) models x 10 k segments, d = 400, rank 200,
device-resident operands, fp32 device scores.  Prints trials/s and the achieved score-write bandwidth."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from lia_ral_b200 import capi, synth
capi.init(0)
NM, NT, d, r = int(os.environ.get("NM", 125000)), int(os.environ.get("NT", 10000)), 400, 200
F, G, Sigma, _, _, _ = synth.make_plda(d=d, rF=r, rG=0, n_models=4, n_test=4, seed=6)
g = torch.Generator(device="cuda"); g.manual_seed(1)
models = torch.randn((d, NM), device="cuda", dtype=torch.float64, generator=g)
segs = torch.randn((d, NT), device="cuda", dtype=torch.float64, generator=g)
out = torch.empty((NM, NT), dtype=torch.float32, device="cuda")
model_of = np.arange(NM, dtype=np.int32)
torch.cuda.synchronize()
for it in range(3):
    t0 = time.perf_counter()
    capi.plda_native_scoring_dev(F, G, Sigma, models.data_ptr(), NM, model_of, segs.data_ptr(), NT, out.data_ptr())
    capi.synchronize()
    dt = time.perf_counter() - t0
print(json.dumps({"models": NM, "segments": NT, "seconds": dt, "trials_per_s": NM * NT / dt,
                  "score_write_GBs": NM * NT * 4 / dt / 1e9, "algorithmic_tflops": 2.0 * NM * NT * r / dt / 1e12,
                  "finite": bool(torch.isfinite(out).all().item())}))
