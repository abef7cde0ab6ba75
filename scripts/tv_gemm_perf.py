"""A/B of the TV contraction kernels on one B200: INT8 digit GEMM (6 / 7 planes) vs cuBLAS fp64.
i-vectors/s of estimateW (2048c/60d, R = 400, 1024 utterances) and utterances/s of estimateAandC
(R = 600, 1280 utterances); also the raw digit GEMM at the L = N TETt shape."""
import json
import sys
import time

import numpy as np
import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from lia_ral_b200 import capi, synth  # noqa: E402

C, D = 2048, 60


def estep_rate(dev, U=1280, R=600):
    w, mean, cov = synth.make_ubm(C, D, seed=1)
    invvar = (1.0 / cov).reshape(-1)
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    occ = torch.zeros((U, C), device=dev, dtype=torch.float64)
    act = torch.randint(0, C, (U, 64), device=dev, generator=g)
    occ.scatter_add_(1, act, torch.rand((U, 64), device=dev, generator=g, dtype=torch.float64))
    occ *= 3000.0 / occ.sum(1, keepdim=True)
    mu = torch.tensor(mean.reshape(-1), device=dev)
    sd = torch.tensor(np.sqrt(cov).reshape(-1), device=dev)
    Nrep = occ.repeat_interleave(D, dim=1)
    F = Nrep * mu + torch.sqrt(Nrep) * sd * torch.randn((U, C * D), device=dev, generator=g, dtype=torch.float64)
    tv = capi.TV(C, D, R, U, mean.reshape(-1), invvar)
    tv.set_stats(occ.cpu().numpy(), F.cpu().numpy())
    del F, Nrep
    tv.set_T(synth.make_T(R, C, D, invvar, seed=4, scale=0.02))
    tv.subtract_m()
    tv.estimate_tett()
    tv.estimate_a_and_c()
    capi.synchronize()
    t0 = time.perf_counter()
    tv.estimate_a_and_c()
    capi.synchronize()
    dt = time.perf_counter() - t0
    W = tv.get_W()
    tv.close()
    return {"utterances_per_s": U / dt, "ms": dt * 1e3, "U": U, "R": R}, W


def main():
    capi.init(0)
    dev = torch.device("cuda:0")
    out = {}
    Ws = {}
    for name, which, planes in (("cublas_fp64", 1, 0), ("digits6", 0, 6), ("digits7", 0, 7)):
        capi.set_tv_gemm(which, planes)
        iv = bench.ivector_rate(torch, capi, dev)
        es, W = estep_rate(dev)
        Ws[name] = W
        out[name] = {"ivectors_per_s": iv["value"], "estimate_w_ms": iv["ms"], "tett_ms": iv["tett_ms_once_per_T"],
                     "estep": es}
        print(name, json.dumps(out[name]), flush=True)
    for name in ("digits6", "digits7"):
        out[name]["estep_W_rel_vs_cublas"] = float(np.abs(Ws[name] - Ws["cublas_fp64"]).max() /
                                                   np.abs(Ws["cublas_fp64"]).max())
    # raw GEMM at the L shape: [1024 x 2048] x [80200 x 2048]^T
    rng = np.random.default_rng(0)
    A = rng.random((1024, 2048))
    B = rng.standard_normal((80200, 2048))
    for planes in (6, 7):
        capi.gemm_digits(A[:128], B[:64], planes=planes)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
