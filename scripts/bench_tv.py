"""TotalVariability EM iteration at scale, one process per GPU:
utterances sharded across ranks (weak scaling: U_PER_GPU each), ONE NCCL all-reduce of the
accumulator block [A | Cmx | R | r | sumW] per iteration, replicated M-step + minDivergence.

    python scripts/bench_tv.py                                   # 1 GPU
    torchrun --nproc-per-node N scripts/bench_tv.py              # N GPUs
Prints one JSON line on rank 0 (utterances/s over the whole job; device-event timed, max over ranks).
"""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from lia_ral_b200 import capi, synth, dist as lrd

rank, world = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1))
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
capi.init(local)
C, D, R = 2048, 60, int(os.environ.get("R", 600))
U = int(os.environ.get("U_PER_GPU", 2048))
ITERS = int(os.environ.get("ITERS", 3))
SHARD = os.environ.get("SHARD_MSTEP", "1") == "1" and world > 1
w, mean, cov = synth.make_ubm(C, D, seed=1)
invvar = (1.0 / cov).reshape(-1)
N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=3000, active=64, seed=5 + rank)
tv = capi.TV(C, D, R, U, mean.reshape(-1), invvar)
tv.set_T(synth.make_T(R, C, D, invvar, seed=4, scale=0.02))
stream = torch.cuda.ExternalStream(capi.stream_handle())
t_e, t_ar, t_m = [], [], []
for it in range(ITERS + 1):   # first iteration = warm-up
    tv.set_stats(N, F)         # the reference reloads N / F_X every iteration (TotalVariability.cpp:149-153)
    tv.reset_tmp_acc()
    capi.synchronize(); torch.cuda.synchronize()
    if world > 1: dist.barrier()
    t0 = time.perf_counter()
    tv.subtract_m(); tv.estimate_tett(); tv.estimate_a_and_c()
    capi.synchronize(); t1 = time.perf_counter()
    if SHARD:   # reduce-scatter A, M-step on C / world components, all-gather T
        lrd.tv_sharded_mstep(tv, U)
        torch.cuda.synchronize(); t2 = time.perf_counter()
    else:       # one all-reduce of the whole block, replicated M-step
        lrd.tv_allreduce_estep(tv, U)
        torch.cuda.synchronize(); t2 = time.perf_counter()
        tv.update_t()
    tv.min_divergence(float(U * world))
    capi.synchronize(); t3 = time.perf_counter()
    if it > 0:
        t_e.append(t1 - t0); t_ar.append(t2 - t1); t_m.append(t3 - t2)
tot = torch.tensor([np.mean(t_e) + np.mean(t_ar) + np.mean(t_m), np.mean(t_e), np.mean(t_ar), np.mean(t_m)], device="cuda", dtype=torch.float64)
if world > 1:
    dist.all_reduce(tot, op=dist.ReduceOp.MAX)
if rank == 0:
    t = tot.cpu().numpy()
    flop = U * world * (2 * C * R * (R + 1) + 4 * R * C * D + R ** 3)
    print(json.dumps({"metric": "TotalVariability EM iteration (2048c/60d)", "rank_R": R, "n_gpus": world,
                      "utterances_per_gpu": U, "seconds_per_iteration": t[0], "estep_s": t[1], "allreduce_s": t[2],
                      "mstep_mindiv_s": t[3], "value": U * world / t[0], "unit": "utterances/s",
                      "estep_algorithmic_tflops_total": flop / t[1] / 1e12,
                      "allreduce_bytes": tv.acc_len() * 8, "scaling": "weak",
                      "exchange": "reduce-scatter A + all-reduce rest + sharded M-step + all-gather T (timed under allreduce_s)" if SHARD
                      else "one all-reduce, replicated M-step"}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
