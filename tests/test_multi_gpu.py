"""Two-GPU checks (skipped on a single-GPU box; run with `gpurun --gpus 2`): utterance-sharded
TotalVariability E-step with ONE all-reduce of the accumulator block -- or the component-sharded
exchange (reduce-scatter A, M-step on C / world components, all-gather T) -- equals the single-GPU
run, and every rank ends the iteration with the same T."""
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r"""
import os, sys
sys.path.insert(0, os.environ["LR_ROOT"])
import numpy as np, torch, torch.distributed as dist
from lia_ral_b200 import capi, synth, dist as lrd
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(rank)
dist.init_process_group("nccl", device_id=torch.device("cuda", rank))
capi.init(rank)
C, D, R, U = 64, 12, 10, 23
w, mean, cov = synth.make_ubm(C, D, seed=21)
invvar = (1.0 / cov).reshape(-1)
N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=300, active=16, seed=22)
T = synth.make_T(R, C, D, invvar, seed=23, scale=0.05)
b, e = lrd.shard_range(U, rank, world)
tv = capi.TV(C, D, R, e - b, mean.reshape(-1), invvar)
tv.set_stats(N[b:e], F[b:e]); tv.set_T(T); tv.reset_tmp_acc(); tv.subtract_m(); tv.estimate_tett()
tv.estimate_a_and_c()
if os.environ.get("LR_SHARD") == "1":
    # component-sharded exchange + M-step: only this rank's components of A are fully reduced
    lrd.tv_sharded_mstep(tv, e - b)
    A, Cmx, Rm, r, mw = tv.get_acc()
    cw = C // world
    A[:rank * cw] = 0; A[(rank + 1) * cw:] = 0
    At = torch.tensor(A, device="cuda"); dist.all_reduce(At); A = At.cpu().numpy()
else:
    lrd.tv_allreduce_estep(tv, e - b)
    A, Cmx, Rm, r, mw = tv.get_acc()   # before minDivergence rewrites Rm / r in place
    tv.update_t()
tv.min_divergence(float(U))
Tall = torch.tensor(tv.get_T(), device="cuda")
Tmax = Tall.clone(); dist.all_reduce(Tmax, op=dist.ReduceOp.MAX)
Tmin = Tall.clone(); dist.all_reduce(Tmin, op=dist.ReduceOp.MIN)
if rank == 0:
    np.savez(os.environ["LR_OUT"], A=A, Cmx=Cmx, Rm=Rm, r=r, mw=mw, T=tv.get_T(),
             rank_spread=float(((Tmax - Tmin).abs().max() / Tall.abs().max()).item()))
dist.barrier(); dist.destroy_process_group()
"""


@pytest.mark.parametrize("shard", ["0", "1"], ids=["allreduce", "sharded-mstep"])
def test_tv_estep_two_gpus(tmp_path, oracle, shard):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    from lia_ral_b200 import synth
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, LR_ROOT=ROOT, LR_OUT=str(tmp_path / "out.npz"), LR_SHARD=shard)
    subprocess.check_call([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                           "--master-addr", "127.0.0.1", "--master-port", "29533", str(script)], env=env, timeout=600)
    z = np.load(tmp_path / "out.npz")
    C, D, R, U = 64, 12, 10, 23
    w, mean, cov = synth.make_ubm(C, D, seed=21)
    invvar = (1.0 / cov).reshape(-1)
    N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=300, active=16, seed=22)
    T = synth.make_T(R, C, D, invvar, seed=23, scale=0.05)
    Fc = oracle.tv_subtract_m(N, F, mean.reshape(-1))
    tett = oracle.tv_tett(T, invvar, C, D)
    _, A, Cmx, Rm, r, mw = oracle.tv_estep(N, Fc, T, invvar, tett)
    rel = lambda a, b: np.abs(a - b).max() / np.abs(b).max()
    assert rel(z["A"], A) < 1e-9 and rel(z["Cmx"], Cmx) < 1e-9 and rel(z["Rm"], Rm) < 1e-9
    assert rel(z["r"], r) < 1e-9 and rel(z["mw"], mw) < 1e-9
    T1 = oracle.tv_mstep(A, Cmx, C, D)
    _, T2 = oracle.tv_mindiv(Rm, r, mw, mean.reshape(-1), T1, float(U), C, D)
    assert rel(z["T"], T2) < 1e-7
    assert float(z["rank_spread"]) < 1e-12   # every rank ends the iteration with the same T
