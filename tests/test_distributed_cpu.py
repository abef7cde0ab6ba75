"""N > 1 host logic on CPU: world_size-2 gloo processes shard the units, compute their share with
the oracle (checker only) and exchange statistics with ONE all-reduce; the result must equal the
single-process statistics.  Covers the sharding helpers bench.py and the TV EM loop use."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world_size, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world_size)
    from lia_ral_b200 import dist as lrd, synth
    from oracle.ffi import Oracle
    orc = Oracle()
    # ---- TrainWorld: frames shard, one all-reduce of {occ, m1, m2, llk, n}
    C, D, T = 24, 10, 1501
    w, mean, cov = synth.make_ubm(C, D, seed=41)
    X = synth.make_frames(w, mean, cov, T, seed=42)
    g = orc.gmm(*synth.perturb_ubm(w, mean, cov, seed=43, frac=1.0, scale=0.4))
    b, e = lrd.shard_range(T, rank, world_size)
    llk, n, occ, m1, m2 = orc.em_accumulate(g, np.ascontiguousarray(X[b:e]))
    stats = torch.from_numpy(np.concatenate([occ, m1.ravel(), m2.ravel(), [llk, n]]))
    lrd.allreduce_stats(stats)
    # ---- TotalVariability E-step: utterance shard, one all-reduce of [A | Cmx | R | r | sumW]
    R, U = 6, 11
    N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=200, active=8, seed=44)
    invvar = (1.0 / cov).reshape(-1)
    Tm = synth.make_T(R, C, D, invvar, seed=45, scale=0.05)
    Fc = orc.tv_subtract_m(N, F, mean.reshape(-1))
    tett = orc.tv_tett(Tm, invvar, C, D)
    cuts = lrd.shard_utterances_by_frames([200] * U, world_size)
    ub, ue = cuts[rank], cuts[rank + 1]
    W, A, Cmx, Rm, r, mw = orc.tv_estep(N[ub:ue], Fc[ub:ue], Tm, invvar, tett)
    acc = torch.from_numpy(np.concatenate([A.ravel(), Cmx.ravel(), Rm.ravel(), r, mw * (ue - ub)]))
    lrd.allreduce_stats(acc)
    if rank == 0:
        np.savez(os.path.join(out_dir, "out.npz"), stats=stats.numpy(), acc=acc.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_statistics_allreduce(tmp_path, oracle):
    from lia_ral_b200 import synth
    port = 29500 + os.getpid() % 2000
    mp.spawn(_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    z = np.load(tmp_path / "out.npz")
    C, D, T, R, U = 24, 10, 1501, 6, 11
    w, mean, cov = synth.make_ubm(C, D, seed=41)
    X = synth.make_frames(w, mean, cov, T, seed=42)
    g = oracle.gmm(*synth.perturb_ubm(w, mean, cov, seed=43, frac=1.0, scale=0.4))
    llk, n, occ, m1, m2 = oracle.em_accumulate(g, X)
    ref = np.concatenate([occ, m1.ravel(), m2.ravel(), [llk, n]])
    assert np.allclose(z["stats"], ref, rtol=1e-12, atol=1e-10)
    N, F = synth.make_bw_stats(U, w, mean, cov, frames_per_utt=200, active=8, seed=44)
    invvar = (1.0 / cov).reshape(-1)
    Tm = synth.make_T(R, C, D, invvar, seed=45, scale=0.05)
    Fc = oracle.tv_subtract_m(N, F, mean.reshape(-1))
    tett = oracle.tv_tett(Tm, invvar, C, D)
    W, A, Cmx, Rm, r, mw = oracle.tv_estep(N, Fc, Tm, invvar, tett)
    ref = np.concatenate([A.ravel(), Cmx.ravel(), Rm.ravel(), r, mw * U])
    assert np.allclose(z["acc"], ref, rtol=1e-11, atol=1e-9)


def test_shard_helpers():
    from lia_ral_b200 import dist as lrd
    for n, wsize in ((10, 3), (7, 8), (1_000_003, 8), (0, 2)):
        spans = [lrd.shard_range(n, r, wsize) for r in range(wsize)]
        assert spans[0][0] == 0 and spans[-1][1] == n
        assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))
        assert max(e - b for b, e in spans) - min(e - b for b, e in spans) <= 1
    cuts = lrd.shard_utterances_by_frames([100, 3000, 100, 100, 2800, 50], 2)
    assert cuts[0] == 0 and cuts[-1] == 6 and len(cuts) == 3
    loads = [sum([100, 3000, 100, 100, 2800, 50][cuts[i]:cuts[i + 1]]) for i in range(2)]
    assert abs(loads[0] - loads[1]) <= 3000
