"""One-process-per-GPU plumbing for the two EM loops that exchange statistics.

The data path shards embarrassingly (frames for TrainWorld, utterances for TotalVariability /
IvExtractor); the ONLY exchange is one all-reduce of sufficient statistics per EM iteration -- the
analogue of `emAcc.addAccEM` merging per-thread accumulators (AccumulateStat.cpp:286-292) and of
the mutex-guarded A / C updates (AccumulateTVStat.cpp:1920-1937) -- or, for TotalVariability, its
component-sharded form (reduce-scatter A, M-step on C / world components, all-gather T).  `torch.distributed` is plumbing:
NCCL on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    return (dist.get_rank(), dist.get_world_size()) if dist.is_initialized() else (0, 1)


def shard_range(n, rank, world_size):
    """Contiguous, balanced [begin, end) of n units for `rank` -- the split the reference uses for
    NDX lines and speakers (AccumulateTVStat.cpp:498-507, :1989-2023)."""
    per, rem = divmod(n, world_size)
    begin = rank * per + min(rank, rem)
    return begin, begin + per + (1 if rank < rem else 0)


def shard_utterances_by_frames(frame_counts, world_size):
    """Contiguous utterance ranges balanced by frame count (BW statistics cost ~ frames).
    Returns world_size + 1 cut points."""
    total = float(sum(frame_counts))
    cuts, acc, nxt = [0], 0.0, 1
    for i, f in enumerate(frame_counts):
        acc += f
        while nxt < world_size and acc >= total * nxt / world_size:
            cuts.append(i + 1)
            nxt += 1
    while len(cuts) < world_size + 1:
        cuts.append(len(frame_counts))
    cuts[-1] = len(frame_counts)
    return cuts


def allreduce_stats(stats, stream=None):
    """In-place SUM all-reduce of a statistics tensor ([occ | m1 | m2 | llk | n] or the TV block
    [A | Cmx | R | r | sumW]).  `stream`: torch stream the producer kernels were enqueued on."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return stats
    if stream is not None:
        with torch.cuda.stream(stream):
            dist.all_reduce(stats)
    else:
        dist.all_reduce(stats)
    return stats


def tv_allreduce_estep(tv, n_speakers_local, stream=None):
    """After lr_tv_estimate_a_and_c on this rank's utterance shard: one all-reduce of the
    contiguous accumulator block, then meanW = sumW / total speakers (lr_tv_finish_estep)."""
    n_total = torch.tensor([float(n_speakers_local)], dtype=torch.float64)
    if dist.is_initialized() and dist.get_world_size() > 1:
        import ctypes as ct
        n = tv.acc_len()
        # wrap the library's device block as a tensor without copying
        from torch.utils.dlpack import from_dlpack  # noqa: F401  (kept for clarity of intent)
        acc = _device_tensor(tv.dev_acc(), n)
        allreduce_stats(acc, stream)
        n_total = n_total.cuda()
        dist.all_reduce(n_total)
        torch.cuda.synchronize()
    tv.finish_estep(float(n_total.item()))
    return float(n_total.item())


def tv_sharded_mstep(tv, n_speakers_local, stream=None):
    """Component-sharded exchange + M-step of one TotalVariability iteration: after
    lr_tv_estimate_a_and_c on this rank's utterances,
      reduce-scatter A by component (each rank receives the sum of ITS C / world components),
      all-reduce [Cmx | R | r | sumW], meanW = sumW / total speakers,
      updateTestimate on the rank's components only (it is independent per component, :981-1000),
      all-gather the new columns of T.
    minDivergence (replicated, O(R^2 C D)) follows on every rank.  Needs C % world == 0; falls back to
    the single all-reduce + replicated M-step otherwise.  Returns the total speaker count."""
    rank, ws = world()
    C = tv.C
    if ws == 1 or C % ws != 0:
        n_total = tv_allreduce_estep(tv, n_speakers_local, stream)
        tv.update_t()
        return n_total
    ctx = torch.cuda.stream(stream) if stream is not None else _NullCtx()
    stride, cw = tv.acc_a_stride(), C // ws
    acc = _device_tensor(tv.dev_acc(), tv.acc_len())
    a_part, rest = acc[:C * stride], acc[C * stride:]
    with ctx:
        mine = torch.empty(cw * stride, dtype=torch.float64, device="cuda")
        dist.reduce_scatter_tensor(mine, a_part)
        a_part[rank * cw * stride:(rank + 1) * cw * stride].copy_(mine)
        dist.all_reduce(rest)
        n_total = torch.tensor([float(n_speakers_local)], dtype=torch.float64, device="cuda")
        dist.all_reduce(n_total)
    torch.cuda.synchronize()
    tv.finish_estep(float(n_total.item()))
    tv.update_t_range(rank * cw, (rank + 1) * cw)
    blk = tv.R * cw * tv.D
    with ctx:
        send = torch.empty(blk, dtype=torch.float64, device="cuda")
        recv = torch.empty(ws * blk, dtype=torch.float64, device="cuda")
    torch.cuda.synchronize()
    from . import capi
    tv.pack_t(rank * cw, (rank + 1) * cw, send.data_ptr())
    capi.synchronize()   # the library enqueues on its own stream
    with ctx:
        dist.all_gather_into_tensor(recv, send)
    torch.cuda.synchronize()
    for r in range(ws):
        if r != rank:
            tv.unpack_t(r * cw, (r + 1) * cw, recv.data_ptr() + r * blk * 8)
    capi.synchronize()
    return float(n_total.item())


class _NullCtx:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


def _device_tensor(ptr, n_doubles):
    """float64 CUDA tensor aliasing device memory owned by the C library (no copy)."""
    class _Holder:
        pass
    h = _Holder()
    h.__cuda_array_interface__ = {"shape": (int(n_doubles),), "typestr": "<f8",
                                  "data": (int(ptr), False), "version": 3}
    return torch.as_tensor(h, device="cuda")
