// comm.cu -- the collective side of the C ABI: one process per GPU, NCCL over NVLink / NVSwitch.
//
// The path has exactly one exchange step per EM iteration: the sum of the sufficient statistics of
// the ranks' shards -- the analogue of emAcc.addAccEM merging the per-thread accumulators
// (LIA_SpkTools/src/AccumulateStat.cpp:286-292) and of the mutex-guarded A / C updates of the threaded
// E-step (AccumulateTVStat.cpp:1920-1937).  TrainWorld all-reduces {occ, m1, m2, llk, n};
// TotalVariability exchanges component-sharded: reduce-scatter of A by component, all-reduce of
// [Cmx | R | r | sumW], updateTestimate (:974-1005, independent per component) on C / world components,
// all-gather of the new T columns.  BW statistics / ComputeTest / PLDA scoring shard with no
// collective; lr_allgather only assembles their outputs.
//
// NCCL is bound at run time (dlopen of libnccl.so.2): inside a torch process that is the NCCL torch
// already loaded, in the C++ host programs the system library; the CUDA library itself carries no
// link-time dependency on it.
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>

#include <chrono>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <thread>

#include "common.cuh"

namespace lr {
namespace {

struct Nccl {
  void *so = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*AllReduce)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*ReduceScatter)(const void *, void *, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

Nccl &nc() {
  static Nccl n;
  return n;
}

lr_status load_nccl() {
  Nccl &n = nc();
  if (n.so) return LR_OK;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  for (const char *nm : names) {
    n.so = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
    if (n.so) break;
  }
  if (!n.so) return fail(LR_ERR_CUDA, "cannot load libnccl.so.2: %s", dlerror());
#define LR_SYM(field, name)                                                   \
  do {                                                                        \
    *(void **)(&n.field) = dlsym(n.so, name);                                 \
    if (!n.field) return fail(LR_ERR_CUDA, "libnccl: symbol %s missing", name); \
  } while (0)
  LR_SYM(GetUniqueId, "ncclGetUniqueId");
  LR_SYM(CommInitRank, "ncclCommInitRank");
  LR_SYM(CommDestroy, "ncclCommDestroy");
  LR_SYM(GetErrorString, "ncclGetErrorString");
  LR_SYM(AllReduce, "ncclAllReduce");
  LR_SYM(ReduceScatter, "ncclReduceScatter");
  LR_SYM(AllGather, "ncclAllGather");
#undef LR_SYM
  return LR_OK;
}

#define LR_NCCL(expr)                                                                            \
  do {                                                                                           \
    ncclResult_t r__ = (expr);                                                                   \
    if (r__ != ncclSuccess)                                                                      \
      return lr::fail(LR_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr, nc().GetErrorString(r__)); \
  } while (0)

}  // namespace
}  // namespace lr

using namespace lr;

extern "C" {

lr_status lr_comm_unique_id(void *id128) {
  LR_REQUIRE(id128, "lr_comm_unique_id: null argument");
  lr_status st = load_nccl();
  if (st != LR_OK) return st;
  ncclUniqueId id;
  LR_NCCL(nc().GetUniqueId(&id));
  static_assert(sizeof(ncclUniqueId) == LR_COMM_ID_BYTES, "ncclUniqueId size");
  std::memcpy(id128, &id, sizeof(id));
  return LR_OK;
}

lr_status lr_comm_init(int rank, int world, const void *id128) {
  LR_READY();
  LR_REQUIRE(world >= 1 && rank >= 0 && rank < world, "lr_comm_init: rank %d of %d", rank, world);
  Nccl &n = nc();
  if (n.comm) return fail(LR_ERR_ARG, "lr_comm_init: a communicator already exists (lr_comm_destroy first)");
  n.rank = rank;
  n.world = world;
  if (world == 1) return LR_OK;
  LR_REQUIRE(id128, "lr_comm_init: null id");
  lr_status st = load_nccl();
  if (st != LR_OK) return st;
  ncclUniqueId id;
  std::memcpy(&id, id128, sizeof(id));
  LR_NCCL(n.CommInitRank(&n.comm, world, id, rank));
  return LR_OK;
}

// Bootstrap through a file on a filesystem every rank sees: rank 0 writes the id (atomically, via
// rename), the others wait for it.  For the C++ programs, which have no other rendezvous.
lr_status lr_comm_init_file(int rank, int world, const char *path) {
  LR_REQUIRE(path && *path, "lr_comm_init_file: empty path");
  unsigned char id[LR_COMM_ID_BYTES];
  if (world > 1) {
    if (rank == 0) {
      lr_status st = lr_comm_unique_id(id);
      if (st != LR_OK) return st;
      const std::string tmp = std::string(path) + ".tmp";
      {
        std::ofstream f(tmp, std::ios::binary | std::ios::trunc);
        f.write((const char *)id, sizeof(id));
        if (!f) return fail(LR_ERR_IO, "lr_comm_init_file: cannot write %s", tmp.c_str());
      }
      if (std::rename(tmp.c_str(), path) != 0) return fail(LR_ERR_IO, "lr_comm_init_file: cannot rename to %s", path);
    } else {
      bool got = false;
      for (int i = 0; i < 6000 && !got; i++) {  // up to 10 minutes
        std::ifstream f(path, std::ios::binary);
        if (f && f.read((char *)id, sizeof(id)) && f.gcount() == (std::streamsize)sizeof(id)) got = true;
        if (!got) std::this_thread::sleep_for(std::chrono::milliseconds(100));
      }
      if (!got) return fail(LR_ERR_IO, "lr_comm_init_file: rank %d never saw %s", rank, path);
    }
  }
  return lr_comm_init(rank, world, id);
}

int lr_comm_rank(void) { return nc().rank; }
int lr_comm_world(void) { return nc().world; }

lr_status lr_comm_destroy(void) {
  Nccl &n = nc();
  if (n.comm) {
    cudaStreamSynchronize(engine().stream);
    n.CommDestroy(n.comm);
  }
  n.comm = nullptr;
  n.rank = 0;
  n.world = 1;
  return LR_OK;
}

// In-place sum of n doubles (device) over the ranks, enqueued on the engine stream behind the kernels
// that produced them: the single exchange step of an EM iteration.
lr_status lr_allreduce_stats(double *d_buf, size_t n) {
  LR_READY();
  Nccl &c = nc();
  if (c.world == 1 || n == 0) return LR_OK;
  LR_REQUIRE(c.comm && d_buf, "lr_allreduce_stats: no communicator (lr_comm_init) or null buffer");
  LR_NCCL(c.AllReduce(d_buf, d_buf, n, ncclDouble, ncclSum, c.comm, engine().stream));
  return LR_OK;
}

// host-buffer convenience of the same step (the C++ programs keep their accumulators on the host)
lr_status lr_allreduce_host(double *buf, size_t n) {
  LR_READY();
  Nccl &c = nc();
  if (c.world == 1 || n == 0) return LR_OK;
  LR_REQUIRE(buf, "lr_allreduce_host: null buffer");
  Engine &e = engine();
  DevBuf<double> d;
  LR_CUDA(d.alloc(n));
  LR_CUDA(cudaMemcpyAsync(d.p, buf, n * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  lr_status st = lr_allreduce_stats(d.p, n);
  if (st != LR_OK) return st;
  LR_CUDA(cudaMemcpyAsync(buf, d.p, n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

// d_dst[world * n] = the ranks' d_src[n] in rank order (device)
lr_status lr_allgather(const double *d_src, size_t n, double *d_dst) {
  LR_READY();
  Nccl &c = nc();
  LR_REQUIRE(d_src && d_dst, "lr_allgather: null buffer");
  if (c.world == 1) {
    if (d_dst != d_src)
      LR_CUDA(cudaMemcpyAsync(d_dst, d_src, n * sizeof(double), cudaMemcpyDeviceToDevice, engine().stream));
    return LR_OK;
  }
  LR_REQUIRE(c.comm, "lr_allgather: no communicator (lr_comm_init)");
  LR_NCCL(c.AllGather(d_src, d_dst, n, ncclDouble, c.comm, engine().stream));
  return LR_OK;
}

lr_status lr_allgather_host(const double *src, size_t n, double *dst) {
  LR_READY();
  Nccl &c = nc();
  LR_REQUIRE(src && dst, "lr_allgather_host: null buffer");
  if (c.world == 1) {
    std::memcpy(dst, src, n * sizeof(double));
    return LR_OK;
  }
  Engine &e = engine();
  DevBuf<double> a, b;
  LR_CUDA(a.alloc(n));
  LR_CUDA(b.alloc(n * c.world));
  LR_CUDA(cudaMemcpyAsync(a.p, src, n * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  lr_status st = lr_allgather(a.p, n, b.p);
  if (st != LR_OK) return st;
  LR_CUDA(cudaMemcpyAsync(dst, b.p, n * c.world * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

// The exchange + M-step of one TotalVariability iteration after lr_tv_estimate_a_and_c on this rank's
// utterances.  C % world != 0 falls back to one all-reduce + the replicated M-step.
// *n_speakers_total = sum over ranks of n_speakers_local.
lr_status lr_tv_exchange_sharded(lr_tv *tv, double n_speakers_local, double *n_speakers_total) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_exchange_sharded: null handle");
  Nccl &c = nc();
  Engine &e = engine();
  int C = 0, D = 0, R = 0;
  lr_status st = lr_tv_dims(tv, &C, &D, &R);
  if (st != LR_OK) return st;
  double n_total = n_speakers_local;
  if (c.world == 1) {
    if ((st = lr_tv_finish_estep(tv, n_total)) != LR_OK) return st;
    if (n_speakers_total) *n_speakers_total = n_total;
    return lr_tv_update_t(tv);
  }
  LR_REQUIRE(c.comm, "lr_tv_exchange_sharded: no communicator (lr_comm_init)");
  double *acc = lr_tv_dev_acc(tv);
  const size_t len = lr_tv_acc_len(tv), stride = lr_tv_acc_a_stride(tv);
  DevBuf<double> dn;
  LR_CUDA(dn.alloc(1));
  LR_CUDA(cudaMemcpyAsync(dn.p, &n_speakers_local, sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_NCCL(c.AllReduce(dn.p, dn.p, 1, ncclDouble, ncclSum, c.comm, e.stream));
  if (C % c.world != 0) {
    LR_NCCL(c.AllReduce(acc, acc, len, ncclDouble, ncclSum, c.comm, e.stream));
    LR_CUDA(cudaMemcpyAsync(&n_total, dn.p, sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    LR_CUDA(cudaStreamSynchronize(e.stream));
    if ((st = lr_tv_finish_estep(tv, n_total)) != LR_OK) return st;
    if (n_speakers_total) *n_speakers_total = n_total;
    return lr_tv_update_t(tv);
  }
  const int cw = C / c.world, c0 = c.rank * cw;
  // reduce-scatter A by component: rank r receives the summed packed triangles of its cw components,
  // in place (the receive block is this rank's own slice of the send buffer)
  LR_NCCL(c.ReduceScatter(acc, acc + (size_t)c0 * stride, (size_t)cw * stride, ncclDouble, ncclSum, c.comm, e.stream));
  LR_NCCL(c.AllReduce(acc + (size_t)C * stride, acc + (size_t)C * stride, len - (size_t)C * stride, ncclDouble,
                      ncclSum, c.comm, e.stream));
  LR_CUDA(cudaMemcpyAsync(&n_total, dn.p, sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  if ((st = lr_tv_finish_estep(tv, n_total)) != LR_OK) return st;
  if ((st = lr_tv_update_t_range(tv, c0, c0 + cw)) != LR_OK) return st;
  // all-gather the new T columns of every rank's components
  const size_t blk = (size_t)R * cw * D;
  DevBuf<double> send, recv;
  LR_CUDA(send.alloc(blk));
  LR_CUDA(recv.alloc(blk * c.world));
  if ((st = lr_tv_pack_t(tv, c0, c0 + cw, send.p)) != LR_OK) return st;
  LR_NCCL(c.AllGather(send.p, recv.p, blk, ncclDouble, c.comm, e.stream));
  for (int r = 0; r < c.world; r++)
    if (r != c.rank && (st = lr_tv_unpack_t(tv, r * cw, (r + 1) * cw, recv.p + (size_t)r * blk)) != LR_OK) return st;
  LR_CUDA(cudaStreamSynchronize(e.stream));
  if (n_speakers_total) *n_speakers_total = n_total;
  return LR_OK;
}

}  // extern "C"
