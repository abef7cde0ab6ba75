// engine.cu -- process-wide engine state: device binding, streams, cuBLAS handle, errors.
#include "common.cuh"

#include <cusolverDn.h>

#include <map>
#include <unordered_map>

namespace lr {

static thread_local std::string g_error;

Engine &engine() {
  static Engine e;
  return e;
}

void set_error(const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
}

lr_status fail(lr_status code, const char *fmt, ...) {
  char buf[1024];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_error = buf;
  return code;
}

void *scratch_get(int slot, size_t bytes) {
  Engine &e = engine();
  if (bytes == 0) bytes = 16;
  if (e.scratch_cap[slot] >= bytes) return e.scratch[slot];
  if (e.scratch[slot]) {
    cudaStreamSynchronize(e.stream);
    cudaStreamSynchronize(e.copy_stream);
    cudaFree(e.scratch[slot]);
    e.scratch[slot] = nullptr;
    e.scratch_cap[slot] = 0;
  }
  size_t cap = bytes + bytes / 8;
  if (cudaMalloc(&e.scratch[slot], cap) != cudaSuccess) {
    cudaGetLastError();
    if (cudaMalloc(&e.scratch[slot], bytes) != cudaSuccess) {
      cudaGetLastError();
      set_error("device scratch allocation of %zu bytes failed (slot %d)", bytes, slot);
      e.scratch[slot] = nullptr;
      return nullptr;
    }
    cap = bytes;
  }
  e.scratch_cap[slot] = cap;
  return e.scratch[slot];
}

ProfileScope::ProfileScope(int k) : kind(k) {
  Engine &e = engine();
  if (!e.profile) return;
  if (cudaEventCreate(&a) != cudaSuccess || cudaEventCreate(&b) != cudaSuccess) {
    a = b = nullptr;
    return;
  }
  cudaEventRecord(a, e.stream);
}
ProfileScope::~ProfileScope() {
  if (!a) return;
  Engine &e = engine();
  cudaEventRecord(b, e.stream);
  e.prof_events[kind].push_back({a, b});
}

static void profile_clear() {
  Engine &e = engine();
  for (auto &v : e.prof_events) {
    for (auto &p : v) {
      cudaEventDestroy(p.first);
      cudaEventDestroy(p.second);
    }
    v.clear();
  }
}

bool ensure_ready() {
  Engine &e = engine();
  if (e.ready) return true;
  // lazily bind to the current device (device 0 unless lr_init chose another)
  return lr_init(e.device < 0 ? 0 : e.device) == LR_OK;
}

}  // namespace lr

using namespace lr;

// ---- caching pool behind DevBuf
namespace lr {
namespace {
struct Pool {
  std::multimap<size_t, void *> free_blocks;
  std::unordered_map<void *, size_t> live;
  size_t pooled_bytes = 0;
};
Pool &pool() {
  static Pool p;
  return p;
}
constexpr size_t kPoolCap = (size_t)8 << 30;
}  // namespace

cudaError_t pool_alloc(void **p, size_t bytes) {
  Pool &pl = pool();
  const size_t want = bytes <= (1u << 20) ? (bytes + 255) / 256 * 256 : (bytes + (1u << 20) - 1) >> 20 << 20;
  auto it = pl.free_blocks.lower_bound(want);
  if (it != pl.free_blocks.end() && it->first <= 2 * want + (1u << 20)) {
    *p = it->second;
    pl.live[*p] = it->first;
    pl.pooled_bytes -= it->first;
    pl.free_blocks.erase(it);
    return cudaSuccess;
  }
  cudaError_t err = cudaMalloc(p, want);
  if (err != cudaSuccess) {  // give the pooled blocks back to the driver and try once more
    cudaGetLastError();
    pool_release_all();
    err = cudaMalloc(p, want);
  }
  if (err == cudaSuccess) pl.live[*p] = want;
  return err;
}

void pool_free(void *p) {
  Pool &pl = pool();
  auto it = pl.live.find(p);
  if (it == pl.live.end()) {
    cudaFree(p);
    return;
  }
  const size_t sz = it->second;
  pl.live.erase(it);
  if (pl.pooled_bytes + sz > kPoolCap) {
    cudaFree(p);
    return;
  }
  pl.free_blocks.emplace(sz, p);
  pl.pooled_bytes += sz;
}

void pool_release_all() {
  Pool &pl = pool();
  for (auto &kv : pl.free_blocks) cudaFree(kv.second);
  pl.free_blocks.clear();
  pl.pooled_bytes = 0;
}
}  // namespace lr


extern "C" {

const char *lr_last_error(void) { return g_error.c_str(); }
const char *lr_version(void) { return "lia_ral_b200 0.1 (sm_100a)"; }

lr_status lr_init(int device) {
  Engine &e = engine();
  int n = 0;
  cudaError_t err = cudaGetDeviceCount(&n);
  if (err != cudaSuccess || n == 0)
    return fail(LR_ERR_CUDA, "no CUDA device (%s): this engine has no CPU fallback",
                err == cudaSuccess ? "device count 0" : cudaGetErrorString(err));
  LR_REQUIRE(device >= 0 && device < n, "device %d out of range (have %d)", device, n);
  if (e.ready && e.device == device) return LR_OK;
  if (e.ready) lr_shutdown();
  LR_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  LR_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(LR_ERR_CUDA, "device %d is sm_%d%d; this library is built for sm_100a only",
                device, prop.major, prop.minor);
  e.device = device;
  e.sm_count = prop.multiProcessorCount;
  LR_CUDA(cudaStreamCreateWithFlags(&e.stream, cudaStreamNonBlocking));
  LR_CUDA(cudaStreamCreateWithFlags(&e.copy_stream, cudaStreamNonBlocking));
  for (int i = 0; i < 2; i++) {
    LR_CUDA(cudaEventCreateWithFlags(&e.ev_copied[i], cudaEventDisableTiming));
    LR_CUDA(cudaEventCreateWithFlags(&e.ev_consumed[i], cudaEventDisableTiming));
  }
  LR_CUBLAS(cublasCreate(&e.blas));
  LR_CUBLAS(cublasSetStream(e.blas, e.stream));
  e.ready = true;
  return LR_OK;
}

lr_status lr_shutdown(void) {
  Engine &e = engine();
  if (!e.ready) return LR_OK;
  cudaStreamSynchronize(e.stream);
  if (e.copy_stream) cudaStreamSynchronize(e.copy_stream);
  profile_clear();
  e.profile = false;
  for (bool &b : e.attr_set) b = false;
  pool_release_all();
  for (int i = 0; i < Engine::kScratchSlots; i++) {
    if (e.scratch[i]) cudaFree(e.scratch[i]);
    e.scratch[i] = nullptr;
    e.scratch_cap[i] = 0;
  }
  for (int i = 0; i < 2; i++) {
    if (e.ev_copied[i]) cudaEventDestroy(e.ev_copied[i]);
    if (e.ev_consumed[i]) cudaEventDestroy(e.ev_consumed[i]);
    e.ev_copied[i] = e.ev_consumed[i] = nullptr;
  }
  if (e.solver) cusolverDnDestroy((cusolverDnHandle_t)e.solver);
  e.solver = nullptr;
  if (e.blas) cublasDestroy(e.blas);
  if (e.stream) cudaStreamDestroy(e.stream);
  if (e.copy_stream) cudaStreamDestroy(e.copy_stream);
  e.blas = nullptr;
  e.stream = e.copy_stream = nullptr;
  e.ready = false;
  return LR_OK;
}

lr_status lr_synchronize(void) {
  LR_READY();
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

uint64_t lr_stream_handle(void) {
  if (!ensure_ready()) return 0;
  return (uint64_t)(uintptr_t)engine().stream;
}

int lr_sm_count(void) { return ensure_ready() ? engine().sm_count : 0; }
uint64_t lr_launch_count(void) { return engine().launches; }
void lr_reset_launch_count(void) { engine().launches = 0; }

lr_status lr_profile(int enable) {
  LR_READY();
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  profile_clear();
  engine().profile = enable != 0;
  return LR_OK;
}

lr_status lr_profile_read(int kind, double *total_ms, uint64_t *n_launches) {
  LR_READY();
  LR_REQUIRE(kind >= 0 && kind < 2 && total_ms && n_launches, "lr_profile_read: bad argument");
  Engine &e = engine();
  LR_CUDA(cudaStreamSynchronize(e.stream));
  double tot = 0.0;
  for (auto &p : e.prof_events[kind]) {
    float ms = 0.f;
    LR_CUDA(cudaEventElapsedTime(&ms, p.first, p.second));
    tot += ms;
  }
  *total_ms = tot;
  *n_launches = e.prof_events[kind].size();
  return LR_OK;
}

lr_status lr_set_gmm_kernel(int which) {
  LR_REQUIRE(which >= 0 && which <= 3,
             "kernel selector must be 0 (auto), 1 (simt), 2 (tcgen05) or 3 (tcgen05, two-pass statistics)");
  engine().gmm_kernel = which;
  return LR_OK;
}
int lr_get_gmm_kernel(void) { return engine().gmm_kernel; }
lr_status lr_set_gmm_products(int level) {
  LR_REQUIRE(level >= 0 && level <= 2, "product level must be 0 (five fp16 products), 1 (four) or 2 (three)");
  engine().gmm_products = level;
  return LR_OK;
}
int lr_get_gmm_products(void) { return engine().gmm_products; }
#ifdef LR_DEBUG_BUILD
// profiling experiments (results are WRONG while set): only in `make DEBUG=1` builds, not part of the product ABI
void lr_debug_flags(int flags) { engine().tc_debug = flags; }
#endif

}  // extern "C"
