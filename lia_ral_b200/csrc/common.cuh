// common.cuh -- engine context, error plumbing and small device helpers shared by every
// translation unit of liblia_ral_b200.so.
#pragma once

#include <cuda_runtime.h>
#include <cublas_v2.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <string>
#include <vector>

#include "../../include/lia_ral_b200.h"

namespace lr {

struct Engine {
  bool ready = false;
  int device = -1;
  int sm_count = 0;
  cudaStream_t stream = nullptr;   // compute
  cudaStream_t copy_stream = nullptr;  // H2D staging for the host-buffer entry points
  cublasHandle_t blas = nullptr;
  void *solver = nullptr;          // cusolverDnHandle_t, created on first use (plda.cu / ivbackend.cu)
  uint64_t launches = 0;
  unsigned long long norm_seq = 0;  // normalisation identifiers of the tensor-core path (gmm_tc.cu)
  int gmm_kernel = 0;  // 0 auto, 1 simt, 2 tcgen05 (one-pass statistics), 3 tcgen05 two-pass
  int tc_debug = 0;    // profiling experiments only (LR_TC_DEBUG builds): results are WRONG when set
  int tv_gemm = 0;     // TV contractions: 0 = INT8 digit GEMM (gemm_i8.cu), 1 = cuBLAS fp64 (cross-check)
  int gmm_products = 0;  // one-pass kernel: 0 = all five fp16 products (default), 1 = statistics GEMM on the hi frame panels only,
                         // 2 = likelihood GEMM too (lr_set_gmm_products)
  int tv_planes = 6;   // digit planes per operand of the INT8 digit GEMM
  // per-device "cudaFuncSetAttribute done" flags (reset by lr_shutdown: the attribute is per context)
  enum { kAttrTc = 0, kAttrSimtLse, kAttrSimtAcc, kAttrTopk, kAttrTvDiag, kAttrTvGemm, kAttrPlda, kAttrGemmI8, kAttrCount };
  bool attr_set[kAttrCount] = {};
  // grow-only device scratch slots reused across calls (freed by lr_shutdown)
  static constexpr int kScratchSlots = 14;
  void *scratch[kScratchSlots] = {};
  size_t scratch_cap[kScratchSlots] = {};
  cudaEvent_t ev_copied[2] = {}, ev_consumed[2] = {};
  // optional per-launch timing (lr_profile): event pairs per kernel kind
  bool profile = false;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> prof_events[2];
};

// RAII bracket used by the two GMM passes: records a start/stop event pair when profiling
struct ProfileScope {
  int kind;
  cudaEvent_t a = nullptr, b = nullptr;
  explicit ProfileScope(int k);
  ~ProfileScope();
};

enum ScratchSlot {
  kSlotX0 = 0, kSlotX1, kSlotLse, kSlotIndex, kSlotChunks, kSlotS, kSlotStats, kSlotLlk,
  kSlotIdx, kSlotRest, kSlotTmpA, kSlotTmpB, kSlotSpans, kSlotXchg
};
// returns nullptr (and sets the error) on allocation failure
void *scratch_get(int slot, size_t bytes);

Engine &engine();
void set_error(const char *fmt, ...);
lr_status fail(lr_status code, const char *fmt, ...);
bool ensure_ready();

inline void count_launch(int n = 1) { engine().launches += (uint64_t)n; }

#define LR_CUDA(expr)                                                                      \
  do {                                                                                     \
    cudaError_t e__ = (expr);                                                              \
    if (e__ != cudaSuccess)                                                                \
      return lr::fail(LR_ERR_CUDA, "%s:%d %s: %s", __FILE__, __LINE__, #expr,              \
                      cudaGetErrorString(e__));                                            \
  } while (0)

#define LR_CUBLAS(expr)                                                                    \
  do {                                                                                     \
    cublasStatus_t s__ = (expr);                                                           \
    if (s__ != CUBLAS_STATUS_SUCCESS)                                                      \
      return lr::fail(LR_ERR_CUDA, "%s:%d %s: cublas status %d", __FILE__, __LINE__,       \
                      #expr, (int)s__);                                                    \
  } while (0)

#define LR_CHECK_LAUNCH()                                                                  \
  do {                                                                                     \
    lr::count_launch();                                                                    \
    LR_CUDA(cudaGetLastError());                                                           \
  } while (0)

#define LR_REQUIRE(cond, ...)                                                              \
  do {                                                                                     \
    if (!(cond)) return lr::fail(LR_ERR_ARG, __VA_ARGS__);                                 \
  } while (0)

#define LR_READY()                                                                         \
  do {                                                                                     \
    if (!lr::ensure_ready()) return LR_ERR_CUDA;                                           \
  } while (0)

// Per-call device scratch comes from a small caching pool (engine.cu): cudaMalloc / cudaFree cost
// 0.1-1 ms each and cudaFree synchronises the device, which dominated calls such as PLDA scoring
// (~25 buffers per call).  Blocks go back to the pool on release and to the driver at lr_shutdown
// (or when the pool holds more than 8 GB).  Everything the library enqueues runs on the engine's
// streams and every entry point that uses the copy stream joins it before returning, so reuse of a
// released block is stream-ordered.
cudaError_t pool_alloc(void **p, size_t bytes);
void pool_free(void *p);
void pool_release_all();

// RAII device buffer (typed), returned to the pool on scope exit; used for per-call scratch.
template <typename T>
struct DevBuf {
  T *p = nullptr;
  size_t n = 0;
  DevBuf() = default;
  DevBuf(const DevBuf &) = delete;
  DevBuf &operator=(const DevBuf &) = delete;
  ~DevBuf() { release(); }
  cudaError_t alloc(size_t count) {
    release();
    n = count;
    if (count == 0) return cudaSuccess;
    return pool_alloc(reinterpret_cast<void **>(&p), count * sizeof(T));
  }
  void release() {
    if (p) pool_free(p);
    p = nullptr;
    n = 0;
  }
};

inline int ceil_div(long a, long b) { return (int)((a + b - 1) / b); }

}  // namespace lr
