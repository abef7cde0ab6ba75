// plda.cu -- PLDA native scoring (PldaTest::pldaNativeScoring + pldaScoring,
// LIA_SpkTools/src/PldaTools.cpp:4489-4519, 4175-4271; PldaModel::preComputation :2950-2972;
// rotateLeft :3770-3790).
//
// The reference recomputes t^T K_1 t and rebuilds (t + m) for every (model, segment) pair --
// O(Nm Nt r^2).  Algebraically
//   score(m, t) = t^T K' m  +  1/2 t^T (K' - K_1) t  +  1/2 m^T (K' - K_L) m  +  const_L
// with K' = K_{L+1}, so the trial matrix is ONE GEMM  [Nt x r] [r x Nm]  plus a rank-1 style
// epilogue; per-L quantities are recomputed only when the session count changes, like the
// reference (:4221-4250).  The small r x r / d x d algebra is fp64 (cuBLAS / cuSOLVER); the trial
// matrix itself -- the 2 Nm Nt r flop and the Nm Nt output write that dominate -- is the repo's own
// tcgen05 split-precision kernel (gemm_split.cu) with the row / column terms fused into its epilogue.
#include <cusolverDn.h>

#include <algorithm>
#include <cmath>

#include "common.cuh"
#include "gemm_split.cuh"

#define LR_CUSOLVER(expr)                                                                  \
  do {                                                                                     \
    cusolverStatus_t s__ = (expr);                                                         \
    if (s__ != CUSOLVER_STATUS_SUCCESS)                                                    \
      return lr::fail(LR_ERR_CUDA, "%s:%d %s: cusolver status %d", __FILE__, __LINE__,     \
                      #expr, (int)s__);                                                    \
  } while (0)

namespace lr {
namespace {

__global__ void k_sym_from_lower(int n, double *__restrict__ M) {
  // column-major lower triangle -> full symmetric
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n * n) {
    int col = e / n, row = e - col * n;
    if (row < col) M[e] = M[(size_t)row * n + col];
  }
}

__global__ void k_axpy_identity(int n, double a, const double *X, double *Y) {
  // Y = a X + I
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n * n) Y[e] = a * X[e] + ((e / n == e % n) ? 1.0 : 0.0);
}

__global__ void k_logdiag_sum(int n, const double *__restrict__ Lf, double *__restrict__ out) {
  // out = 2 * sum_i log L_ii   (log det via Cholesky, PldaTools.cpp:4511-4516)
  __shared__ double sh[256];
  double s = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s += log(Lf[(size_t)i * n + i]);
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *out = 2.0 * sh[0];
}

// q[j] = 1/2 v_j^T M v_j for the columns v_j (length r, contiguous) of V[r x n] column-major;
// one warp per column.
__global__ void k_half_quad(int r, long n, const double *__restrict__ V,
                            const double *__restrict__ MV, double *__restrict__ q) {
  long j = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (j >= n) return;
  int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int i = lane; i < r; i += 32) s += V[(size_t)j * r + i] * MV[(size_t)j * r + i];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if (lane == 0) q[j] = 0.5 * s;
}

// model m = sum of its enrolment columns (pldaScoring :4205-4218); pm column-major [r x n_enrol]
__global__ void k_model_sums(int r, const double *__restrict__ pm, const long *__restrict__ first,
                             long first_offset, const int *__restrict__ count, long n_models,
                             double *__restrict__ M) {
  long m = blockIdx.x;
  if (m >= n_models) return;
  for (int i = threadIdx.x; i < r; i += blockDim.x) {
    double s = 0.0;
    for (int k = 0; k < count[m]; k++) s += pm[(size_t)(first[m] - first_offset + k) * r + i];
    M[(size_t)m * r + i] = s;
  }
}

__global__ void k_add_const(long n, double c, double *__restrict__ v) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) v[i] += c;
}

struct Dense {
  cusolverDnHandle_t solver = nullptr;  // the engine's handle (created once: ~10 ms)
  DevBuf<double> work;
  DevBuf<int> info;
  lr_status init() {
    Engine &e = engine();
    if (!e.solver) {
      cusolverDnHandle_t h = nullptr;
      LR_CUSOLVER(cusolverDnCreate(&h));
      LR_CUSOLVER(cusolverDnSetStream(h, e.stream));
      e.solver = h;
    }
    solver = (cusolverDnHandle_t)e.solver;
    LR_CUDA(info.alloc(1));
    return LR_OK;
  }
  // in-place Cholesky (column-major lower); returns LR_ERR_NUMERIC when not SPD
  lr_status potrf(int n, double *A, const char *what) {
    Engine &e = engine();
    int lwork = 0;
    LR_CUSOLVER(cusolverDnDpotrf_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork));
    if ((size_t)lwork > work.n) LR_CUDA(work.alloc(lwork));
    LR_CUSOLVER(cusolverDnDpotrf(solver, CUBLAS_FILL_MODE_LOWER, n, A, n, work.p, lwork, info.p));
    count_launch();
    int h = 0;
    LR_CUDA(cudaMemcpyAsync(&h, info.p, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
    LR_CUDA(cudaStreamSynchronize(e.stream));
    if (h != 0) return fail(LR_ERR_NUMERIC, "%s is not positive definite (potrf info %d)", what, h);
    return LR_OK;
  }
  // A <- A^-1 for SPD A (full symmetric result); optionally log det A^-1 = -2 sum log L_ii
  lr_status spd_inverse(int n, double *A, const char *what, double *d_neg_logdet = nullptr) {
    Engine &e = engine();
    lr_status st = potrf(n, A, what);
    if (st != LR_OK) return st;
    if (d_neg_logdet) {
      k_logdiag_sum<<<1, 256, 0, e.stream>>>(n, A, d_neg_logdet);
      LR_CHECK_LAUNCH();
    }
    int lwork = 0;
    LR_CUSOLVER(cusolverDnDpotri_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork));
    if ((size_t)lwork > work.n) LR_CUDA(work.alloc(lwork));
    LR_CUSOLVER(cusolverDnDpotri(solver, CUBLAS_FILL_MODE_LOWER, n, A, n, work.p, lwork, info.p));
    count_launch();
    k_sym_from_lower<<<ceil_div((long)n * n, 256), 256, 0, e.stream>>>(n, A);
    LR_CHECK_LAUNCH();
    return LR_OK;
  }
};

// C[m x n] (row-major) = op(A) op(B), row-major operands; thin wrapper over column-major cuBLAS
lr_status gemm_rm(bool ta, bool tb, int m, int n, int k, const double *A, int lda, const double *B,
                  int ldb, double *C, int ldc, double beta = 0.0) {
  const double one = 1.0;
  // row-major C = A B  <=>  column-major C^T = B^T A^T
  LR_CUBLAS(cublasDgemm(engine().blas, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N,
                        n, m, k, &one, B, ldb, A, lda, &beta, C, ldc));
  count_launch();
  return LR_OK;
}

}  // namespace
}  // namespace lr

using namespace lr;

// models / segments: [d x n] row-major (the reference's _models / _segments), in host memory
// (dev_in = false) or device memory (dev_in = true).  Output: scores_host (fp64, host, [n_models x
// n_test]) or d_scores_f32 (fp32, device, leading dimension ld_scores) -- exactly one is non-null.
static lr_status plda_score_impl(int d, int rF, int rG, const double *F, const double *G,
                                 const double *Sigma, const double *models, bool dev_in, size_t n_enrol,
                                 const int32_t *model_of, size_t n_models, const double *segments,
                                 size_t n_test, double *scores_host, float *d_scores_f32,
                                 size_t ld_scores) {
  LR_READY();
  LR_REQUIRE(d > 0 && rF > 0 && rF <= d && rG >= 0 && F && Sigma && models && model_of &&
                 segments && (scores_host || d_scores_f32) && n_enrol > 0 && n_models > 0 && n_test > 0,
             "lr_plda_native_scoring: bad arguments");
  LR_REQUIRE(rG == 0 || G, "lr_plda_native_scoring: G is null but rG = %d", rG);
  LR_REQUIRE(rF <= 256, "lr_plda_native_scoring: rank %d above the 256 the scoring kernel holds in TMEM", rF);
  const cudaMemcpyKind in_kind = dev_in ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  Engine &e = engine();
  const int r = rF;
  // model -> (first enrolment column, count); consecutive columns of a model are adjacent
  std::vector<long> first;
  std::vector<int> count;
  for (size_t s = 0; s < n_enrol;) {
    size_t b = s;
    while (s < n_enrol && model_of[s] == model_of[b]) s++;
    first.push_back((long)b);
    count.push_back((int)(s - b));
  }
  LR_REQUIRE(first.size() == n_models, "lr_plda_native_scoring: model_of describes %zu models, not %zu",
             first.size(), n_models);

  Dense dn;
  lr_status st = dn.init();
  if (st != LR_OK) return st;
  DevBuf<double> dF, dG, dIS, dFtw, dFTJ, dPhi, dPm, dPs, dM, dK1, dKL, dKL1, dTmp, dA, dB, dS, dScal;
  DevBuf<double> dGtw, dGG, dFtwG, dT1;
  DevBuf<long> dFirst;
  DevBuf<int> dCount;
  LR_CUDA(dF.alloc((size_t)d * r));
  LR_CUDA(dIS.alloc((size_t)d * d));
  LR_CUDA(dFtw.alloc((size_t)r * d));
  LR_CUDA(dFTJ.alloc((size_t)r * d));
  LR_CUDA(dPhi.alloc((size_t)r * r));
  LR_CUDA(dK1.alloc((size_t)r * r));
  LR_CUDA(dKL.alloc((size_t)r * r));
  LR_CUDA(dKL1.alloc((size_t)r * r));
  LR_CUDA(dScal.alloc(4));
  LR_CUDA(cudaMemcpyAsync(dF.p, F, (size_t)d * r * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(dIS.p, Sigma, (size_t)d * d * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  // ---- preComputation (:2950-2972): Lambda = Sigma^-1 ; FTJ = F^T L - F^T L G (G^T L G + I)^-1 G^T L
  st = dn.spd_inverse(d, dIS.p, "PLDA Sigma");
  if (st != LR_OK) return st;
  if ((st = gemm_rm(true, false, r, d, d, dF.p, r, dIS.p, d, dFtw.p, d)) != LR_OK) return st;
  LR_CUDA(cudaMemcpyAsync(dFTJ.p, dFtw.p, (size_t)r * d * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  if (rG > 0) {
    LR_CUDA(dG.alloc((size_t)d * rG));
    LR_CUDA(dGtw.alloc((size_t)rG * d));
    LR_CUDA(dGG.alloc((size_t)rG * rG));
    LR_CUDA(dFtwG.alloc((size_t)r * rG));
    LR_CUDA(dT1.alloc((size_t)r * rG));
    LR_CUDA(cudaMemcpyAsync(dG.p, G, (size_t)d * rG * sizeof(double), cudaMemcpyHostToDevice, e.stream));
    if ((st = gemm_rm(true, false, rG, d, d, dG.p, rG, dIS.p, d, dGtw.p, d)) != LR_OK) return st;
    if ((st = gemm_rm(false, false, rG, rG, d, dGtw.p, d, dG.p, rG, dGG.p, rG)) != LR_OK) return st;
    k_axpy_identity<<<ceil_div((long)rG * rG, 256), 256, 0, e.stream>>>(rG, 1.0, dGG.p, dGG.p);
    LR_CHECK_LAUNCH();
    if ((st = dn.spd_inverse(rG, dGG.p, "PLDA (G^T Sigma^-1 G + I)")) != LR_OK) return st;
    if ((st = gemm_rm(false, false, r, rG, d, dFtw.p, d, dG.p, rG, dFtwG.p, rG)) != LR_OK) return st;
    if ((st = gemm_rm(false, false, r, rG, rG, dFtwG.p, rG, dGG.p, rG, dT1.p, rG)) != LR_OK) return st;
    // FTJ -= T1 Gtw : beta = 1 with a negated product -> scale T1 by -1 first
    const double mone = -1.0;
    LR_CUBLAS(cublasDscal(e.blas, r * rG, &mone, dT1.p, 1));
    count_launch();
    if ((st = gemm_rm(false, false, r, d, rG, dT1.p, rG, dGtw.p, d, dFTJ.p, d, 1.0)) != LR_OK) return st;
  }
  // Phi = FTJ F (:4497)
  if ((st = gemm_rm(false, false, r, r, d, dFTJ.p, d, dF.p, r, dPhi.p, r)) != LR_OK) return st;
  // ---- rotateLeft (:3770-3790): project; store projected vectors contiguously (row per vector)
  LR_CUDA(dPs.alloc((size_t)n_test * r));
  {
    DevBuf<double> dSeg;
    LR_CUDA(dSeg.alloc((size_t)d * n_test));
    LR_CUDA(cudaMemcpyAsync(dSeg.p, segments, (size_t)d * n_test * sizeof(double), in_kind, e.stream));
    // Ps[n_test x r] = segments^T[n_test x d] FTJ^T[d x r]
    if ((st = gemm_rm(true, true, (int)n_test, r, d, dSeg.p, (int)n_test, dFTJ.p, d, dPs.p, r)) != LR_OK) return st;
    LR_CUDA(cudaStreamSynchronize(e.stream));
  }
  // K_1 = (Phi + I)^-1, alpha_1 = log det K_1 (:4507-4516)
  k_axpy_identity<<<ceil_div((long)r * r, 256), 256, 0, e.stream>>>(r, 1.0, dPhi.p, dK1.p);
  LR_CHECK_LAUNCH();
  if ((st = dn.spd_inverse(r, dK1.p, "PLDA (Phi + I)", dScal.p)) != LR_OK) return st;
  double h_ld[3];
  LR_CUDA(cudaMemcpyAsync(&h_ld[0], dScal.p, sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  const double alpha1 = -h_ld[0];  // dScal held log det (Phi + I)

  // ---- pldaScoring (:4186-4271).  Segment operand of the trial GEMM (fp16 hi / lo panels) once;
  // models in blocks: u = M K' is the model operand, row term b[m] + const, column term a[t].
  DevBuf<unsigned char> dBp, dAp;
  DevBuf<double> dRow, dOut[2];
  double scaleB = 1.0, scaleA = 1.0;
  LR_CUDA(dBp.alloc(gemm_split_panel_bytes((long)n_test, r)));
  if ((st = gemm_split_prepare(dPs.p, r, (long)n_test, r, dBp.p, dScal.p + 3, &scaleB)) != LR_OK) return st;
  // host output: two device score blocks of <= 1 GB alternate with their D2H copies; device output: the
  // block only bounds the temporaries
  const size_t mblk = scores_host ? std::max<size_t>(128, std::min<size_t>((n_models + 127) / 128 * 128,
                                                                          (((size_t)1 << 27) / n_test + 127) / 128 * 128))
                                  : std::min<size_t>((n_models + 127) / 128 * 128, 32768);
  LR_CUDA(dM.alloc(mblk * r));
  LR_CUDA(dTmp.alloc(std::max(mblk, n_test) * r));
  LR_CUDA(dA.alloc(n_test));
  LR_CUDA(dB.alloc(mblk));
  LR_CUDA(dAp.alloc(gemm_split_panel_bytes((long)mblk, r)));
  if (scores_host) {
    LR_CUDA(dOut[0].alloc(mblk * n_test));
    LR_CUDA(dOut[1].alloc(mblk * n_test));
  }
  LR_CUDA(dFirst.alloc(n_models));
  LR_CUDA(dCount.alloc(n_models));
  LR_CUDA(cudaMemcpyAsync(dFirst.p, first.data(), n_models * sizeof(long), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(dCount.p, count.data(), n_models * sizeof(int), cudaMemcpyHostToDevice, e.stream));
  DevBuf<double> dEnr, dDiff;
  LR_CUDA(dDiff.alloc((size_t)r * r));
  cudaEvent_t ev_done[2] = {nullptr, nullptr}, ev_copied[2] = {nullptr, nullptr};
  struct EvGuard {
    cudaEvent_t *a, *b;
    ~EvGuard() {
      for (int i = 0; i < 2; i++) {
        if (a[i]) cudaEventDestroy(a[i]);
        if (b[i]) cudaEventDestroy(b[i]);
      }
    }
  } guard{ev_done, ev_copied};
  if (scores_host)
    for (int i = 0; i < 2; i++) {
      LR_CUDA(cudaEventCreateWithFlags(&ev_done[i], cudaEventDisableTiming));
      LR_CUDA(cudaEventCreateWithFlags(&ev_copied[i], cudaEventDisableTiming));
    }
  size_t m0 = 0;
  int cur_nb = -1, blk = 0;
  double constant = 0.0;
  size_t pend_m0 = 0, pend_nm = 0;
  int pend_buf = -1;
  // the D2H copy of a finished block runs on the copy stream while the next block is computed
  auto drain = [&]() -> lr_status {
    if (pend_buf < 0) return LR_OK;
    LR_CUDA(cudaStreamWaitEvent(e.copy_stream, ev_done[pend_buf], 0));
    LR_CUDA(cudaMemcpyAsync(scores_host + pend_m0 * n_test, dOut[pend_buf].p, pend_nm * n_test * sizeof(double),
                            cudaMemcpyDeviceToHost, e.copy_stream));
    LR_CUDA(cudaEventRecord(ev_copied[pend_buf], e.copy_stream));
    pend_buf = -1;
    return LR_OK;
  };
  while (m0 < n_models) {
    // a run of models with the same session count, capped at the block size
    size_t m1 = m0;
    while (m1 < n_models && count[m1] == count[m0] && m1 - m0 < mblk) m1++;
    const size_t nm = m1 - m0;
    const int nb = count[m0];
    if (nb != cur_nb) {
      cur_nb = nb;
      // K_L = (L Phi + I)^-1, K_{L+1}; alpha's via Cholesky
      k_axpy_identity<<<ceil_div((long)r * r, 256), 256, 0, e.stream>>>(r, (double)nb, dPhi.p, dKL.p);
      LR_CHECK_LAUNCH();
      if ((st = dn.spd_inverse(r, dKL.p, "PLDA (L Phi + I)", dScal.p + 1)) != LR_OK) return st;
      k_axpy_identity<<<ceil_div((long)r * r, 256), 256, 0, e.stream>>>(r, (double)nb + 1.0, dPhi.p, dKL1.p);
      LR_CHECK_LAUNCH();
      if ((st = dn.spd_inverse(r, dKL1.p, "PLDA ((L+1) Phi + I)", dScal.p + 2)) != LR_OK) return st;
      LR_CUDA(cudaMemcpyAsync(&h_ld[1], dScal.p + 1, 2 * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
      LR_CUDA(cudaStreamSynchronize(e.stream));
      constant = ((-h_ld[2]) - (-h_ld[1]) - alpha1) / 2.0;
      // a[t] = 1/2 t^T (K' - K_1) t : Tmp = Ps (K' - K_1)
      const double mone = -1.0;
      LR_CUDA(cudaMemcpyAsync(dDiff.p, dKL1.p, (size_t)r * r * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
      LR_CUBLAS(cublasDaxpy(e.blas, r * r, &mone, dK1.p, 1, dDiff.p, 1));
      count_launch();
      if ((st = gemm_rm(false, false, (int)n_test, r, r, dPs.p, r, dDiff.p, r, dTmp.p, r)) != LR_OK) return st;
      k_half_quad<<<ceil_div((long)n_test, 8), 256, 0, e.stream>>>(r, (long)n_test, dPs.p, dTmp.p, dA.p);
      LR_CHECK_LAUNCH();
    }
    // project this block's enrolment vectors and sum them per model
    const long e0 = first[m0];
    const long e1 = (m1 < n_models) ? first[m1] : (long)n_enrol;
    const size_t ne = (size_t)(e1 - e0);
    {
      // Pm[ne x r] = (models[:, e0 .. e1))^T FTJ^T.  Device input: the column block is read in place
      // (leading dimension n_enrol); host input: gathered through a strided copy first.
      if (dPm.n < ne * r) LR_CUDA(dPm.alloc(ne * r));
      if (dev_in) {
        if ((st = gemm_rm(true, true, (int)ne, r, d, models + e0, (int)n_enrol, dFTJ.p, d, dPm.p, r)) != LR_OK)
          return st;
      } else {
        if (dEnr.n < (size_t)d * ne) LR_CUDA(dEnr.alloc((size_t)d * ne));
        LR_CUDA(cudaMemcpy2DAsync(dEnr.p, ne * sizeof(double), models + e0, n_enrol * sizeof(double),
                                  ne * sizeof(double), d, in_kind, e.stream));
        if ((st = gemm_rm(true, true, (int)ne, r, d, dEnr.p, (int)ne, dFTJ.p, d, dPm.p, r)) != LR_OK) return st;
      }
    }
    k_model_sums<<<(unsigned)nm, 128, 0, e.stream>>>(r, dPm.p, dFirst.p + m0, e0, dCount.p + m0, (long)nm, dM.p);
    LR_CHECK_LAUNCH();
    // row term b[m] + const = 1/2 m^T (K' - K_L) m + const
    {
      const double mone = -1.0;
      LR_CUDA(cudaMemcpyAsync(dDiff.p, dKL1.p, (size_t)r * r * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
      LR_CUBLAS(cublasDaxpy(e.blas, r * r, &mone, dKL.p, 1, dDiff.p, 1));
      count_launch();
      if ((st = gemm_rm(false, false, (int)nm, r, r, dM.p, r, dDiff.p, r, dTmp.p, r)) != LR_OK) return st;
      k_half_quad<<<ceil_div((long)nm, 8), 256, 0, e.stream>>>(r, (long)nm, dM.p, dTmp.p, dB.p);
      LR_CHECK_LAUNCH();
      k_add_const<<<ceil_div((long)nm, 256), 256, 0, e.stream>>>((long)nm, constant, dB.p);
      LR_CHECK_LAUNCH();
    }
    // cross term: scores[nm x n_test] = (M K') Ps^T + row + column terms, in the tcgen05 kernel
    if ((st = gemm_rm(false, false, (int)nm, r, r, dM.p, r, dKL1.p, r, dTmp.p, r)) != LR_OK) return st;
    if ((st = gemm_split_prepare(dTmp.p, r, (long)nm, r, dAp.p, dScal.p + 3, &scaleA)) != LR_OK) return st;
    if (scores_host) {
      const int buf = blk & 1;
      if (blk >= 2) LR_CUDA(cudaStreamWaitEvent(e.stream, ev_copied[buf], 0));  // its previous copy is out
      if ((st = gemm_split_run<double>(dAp.p, scaleA, (long)nm, dBp.p, scaleB, (long)n_test, r, dOut[buf].p,
                                       n_test, dB.p, dA.p)) != LR_OK)
        return st;
      LR_CUDA(cudaEventRecord(ev_done[buf], e.stream));
      // the D2H copy of the PREVIOUS block is issued now: a copy to pageable host memory holds the
      // host thread, and this block's kernels are already queued behind it on the compute stream
      if ((st = drain()) != LR_OK) return st;
      pend_buf = buf;
      pend_m0 = m0;
      pend_nm = nm;
    } else {
      if ((st = gemm_split_run<float>(dAp.p, scaleA, (long)nm, dBp.p, scaleB, (long)n_test, r,
                                      d_scores_f32 + m0 * ld_scores, ld_scores, dB.p, dA.p)) != LR_OK)
        return st;
    }
    blk++;
    m0 = m1;
  }
  if (scores_host && (st = drain()) != LR_OK) return st;
  LR_CUDA(cudaStreamSynchronize(e.stream));
  LR_CUDA(cudaStreamSynchronize(e.copy_stream));
  return LR_OK;
}

extern "C" lr_status lr_plda_native_scoring(int d, int rF, int rG, const double *F, const double *G,
                                            const double *Sigma, const double *models,
                                            size_t n_enrol, const int32_t *model_of,
                                            size_t n_models, const double *segments, size_t n_test,
                                            double *scores) {
  return plda_score_impl(d, rF, rG, F, G, Sigma, models, false, n_enrol, model_of, n_models, segments, n_test,
                         scores, nullptr, 0);
}

extern "C" lr_status lr_plda_native_scoring_dev(int d, int rF, int rG, const double *F, const double *G,
                                                const double *Sigma, const double *d_models,
                                                size_t n_enrol, const int32_t *model_of, size_t n_models,
                                                const double *d_segments, size_t n_test, float *d_scores,
                                                size_t ld_scores) {
  LR_REQUIRE(d_scores && ld_scores >= n_test, "lr_plda_native_scoring_dev: bad output");
  return plda_score_impl(d, rF, rG, F, G, Sigma, d_models, true, n_enrol, model_of, n_models, d_segments,
                         n_test, nullptr, d_scores, ld_scores);
}
