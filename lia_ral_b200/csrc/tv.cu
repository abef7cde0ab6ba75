// tv.cu -- device twin of the reference's TVAcc object (LIA_SpkTools/src/AccumulateTVStat.cpp):
// Baum-Welch statistics in HBM, i-vector posterior solve, T-matrix EM accumulators and M-step.
//
// Round-1 formulation: every triple loop of the reference is restated as a dense fp64 GEMM
// over a batch of utterances (cuBLAS on the fp64 tensor pipe), the per-utterance R x R systems
// are factorised with batched Cholesky (L = I + sum_c N_c TETt_c is SPD), and the glue
// (centring, rank-1 updates, reductions) is hand-written kernels.  Layout notes use
// "row-major X[a x b]" for the reference's Matrix<double> buffers; cuBLAS sees the same
// memory as the column-major transpose.
#include <cusolverDn.h>

#include <algorithm>

#include "common.cuh"

#define LR_CUSOLVER(expr)                                                                  \
  do {                                                                                     \
    cusolverStatus_t s__ = (expr);                                                         \
    if (s__ != CUSOLVER_STATUS_SUCCESS)                                                    \
      return lr::fail(LR_ERR_CUDA, "%s:%d %s: cusolver status %d", __FILE__, __LINE__,     \
                      #expr, (int)s__);                                                    \
  } while (0)

struct lr_tv {
  int C = 0, D = 0, R = 0;
  size_t U = 0;
  size_t sv = 0;  // C * D
  int batch = 0;  // utterances per batch of the posterior solve
  double *d_N = nullptr, *d_F = nullptr, *d_T = nullptr, *d_Ts = nullptr, *d_W = nullptr;
  double *d_mean = nullptr, *d_invvar = nullptr, *d_tett = nullptr;
  double *d_acc = nullptr;  // [A C*R*R | Cmx R*sv | Rm R*R | r R | sumW R]
  double *d_meanW = nullptr;
  double *d_Lb = nullptr, *d_Eb = nullptr;  // [batch x R*R] work
  double *d_Yb = nullptr;                   // [batch x R*R] triangular inverse (E-step)
  double *d_invD = nullptr;                 // [batch x nblk x 64 x 64] diagonal-block inverses
  double *d_ones = nullptr;                 // [max(batch, R)]
  double **d_ptr_L = nullptr, **d_ptr_E = nullptr, **d_ptr_W = nullptr;  // batch pointers
  double **d_ptr_A = nullptr, **d_ptr_Tc = nullptr;                      // component pointers
  int *d_info = nullptr;
  cusolverDnHandle_t solver = nullptr;
  double *A() const { return d_acc; }
  double *Cmx() const { return d_acc + (size_t)C * R * R; }
  double *Rm() const { return Cmx() + (size_t)R * sv; }
  double *r() const { return Rm() + (size_t)R * R; }
  double *sumW() const { return r() + R; }
  size_t acc_len() const { return (size_t)C * R * R + (size_t)R * sv + (size_t)R * R + 2 * (size_t)R; }
};

namespace lr {
namespace {

// substractM (AccumulateTVStat.cpp:1088-1105): F[s, c, :] -= mean[c, :] * N[s, c]
__global__ void k_subtract_m(size_t U, int C, int D, const double *__restrict__ N,
                             const double *__restrict__ mean, double *__restrict__ F) {
  size_t sv = (size_t)C * D;
  size_t total = U * sv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t s = i / sv, k = i - s * sv;
    F[i] -= mean[k] * N[s * C + k / D];
  }
}

// Ts = T o invvar (row i of T scaled column-wise)
__global__ void k_scale_cols(int R, size_t sv, const double *__restrict__ T,
                             const double *__restrict__ invvar, double *__restrict__ Ts) {
  size_t total = (size_t)R * sv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x)
    Ts[i] = T[i] * invvar[i % sv];
}

// L[b] += I
__global__ void k_add_identity(int nb, int R, double *__restrict__ L) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < nb * R) {
    int b = i / R, d = i - b * R;
    L[(size_t)b * R * R + (size_t)d * R + d] += 1.0;
  }
}

// E[b] += w_b w_b^T  (Linv += y y^T, AccumulateTVStat.cpp:1766-1768)
__global__ void k_rank1(int nb, int R, const double *__restrict__ W, double *__restrict__ E) {
  size_t total = (size_t)nb * R * R;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t b = i / ((size_t)R * R), e = i - b * (size_t)R * R;
    E[i] += W[b * R + e / R] * W[b * R + e % R];
  }
}

__global__ void k_check_info(int n, const int *__restrict__ info, int *__restrict__ bad) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && info[i] != 0) atomicExch(bad, i + 1);
}

__global__ void k_scale(size_t n, double a, const double *__restrict__ x, double *__restrict__ y) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = a * x[i];
}

// Rm <- Rm / n - r r^T with r <- r / n first (minDivergence :2061-2070)
__global__ void k_mindiv_prep(int R, double n, double *__restrict__ Rm, double *__restrict__ r) {
  __shared__ double rs[1024];
  for (int i = threadIdx.x; i < R; i += blockDim.x) rs[i] = r[i] / n;
  __syncthreads();
  for (int e = threadIdx.x; e < R * R; e += blockDim.x) {
    int i = e / R, j = e - i * R;
    Rm[e] = Rm[e] / n - rs[i] * rs[j];
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R; i += blockDim.x) r[i] = rs[i];
}

// keep the row-major UPPER triangle (= column-major lower) of an R x R factor, zero the rest
__global__ void k_keep_upper_rowmajor(int R, double *__restrict__ M) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < R * R) {
    int i = e / R, j = e - i * R;
    if (j < i) M[e] = 0.0;
  }
}

__global__ void k_scale_row(size_t n, const double *__restrict__ nrm, double *__restrict__ v) {
  // v /= nrm (or 0 when the norm vanished), orthonormalizeT :1585-1592
  double d = *nrm;
  double s = d > 0.0 ? 1.0 / d : 0.0;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    v[i] *= s;
}

// ---- batched dense helpers --------------------------------------------------------------
// cusolverDnDpotrsBatched and cublasDtrsmBatched run at < 1 TFLOP/s for R = 400..600 (measured,
// profiles/r01_extra.md); the two routines below replace them.
//
// Solve L L^T x = b for ONE right-hand side per matrix (column-major lower factor, ld = n).
// One CTA per matrix; x lives in shared memory.  Forward substitution is column oriented
// (axpy over the contiguous column j), the backward pass is a dot product with column j.
constexpr int kSolveThreads = 256;
__global__ void __launch_bounds__(kSolveThreads)
k_chol_solve(int n, const double *__restrict__ Lall, size_t stride, double *__restrict__ rhs) {
  extern __shared__ double xs[];
  __shared__ double red[kSolveThreads / 32];
  const double *L = Lall + (size_t)blockIdx.x * stride;
  double *b = rhs + (size_t)blockIdx.x * n;
  const int tid = threadIdx.x;
  for (int i = tid; i < n; i += kSolveThreads) xs[i] = b[i];
  __syncthreads();
  for (int j = 0; j < n; j++) {  // L y = b
    const double xj = xs[j] / L[(size_t)j * n + j];
    __syncthreads();
    if (tid == 0) xs[j] = xj;
    for (int i = j + 1 + tid; i < n; i += kSolveThreads) xs[i] -= L[(size_t)j * n + i] * xj;
    __syncthreads();
  }
  for (int j = n - 1; j >= 0; j--) {  // L^T x = y :  x_j = (y_j - sum_{i>j} L_ij x_i) / L_jj
    double part = 0.0;
    for (int i = j + 1 + tid; i < n; i += kSolveThreads) part += L[(size_t)j * n + i] * xs[i];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      double t = 0.0;
      for (int w = 0; w < kSolveThreads / 32; w++) t += red[w];
      xs[j] = (xs[j] - t) / L[(size_t)j * n + j];
    }
    __syncthreads();
  }
  for (int i = tid; i < n; i += kSolveThreads) b[i] = xs[i];
}

// Inverses of the NB x NB diagonal blocks of a batch of lower factors.  One CTA per
// (block, matrix); thread j computes column j of the inverse by forward substitution.
constexpr int kNB = 64;
__global__ void __launch_bounds__(kNB)
k_diag_inv(int n, const double *__restrict__ Lall, size_t stride, double *__restrict__ invD,
           int nblk) {
  extern __shared__ double dsm[];
  double (*Ls)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(dsm);
  double (*Xs)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(dsm + kNB * (kNB + 1));
  const int blk = blockIdx.x, mat = blockIdx.y;
  const int k0 = blk * kNB, nb = min(kNB, n - k0);
  const double *L = Lall + (size_t)mat * stride;
  const int j = threadIdx.x;
  for (int c = 0; c < nb; c++)
    if (j < nb) Ls[j][c] = (j >= c) ? L[(size_t)(k0 + c) * n + k0 + j] : 0.0;  // Ls[row][col]
  __syncthreads();
  if (j < nb) {
    for (int i = 0; i < nb; i++) {
      double v = (i == j) ? 1.0 : 0.0;
      if (i < j) {
        Xs[i][j] = 0.0;
        continue;
      }
      for (int k = j; k < i; k++) v -= Ls[i][k] * Xs[k][j];
      Xs[i][j] = v / Ls[i][i];
    }
  }
  __syncthreads();
  double *out = invD + ((size_t)mat * nblk + blk) * kNB * kNB;  // column-major, ld = kNB
  for (int c = 0; c < kNB; c++) out[(size_t)c * kNB + j] = (j < nb && c < nb) ? Xs[j][c] : 0.0;
}

// dst[mat][0:nb, 0:nb] = src[mat][0:nb, 0:nb]  (column-major blocks with their own ld / stride)
__global__ void k_copy_block(int nb, const double *__restrict__ src, int lds, size_t ss,
                             double *__restrict__ dst, int ldd, size_t sd) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= nb * nb) return;
  int col = e / nb, row = e - col * nb;
  dst[(size_t)blockIdx.y * sd + (size_t)col * ldd + row] =
      src[(size_t)blockIdx.y * ss + (size_t)col * lds + row];
}

// zero the strictly upper triangle (column-major) of a batch of n x n matrices
__global__ void k_zero_upper(int n, size_t stride, double *__restrict__ A, int batch) {
  size_t total = (size_t)batch * n * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t m = i / ((size_t)n * n), e = i - m * (size_t)n * n;
    int col = (int)(e / n), row = (int)(e - (size_t)col * n);
    if (row < col) A[m * stride + e] = 0.0;
  }
}

int grid_for(size_t n, int threads = 256) {
  size_t g = (n + threads - 1) / threads;
  return (int)std::min<size_t>(g, (size_t)engine().sm_count * 16);
}

void tv_free(lr_tv *tv) {
  if (!tv) return;
  cudaFree(tv->d_N);
  cudaFree(tv->d_F);
  cudaFree(tv->d_T);
  cudaFree(tv->d_Ts);
  cudaFree(tv->d_W);
  cudaFree(tv->d_mean);
  cudaFree(tv->d_invvar);
  cudaFree(tv->d_tett);
  cudaFree(tv->d_acc);
  cudaFree(tv->d_meanW);
  cudaFree(tv->d_Lb);
  cudaFree(tv->d_Eb);
  cudaFree(tv->d_Yb);
  cudaFree(tv->d_invD);
  cudaFree(tv->d_ones);
  cudaFree(tv->d_ptr_L);
  cudaFree(tv->d_ptr_E);
  cudaFree(tv->d_ptr_W);
  cudaFree(tv->d_ptr_A);
  cudaFree(tv->d_ptr_Tc);
  cudaFree(tv->d_info);
  if (tv->solver) cusolverDnDestroy(tv->solver);
  delete tv;
}

lr_status upload_ptrs(double **dst, double *base, size_t stride, int n) {
  std::vector<double *> h(n);
  for (int i = 0; i < n; i++) h[i] = base + (size_t)i * stride;
  LR_CUDA(cudaMemcpy(dst, h.data(), n * sizeof(double *), cudaMemcpyHostToDevice));
  return LR_OK;
}

lr_status check_factor(lr_tv *tv, int n, const char *what) {
  Engine &e = engine();
  int *bad = tv->d_info + std::max(tv->batch, tv->C);
  LR_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), e.stream));
  k_check_info<<<ceil_div(n, 256), 256, 0, e.stream>>>(n, tv->d_info, bad);
  LR_CHECK_LAUNCH();
  int h = 0;
  LR_CUDA(cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  if (h != 0) return fail(LR_ERR_NUMERIC, "%s: matrix %d of the batch is not positive definite", what, h - 1);
  return LR_OK;
}

// Posterior of a batch of utterances [u0, u0+nb): L = I + N TETt (in d_Lb), Cholesky, then
//   want_inverse == false: W[u] = L^-1 aux (potrs)
//   want_inverse == true : d_Eb = L^-1 (two triangular solves on I), W[u] = L^-1 aux
// (estimateW :2126-2168 / estimateAandC :1722-1760)
lr_status posterior_batch(lr_tv *tv, size_t u0, int nb, bool want_inverse) {
  Engine &e = engine();
  const int R = tv->R, C = tv->C;
  const size_t rr = (size_t)R * R;
  const double one = 1.0, zero = 0.0;
  // Lb[nb x R*R] = N_b[nb x C] * TETt[C x R*R]
  LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_N, CUBLAS_OP_N, (int)rr, nb, C, &one, tv->d_tett, (int)rr,
                        tv->d_N + u0 * C, C, &zero, tv->d_Lb, (int)rr));
  count_launch();
  k_add_identity<<<ceil_div((long)nb * R, 256), 256, 0, e.stream>>>(nb, R, tv->d_Lb);
  LR_CHECK_LAUNCH();
  // aux[nb x R] = Fc_b[nb x sv] * Ts^T  -> written straight into W
  LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_T, CUBLAS_OP_N, R, nb, (int)tv->sv, &one, tv->d_Ts,
                        (int)tv->sv, tv->d_F + u0 * tv->sv, (int)tv->sv, &zero, tv->d_W + u0 * R, R));
  count_launch();
  LR_CUSOLVER(cusolverDnDpotrfBatched(tv->solver, CUBLAS_FILL_MODE_LOWER, R, tv->d_ptr_L, R,
                                      tv->d_info, nb));
  count_launch();
  lr_status st = check_factor(tv, nb, "i-vector posterior precision L");
  if (st != LR_OK) return st;
  if (!want_inverse) {
    // w = L^-1 aux: one CTA per utterance, aux sits in W and is overwritten by the i-vector
    k_chol_solve<<<nb, kSolveThreads, R * sizeof(double), e.stream>>>(R, tv->d_Lb, rr,
                                                                      tv->d_W + u0 * R);
    LR_CHECK_LAUNCH();
    return LR_OK;
  }
  // Explicit inverse Linv = Y^T Y with Y = Lfac^-1, built block row by block row from the
  // inverses of the 64 x 64 diagonal blocks -- everything heavy is a strided-batched DGEMM:
  //   Y[i, 0:i] = -invD_ii (Lfac[i, 0:i] Y[0:i, 0:i]),   Y[i, i] = invD_ii
  const int nblk = (R + kNB - 1) / kNB;
  const double mone = -1.0;
  k_zero_upper<<<grid_for((size_t)nb * rr), 256, 0, e.stream>>>(R, rr, tv->d_Lb, nb);
  LR_CHECK_LAUNCH();
  static bool diag_attr = false;
  const size_t diag_smem = 2 * kNB * (kNB + 1) * sizeof(double);
  if (!diag_attr) {
    LR_CUDA(cudaFuncSetAttribute(k_diag_inv, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)diag_smem));
    diag_attr = true;
  }
  k_diag_inv<<<dim3(nblk, nb), kNB, diag_smem, e.stream>>>(R, tv->d_Lb, rr, tv->d_invD, nblk);
  LR_CHECK_LAUNCH();
  LR_CUDA(cudaMemsetAsync(tv->d_Yb, 0, (size_t)nb * rr * sizeof(double), e.stream));
  const long long sD = (long long)nblk * kNB * kNB;
  for (int i = 0; i < nblk; i++) {
    const int r0 = i * kNB, nbi = std::min(kNB, R - r0);
    if (i > 0) {
      // Tmp[nbi x r0] = Lfac[r0:r0+nbi, 0:r0] * Y[0:r0, 0:r0]   (Tmp lives in Eb)
      LR_CUBLAS(cublasDgemmStridedBatched(e.blas, CUBLAS_OP_N, CUBLAS_OP_N, nbi, r0, r0, &one,
                                          tv->d_Lb + r0, R, (long long)rr, tv->d_Yb, R,
                                          (long long)rr, &zero, tv->d_Eb, R, (long long)rr, nb));
      count_launch();
      // Y[r0:, 0:r0] = -invD_ii * Tmp
      LR_CUBLAS(cublasDgemmStridedBatched(e.blas, CUBLAS_OP_N, CUBLAS_OP_N, nbi, r0, nbi, &mone,
                                          tv->d_invD + (size_t)i * kNB * kNB, kNB, sD, tv->d_Eb, R,
                                          (long long)rr, &zero, tv->d_Yb + r0, R, (long long)rr, nb));
      count_launch();
    }
    // Y[r0:, r0:] = invD_ii
    k_copy_block<<<dim3(ceil_div((long)nbi * nbi, 256), nb), 256, 0, e.stream>>>(
        nbi, tv->d_invD + (size_t)i * kNB * kNB, kNB, (size_t)sD, tv->d_Yb + (size_t)r0 * R + r0, R, rr);
    LR_CHECK_LAUNCH();
  }
  // Linv = Y^T Y  -> Eb
  LR_CUBLAS(cublasDgemmStridedBatched(e.blas, CUBLAS_OP_T, CUBLAS_OP_N, R, R, R, &one, tv->d_Yb, R,
                                      (long long)rr, tv->d_Yb, R, (long long)rr, &zero, tv->d_Eb, R,
                                      (long long)rr, nb));
  count_launch();
  // W_b = Linv_b aux_b : aux currently sits in W; go through Lb's first nb*R doubles as scratch
  LR_CUDA(cudaMemcpyAsync(tv->d_Lb, tv->d_W + u0 * R, (size_t)nb * R * sizeof(double),
                          cudaMemcpyDeviceToDevice, e.stream));
  LR_CUBLAS(cublasDgemmStridedBatched(e.blas, CUBLAS_OP_N, CUBLAS_OP_N, R, 1, R, &one, tv->d_Eb, R,
                                      (long long)rr, tv->d_Lb, R, R, &zero, tv->d_W + u0 * R, R, R,
                                      nb));
  count_launch();
  return LR_OK;
}

}  // namespace
}  // namespace lr

using namespace lr;

extern "C" {

lr_tv *lr_tv_create(int C, int D, int R, size_t U, const double *ubm_mean,
                    const double *ubm_invvar) {
  if (!ensure_ready()) return nullptr;
  if (C < 1 || D < 1 || R < 1 || R > 1024 || U < 1 || !ubm_mean || !ubm_invvar) {
    fail(LR_ERR_ARG, "lr_tv_create: bad arguments (C=%d D=%d R=%d U=%zu; R <= 1024)", C, D, R, U);
    return nullptr;
  }
  lr_tv *tv = new lr_tv();
  tv->C = C;
  tv->D = D;
  tv->R = R;
  tv->U = U;
  tv->sv = (size_t)C * D;
  size_t rr = (size_t)R * R;
  tv->batch = (int)std::min<size_t>(U, std::max<size_t>(32, ((size_t)1 << 29) / (rr * sizeof(double))));
  const int nbmax = tv->batch;
  auto A = [&](double **p, size_t n) { return cudaMalloc(p, n * sizeof(double)) == cudaSuccess; };
  bool ok = A(&tv->d_N, U * C) && A(&tv->d_F, U * tv->sv) && A(&tv->d_T, R * tv->sv) &&
            A(&tv->d_Ts, R * tv->sv) && A(&tv->d_W, U * R) && A(&tv->d_mean, tv->sv) &&
            A(&tv->d_invvar, tv->sv) && A(&tv->d_tett, (size_t)C * rr) && A(&tv->d_acc, tv->acc_len()) &&
            A(&tv->d_meanW, R) && A(&tv->d_Lb, (size_t)nbmax * rr) && A(&tv->d_Eb, (size_t)nbmax * rr) &&
            A(&tv->d_Yb, (size_t)nbmax * rr) &&
            A(&tv->d_invD, (size_t)nbmax * ((R + kNB - 1) / kNB) * kNB * kNB) &&
            A(&tv->d_ones, std::max<size_t>(nbmax, R)) &&
            cudaMalloc(&tv->d_ptr_L, nbmax * sizeof(double *)) == cudaSuccess &&
            cudaMalloc(&tv->d_ptr_E, nbmax * sizeof(double *)) == cudaSuccess &&
            cudaMalloc(&tv->d_ptr_W, nbmax * sizeof(double *)) == cudaSuccess &&
            cudaMalloc(&tv->d_ptr_A, C * sizeof(double *)) == cudaSuccess &&
            cudaMalloc(&tv->d_ptr_Tc, C * sizeof(double *)) == cudaSuccess &&
            cudaMalloc(&tv->d_info, (std::max(nbmax, C) + 1) * sizeof(int)) == cudaSuccess;
  if (!ok) {
    fail(LR_ERR_CUDA, "lr_tv_create: cudaMalloc failed: %s", cudaGetErrorString(cudaGetLastError()));
    tv_free(tv);
    return nullptr;
  }
  Engine &e = engine();
  bool good = cusolverDnCreate(&tv->solver) == CUSOLVER_STATUS_SUCCESS &&
              cusolverDnSetStream(tv->solver, e.stream) == CUSOLVER_STATUS_SUCCESS &&
              cudaMemcpy(tv->d_mean, ubm_mean, tv->sv * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(tv->d_invvar, ubm_invvar, tv->sv * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemset(tv->d_acc, 0, tv->acc_len() * sizeof(double)) == cudaSuccess &&
              cudaMemset(tv->d_N, 0, U * C * sizeof(double)) == cudaSuccess &&
              cudaMemset(tv->d_F, 0, U * tv->sv * sizeof(double)) == cudaSuccess &&
              cudaMemset(tv->d_W, 0, U * R * sizeof(double)) == cudaSuccess &&
              cudaMemset(tv->d_T, 0, R * tv->sv * sizeof(double)) == cudaSuccess &&
              cudaMemset(tv->d_meanW, 0, R * sizeof(double)) == cudaSuccess;
  if (good) {
    std::vector<double> ones(std::max<size_t>(nbmax, R), 1.0);
    good = cudaMemcpy(tv->d_ones, ones.data(), ones.size() * sizeof(double), cudaMemcpyHostToDevice) == cudaSuccess &&
           upload_ptrs(tv->d_ptr_L, tv->d_Lb, rr, nbmax) == LR_OK &&
           upload_ptrs(tv->d_ptr_E, tv->d_Eb, rr, nbmax) == LR_OK &&
           upload_ptrs(tv->d_ptr_A, tv->d_tett, rr, C) == LR_OK &&  // M-step factors a COPY of A
           upload_ptrs(tv->d_ptr_Tc, tv->d_T, (size_t)D, C) == LR_OK;
  }
  if (!good) {
    fail(LR_ERR_CUDA, "lr_tv_create: initialisation failed: %s", cudaGetErrorString(cudaGetLastError()));
    tv_free(tv);
    return nullptr;
  }
  return tv;
}

void lr_tv_destroy(lr_tv *tv) { tv_free(tv); }

#define TV_COPY(dst, src, n, kind)                                                         \
  LR_CUDA(cudaMemcpyAsync(dst, src, (n) * sizeof(double), kind, engine().stream))

lr_status lr_tv_set_stats(lr_tv *tv, const double *N, const double *F) {
  LR_READY();
  LR_REQUIRE(tv && N && F, "lr_tv_set_stats: null argument");
  TV_COPY(tv->d_N, N, tv->U * tv->C, cudaMemcpyHostToDevice);
  TV_COPY(tv->d_F, F, tv->U * tv->sv, cudaMemcpyHostToDevice);
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

lr_status lr_tv_get_stats(lr_tv *tv, double *N, double *F) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_get_stats: null handle");
  if (N) TV_COPY(N, tv->d_N, tv->U * tv->C, cudaMemcpyDeviceToHost);
  if (F) TV_COPY(F, tv->d_F, tv->U * tv->sv, cudaMemcpyDeviceToHost);
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

double *lr_tv_dev_N(lr_tv *tv) { return tv ? tv->d_N : nullptr; }
double *lr_tv_dev_F(lr_tv *tv) { return tv ? tv->d_F : nullptr; }

lr_status lr_tv_set_T(lr_tv *tv, const double *T) {
  LR_READY();
  LR_REQUIRE(tv && T, "lr_tv_set_T: null argument");
  TV_COPY(tv->d_T, T, (size_t)tv->R * tv->sv, cudaMemcpyHostToDevice);
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

lr_status lr_tv_get_T(lr_tv *tv, double *T) {
  LR_READY();
  LR_REQUIRE(tv && T, "lr_tv_get_T: null argument");
  TV_COPY(T, tv->d_T, (size_t)tv->R * tv->sv, cudaMemcpyDeviceToHost);
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

lr_status lr_tv_get_mean(lr_tv *tv, double *ubm_mean) {
  LR_READY();
  LR_REQUIRE(tv && ubm_mean, "lr_tv_get_mean: null argument");
  TV_COPY(ubm_mean, tv->d_mean, tv->sv, cudaMemcpyDeviceToHost);
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

lr_status lr_tv_set_mean(lr_tv *tv, const double *ubm_mean) {
  LR_READY();
  LR_REQUIRE(tv && ubm_mean, "lr_tv_set_mean: null argument");
  TV_COPY(tv->d_mean, ubm_mean, tv->sv, cudaMemcpyHostToDevice);
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

lr_status lr_tv_get_W(lr_tv *tv, double *W) {
  LR_READY();
  LR_REQUIRE(tv && W, "lr_tv_get_W: null argument");
  TV_COPY(W, tv->d_W, tv->U * tv->R, cudaMemcpyDeviceToHost);
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

lr_status lr_tv_get_acc(lr_tv *tv, double *A, double *Cmx, double *Rm, double *r, double *meanW) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_get_acc: null handle");
  size_t rr = (size_t)tv->R * tv->R;
  if (A) TV_COPY(A, tv->A(), (size_t)tv->C * rr, cudaMemcpyDeviceToHost);
  if (Cmx) TV_COPY(Cmx, tv->Cmx(), (size_t)tv->R * tv->sv, cudaMemcpyDeviceToHost);
  if (Rm) TV_COPY(Rm, tv->Rm(), rr, cudaMemcpyDeviceToHost);
  if (r) TV_COPY(r, tv->r(), tv->R, cudaMemcpyDeviceToHost);
  if (meanW) TV_COPY(meanW, tv->d_meanW, tv->R, cudaMemcpyDeviceToHost);
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

lr_status lr_tv_reset_tmp_acc(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_reset_tmp_acc: null handle");
  // resetTmpAcc (TotalVariability.cpp:149): the only place the reference zeroes _Cmx
  LR_CUDA(cudaMemsetAsync(tv->d_acc, 0, tv->acc_len() * sizeof(double), engine().stream));
  return LR_OK;
}

lr_status lr_tv_subtract_m(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_subtract_m: null handle");
  k_subtract_m<<<grid_for(tv->U * tv->sv), 256, 0, engine().stream>>>(tv->U, tv->C, tv->D, tv->d_N,
                                                                      tv->d_mean, tv->d_F);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

lr_status lr_tv_estimate_tett(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_estimate_tett: null handle");
  Engine &e = engine();
  const double one = 1.0, zero = 0.0;
  k_scale_cols<<<grid_for((size_t)tv->R * tv->sv), 256, 0, e.stream>>>(tv->R, tv->sv, tv->d_T,
                                                                       tv->d_invvar, tv->d_Ts);
  LR_CHECK_LAUNCH();
  // TETt_c = (T_c o invvar_c) T_c^T : column-major view of the row-major slice T[:, cD:(c+1)D]
  // is the D x R matrix T_c^T with leading dimension C*D.
  LR_CUBLAS(cublasDgemmStridedBatched(e.blas, CUBLAS_OP_T, CUBLAS_OP_N, tv->R, tv->R, tv->D, &one,
                                      tv->d_Ts, (int)tv->sv, tv->D, tv->d_T, (int)tv->sv, tv->D,
                                      &zero, tv->d_tett, tv->R, (long long)tv->R * tv->R, tv->C));
  count_launch();
  return LR_OK;
}

lr_status lr_tv_estimate_w(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_estimate_w: null handle");
  for (size_t u0 = 0; u0 < tv->U; u0 += tv->batch) {
    int nb = (int)std::min<size_t>(tv->batch, tv->U - u0);
    lr_status st = posterior_batch(tv, u0, nb, false);
    if (st != LR_OK) return st;
  }
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

lr_status lr_tv_estimate_a_and_c(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_estimate_a_and_c: null handle");
  Engine &e = engine();
  const int R = tv->R, C = tv->C;
  const size_t rr = (size_t)R * R;
  const double one = 1.0;
  // _A, _R, _r, _meanW are zeroed here; _Cmx is NOT (AccumulateTVStat.cpp:1719-1721)
  LR_CUDA(cudaMemsetAsync(tv->A(), 0, (size_t)C * rr * sizeof(double), e.stream));
  LR_CUDA(cudaMemsetAsync(tv->Rm(), 0, (rr + 2 * (size_t)R) * sizeof(double), e.stream));
  for (size_t u0 = 0; u0 < tv->U; u0 += tv->batch) {
    int nb = (int)std::min<size_t>(tv->batch, tv->U - u0);
    lr_status st = posterior_batch(tv, u0, nb, true);
    if (st != LR_OK) return st;
    const double *Wb = tv->d_W + u0 * R;
    // r += sum_b w_b ; sumW likewise (:1762, :1772)
    LR_CUBLAS(cublasDgemv(e.blas, CUBLAS_OP_N, R, nb, &one, Wb, R, tv->d_ones, 1, &one, tv->r(), 1));
    count_launch();
    // E_b = Linv_b + w_b w_b^T
    k_rank1<<<grid_for((size_t)nb * rr), 256, 0, e.stream>>>(nb, R, Wb, tv->d_Eb);
    LR_CHECK_LAUNCH();
    // Rm += sum_b E_b (:1770)
    LR_CUBLAS(cublasDgemv(e.blas, CUBLAS_OP_N, (int)rr, nb, &one, tv->d_Eb, (int)rr, tv->d_ones, 1,
                          &one, tv->Rm(), 1));
    count_launch();
    // A[C x R*R] += N_b^T E_b (:1775-1782)
    LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_N, CUBLAS_OP_T, (int)rr, C, nb, &one, tv->d_Eb, (int)rr,
                          tv->d_N + u0 * C, C, &one, tv->A(), (int)rr));
    count_launch();
    // Cmx[R x sv] += W_b^T Fc_b (:1784-1788)
    LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_N, CUBLAS_OP_T, (int)tv->sv, R, nb, &one,
                          tv->d_F + u0 * tv->sv, (int)tv->sv, Wb, R, &one, tv->Cmx(), (int)tv->sv));
    count_launch();
  }
  LR_CUDA(cudaMemcpyAsync(tv->sumW(), tv->r(), R * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  return lr_tv_finish_estep(tv, (double)tv->U);
}

lr_status lr_tv_finish_estep(lr_tv *tv, double n_speakers_total) {
  LR_READY();
  LR_REQUIRE(tv && n_speakers_total > 0, "lr_tv_finish_estep: bad argument");
  Engine &e = engine();
  k_scale<<<ceil_div(tv->R, 256), 256, 0, e.stream>>>((size_t)tv->R, 1.0 / n_speakers_total,
                                                      tv->sumW(), tv->d_meanW);
  LR_CHECK_LAUNCH();
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_tv_update_t(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_update_t: null handle");
  Engine &e = engine();
  const int R = tv->R, C = tv->C, D = tv->D;
  const size_t rr = (size_t)R * R;
  const double one = 1.0;
  // factor a copy of A in the TETt buffer (TETt is re-estimated from the new T anyway)
  LR_CUDA(cudaMemcpyAsync(tv->d_tett, tv->A(), (size_t)C * rr * sizeof(double),
                          cudaMemcpyDeviceToDevice, e.stream));
  LR_CUSOLVER(cusolverDnDpotrfBatched(tv->solver, CUBLAS_FILL_MODE_LOWER, R, tv->d_ptr_A, R,
                                      tv->d_info, C));
  count_launch();
  lr_status st = check_factor(tv, C, "M-step accumulator A_c");
  if (st != LR_OK) return st;
  // T_c = A_c^-1 Cmx_c: column-major we hold T_c^T (D x R, ld sv): X^T L L^T = Cmx_c^T
  LR_CUDA(cudaMemcpyAsync(tv->d_T, tv->Cmx(), (size_t)R * tv->sv * sizeof(double),
                          cudaMemcpyDeviceToDevice, e.stream));
  LR_CUBLAS(cublasDtrsmBatched(e.blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T,
                               CUBLAS_DIAG_NON_UNIT, D, R, &one, tv->d_ptr_A, R, tv->d_ptr_Tc,
                               (int)tv->sv, C));
  count_launch();
  LR_CUBLAS(cublasDtrsmBatched(e.blas, CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N,
                               CUBLAS_DIAG_NON_UNIT, D, R, &one, tv->d_ptr_A, R, tv->d_ptr_Tc,
                               (int)tv->sv, C));
  count_launch();
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_tv_min_divergence(lr_tv *tv, double n_sessions) {
  LR_READY();
  LR_REQUIRE(tv && n_sessions > 0, "lr_tv_min_divergence: bad argument");
  Engine &e = engine();
  const int R = tv->R;
  const double one = 1.0, zero = 0.0;
  k_mindiv_prep<<<1, 1024, 0, e.stream>>>(R, n_sessions, tv->Rm(), tv->r());
  LR_CHECK_LAUNCH();
  // Ch = upperCholesky(Rm), Rm = Ch^T Ch.  Column-major LOWER factor of the same buffer is Ch^T,
  // i.e. exactly the row-major upper factor.
  double *Ch = tv->d_Lb;
  LR_CUDA(cudaMemcpyAsync(Ch, tv->Rm(), (size_t)R * R * sizeof(double), cudaMemcpyDeviceToDevice,
                          e.stream));
  int lwork = 0;
  LR_CUSOLVER(cusolverDnDpotrf_bufferSize(tv->solver, CUBLAS_FILL_MODE_LOWER, R, Ch, R, &lwork));
  double *work = (double *)scratch_get(kSlotTmpA, (size_t)lwork * sizeof(double));
  if (!work) return LR_ERR_CUDA;
  LR_CUSOLVER(cusolverDnDpotrf(tv->solver, CUBLAS_FILL_MODE_LOWER, R, Ch, R, work, lwork, tv->d_info));
  count_launch();
  lr_status st = check_factor(tv, 1, "minDivergence covariance R");
  if (st != LR_OK) return st;
  k_keep_upper_rowmajor<<<ceil_div((long)R * R, 256), 256, 0, e.stream>>>(R, Ch);
  LR_CHECK_LAUNCH();
  // mean += meanW^T T (T after the M-step, before the rotation; :2074-2085)
  LR_CUBLAS(cublasDgemv(e.blas, CUBLAS_OP_N, (int)tv->sv, R, &one, tv->d_T, (int)tv->sv, tv->d_meanW,
                        1, &one, tv->d_mean, 1));
  count_launch();
  // T <- Ch T  (:2087-2097)
  LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_N, CUBLAS_OP_N, (int)tv->sv, R, R, &one, tv->d_T,
                        (int)tv->sv, Ch, R, &zero, tv->d_Ts, (int)tv->sv));
  count_launch();
  LR_CUDA(cudaMemcpyAsync(tv->d_T, tv->d_Ts, (size_t)R * tv->sv * sizeof(double),
                          cudaMemcpyDeviceToDevice, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_tv_orthonormalize_t(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_orthonormalize_t: null handle");
  Engine &e = engine();
  const int R = tv->R;
  const int sv = (int)tv->sv;
  const double one = 1.0, zero = 0.0, mone = -1.0;
  // classical Gram-Schmidt over rows, projections against the ORIGINAL row (:1548-1596);
  // Q is built in d_Ts, coefficients in d_Lb
  double *Q = tv->d_Ts, *coef = tv->d_Lb, *nrm = tv->d_Lb + R;
  LR_CUBLAS(cublasSetPointerMode(e.blas, CUBLAS_POINTER_MODE_HOST));
  for (int j = 0; j < R; j++) {
    double *qj = Q + (size_t)j * sv;
    const double *tj = tv->d_T + (size_t)j * sv;
    LR_CUDA(cudaMemcpyAsync(qj, tj, (size_t)sv * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
    if (j > 0) {
      // coef = Q[0:j] t_j ; q_j -= Q[0:j]^T coef   (Q rows are column-major columns, ld sv)
      LR_CUBLAS(cublasDgemv(e.blas, CUBLAS_OP_T, sv, j, &one, Q, sv, tj, 1, &zero, coef, 1));
      LR_CUBLAS(cublasDgemv(e.blas, CUBLAS_OP_N, sv, j, &mone, Q, sv, coef, 1, &one, qj, 1));
      count_launch(2);
    }
    LR_CUBLAS(cublasSetPointerMode(e.blas, CUBLAS_POINTER_MODE_DEVICE));
    cublasStatus_t cs = cublasDnrm2(e.blas, sv, qj, 1, nrm);
    cublasSetPointerMode(e.blas, CUBLAS_POINTER_MODE_HOST);
    if (cs != CUBLAS_STATUS_SUCCESS) return fail(LR_ERR_CUDA, "cublasDnrm2 failed (%d)", (int)cs);
    k_scale_row<<<grid_for((size_t)sv), 256, 0, e.stream>>>((size_t)sv, nrm, qj);
    LR_CHECK_LAUNCH();
  }
  LR_CUDA(cudaMemcpyAsync(tv->d_T, Q, (size_t)R * sv * sizeof(double), cudaMemcpyDeviceToDevice,
                          e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

double *lr_tv_dev_acc(lr_tv *tv) { return tv ? tv->d_acc : nullptr; }
size_t lr_tv_acc_len(const lr_tv *tv) { return tv ? tv->acc_len() : 0; }

}  // extern "C"
