// ivbackend.cu -- i-vector back-end around the PLDA scorer: the development-set
// statistics and normalisations of PldaDev and the cosine / Mahalanobis / two-covariance scorings
// of PldaTest (LIA_SpkTools/src/PldaTools.cpp), fp64.
//
// The reference walks these as scalar triple loops over (dimension, dimension, session) or
// (model, segment, dimension, dimension).  Here every covariance is a rank-n symmetric update
// (cuBLAS DSYRK on centred copies of the data) and every scoring is one [n_models x d] x [d x n_test]
// GEMM plus per-model / per-segment quadratic forms; the glue is hand-written kernels.
// Vectors are COLUMNS of row-major [d x n] matrices like the reference's _data / _models /
// _segments, i.e. cuBLAS sees a column-major [n x d] matrix with leading dimension n.
#include <cusolverDn.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

#define LR_CUSOLVER(expr)                                                                  \
  do {                                                                                     \
    cusolverStatus_t s__ = (expr);                                                         \
    if (s__ != CUSOLVER_STATUS_SUCCESS)                                                    \
      return lr::fail(LR_ERR_CUDA, "%s:%d %s: cusolver status %d", __FILE__, __LINE__,     \
                      #expr, (int)s__);                                                    \
  } while (0)

namespace lr {
namespace {

constexpr int kThreads = 256;

inline int grid_for(size_t n) {
  size_t g = (n + kThreads - 1) / kThreads;
  return (int)std::min<size_t>(std::max<size_t>(g, 1), (size_t)engine().sm_count * 16);
}

// ---- kernels -------------------------------------------------------------------------------
// row sums of a row-major [d x n] matrix, scaled: out[k] = scale * sum_s X[k, s]
__global__ void k_row_sum(size_t n, const double *__restrict__ X, double scale, double *__restrict__ out) {
  const double *row = X + (size_t)blockIdx.x * n;
  double p = 0.0;
  for (size_t s = threadIdx.x; s < n; s += blockDim.x) p += row[s];
  __shared__ double red[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < kThreads / 32; w++) t += red[w];
    out[blockIdx.x] = t * scale;
  }
}
// per-speaker sums: sums[k, class_of[s]] += X[k, s]; counts from row 0 only
__global__ void k_class_sums(int d, size_t n, size_t n_spk, const double *__restrict__ X,
                             const int *__restrict__ class_of, double *__restrict__ sums,
                             double *__restrict__ cnt) {
  size_t total = (size_t)d * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t k = i / n, s = i - k * n;
    atomicAdd(&sums[k * n_spk + class_of[s]], X[i]);
    if (k == 0) atomicAdd(&cnt[class_of[s]], 1.0);
  }
}
__global__ void k_class_div(int d, size_t n_spk, const double *__restrict__ cnt, double *__restrict__ sums) {
  size_t total = (size_t)d * n_spk;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x)
    sums[i] /= cnt[i % n_spk];
}
// Y[k, s] = (X[k, s] - mu[k]) : mode 0;  (X[k, s] - sm[k, class s]) * (wccn ? 1/sqrt(cnt) : 1) : mode 1
__global__ void k_center(int d, size_t n, size_t n_spk, const double *__restrict__ X,
                         const double *__restrict__ mu, const double *__restrict__ sm,
                         const int *__restrict__ class_of, const double *__restrict__ cnt, int mode,
                         int wccn, double *__restrict__ Y) {
  size_t total = (size_t)d * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t k = i / n, s = i - k * n;
    if (mode == 0) {
      Y[i] = X[i] - mu[k];
    } else {
      int c = class_of[s];
      double v = X[i] - sm[k * n_spk + c];
      Y[i] = wccn ? v / sqrt(cnt[c]) : v;
    }
  }
}
// Y[k, c] = sqrt(cnt[c]) (sm[k, c] - mu[k])   (between-class term, :544-546)
__global__ void k_between(int d, size_t n_spk, const double *__restrict__ sm, const double *__restrict__ mu,
                          const double *__restrict__ cnt, double *__restrict__ Y) {
  size_t total = (size_t)d * n_spk;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t k = i / n_spk, c = i - k * n_spk;
    Y[i] = sqrt(cnt[c]) * (sm[i] - mu[k]);
  }
}
// mirror the column-major LOWER triangle (= row-major upper) of an n x n matrix into the other half
__global__ void k_sym_from_lower(int n, double *__restrict__ A) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n * n) {
    int col = e / n, row = e - col * n;
    if (row < col) A[e] = A[(size_t)row * n + col];
  }
}
// keep the row-major UPPER triangle (column-major lower), zero the rest
__global__ void k_keep_upper_rm(int n, double *__restrict__ M) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n * n) {
    int i = e / n, j = e - i * n;
    if (j < i) M[e] = 0.0;
  }
}
// out[j, :] = eigenvector (n - 1 - j) of column-major V (ascending eigenvalues), scaled by
// `scale_mode`: 0 -> unit Euclidean norm, 1 -> 1 / sqrt(lambda); sign: largest |component| > 0.
__global__ void k_eig_rows(int n, int rank, const double *__restrict__ V, const double *__restrict__ lam,
                           int scale_mode, double *__restrict__ out, int *__restrict__ bad) {
  int j = blockIdx.x;
  if (j >= rank) return;
  const double *col = V + (size_t)(n - 1 - j) * n;
  __shared__ double best[kThreads], nrm2[kThreads];
  __shared__ int besti[kThreads];
  double b = -1.0, q = 0.0;
  int bi = 0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double a = fabs(col[k]);
    q += col[k] * col[k];
    if (a > b) {
      b = a;
      bi = k;
    }
  }
  best[threadIdx.x] = b;
  besti[threadIdx.x] = bi;
  nrm2[threadIdx.x] = q;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      double ob = best[threadIdx.x + o];
      int oi = besti[threadIdx.x + o];
      if (ob > best[threadIdx.x] || (ob == best[threadIdx.x] && oi < besti[threadIdx.x])) {
        best[threadIdx.x] = ob;
        besti[threadIdx.x] = oi;
      }
      nrm2[threadIdx.x] += nrm2[threadIdx.x + o];
    }
    __syncthreads();
  }
  double sc = col[besti[0]] < 0.0 ? -1.0 : 1.0;
  if (scale_mode == 0) {
    sc /= sqrt(nrm2[0]);
  } else {
    double l = lam[n - 1 - j];
    if (!(l > 0.0)) {
      if (threadIdx.x == 0) atomicExch(bad, j + 1);
      l = 1.0;
    }
    sc /= sqrt(l) * sqrt(nrm2[0]);
  }
  for (int k = threadIdx.x; k < n; k += blockDim.x) out[(size_t)j * n + k] = sc * col[k];
}
// X[k, s] -= mu[k]
__global__ void k_sub_mu(int d, size_t n, const double *__restrict__ mu, double *__restrict__ X) {
  size_t total = (size_t)d * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x)
    X[i] -= mu[i / n];
}
// column norms of a row-major [d x n] matrix: nrm[s] = sqrt(sum_k X[k, s]^2)
__global__ void k_col_norm(int d, size_t n, const double *__restrict__ X, double *__restrict__ nrm) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double t = 0.0;
  for (int k = 0; k < d; k++) t += X[(size_t)k * n + s] * X[(size_t)k * n + s];
  nrm[s] = sqrt(t);
}
__global__ void k_div_cols(int d, size_t n, const double *__restrict__ nrm, double *__restrict__ X) {
  size_t total = (size_t)d * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x)
    X[i] /= nrm[i % n];
}
// q[s] = sum_k X[k, s] Y[k, s]  (quadratic forms x^T A x with Y = A X)
__global__ void k_col_dot(int d, size_t n, const double *__restrict__ X, const double *__restrict__ Y,
                          double *__restrict__ q) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double t = 0.0;
  for (int k = 0; k < d; k++) t += X[(size_t)k * n + s] * Y[(size_t)k * n + s];
  q[s] = t;
}
// finish a block of scores S[mb x nt] (models m0..m0+mb):
//   mode 0 cosine      : S = cross / (a[m] b[s])
//   mode 1 mahalanobis : S = -0.5 (a[m] + b[s] - cross)
//   mode 2 two-cov     : S = cross + a[m] + b[s]
// trials (may be NULL): pairs outside the mask score 0 like the reference's untouched _scores
__global__ void k_finish_scores(size_t mb, size_t nt, size_t m0, const double *__restrict__ a,
                                const double *__restrict__ b, const unsigned char *__restrict__ trials,
                                int mode, double *__restrict__ S) {
  size_t total = mb * nt;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    size_t m = e / nt, s = e - m * nt;
    double v;
    if (trials && !trials[(m0 + m) * nt + s]) {
      v = 0.0;
    } else if (mode == 0) {
      v = S[e] / (a[m0 + m] * b[s]);
    } else if (mode == 1) {
      v = -0.5 * (a[m0 + m] + b[s] - S[e]);
    } else {
      v = S[e] + a[m0 + m] + b[s];
    }
    S[e] = v;
  }
}
// C = alpha A + beta B^T-or-B (n x n, elementwise helpers for the small d x d algebra)
__global__ void k_axpby(int n, double alpha, const double *__restrict__ A, double beta,
                        const double *__restrict__ B, int transpose_b, double *__restrict__ C) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < n * n) {
    int i = e / n, j = e - i * n;
    C[e] = alpha * A[e] + beta * (transpose_b ? B[(size_t)j * n + i] : B[e]);
  }
}

// ---- PLDA EM kernels -------------------------------------------------------------------------
// per-class column sums of a row-major [r x n] matrix: out[i, class_of[s]] += X[i, s]
__global__ void k_class_colsum(int r, size_t n, size_t n_spk, const double *__restrict__ X,
                               const int *__restrict__ class_of, double *__restrict__ out) {
  size_t total = (size_t)r * n;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t k = i / n, s = i - k * n;
    atomicAdd(&out[k * n_spk + class_of[s]], X[i]);
  }
}
__global__ void k_count_classes(size_t n, const int *__restrict__ class_of, double *__restrict__ cnt) {
  size_t s = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (s < n) atomicAdd(&cnt[class_of[s]], 1.0);
}
// Z[i, spk] /= (cnt[spk] * lam[i] + 1)   (M_n = V (n D + I)^-1 V^T applied in the eigenbasis, :2417-2424)
__global__ void k_scale_posterior(int rF, size_t n_spk, const double *__restrict__ cnt,
                                  const double *__restrict__ lam, double *__restrict__ Z) {
  size_t total = (size_t)rF * n_spk;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    size_t i = e / n_spk, c = e - i * n_spk;
    Z[e] /= (cnt[c] * lam[i] + 1.0);
  }
}
// w[i] = sum_spk cnt / (cnt lam_i + 1)  -> sum over speakers of n_spk M_{n_spk} in the eigenbasis
__global__ void k_mbar_weights(int rF, size_t n_spk, const double *__restrict__ cnt,
                               const double *__restrict__ lam, double *__restrict__ w) {
  int i = blockIdx.x;
  if (i >= rF) return;
  double p = 0.0;
  for (size_t c = threadIdx.x; c < n_spk; c += blockDim.x) p += cnt[c] / (cnt[c] * lam[i] + 1.0);
  __shared__ double red[kThreads / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < kThreads / 32; k++) t += red[k];
    w[i] = t;
  }
}
// Y[i, :] = w[i] X[i, :]  (row scaling of a row-major [r x c] matrix)
__global__ void k_scale_rows(int r, int c, const double *__restrict__ w, const double *__restrict__ X,
                             double *__restrict__ Y) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < r * c) Y[e] = w[e / c] * X[e];
}
// Eh[(rF + rG) x n]: rows 0..rF = thisEh[:, class], rows rF.. = low[:, s] - SE[:, class]  (:2455-2464)
__global__ void k_build_eh(int rF, int rG, size_t n, size_t n_spk, const int *__restrict__ class_of,
                           const double *__restrict__ thisEh, const double *__restrict__ low,
                           const double *__restrict__ SE, double *__restrict__ Eh) {
  size_t total = (size_t)(rF + rG) * n;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    size_t i = e / n, s = e - i * n;
    int c = class_of[s];
    Eh[e] = i < (size_t)rF ? thisEh[i * n_spk + c]
                           : low[(i - rF) * n + s] - SE[(i - rF) * n_spk + c];
  }
}
// dst[row0 + i, col0 + j] += alpha * (transpose ? src[j, i] : src[i, j])  (row-major, src is rows x cols
// as seen AFTER the optional transposition)
__global__ void k_add_block(double *__restrict__ dst, int ldd, int row0, int col0,
                            const double *__restrict__ src, int lds, int rows, int cols, double alpha,
                            int transpose) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < rows * cols) {
    int i = e / cols, j = e - i * cols;
    dst[(size_t)(row0 + i) * ldd + col0 + j] += alpha * (transpose ? src[(size_t)j * lds + i] : src[(size_t)i * lds + j]);
  }
}
__global__ void k_add_diag(int n, double v, double *__restrict__ A) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) A[(size_t)i * n + i] += v;
}
// c = Ehh / n - u u^T with u = U / n (U scaled in place)   (:2806-2807)
__global__ void k_mindiv_cov(int r, double n, const double *__restrict__ Ehh, double *__restrict__ U,
                             double *__restrict__ c) {
  __shared__ double u[1024];
  for (int i = threadIdx.x; i < r; i += blockDim.x) u[i] = U[i] / n;
  __syncthreads();
  for (int e = threadIdx.x; e < r * r; e += blockDim.x) c[e] = Ehh[e] / n - u[e / r] * u[e % r];
  __syncthreads();
  for (int i = threadIdx.x; i < r; i += blockDim.x) U[i] = u[i];
}
// out[i, j] = src[row0 + i, col0 + j] (row-major block copy)
__global__ void k_copy_block_rm(const double *__restrict__ src, int lds, int row0, int col0, int rows,
                                int cols, double *__restrict__ out) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < rows * cols) out[e] = src[(size_t)(row0 + e / cols) * lds + col0 + e % cols];
}
// Sigma = (sigmaObs - SL) / n
__global__ void k_sigma_update(int dd, double n, const double *__restrict__ obs, const double *__restrict__ SL,
                               double *__restrict__ Sigma) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e < dd) Sigma[e] = (obs[e] - SL[e]) / n;
}

// ---- dense helpers -------------------------------------------------------------------------
struct Dense {
  cusolverDnHandle_t solver = nullptr;
  DevBuf<double> work;
  DevBuf<int> info;
  ~Dense() {
    if (solver) cusolverDnDestroy(solver);
  }
  lr_status init() {
    LR_CUSOLVER(cusolverDnCreate(&solver));
    LR_CUSOLVER(cusolverDnSetStream(solver, engine().stream));
    LR_CUDA(info.alloc(2));
    return LR_OK;
  }
  lr_status check(const char *what, const char *routine) {
    int h = 0;
    LR_CUDA(cudaMemcpyAsync(&h, info.p, sizeof(int), cudaMemcpyDeviceToHost, engine().stream));
    LR_CUDA(cudaStreamSynchronize(engine().stream));
    if (h != 0) return fail(LR_ERR_NUMERIC, "%s: %s failed (info %d: not positive definite / no convergence)", what, routine, h);
    return LR_OK;
  }
  lr_status potrf(int n, double *A, const char *what) {
    int lwork = 0;
    LR_CUSOLVER(cusolverDnDpotrf_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork));
    if ((size_t)lwork > work.n) LR_CUDA(work.alloc(lwork));
    LR_CUSOLVER(cusolverDnDpotrf(solver, CUBLAS_FILL_MODE_LOWER, n, A, n, work.p, lwork, info.p));
    count_launch();
    return check(what, "Cholesky");
  }
  // A <- A^-1 for symmetric positive definite A (full symmetric result)
  lr_status spd_inverse(int n, double *A, const char *what) {
    lr_status st = potrf(n, A, what);
    if (st != LR_OK) return st;
    int lwork = 0;
    LR_CUSOLVER(cusolverDnDpotri_bufferSize(solver, CUBLAS_FILL_MODE_LOWER, n, A, n, &lwork));
    if ((size_t)lwork > work.n) LR_CUDA(work.alloc(lwork));
    LR_CUSOLVER(cusolverDnDpotri(solver, CUBLAS_FILL_MODE_LOWER, n, A, n, work.p, lwork, info.p));
    count_launch();
    k_sym_from_lower<<<ceil_div((long)n * n, kThreads), kThreads, 0, engine().stream>>>(n, A);
    LR_CHECK_LAUNCH();
    return LR_OK;
  }
  // eigen-decomposition of symmetric A (overwritten by the eigenvectors, ascending eigenvalues)
  lr_status syevd(int n, double *A, double *lam, const char *what) {
    int lwork = 0;
    LR_CUSOLVER(cusolverDnDsyevd_bufferSize(solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A,
                                            n, lam, &lwork));
    if ((size_t)lwork > work.n) LR_CUDA(work.alloc(lwork));
    LR_CUSOLVER(cusolverDnDsyevd(solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, A, n, lam,
                                 work.p, lwork, info.p));
    count_launch();
    return check(what, "eigen-decomposition");
  }
  // generalised symmetric-definite problem A x = lambda B x (A overwritten by the eigenvectors)
  lr_status sygvd(int n, double *A, double *B, double *lam, const char *what) {
    int lwork = 0;
    LR_CUSOLVER(cusolverDnDsygvd_bufferSize(solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR,
                                            CUBLAS_FILL_MODE_LOWER, n, A, n, B, n, lam, &lwork));
    if ((size_t)lwork > work.n) LR_CUDA(work.alloc(lwork));
    LR_CUSOLVER(cusolverDnDsygvd(solver, CUSOLVER_EIG_TYPE_1, CUSOLVER_EIG_MODE_VECTOR,
                                 CUBLAS_FILL_MODE_LOWER, n, A, n, B, n, lam, work.p, lwork, info.p));
    count_launch();
    return check(what, "generalised eigen-decomposition");
  }
};

// C[m x n] (row-major) = op(A) op(B) on row-major operands
lr_status gemm_rm(bool ta, bool tb, int m, int n, int k, const double *A, int lda, const double *B,
                  int ldb, double *C, int ldc, double alpha = 1.0, double beta = 0.0) {
  LR_CUBLAS(cublasDgemm(engine().blas, tb ? CUBLAS_OP_T : CUBLAS_OP_N, ta ? CUBLAS_OP_T : CUBLAS_OP_N,
                        n, m, k, &alpha, B, ldb, A, lda, &beta, C, ldc));
  count_launch();
  return LR_OK;
}

// S[d x d] = scale * Y Y^T for row-major Y[d x n]: DSYRK on the column-major view (n x d, ld n),
// then the missing triangle is mirrored so the result is exactly symmetric like the reference's.
lr_status outer_rm(int d, size_t n, const double *Y, double scale, double *S) {
  Engine &e = engine();
  const double zero = 0.0;
  LR_CUBLAS(cublasDsyrk(e.blas, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, d, (int)n, &scale, Y, (int)n, &zero, S, d));
  count_launch();
  k_sym_from_lower<<<ceil_div((long)d * d, kThreads), kThreads, 0, e.stream>>>(d, S);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

// development-set state on the device
struct DevSet {
  int d = 0;
  size_t n = 0, n_spk = 0;
  DevBuf<double> X, Y, mu, sm, cnt;
  DevBuf<int> cls;
  lr_status load(int d_, size_t n_, const double *data, const int32_t *class_of, size_t n_spk_) {
    Engine &e = engine();
    d = d_;
    n = n_;
    n_spk = n_spk_;
    for (size_t s = 0; s < n; s++)
      if (class_of[s] < 0 || (size_t)class_of[s] >= n_spk)
        return fail(LR_ERR_ARG, "class_of[%zu] = %d outside [0, %zu)", s, class_of[s], n_spk);
    LR_CUDA(X.alloc((size_t)d * n));
    LR_CUDA(Y.alloc((size_t)d * std::max(n, n_spk)));
    LR_CUDA(mu.alloc(d));
    LR_CUDA(sm.alloc((size_t)d * n_spk));
    LR_CUDA(cnt.alloc(n_spk));
    LR_CUDA(cls.alloc(n));
    LR_CUDA(cudaMemcpyAsync(X.p, data, (size_t)d * n * sizeof(double), cudaMemcpyHostToDevice, e.stream));
    LR_CUDA(cudaMemcpyAsync(cls.p, class_of, n * sizeof(int), cudaMemcpyHostToDevice, e.stream));
    // computeAll (:353-385)
    k_row_sum<<<d, kThreads, 0, e.stream>>>(n, X.p, 1.0 / (double)n, mu.p);
    LR_CHECK_LAUNCH();
    LR_CUDA(cudaMemsetAsync(sm.p, 0, (size_t)d * n_spk * sizeof(double), e.stream));
    LR_CUDA(cudaMemsetAsync(cnt.p, 0, n_spk * sizeof(double), e.stream));
    k_class_sums<<<grid_for((size_t)d * n), kThreads, 0, e.stream>>>(d, n, n_spk, X.p, cls.p, sm.p, cnt.p);
    LR_CHECK_LAUNCH();
    k_class_div<<<grid_for((size_t)d * n_spk), kThreads, 0, e.stream>>>(d, n_spk, cnt.p, sm.p);
    LR_CHECK_LAUNCH();
    return LR_OK;
  }
  // within-class scatter of the sessions about their speaker mean; wccn: each session weighted by
  // 1 / sessions-of-its-speaker and the sum divided by the speaker count (:1124-1165), else / n
  lr_status within(bool wccn, double *W) {
    Engine &e = engine();
    k_center<<<grid_for((size_t)d * n), kThreads, 0, e.stream>>>(d, n, n_spk, X.p, mu.p, sm.p, cls.p, cnt.p, 1,
                                                                wccn ? 1 : 0, Y.p);
    LR_CHECK_LAUNCH();
    return outer_rm(d, n, Y.p, wccn ? 1.0 / (double)n_spk : 1.0 / (double)n, W);
  }
};

lr_status copy_out(double *host, const double *dev, size_t count) {
  if (!host) return LR_OK;
  LR_CUDA(cudaMemcpyAsync(host, dev, count * sizeof(double), cudaMemcpyDeviceToHost, engine().stream));
  return LR_OK;
}

// shared scoring driver: cross[nm x nt] = models^T P with P = Mx segments (Mx may be NULL = identity),
// finished block by block of models so the device score block stays below ~1 GB
lr_status score_blocks(int d, size_t nm, size_t nt, const double *d_models, const double *d_P,
                       const double *d_a, const double *d_b, const unsigned char *trials, int mode,
                       double *scores) {
  Engine &e = engine();
  const size_t mb_max = std::max<size_t>(1, std::min<size_t>(nm, ((size_t)1 << 27) / std::max<size_t>(nt, 1)));
  DevBuf<double> S;
  DevBuf<unsigned char> d_tr;
  LR_CUDA(S.alloc(mb_max * nt));
  if (trials) {
    LR_CUDA(d_tr.alloc(nm * nt));
    LR_CUDA(cudaMemcpyAsync(d_tr.p, trials, nm * nt, cudaMemcpyHostToDevice, e.stream));
  }
  for (size_t m0 = 0; m0 < nm; m0 += mb_max) {
    const size_t mb = std::min(mb_max, nm - m0);
    // S[mb x nt] = models[:, m0:m0+mb]^T P : row-major A = models block viewed [d x mb] (ld nm), transposed
    lr_status st = gemm_rm(true, false, (int)mb, (int)nt, d, d_models + m0, (int)nm, d_P, (int)nt, S.p, (int)nt);
    if (st != LR_OK) return st;
    k_finish_scores<<<grid_for(mb * nt), kThreads, 0, e.stream>>>(mb, nt, m0, d_a, d_b, trials ? d_tr.p : nullptr,
                                                                  mode, S.p);
    LR_CHECK_LAUNCH();
    LR_CUDA(cudaMemcpyAsync(scores + m0 * nt, S.p, mb * nt * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
    LR_CUDA(cudaStreamSynchronize(e.stream));
  }
  return LR_OK;
}

}  // namespace
}  // namespace lr

using namespace lr;

extern "C" {

lr_status lr_iv_cov_mat(int d, size_t n, const double *data, const int32_t *class_of, size_t n_spk,
                        double *mean, double *spk_means, double *Sigma, double *W, double *B) {
  LR_READY();
  LR_REQUIRE(d >= 1 && n >= 1 && n_spk >= 1 && data && class_of, "lr_iv_cov_mat: bad arguments");
  Engine &e = engine();
  DevSet ds;
  lr_status st = ds.load(d, n, data, class_of, n_spk);
  if (st != LR_OK) return st;
  DevBuf<double> M;
  LR_CUDA(M.alloc((size_t)d * d));
  st = copy_out(mean, ds.mu.p, d);
  if (st == LR_OK) st = copy_out(spk_means, ds.sm.p, (size_t)d * n_spk);
  if (st != LR_OK) return st;
  if (Sigma) {  // total covariance (:537)
    k_center<<<grid_for((size_t)d * n), kThreads, 0, e.stream>>>(d, n, n_spk, ds.X.p, ds.mu.p, ds.sm.p, ds.cls.p,
                                                                ds.cnt.p, 0, 0, ds.Y.p);
    LR_CHECK_LAUNCH();
    st = outer_rm(d, n, ds.Y.p, 1.0 / (double)n, M.p);
    if (st == LR_OK) st = copy_out(Sigma, M.p, (size_t)d * d);
    if (st != LR_OK) return st;
    LR_CUDA(cudaStreamSynchronize(e.stream));
  }
  if (W) {  // within-class covariance (:538)
    st = ds.within(false, M.p);
    if (st == LR_OK) st = copy_out(W, M.p, (size_t)d * d);
    if (st != LR_OK) return st;
    LR_CUDA(cudaStreamSynchronize(e.stream));
  }
  if (B) {  // between-class covariance (:541-543)
    k_between<<<grid_for((size_t)d * n_spk), kThreads, 0, e.stream>>>(d, n_spk, ds.sm.p, ds.mu.p, ds.cnt.p, ds.Y.p);
    LR_CHECK_LAUNCH();
    st = outer_rm(d, n_spk, ds.Y.p, 1.0 / (double)n, M.p);
    if (st == LR_OK) st = copy_out(B, M.p, (size_t)d * d);
    if (st != LR_OK) return st;
  }
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_iv_wccn_chol(int d, size_t n, const double *data, const int32_t *class_of, size_t n_spk,
                          double *WCCN) {
  LR_READY();
  LR_REQUIRE(d >= 1 && n >= 1 && n_spk >= 1 && data && class_of && WCCN, "lr_iv_wccn_chol: bad arguments");
  Engine &e = engine();
  DevSet ds;
  Dense dn;
  lr_status st = ds.load(d, n, data, class_of, n_spk);
  if (st == LR_OK) st = dn.init();
  if (st != LR_OK) return st;
  DevBuf<double> M;
  LR_CUDA(M.alloc((size_t)d * d));
  st = ds.within(true, M.p);
  if (st == LR_OK) st = dn.spd_inverse(d, M.p, "WCCN within-class covariance");
  // upperCholesky(invW): the column-major LOWER factor of the buffer is the row-major UPPER one
  if (st == LR_OK) st = dn.potrf(d, M.p, "inverse WCCN covariance");
  if (st != LR_OK) return st;
  k_keep_upper_rm<<<ceil_div((long)d * d, kThreads), kThreads, 0, e.stream>>>(d, M.p);
  LR_CHECK_LAUNCH();
  st = copy_out(WCCN, M.p, (size_t)d * d);
  if (st != LR_OK) return st;
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_iv_mahalanobis_matrix(int d, size_t n, const double *data, const int32_t *class_of,
                                   size_t n_spk, double *M) {
  LR_READY();
  LR_REQUIRE(d >= 1 && n >= 1 && n_spk >= 1 && data && class_of && M, "lr_iv_mahalanobis_matrix: bad arguments");
  DevSet ds;
  Dense dn;
  lr_status st = ds.load(d, n, data, class_of, n_spk);
  if (st == LR_OK) st = dn.init();
  if (st != LR_OK) return st;
  DevBuf<double> W;
  LR_CUDA(W.alloc((size_t)d * d));
  st = ds.within(false, W.p);
  if (st == LR_OK) st = dn.spd_inverse(d, W.p, "Mahalanobis within-class covariance");
  if (st == LR_OK) st = copy_out(M, W.p, (size_t)d * d);
  if (st != LR_OK) return st;
  LR_CUDA(cudaStreamSynchronize(engine().stream));
  return LR_OK;
}

lr_status lr_iv_efr_matrix(int d, const double *cov, double *mat) {
  LR_READY();
  LR_REQUIRE(d >= 1 && cov && mat, "lr_iv_efr_matrix: bad arguments");
  Engine &e = engine();
  Dense dn;
  lr_status st = dn.init();
  if (st != LR_OK) return st;
  DevBuf<double> V, lam, out;
  LR_CUDA(V.alloc((size_t)d * d));
  LR_CUDA(lam.alloc(d));
  LR_CUDA(out.alloc((size_t)d * d));
  LR_CUDA(cudaMemcpyAsync(V.p, cov, (size_t)d * d * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  st = dn.syevd(d, V.p, lam.p, "EFR covariance");
  if (st != LR_OK) return st;
  LR_CUDA(cudaMemsetAsync(dn.info.p + 1, 0, sizeof(int), e.stream));
  k_eig_rows<<<d, kThreads, 0, e.stream>>>(d, d, V.p, lam.p, 1, out.p, dn.info.p + 1);
  LR_CHECK_LAUNCH();
  int h = 0;
  LR_CUDA(cudaMemcpyAsync(&h, dn.info.p + 1, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaMemcpyAsync(mat, out.p, (size_t)d * d * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  if (h != 0) return fail(LR_ERR_NUMERIC, "lr_iv_efr_matrix: eigenvalue %d of the covariance is not positive", h - 1);
  return LR_OK;
}

lr_status lr_iv_lda(int d, const double *W, const double *B, int rank, double *ldaMat) {
  LR_READY();
  LR_REQUIRE(d >= 1 && rank >= 1 && rank <= d && W && B && ldaMat, "lr_iv_lda: bad arguments");
  Engine &e = engine();
  Dense dn;
  lr_status st = dn.init();
  if (st != LR_OK) return st;
  DevBuf<double> A, Bw, lam, out;
  LR_CUDA(A.alloc((size_t)d * d));
  LR_CUDA(Bw.alloc((size_t)d * d));
  LR_CUDA(lam.alloc(d));
  LR_CUDA(out.alloc((size_t)rank * d));
  // eigenvectors of W^-1 B  <=>  B v = lambda W v (W symmetric positive definite)
  LR_CUDA(cudaMemcpyAsync(A.p, B, (size_t)d * d * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(Bw.p, W, (size_t)d * d * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  st = dn.sygvd(d, A.p, Bw.p, lam.p, "LDA (within-class covariance)");
  if (st != LR_OK) return st;
  k_eig_rows<<<rank, kThreads, 0, e.stream>>>(d, rank, A.p, lam.p, 0, out.p, dn.info.p + 1);
  LR_CHECK_LAUNCH();
  LR_CUDA(cudaMemcpyAsync(ldaMat, out.p, (size_t)rank * d * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_iv_normalize(int d, size_t n, const double *data, const double *mu, const double *M, int r,
                          int length_norm, double *out) {
  LR_READY();
  LR_REQUIRE(d >= 1 && n >= 1 && data && out && (!M || r >= 1), "lr_iv_normalize: bad arguments");
  Engine &e = engine();
  const int dout = M ? r : d;
  DevBuf<double> X, Y, dmu, dM, nrm;
  LR_CUDA(X.alloc((size_t)d * n));
  LR_CUDA(cudaMemcpyAsync(X.p, data, (size_t)d * n * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  if (mu) {  // center (:466-474 / :3754-3767)
    LR_CUDA(dmu.alloc(d));
    LR_CUDA(cudaMemcpyAsync(dmu.p, mu, d * sizeof(double), cudaMemcpyHostToDevice, e.stream));
    k_sub_mu<<<grid_for((size_t)d * n), kThreads, 0, e.stream>>>(d, n, dmu.p, X.p);
    LR_CHECK_LAUNCH();
  }
  double *cur = X.p;
  if (M) {  // rotateLeft (:498-514 / :3770-3790): out[r x n] = M[r x d] X[d x n]
    LR_CUDA(dM.alloc((size_t)r * d));
    LR_CUDA(Y.alloc((size_t)r * n));
    LR_CUDA(cudaMemcpyAsync(dM.p, M, (size_t)r * d * sizeof(double), cudaMemcpyHostToDevice, e.stream));
    lr_status st = gemm_rm(false, false, r, (int)n, d, dM.p, d, X.p, (int)n, Y.p, (int)n);
    if (st != LR_OK) return st;
    cur = Y.p;
  }
  if (length_norm) {  // lengthNorm (:436-464 / :3706-3750)
    LR_CUDA(nrm.alloc(n));
    k_col_norm<<<ceil_div((long)n, kThreads), kThreads, 0, e.stream>>>(dout, n, cur, nrm.p);
    LR_CHECK_LAUNCH();
    k_div_cols<<<grid_for((size_t)dout * n), kThreads, 0, e.stream>>>(dout, n, nrm.p, cur);
    LR_CHECK_LAUNCH();
  }
  LR_CUDA(cudaMemcpyAsync(out, cur, (size_t)dout * n * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

static lr_status upload_pair(int d, size_t nm, size_t nt, const double *models, const double *segments,
                             DevBuf<double> &dM, DevBuf<double> &dS) {
  Engine &e = engine();
  LR_CUDA(dM.alloc((size_t)d * nm));
  LR_CUDA(dS.alloc((size_t)d * nt));
  LR_CUDA(cudaMemcpyAsync(dM.p, models, (size_t)d * nm * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(dS.p, segments, (size_t)d * nt * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  return LR_OK;
}

lr_status lr_iv_cosine_scoring(int d, size_t nm, size_t nt, const double *models, const double *segments,
                               const uint8_t *trials, double *scores) {
  LR_READY();
  LR_REQUIRE(d >= 1 && nm >= 1 && nt >= 1 && models && segments && scores, "lr_iv_cosine_scoring: bad arguments");
  Engine &e = engine();
  DevBuf<double> dM, dS, a, b;
  lr_status st = upload_pair(d, nm, nt, models, segments, dM, dS);
  if (st != LR_OK) return st;
  LR_CUDA(a.alloc(nm));
  LR_CUDA(b.alloc(nt));
  k_col_norm<<<ceil_div((long)nm, kThreads), kThreads, 0, e.stream>>>(d, nm, dM.p, a.p);
  LR_CHECK_LAUNCH();
  k_col_norm<<<ceil_div((long)nt, kThreads), kThreads, 0, e.stream>>>(d, nt, dS.p, b.p);
  LR_CHECK_LAUNCH();
  return score_blocks(d, nm, nt, dM.p, dS.p, a.p, b.p, trials, 0, scores);
}

lr_status lr_iv_mahalanobis_scoring(int d, size_t nm, size_t nt, const double *models, const double *segments,
                                    const double *Mah, const uint8_t *trials, double *scores) {
  LR_READY();
  LR_REQUIRE(d >= 1 && nm >= 1 && nt >= 1 && models && segments && Mah && scores,
             "lr_iv_mahalanobis_scoring: bad arguments");
  Engine &e = engine();
  DevBuf<double> dM, dS, a, b, dA, dA2, P, Q;
  lr_status st = upload_pair(d, nm, nt, models, segments, dM, dS);
  if (st != LR_OK) return st;
  const size_t dd = (size_t)d * d;
  LR_CUDA(a.alloc(nm));
  LR_CUDA(b.alloc(nt));
  LR_CUDA(dA.alloc(dd));
  LR_CUDA(dA2.alloc(dd));
  LR_CUDA(P.alloc((size_t)d * nt));
  LR_CUDA(Q.alloc((size_t)d * nm));
  LR_CUDA(cudaMemcpyAsync(dA.p, Mah, dd * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  // -0.5 (m - s)^T Mah (m - s) = -0.5 [ m^T Mah m + s^T Mah s - m^T (Mah + Mah^T) s ]
  st = gemm_rm(false, false, d, (int)nm, d, dA.p, d, dM.p, (int)nm, Q.p, (int)nm);
  if (st != LR_OK) return st;
  k_col_dot<<<ceil_div((long)nm, kThreads), kThreads, 0, e.stream>>>(d, nm, dM.p, Q.p, a.p);
  LR_CHECK_LAUNCH();
  st = gemm_rm(false, false, d, (int)nt, d, dA.p, d, dS.p, (int)nt, P.p, (int)nt);
  if (st != LR_OK) return st;
  k_col_dot<<<ceil_div((long)nt, kThreads), kThreads, 0, e.stream>>>(d, nt, dS.p, P.p, b.p);
  LR_CHECK_LAUNCH();
  k_axpby<<<ceil_div((long)dd, kThreads), kThreads, 0, e.stream>>>(d, 1.0, dA.p, 1.0, dA.p, 1, dA2.p);
  LR_CHECK_LAUNCH();
  st = gemm_rm(false, false, d, (int)nt, d, dA2.p, d, dS.p, (int)nt, P.p, (int)nt);
  if (st != LR_OK) return st;
  return score_blocks(d, nm, nt, dM.p, P.p, a.p, b.p, trials, 1, scores);
}

lr_status lr_iv_two_cov_scoring(int d, size_t nm, size_t nt, const double *models, const double *segments,
                                const double *W, const double *B, double *scores) {
  LR_READY();
  LR_REQUIRE(d >= 1 && nm >= 1 && nt >= 1 && models && segments && W && B && scores,
             "lr_iv_two_cov_scoring: bad arguments");
  Engine &e = engine();
  Dense dn;
  lr_status st = dn.init();
  if (st != LR_OK) return st;
  DevBuf<double> dM, dS, a, b, iW, iB, sG, sH, T1, G, H, GH, P, Q;
  st = upload_pair(d, nm, nt, models, segments, dM, dS);
  if (st != LR_OK) return st;
  const size_t dd = (size_t)d * d;
  for (DevBuf<double> *p : {&iW, &iB, &sG, &sH, &T1, &G, &H, &GH}) LR_CUDA(p->alloc(dd));
  LR_CUDA(a.alloc(nm));
  LR_CUDA(b.alloc(nt));
  LR_CUDA(P.alloc((size_t)d * nt));
  LR_CUDA(Q.alloc((size_t)d * nm));
  LR_CUDA(cudaMemcpyAsync(iW.p, W, dd * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(iB.p, B, dd * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  // G = W^-1 (B^-1 + 2 W^-1)^-1 W^-1,  H = W^-1 (B^-1 + W^-1)^-1 W^-1   (:4088-4125)
  st = dn.spd_inverse(d, iW.p, "two-covariance W");
  if (st == LR_OK) st = dn.spd_inverse(d, iB.p, "two-covariance B");
  if (st != LR_OK) return st;
  const int g2 = ceil_div((long)dd, kThreads);
  k_axpby<<<g2, kThreads, 0, e.stream>>>(d, 1.0, iB.p, 2.0, iW.p, 0, sG.p);
  LR_CHECK_LAUNCH();
  k_axpby<<<g2, kThreads, 0, e.stream>>>(d, 1.0, iB.p, 1.0, iW.p, 0, sH.p);
  LR_CHECK_LAUNCH();
  st = dn.spd_inverse(d, sG.p, "two-covariance B^-1 + 2 W^-1");
  if (st == LR_OK) st = dn.spd_inverse(d, sH.p, "two-covariance B^-1 + W^-1");
  if (st == LR_OK) st = gemm_rm(false, false, d, d, d, iW.p, d, sG.p, d, T1.p, d);
  if (st == LR_OK) st = gemm_rm(false, false, d, d, d, T1.p, d, iW.p, d, G.p, d);
  if (st == LR_OK) st = gemm_rm(false, false, d, d, d, iW.p, d, sH.p, d, T1.p, d);
  if (st == LR_OK) st = gemm_rm(false, false, d, d, d, T1.p, d, iW.p, d, H.p, d);
  if (st != LR_OK) return st;
  // (m + s)^T G (m + s) - m^T H m - s^T H s = m^T (G - H) m + s^T (G - H) s + m^T (G + G^T) s
  k_axpby<<<g2, kThreads, 0, e.stream>>>(d, 1.0, G.p, -1.0, H.p, 0, GH.p);
  LR_CHECK_LAUNCH();
  st = gemm_rm(false, false, d, (int)nm, d, GH.p, d, dM.p, (int)nm, Q.p, (int)nm);
  if (st != LR_OK) return st;
  k_col_dot<<<ceil_div((long)nm, kThreads), kThreads, 0, e.stream>>>(d, nm, dM.p, Q.p, a.p);
  LR_CHECK_LAUNCH();
  st = gemm_rm(false, false, d, (int)nt, d, GH.p, d, dS.p, (int)nt, P.p, (int)nt);
  if (st != LR_OK) return st;
  k_col_dot<<<ceil_div((long)nt, kThreads), kThreads, 0, e.stream>>>(d, nt, dS.p, P.p, b.p);
  LR_CHECK_LAUNCH();
  k_axpby<<<g2, kThreads, 0, e.stream>>>(d, 1.0, G.p, 1.0, G.p, 1, T1.p);
  LR_CHECK_LAUNCH();
  st = gemm_rm(false, false, d, (int)nt, d, T1.p, d, dS.p, (int)nt, P.p, (int)nt);
  if (st != LR_OK) return st;
  return score_blocks(d, nm, nt, dM.p, P.p, a.p, b.p, nullptr, 2, scores);
}

// ---- PLDA EM training: one PldaModel::em_iteration (PldaTools.cpp:2329-2343) -------------------
// center by Delta, scatter matrix (computeCovMatEigen :931-950), E-step (getExpectedValues
// :2359-2485), M-step with minimum divergence (:2790-2813).  The reference walks the speakers one
// at a time with Eigen; here every per-speaker product is one GEMM over ALL sessions / speakers:
//   fi = F^T S^-1 X, gi = G^T S^-1 X (all sessions), f / g = per-speaker column sums,
//   E[h_spk] = V diag(1 / (n_spk lam + 1)) V^T (f - S^T g)        (M_n in the eigenbasis of A)
//   Eh = [E[h_spk] per session ; iGG gi - S E[h_spk]],  EhhSum = Eh Eh^T + sum_spk n_spk tmpM_{n_spk},
//   xhSum = X Eh^T, Umx = row sums of Eh.
lr_status lr_plda_em_iteration(int d, int rF, int rG, size_t n, double *data, const int32_t *class_of,
                               size_t n_spk, double *F, double *G, double *Sigma, double *Delta) {
  LR_READY();
  LR_REQUIRE(d >= 1 && rF >= 1 && rG >= 0 && rF + rG <= 1024 && n >= 1 && n_spk >= 1 && data && class_of && F &&
                 (G || rG == 0) && Sigma && Delta,
             "lr_plda_em_iteration: bad arguments (d=%d rF=%d rG=%d)", d, rF, rG);
  for (size_t s = 0; s < n; s++)
    LR_REQUIRE(class_of[s] >= 0 && (size_t)class_of[s] < n_spk && (s == 0 || class_of[s] >= class_of[s - 1]),
               "lr_plda_em_iteration: class_of must be non-decreasing in [0, %zu)", n_spk);
  Engine &e = engine();
  Dense dn;
  lr_status st = dn.init();
  if (st != LR_OK) return st;
  const int r = rF + rG;
  const size_t dd = (size_t)d * d;
  const double nn = (double)n;
  DevBuf<double> X, dF, dG, dS, dDelta, obs, iS, Ftw, Gtw, iGG, FtwG, Sm, A, lam, fi, gi, fs, gs, cnt, R, Z, eh, SE,
      low, Eh, Ehh, Ehh0, xh, U, w, T1, Mbar, MsT, SMsT, FG, SL, cm, cF, cG, Fn, Gn;
  DevBuf<int> cls;
#define ALLOC(buf, count) LR_CUDA(buf.alloc(std::max<size_t>((size_t)(count), 1)))
  ALLOC(X, (size_t)d * n); ALLOC(dF, (size_t)d * rF); ALLOC(dG, (size_t)d * rG); ALLOC(dS, dd); ALLOC(dDelta, d);
  ALLOC(obs, dd); ALLOC(iS, dd); ALLOC(Ftw, (size_t)rF * d); ALLOC(Gtw, (size_t)rG * d);
  ALLOC(iGG, (size_t)rG * rG); ALLOC(FtwG, (size_t)rF * rG); ALLOC(Sm, (size_t)rG * rF); ALLOC(A, (size_t)rF * rF);
  ALLOC(lam, rF); ALLOC(fi, (size_t)rF * n); ALLOC(gi, (size_t)rG * n); ALLOC(fs, (size_t)rF * n_spk);
  ALLOC(gs, (size_t)rG * n_spk); ALLOC(cnt, n_spk); ALLOC(R, (size_t)rF * n_spk); ALLOC(Z, (size_t)rF * n_spk);
  ALLOC(eh, (size_t)rF * n_spk); ALLOC(SE, (size_t)rG * n_spk); ALLOC(low, (size_t)rG * n); ALLOC(Eh, (size_t)r * n);
  ALLOC(Ehh, (size_t)r * r); ALLOC(Ehh0, (size_t)r * r); ALLOC(xh, (size_t)d * r); ALLOC(U, r); ALLOC(w, rF);
  ALLOC(T1, (size_t)rF * rF); ALLOC(Mbar, (size_t)rF * rF); ALLOC(MsT, (size_t)rF * rG); ALLOC(SMsT, (size_t)rG * rG);
  ALLOC(FG, (size_t)d * r); ALLOC(SL, dd); ALLOC(cm, (size_t)r * r); ALLOC(cF, (size_t)rF * rF);
  ALLOC(cG, (size_t)rG * rG); ALLOC(Fn, (size_t)d * rF); ALLOC(Gn, (size_t)d * rG); ALLOC(cls, n);
#undef ALLOC
  auto up = [&](double *dst, const double *src, size_t count) {
    return count ? cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyHostToDevice, e.stream) : cudaSuccess;
  };
  LR_CUDA(up(X.p, data, (size_t)d * n));
  LR_CUDA(up(dF.p, F, (size_t)d * rF));
  LR_CUDA(up(dG.p, G, (size_t)d * rG));
  LR_CUDA(up(dS.p, Sigma, dd));
  LR_CUDA(up(dDelta.p, Delta, d));
  LR_CUDA(cudaMemcpyAsync(cls.p, class_of, n * sizeof(int), cudaMemcpyHostToDevice, e.stream));
  auto g1 = [](size_t count) { return ceil_div((long)std::max<size_t>(count, 1), kThreads); };
#define GEMM(...)                   \
  do {                              \
    st = gemm_rm(__VA_ARGS__);      \
    if (st != LR_OK) return st;     \
  } while (0)
#define LAUNCHED() LR_CHECK_LAUNCH()
  // _Dev.center(_Delta); sigmaObs = X X^T
  k_sub_mu<<<grid_for((size_t)d * n), kThreads, 0, e.stream>>>(d, n, dDelta.p, X.p);
  LAUNCHED();
  st = outer_rm(d, n, X.p, 1.0, obs.p);
  if (st != LR_OK) return st;
  // preComputation (:2950-2972)
  LR_CUDA(cudaMemcpyAsync(iS.p, dS.p, dd * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  st = dn.spd_inverse(d, iS.p, "PLDA Sigma");
  if (st != LR_OK) return st;
  GEMM(true, false, rF, d, d, dF.p, rF, iS.p, d, Ftw.p, d);            // Ftw = F^T S^-1
  GEMM(false, false, rF, rF, d, Ftw.p, d, dF.p, rF, A.p, rF);          // A = Ftw F (- FtwG iGG FtwG^T below)
  if (rG > 0) {
    GEMM(true, false, rG, d, d, dG.p, rG, iS.p, d, Gtw.p, d);          // Gtw = G^T S^-1
    GEMM(false, false, rG, rG, d, Gtw.p, d, dG.p, rG, iGG.p, rG);      // GtwG
    k_add_diag<<<g1(rG), kThreads, 0, e.stream>>>(rG, 1.0, iGG.p);
    LAUNCHED();
    st = dn.spd_inverse(rG, iGG.p, "PLDA G^T S^-1 G + I");
    if (st != LR_OK) return st;
    GEMM(false, false, rF, rG, d, Ftw.p, d, dG.p, rG, FtwG.p, rG);     // FtwG
    GEMM(false, true, rG, rF, rG, iGG.p, rG, FtwG.p, rG, Sm.p, rF);    // S = iGG FtwG^T  [rG x rF]
    GEMM(false, false, rF, rF, rG, FtwG.p, rG, Sm.p, rF, A.p, rF, -1.0, 1.0);  // A -= FtwG S
  }
  // A = V diag(lam) V^T.  Column-major eigenvectors: the row-major view of the buffer is V^T.
  st = dn.syevd(rF, A.p, lam.p, "PLDA E-step matrix A");
  if (st != LR_OK) return st;
  const double *Vt = A.p;
  // per-session / per-speaker first-order terms
  GEMM(false, false, rF, (int)n, d, Ftw.p, d, X.p, (int)n, fi.p, (int)n);  // fi = F^T S^-1 X
  LR_CUDA(cudaMemsetAsync(fs.p, 0, (size_t)rF * n_spk * sizeof(double), e.stream));
  LR_CUDA(cudaMemsetAsync(cnt.p, 0, n_spk * sizeof(double), e.stream));
  k_class_colsum<<<grid_for((size_t)rF * n), kThreads, 0, e.stream>>>(rF, n, n_spk, fi.p, cls.p, fs.p);
  LAUNCHED();
  k_count_classes<<<g1(n), kThreads, 0, e.stream>>>(n, cls.p, cnt.p);
  LAUNCHED();
  LR_CUDA(cudaMemcpyAsync(R.p, fs.p, (size_t)rF * n_spk * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  if (rG > 0) {
    GEMM(false, false, rG, (int)n, d, Gtw.p, d, X.p, (int)n, gi.p, (int)n);  // gi = G^T S^-1 X
    LR_CUDA(cudaMemsetAsync(gs.p, 0, (size_t)rG * n_spk * sizeof(double), e.stream));
    k_class_colsum<<<grid_for((size_t)rG * n), kThreads, 0, e.stream>>>(rG, n, n_spk, gi.p, cls.p, gs.p);
    LAUNCHED();
    GEMM(true, false, rF, (int)n_spk, rG, Sm.p, rF, gs.p, (int)n_spk, R.p, (int)n_spk, -1.0, 1.0);  // R = f - S^T g
  }
  // E[h_spk] = V diag(1 / (n_spk lam + 1)) V^T R    (:2417-2424, :2453)
  GEMM(false, false, rF, (int)n_spk, rF, Vt, rF, R.p, (int)n_spk, Z.p, (int)n_spk);
  k_scale_posterior<<<grid_for((size_t)rF * n_spk), kThreads, 0, e.stream>>>(rF, n_spk, cnt.p, lam.p, Z.p);
  LAUNCHED();
  GEMM(true, false, rF, (int)n_spk, rF, Vt, rF, Z.p, (int)n_spk, eh.p, (int)n_spk);
  if (rG > 0) {
    GEMM(false, false, rG, (int)n_spk, rF, Sm.p, rF, eh.p, (int)n_spk, SE.p, (int)n_spk);  // S E[h_spk]
    GEMM(false, false, rG, (int)n, rG, iGG.p, rG, gi.p, (int)n, low.p, (int)n);            // iGG gi
  }
  k_build_eh<<<grid_for((size_t)r * n), kThreads, 0, e.stream>>>(rF, rG, n, n_spk, cls.p, eh.p, low.p, SE.p, Eh.p);
  LAUNCHED();
  // EhhSum = Eh Eh^T + sum_spk n_spk tmpM_{n_spk}   (:2467-2473), with Mbar = sum_spk n_spk M_{n_spk}
  st = outer_rm(r, n, Eh.p, 1.0, Ehh.p);
  if (st != LR_OK) return st;
  k_mbar_weights<<<rF, kThreads, 0, e.stream>>>(rF, n_spk, cnt.p, lam.p, w.p);
  LAUNCHED();
  k_scale_rows<<<g1((size_t)rF * rF), kThreads, 0, e.stream>>>(rF, rF, w.p, Vt, T1.p);
  LAUNCHED();
  GEMM(true, false, rF, rF, rF, Vt, rF, T1.p, rF, Mbar.p, rF);  // Mbar = V diag(w) V^T
  k_add_block<<<g1((size_t)rF * rF), kThreads, 0, e.stream>>>(Ehh.p, r, 0, 0, Mbar.p, rF, rF, rF, 1.0, 0);
  LAUNCHED();
  if (rG > 0) {
    GEMM(false, true, rF, rG, rF, Mbar.p, rF, Sm.p, rF, MsT.p, rG);   // Mbar S^T
    GEMM(false, false, rG, rG, rF, Sm.p, rF, MsT.p, rG, SMsT.p, rG);  // S Mbar S^T
    k_add_block<<<g1((size_t)rF * rG), kThreads, 0, e.stream>>>(Ehh.p, r, 0, rF, MsT.p, rG, rF, rG, -1.0, 0);
    LAUNCHED();
    k_add_block<<<g1((size_t)rF * rG), kThreads, 0, e.stream>>>(Ehh.p, r, rF, 0, MsT.p, rG, rG, rF, -1.0, 1);
    LAUNCHED();
    k_add_block<<<g1((size_t)rG * rG), kThreads, 0, e.stream>>>(Ehh.p, r, rF, rF, iGG.p, rG, rG, rG, nn, 0);
    LAUNCHED();
    k_add_block<<<g1((size_t)rG * rG), kThreads, 0, e.stream>>>(Ehh.p, r, rF, rF, SMsT.p, rG, rG, rG, 1.0, 0);
    LAUNCHED();
  }
  // xhSum = X Eh^T (:2476-2479), Umx = row sums of Eh (:2482-2483)
  GEMM(false, true, d, r, (int)n, X.p, (int)n, Eh.p, (int)n, xh.p, r);
  k_row_sum<<<r, kThreads, 0, e.stream>>>(n, Eh.p, 1.0, U.p);
  LAUNCHED();
  // ---- M-step (:2790-2813)
  LR_CUDA(cudaMemcpyAsync(Ehh0.p, Ehh.p, (size_t)r * r * sizeof(double), cudaMemcpyDeviceToDevice, e.stream));
  st = dn.spd_inverse(r, Ehh.p, "PLDA EhhSum");
  if (st != LR_OK) return st;
  GEMM(false, false, d, r, r, xh.p, r, Ehh.p, r, FG.p, r);       // FGEst = xhSum EhhSum^-1
  GEMM(false, true, d, d, r, FG.p, r, xh.p, r, SL.p, d);         // SigmaLat = FGEst xhSum^T
  k_sigma_update<<<g1(dd), kThreads, 0, e.stream>>>((int)dd, nn, obs.p, SL.p, dS.p);
  LAUNCHED();
  k_mindiv_cov<<<1, 1024, 0, e.stream>>>(r, nn, Ehh0.p, U.p, cm.p);  // U <- Umx / n, c = EhhSum / n - u u^T
  LAUNCHED();
  // Rh = upper Cholesky factor of c[0:rF, 0:rF]; F = FGEst[:, 0:rF] Rh^T.  potrf LOWER on the
  // column-major buffer leaves, read row-major, exactly Rh in the upper triangle.
  k_copy_block_rm<<<g1((size_t)rF * rF), kThreads, 0, e.stream>>>(cm.p, r, 0, 0, rF, rF, cF.p);
  LAUNCHED();
  st = dn.potrf(rF, cF.p, "PLDA minimum-divergence covariance (speaker factors)");
  if (st != LR_OK) return st;
  k_keep_upper_rm<<<g1((size_t)rF * rF), kThreads, 0, e.stream>>>(rF, cF.p);
  LAUNCHED();
  GEMM(false, true, d, rF, rF, FG.p, r, cF.p, rF, Fn.p, rF);
  if (rG > 0) {
    k_copy_block_rm<<<g1((size_t)rG * rG), kThreads, 0, e.stream>>>(cm.p, r, rF, rF, rG, rG, cG.p);
    LAUNCHED();
    st = dn.potrf(rG, cG.p, "PLDA minimum-divergence covariance (channel factors)");
    if (st != LR_OK) return st;
    k_keep_upper_rm<<<g1((size_t)rG * rG), kThreads, 0, e.stream>>>(rG, cG.p);
    LAUNCHED();
    GEMM(false, true, d, rG, rG, FG.p + rF, r, cG.p, rG, Gn.p, rG);
  }
  // Delta += FGEst Umx
  {
    const double one = 1.0;
    LR_CUBLAS(cublasDgemv(e.blas, CUBLAS_OP_T, r, d, &one, FG.p, r, U.p, 1, &one, dDelta.p, 1));
    count_launch();
  }
#undef GEMM
#undef LAUNCHED
  auto down = [&](double *dst, const double *src, size_t count) {
    return count ? cudaMemcpyAsync(dst, src, count * sizeof(double), cudaMemcpyDeviceToHost, e.stream) : cudaSuccess;
  };
  LR_CUDA(down(data, X.p, (size_t)d * n));
  LR_CUDA(down(F, Fn.p, (size_t)d * rF));
  LR_CUDA(down(G, Gn.p, (size_t)d * rG));
  LR_CUDA(down(Sigma, dS.p, dd));
  LR_CUDA(down(Delta, dDelta.p, d));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

}  // extern "C"
