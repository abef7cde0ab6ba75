// tv.cu -- device twin of the reference's TVAcc object (LIA_SpkTools/src/AccumulateTVStat.cpp):
// Baum-Welch statistics in HBM, i-vector posterior solve, T-matrix EM accumulators and M-step.
//
// Round-1 formulation: every triple loop of the reference is restated as a dense fp64 GEMM
// over a batch of utterances (cuBLAS on the fp64 tensor pipe), the per-utterance R x R systems
// are factorised with a blocked batched Cholesky (L = I + sum_c N_c TETt_c is SPD), and the glue
// (centring, rank-1 updates, reductions, triangle packing) is hand-written kernels.
// Symmetric R x R quantities that only ever enter a GEMM as a flat vector -- TETt_c, L_s, E_s,
// A_c -- are held as PACKED lower triangles (R (R + 1) / 2 doubles), which halves the two
// dominant GEMMs (L = N TETt, A += N^T E) and the E-step all-reduce.  Layout notes use
// "row-major X[a x b]" for the reference's Matrix<double> buffers; cuBLAS sees the same
// memory as the column-major transpose.
#include <cusolverDn.h>

#include <algorithm>

#include "common.cuh"
#include "gemm_i8.cuh"

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
  double *d_tettp = nullptr;  // [C x Rp] packed lower triangles of TETt_c
  double *d_acc = nullptr;    // [A C*Rp (packed lower triangles) | Cmx R*sv | Rm R*R | r R | sumW R]
  double *d_meanW = nullptr;
  double *d_Lb = nullptr, *d_Eb = nullptr;  // [batch x R*R] work
  double *d_Yb = nullptr;                   // [batch x R*R] triangular inverse (E-step)
  double *d_invD = nullptr;                 // [batch x nblk x 64 x 64] diagonal-block inverses
  double *d_ones = nullptr;                 // [max(batch, R)]
  double **d_ptr_L = nullptr, **d_ptr_E = nullptr, **d_ptr_W = nullptr;  // batch pointers
  double **d_ptr_A = nullptr, **d_ptr_Tc = nullptr;                      // component pointers
  int *d_info = nullptr;
  cusolverDnHandle_t solver = nullptr;
  // digit planes (gemm_i8.cu) of the two operands that only change with T: TETt^T [Rp x C] and Ts [R x sv]
  unsigned char *d_tett_planes = nullptr, *d_ts_planes = nullptr;
  double *d_tett_scale = nullptr, *d_ts_scale = nullptr;
  int planes = 0;  // digit planes the two were cut into (0: not prepared, cuBLAS path)
  // max |F[s, :]| per utterance (bit pattern of the double), a by-product of substractM: the digit GEMM's
  // row scale of the aux operand.  Valid until F is written again (set_stats, dev_F hand-out, normStatistics).
  unsigned long long *d_fmax = nullptr;
  bool fmax_valid = false;
  size_t Rp() const { return (size_t)R * (R + 1) / 2; }
  double *A() const { return d_acc; }  // packed: A_c at d_acc + c * Rp
  double *Cmx() const { return d_acc + (size_t)C * Rp(); }
  double *Rm() const { return Cmx() + (size_t)R * sv; }
  double *r() const { return Rm() + (size_t)R * R; }
  double *sumW() const { return r() + R; }
  size_t acc_len() const { return (size_t)C * Rp() + (size_t)R * sv + (size_t)R * R + 2 * (size_t)R; }
};

namespace lr {
namespace {

// substractM (AccumulateTVStat.cpp:1088-1105): F[s, c, :] -= mean[c, :] * N[s, c]; one block per
// (utterance, 8192-element segment), which also leaves max |F[s, :]| behind (rowmax, bit pattern of the double)
constexpr int kSubSeg = 8192;
__global__ void __launch_bounds__(256)
k_subtract_m(size_t U, int C, int D, const double *__restrict__ N, const double *__restrict__ mean,
             double *__restrict__ F, unsigned long long *__restrict__ rowmax) {
  const size_t sv = (size_t)C * D;
  const size_t k0 = (size_t)blockIdx.x * kSubSeg, k1 = min(sv, k0 + kSubSeg);
  for (size_t s = blockIdx.y; s < U; s += gridDim.y) {
    double m = 0.0;
    for (size_t k = k0 + threadIdx.x; k < k1; k += blockDim.x) {
      const double v = F[s * sv + k] - mean[k] * N[s * C + k / D];
      F[s * sv + k] = v;
      m = fmax(m, fabs(v));
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.0) atomicMax(rowmax + s, (unsigned long long)__double_as_longlong(m));
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

// Packed lower triangle of a symmetric R x R matrix: column j holds rows j..R-1 at
// off(j) = j R - j (j - 1) / 2 (the column-major lower triangle, columns concatenated).
__device__ __forceinline__ size_t packed_off(int R, int col) {
  return (size_t)col * R - (size_t)col * (col - 1) / 2;
}
// full[m] (column-major, ld R): lower triangle from the packed form (+ diag_add on the diagonal);
// the strict upper triangle is zeroed (sym == 0) or mirrored (sym != 0).
__global__ void k_unpack_lower(size_t n, int R, const double *__restrict__ packed,
                               double *__restrict__ full, double diag_add, int sym) {
  const size_t rr = (size_t)R * R, rp = (size_t)R * (R + 1) / 2, total = n * rr;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t m = i / rr, e = i - m * rr;
    int col = (int)(e / R), row = (int)(e - (size_t)col * R);
    double v = 0.0;
    if (row >= col)
      v = packed[m * rp + packed_off(R, col) + (row - col)] + (row == col ? diag_add : 0.0);
    else if (sym)
      v = packed[m * rp + packed_off(R, row) + (col - row)];
    full[i] = v;
  }
}

// estimateTETt (AccumulateTVStat.cpp:777-805): TETt_c[i, j] = sum_d (T o invvar)[i, cD + d] T[j, cD + d],
// written straight into the packed lower triangles [C x Rp] (only the block pairs bj <= bi are
// computed: half the reference's R x R loop, no full-matrix round trip).  One CTA per (64 x 64 tile
// pair, component); 4 x 4 outputs per thread, K chunks of 32 through shared memory.
constexpr int kTtTile = 64, kTtK = 32;
__global__ void __launch_bounds__(256)
k_tett_packed(int R, int D, size_t sv, const double *__restrict__ Ts, const double *__restrict__ T,
              double *__restrict__ out /*[C x Rp]*/) {
  __shared__ double As[kTtTile][kTtK + 1], Bs[kTtTile][kTtK + 1];
  const int c = blockIdx.y;
  int bi = (int)((sqrt(8.0 * blockIdx.x + 1.0) - 1.0) * 0.5);
  while ((bi + 1) * (bi + 2) / 2 <= (int)blockIdx.x) bi++;
  while (bi * (bi + 1) / 2 > (int)blockIdx.x) bi--;
  const int bj = blockIdx.x - bi * (bi + 1) / 2;
  const int i0 = bi * kTtTile, j0 = bj * kTtTile;
  const int ti = threadIdx.x & 15, tj = threadIdx.x >> 4;
  const int lr = threadIdx.x >> 2, lk = (threadIdx.x & 3) * 8;
  double acc[4][4] = {};
  for (int k0 = 0; k0 < D; k0 += kTtK) {
#pragma unroll
    for (int e = 0; e < 8; e++) {
      const int k = k0 + lk + e;
      const bool kin = k < D;
      As[lr][lk + e] = (kin && i0 + lr < R) ? Ts[(size_t)(i0 + lr) * sv + (size_t)c * D + k] : 0.0;
      Bs[lr][lk + e] = (kin && j0 + lr < R) ? T[(size_t)(j0 + lr) * sv + (size_t)c * D + k] : 0.0;
    }
    __syncthreads();
#pragma unroll 8
    for (int k = 0; k < kTtK; k++) {
      double a[4], b[4];
#pragma unroll
      for (int x = 0; x < 4; x++) {
        a[x] = As[ti + 16 * x][k];
        b[x] = Bs[tj * 4 + x][k];
      }
#pragma unroll
      for (int x = 0; x < 4; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) acc[x][y] = fma(a[x], b[y], acc[x][y]);
    }
    __syncthreads();
  }
  const size_t rp = (size_t)R * (R + 1) / 2;
#pragma unroll
  for (int y = 0; y < 4; y++) {
    const int j = j0 + tj * 4 + y;
    if (j >= R) continue;
    double *col = out + (size_t)c * rp + packed_off(R, j);
#pragma unroll
    for (int x = 0; x < 4; x++) {
      const int i = i0 + ti + 16 * x;
      if (i < R && i >= j) col[i - j] = acc[x][y];
    }
  }
}

// packed[b] = lower triangle of (E[b] + w_b w_b^T), E[b] column-major R x R with a valid lower triangle
// (Linv += y y^T, AccumulateTVStat.cpp:1766-1768, fused with the packing)
__global__ void k_pack_lower_rank1(size_t n, int R, const double *__restrict__ full, const double *__restrict__ W,
                                   double *__restrict__ packed) {
  const size_t rp = (size_t)R * (R + 1) / 2, rr = (size_t)R * R, total = n * rp;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    const size_t m = i / rp, e = i - m * rp;
    // column of the packed index e: largest col with col R - col (col - 1) / 2 <= e
    int col = (int)(((2.0 * R + 1.0) - sqrt((2.0 * R + 1.0) * (2.0 * R + 1.0) - 8.0 * (double)e)) * 0.5);
    while (col > 0 && packed_off(R, col) > e) col--;
    while (col + 1 < R && packed_off(R, col + 1) <= e) col++;
    const int row = col + (int)(e - packed_off(R, col));
    packed[i] = full[m * rr + (size_t)col * R + row] + W[m * R + row] * W[m * R + col];
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
// Blocked batched Cholesky (cusolverDnDpotrfBatched runs at ~1.4 TFLOP/s for R = 400..600: measured
// 52 us per 600 x 600 matrix, profiles/r01_tv_breakdown.md): left-looking over 64-wide block columns,
// one CTA per matrix for the whole factorization (k_chol_fused below).  The inverses of the diagonal
// blocks are kept: the blocked solve, the E-step's explicit inverse and the M-step reuse them.
constexpr int kNB = 64;
// Batched fp64 GEMM on the DMMA pipe (the M-step's block substitutions; the tile loop of k_chol_fused):
//     C[b] = alpha A[b] B[b]^T + beta C[b]        (all column-major: A m x K, B n x K, C m x n)
// One CTA per 64 x 64 tile of one matrix of the batch; 8 warps, each a 16 x 32 warp tile of eight
// mma.m8n8k4 accumulators; K in slabs of 16 through shared memory, the next slab fetched into registers
// while the current one is multiplied (k-major slabs, leading dimension
// 72 doubles: the four k rows a fragment load touches fall on two disjoint bank halves).  The tile is
// held in registers until the end, so C may alias A when a CTA's rows of A are read by nobody else
// (the panel solve X = P invD^T overwrites P).
constexpr int kBgTile = 64, kBgSlab = 16, kBgLd = 72;
__device__ __forceinline__ void dmma_8x8x4(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1)
               : "d"(a), "d"(b));
}
// AT / BT: the operand is given TRANSPOSED in memory (k contiguous instead of m / n):
//   !AT: A(m, k) at A[k lda + m]      AT: A(m, k) at A[m lda + k]      (same for B with n)
template <bool AT, bool BT>
__global__ void __launch_bounds__(256)
k_bgemm(int M, int N, int K, double alpha, const double *A, int lda, size_t strideA, const double *B, int ldb,
        size_t strideB, double beta, double *C, int ldc, size_t strideC) {
  __shared__ double As[kBgSlab][kBgLd], Bs[kBgSlab][kBgLd];
  const int m0 = blockIdx.x * kBgTile, n0 = blockIdx.y * kBgTile;
  A += (size_t)blockIdx.z * strideA;
  B += (size_t)blockIdx.z * strideB;
  C += (size_t)blockIdx.z * strideC;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;  // warp tile origin inside the CTA tile
  const int fr = lane >> 2, fk = lane & 3;                // fragment row (m or n) and k
  // slab loaders: 4 consecutive elements along the contiguous index per thread
  const int lk = threadIdx.x >> 4, l4 = (threadIdx.x & 15) * 4;  // plain: k row, first of 4 m / n
  const int tr = threadIdx.x >> 2, tk = (threadIdx.x & 3) * 4;   // transposed: m / n row, first of 4 k
  double acc[2][4][2] = {};
  double ra[4], rb[4];  // next slab, fetched while the current one is multiplied
  auto fetch = [&](int k0) {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      if (AT) {
        const int m = m0 + tr, k = k0 + tk + e;
        ra[e] = (k < K && m < M) ? A[(size_t)m * lda + k] : 0.0;
      } else {
        const int m = m0 + l4 + e, k = k0 + lk;
        ra[e] = (k < K && m < M) ? A[(size_t)k * lda + m] : 0.0;
      }
      if (BT) {
        const int n = n0 + tr, k = k0 + tk + e;
        rb[e] = (k < K && n < N) ? B[(size_t)n * ldb + k] : 0.0;
      } else {
        const int n = n0 + l4 + e, k = k0 + lk;
        rb[e] = (k < K && n < N) ? B[(size_t)k * ldb + n] : 0.0;
      }
    }
  };
  fetch(0);
  for (int k0 = 0; k0 < K; k0 += kBgSlab) {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      if (AT) As[tk + e][tr] = ra[e];
      else As[lk][l4 + e] = ra[e];
      if (BT) Bs[tk + e][tr] = rb[e];
      else Bs[lk][l4 + e] = rb[e];
    }
    __syncthreads();
    if (k0 + kBgSlab < K) fetch(k0 + kBgSlab);
#pragma unroll
    for (int kk = 0; kk < kBgSlab; kk += 4) {
      double a[2], b[4];
#pragma unroll
      for (int x = 0; x < 2; x++) a[x] = As[kk + fk][wm + 8 * x + fr];
#pragma unroll
      for (int y = 0; y < 4; y++) b[y] = Bs[kk + fk][wn + 8 * y + fr];
#pragma unroll
      for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 4; y++) dmma_8x8x4(acc[x][y][0], acc[x][y][1], a[x], b[y]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int x = 0; x < 2; x++) {
    const int m = m0 + wm + 8 * x + fr;
    if (m >= M) continue;
#pragma unroll
    for (int y = 0; y < 4; y++)
#pragma unroll
      for (int z = 0; z < 2; z++) {
        const int n = n0 + wn + 8 * y + 2 * fk + z;
        if (n >= N) continue;
        double *dst = C + (size_t)n * ldc + m;
        *dst = alpha * acc[x][y][z] + (beta != 0.0 ? beta * *dst : 0.0);
      }
  }
}
// C = alpha op(A) op(B)^T + beta C over a batch; ta / tb: operand stored with k contiguous
inline lr_status bgemm(bool ta, bool tb, int M, int N, int K, double alpha, const double *A, int lda, size_t sA,
                       const double *B, int ldb, size_t sB, double beta, double *C, int ldc, size_t sC, int batch) {
  if (M <= 0 || N <= 0 || batch <= 0) return LR_OK;
  dim3 grid((unsigned)ceil_div(M, kBgTile), (unsigned)ceil_div(N, kBgTile), (unsigned)batch);
  cudaStream_t st = engine().stream;
  if (!ta && !tb) k_bgemm<false, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC);
  else if (!ta && tb) k_bgemm<false, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC);
  else if (ta && !tb) k_bgemm<true, false><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC);
  else k_bgemm<true, true><<<grid, 256, 0, st>>>(M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC);
  LR_CHECK_LAUNCH();
  return LR_OK;
}
inline lr_status bgemm_nt(int M, int N, int K, double alpha, const double *A, int lda, size_t sA, const double *B,
                          int ldb, size_t sB, double beta, double *C, int ldc, size_t sC, int batch) {
  return bgemm(false, false, M, N, K, alpha, A, lda, sA, B, ldb, sB, beta, C, ldc, sC, batch);
}

// ---- fused batched Cholesky: ONE launch factors every matrix of the batch -------------------------
// A CTA owns one matrix and runs the whole left-looking factorization on it (the launches-per-block-step
// version above it spent more time in launch gaps and half-empty waves than in arithmetic, and its
// latency-bound 64 x 64 diagonal factorizations ran alone on the machine; here they overlap with the DMMA
// work of the other CTA on the SM).  Per block column k and 64-row tile t of it:
//   acc = L[tile rows, 0:k0] L[k rows, 0:k0]^T          (DMMA, slabs through shared memory)
//   P   = A[tile, k] - acc                                (fragments)
//   t = 0: L_kk = chol(P), X = L_kk^-1 (shared memory, 16-wide sub-blocks) -> global + invD
//   t > 0: L[tile, k] = P X^T                             (DMMA: P through shared memory, X from Ls)
// The factor is read back by the same CTA only: __syncthreads orders its global writes and reads.
// (the panel tile Ps aliases the K slabs As / Bs: they are never live together)
constexpr size_t kCholFusedSmem = (kNB * kBgLd + kNB * (kNB + 1) + kNB) * sizeof(double);

// acc += A_tile B_tile^T over K (A(m, k) at A[k lda + m], B(n, k) at B[k ldb + n]; rows >= mv / nv are zero).
// Two shared-memory slab buffers (buf: [2][A | B][16][72] doubles) and ONE barrier per 16-deep slab: the next
// slab is fetched into registers before the current one is multiplied and stored into the other buffer after.
// `live` (bit 4 x + y): the 8 x 8 sub-tiles of this warp that are needed at all -- rows below mv, columns below nv
// and, on a diagonal tile, not strictly above the diagonal; the others are skipped (a 400-row matrix ends in a
// 16-row block whose tiles would otherwise cost as much as full ones)
__device__ __forceinline__ unsigned chol_live_mask(int wm, int wn, int mv, int nv, bool diag) {
  unsigned live = 0;
#pragma unroll
  for (int x = 0; x < 2; x++)
#pragma unroll
    for (int y = 0; y < 4; y++)
      if (wm + 8 * x < mv && wn + 8 * y < nv && !(diag && wn + 8 * y >= wm + 8 * x + 8)) live |= 1u << (4 * x + y);
  return live;
}

// 8-byte asynchronous copy global -> shared; !valid copies nothing and zero-fills (src-size 0)
__device__ __forceinline__ void cp_async_f64(double *dst_smem, const double *src, bool valid) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst_smem);
  const int bytes = valid ? 8 : 0;
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// acc += A^T B over K rows: A[k][0 .. mv), B[k][0 .. nv) (leading dimensions lda / ldb), the K range streamed through
// a ring of kCsStages shared-memory slabs of kCsSlab rows filled by cp.async (no register staging): the copies of
// three slabs are in flight while one is multiplied, one barrier per slab.  The ring occupies exactly the
// 64 x 72 doubles of Ps.  Ends with a barrier: the caller may reuse the memory.
constexpr int kCsSlab = 8, kCsStages = 4;
static_assert(kCsStages * 2 * kCsSlab * kBgLd == kNB * kBgLd, "slab ring == Ps");
__device__ __forceinline__ void chol_tile_mma(double (&acc)[2][4][2], const double *A, int lda, int mv,
                                              const double *B, int ldb, int nv, int K, double *buf,
                                              unsigned live = 0xFFu) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  constexpr int kHalf = kCsSlab * kBgLd;  // doubles per operand slab
  if (K <= 0) return;
  const int nslab = (K + kCsSlab - 1) / kCsSlab;
  // slab s -> stage s % kCsStages; 2 operands x kCsSlab rows x 64 columns = 4 elements per thread
  auto issue = [&](int s) {
    if (s < nslab) {
      double *As = buf + (s % kCsStages) * 2 * kHalf;
#pragma unroll
      for (int j = 0; j < 2 * kCsSlab * 64 / 256; j++) {
        const int idx = threadIdx.x + 256 * j;
        const int op = idx / (kCsSlab * 64), rem = idx % (kCsSlab * 64);
        const int kr = rem >> 6, c = rem & 63, k = s * kCsSlab + kr;
        const bool ok = k < K && c < (op ? nv : mv);
        const double *src = op ? B + (size_t)k * ldb + c : A + (size_t)k * lda + c;
        cp_async_f64(As + op * kHalf + kr * kBgLd + c, ok ? src : A, ok);
      }
    }
    cp_async_commit();  // (empty groups keep the wait count uniform)
  };
#pragma unroll
  for (int s = 0; s < kCsStages - 1; s++) issue(s);
  for (int s = 0; s < nslab; s++) {
    cp_async_wait<kCsStages - 2>();  // slab s has landed (this thread's copies)
    __syncthreads();                 // ... everyone's, and stage (s - 1) % kCsStages is free
    issue(s + kCsStages - 1);
    const double *As = buf + (s % kCsStages) * 2 * kHalf, *Bs = As + kHalf;
    if (live) {
#pragma unroll
      for (int kk = 0; kk < kCsSlab; kk += 4) {
        double a[2], b[4];
#pragma unroll
        for (int x = 0; x < 2; x++) a[x] = As[(kk + fk) * kBgLd + wm + 8 * x + fr];
#pragma unroll
        for (int y = 0; y < 4; y++) b[y] = Bs[(kk + fk) * kBgLd + wn + 8 * y + fr];
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
          for (int y = 0; y < 4; y++)
            if (live >> (4 * x + y) & 1) dmma_8x8x4(acc[x][y][0], acc[x][y][1], a[x], b[y]);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
}

// packed != nullptr: the input matrices are PACKED lower triangles (column j holds rows j..n-1) with
// diag_add on the diagonal -- the output of the L = N TETt digit GEMM, or the M-step's A accumulators;
// the factor is still written to the full column-major Lall.
__global__ void __launch_bounds__(256, 3)
k_chol_fused(int n, double *Lall, size_t stride, const double *__restrict__ packed, double diag_add,
             double *__restrict__ invD, int nblk, int *__restrict__ bad, unsigned long long *prof) {
  extern __shared__ double csm[];
  // phase clocks of thread 0, summed over the grid (LR_CHOL_PROF=1; prof == nullptr otherwise)
  long long tlast = prof ? clock64() : 0;
#define CH_T(i)                                                    \
  do {                                                             \
    if (prof && threadIdx.x == 0) {                                \
      const long long t_ = clock64();                              \
      atomicAdd(prof + (i), (unsigned long long)(t_ - tlast));     \
      tlast = t_;                                                  \
    }                                                              \
  } while (0)
  double *slabs = csm;  // cp.async ring of chol_tile_mma: exactly the 64 x 72 doubles of Ps
  double (*Ps)[kBgLd] = reinterpret_cast<double (*)[kBgLd]>(csm);
  double (*Ls)[kNB + 1] = reinterpret_cast<double (*)[kNB + 1]>(csm + kNB * kBgLd);
  double *rdiag = csm + kNB * kBgLd + kNB * (kNB + 1);
  __shared__ int fail_flag;
  const int mat = blockIdx.x;
  double *L = Lall + (size_t)mat * stride;
  const size_t rp = (size_t)n * (n + 1) / 2;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int wm = (warp & 3) * 16, wn = (warp >> 2) * 32;
  const int fr = lane >> 2, fk = lane & 3;
  const int j = threadIdx.x >> 2, q = threadIdx.x & 3;  // diagonal-block roles: row / column j, lane q of its quad
  if (threadIdx.x == 0) fail_flag = 0;
  for (int k = 0; k < nblk; k++) {
    const int k0 = k * kNB, nb = min(kNB, n - k0);
    for (int m0 = k0; m0 < n; m0 += kNB) {
      const int mv = min(kNB, n - m0);
      double acc[2][4][2] = {};
      const unsigned live = chol_live_mask(wm, wn, mv, nb, m0 == k0);
      // the tile of the INPUT matrix this update is subtracted from: element (x, y, z) of the fragment layout.
      // Its 16 addresses are prefetched into L2 before the update loop and loaded as one batch after it
      // (a clamped, always-valid address + a select, no branch between the loads: one memory latency, not 16).
      const double *in_base = packed ? packed + (size_t)mat * rp : L;
      auto in_off = [&](int x, int y, int z) -> long long {  // < 0: outside the tile / above the diagonal
        const int r = wm + 8 * x + fr, c = wn + 8 * y + 2 * fk + z;
        const int row = m0 + r, col = k0 + c;
        if (r >= mv || c >= nb || (packed && row < col)) return -1;
        return packed ? (long long)(packed_off(n, col) + (size_t)(row - col)) : (long long)col * n + row;
      };
      if (k0 > 0) {
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
          for (int y = 0; y < 4; y++) {
#pragma unroll
            for (int z = 0; z < 2; z++) {  // z = 1 is the next column: another 32-byte sector
              const long long o = in_off(x, y, z);
              asm volatile("prefetch.global.L2 [%0];" ::"l"(in_base + (o < 0 ? 0 : o)));
            }
          }
      }
      chol_tile_mma(acc, L + m0, n, mv, L + k0, n, nb, k0, slabs, live);
      CH_T(0);
      // P = A[tile, k] - acc, in fragment layout: (row, col) = (wm + 8x + fr, wn + 8y + 2fk + z)
#pragma unroll
      for (int x = 0; x < 2; x++)
#pragma unroll
        for (int y = 0; y < 4; y++)
#pragma unroll
          for (int z = 0; z < 2; z++) {
            const long long o = in_off(x, y, z);
            const double a = in_base[o < 0 ? 0 : o];
            const bool dg = packed && m0 + wm + 8 * x + fr == k0 + wn + 8 * y + 2 * fk + z;
            acc[x][y][z] = (o < 0 ? 0.0 : a + (dg ? diag_add : 0.0)) - acc[x][y][z];
          }
      CH_T(1);
      if (m0 == k0) {
        // ---- diagonal block: factor + inverse in shared memory
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
          for (int y = 0; y < 4; y++)
#pragma unroll
            for (int z = 0; z < 2; z++) Ls[wm + 8 * x + fr][wn + 8 * y + 2 * fk + z] = acc[x][y][z];
        __syncthreads();
        // rows / columns past the matrix edge: identity, so the arithmetic below needs no masks
        if (threadIdx.x >= nb && threadIdx.x < kNB) {
          for (int c = 0; c < (int)threadIdx.x; c++) Ls[threadIdx.x][c] = 0.0;
          Ls[threadIdx.x][threadIdx.x] = 1.0;
        }
        __syncthreads();
        // Blocked in 16-wide sub-blocks so that the latency-bound column recurrence is 16 columns of ONE
        // warp (barriers = __syncwarp) instead of 64 columns of the whole CTA; everything else is
        // element-parallel over the 256 threads.  X[i][j] (i >= j) lives at Ls[j][i + 1].
        constexpr int kSB = 16;
        CH_T(2);
        for (int s0 = 0; s0 < kNB; s0 += kSB) {
          if (warp == 0) {
            const int rr = lane >> 1, qq = lane & 1, r = s0 + rr;
            // (a) factor the 16 x 16 diagonal sub-block
            // left-looking: column cc of row r is A[r][c] - sum_{kk < cc} L[r][kk] L[c][kk] -- independent
            // shared-memory loads feeding one FMA chain per lane (the right-looking form it replaces was a
            // read-modify-write of shared memory per term), the two lanes of a row split the sum
            for (int cc = 0; cc < kSB; cc++) {
              const int c = s0 + cc;
              // fixed trip count + predication: the 16 loads issue back to back, then two FMA chains of four
              double v = 0.0, v2 = 0.0;
#pragma unroll
              for (int i = 0; i < kSB / 2; i += 2) {
                const int ka = qq + 2 * i, kb = ka + 2;
                const bool oa = rr >= cc && ka < cc, ob = rr >= cc && kb < cc;
                const double la = Ls[r][s0 + (oa ? ka : 0)], pa = Ls[c][s0 + (oa ? ka : 0)];
                const double lb = Ls[r][s0 + (ob ? kb : 0)], pb = Ls[c][s0 + (ob ? kb : 0)];
                v = fma(oa ? -la : 0.0, pa, v);
                v2 = fma(ob ? -lb : 0.0, pb, v2);
              }
              v += v2;
              v += __shfl_xor_sync(0xffffffffu, v, 1);
              if (rr >= cc) v += Ls[r][c];
              double d = __shfl_sync(0xffffffffu, v, 2 * cc);  // the pivot: row cc
              if (!(d > 0.0)) {
                fail_flag = 1;
                d = 1.0;
              }
              const double rs = rsqrt(d);
              if (qq == 0 && rr >= cc) Ls[r][c] = rr == cc ? d * rs : v * rs;
              if (lane == 2 * cc) rdiag[c] = rs;  // 1 / L[c][c]
              __syncwarp();
            }
            // ... and its inverse: column rr by the lane pair (rr, qq)
            for (int ii = 0; ii < kSB; ii++) {
              double v = 0.0, v2 = 0.0;
#pragma unroll
              for (int i = 0; i < kSB / 2; i += 2) {
                const int ka = rr + qq + 2 * i, kb = ka + 2;
                const bool oa = ka < ii, ob = kb < ii;
                const double la = Ls[s0 + ii][s0 + (oa ? ka : 0)], xa = Ls[r][s0 + (oa ? ka : 0) + 1];
                const double lb = Ls[s0 + ii][s0 + (ob ? kb : 0)], xb = Ls[r][s0 + (ob ? kb : 0) + 1];
                v = fma(oa ? -la : 0.0, xa, v);
                v2 = fma(ob ? -lb : 0.0, xb, v2);
              }
              v += v2;
              v += __shfl_xor_sync(0xffffffffu, v, 1);
              if (qq == 0 && ii >= rr) Ls[r][s0 + ii + 1] = (ii == rr) ? rdiag[r] : v * rdiag[s0 + ii];
              __syncwarp();
            }
          }
          __syncthreads();
          CH_T(3);
          const int nr = kNB - s0 - kSB;  // rows below the sub-block
          if (nr > 0) {
            // (b) panel below it: L[r, s0 + cc] = sum_{k <= cc} P[r, s0 + k] X[s0 + cc][s0 + k]
            double pv[3];
#pragma unroll
            for (int u = 0; u < 3; u++) {
              const int e = threadIdx.x + 256 * u;
              pv[u] = 0.0;
              if (e < nr * kSB) {
                const int r = s0 + kSB + e / kSB, cc = e % kSB;
                double v = 0.0;
                for (int kk = 0; kk <= cc; kk++) v = fma(Ls[r][s0 + kk], Ls[s0 + kk][s0 + cc + 1], v);
                pv[u] = v;
              }
            }
            __syncthreads();
#pragma unroll
            for (int u = 0; u < 3; u++) {
              const int e = threadIdx.x + 256 * u;
              if (e < nr * kSB) Ls[s0 + kSB + e / kSB][s0 + e % kSB] = pv[u];
            }
            __syncthreads();
            // (c) trailing update of the lower triangle: A[r][c] -= sum_cc L[r, s0 + cc] L[c, s0 + cc]
            for (int e = threadIdx.x; e < nr * nr; e += 256) {
              const int r = s0 + kSB + e / nr, c = s0 + kSB + e % nr;
              if (c > r) continue;
              double v = Ls[r][c];
#pragma unroll
              for (int cc = 0; cc < kSB; cc++) v = fma(-Ls[r][s0 + cc], Ls[c][s0 + cc], v);
              Ls[r][c] = v;
            }
            __syncthreads();
          }
          CH_T(4);
        }
        for (int c = q; c < nb; c += 4)
          if (j < nb && j >= c) L[(size_t)(k0 + c) * n + k0 + j] = Ls[j][c];
        // off-diagonal 16 x 16 blocks of X = L^-1, one block diagonal at a time:
        //   X_ij = -X_ii (sum_{kb = j .. i-1} L_i,kb X_kb,j);   T goes through Ps (free during this tile)
        double *Tm = &Ps[0][0];
        for (int dlev = 1; dlev < kNB / kSB; dlev++) {
          const int nblkd = kNB / kSB - dlev;
          const int ea = threadIdx.x >> 4, eb = threadIdx.x & 15;
          for (int u = 0; u < nblkd; u++) {
            const int bi = dlev + u, bj = u;  // block row / column
            double v = 0.0;
            const int row = kSB * bi + ea, col = kSB * bj + eb;
            for (int m = col; m < kSB * bi; m++) v = fma(Ls[row][m], Ls[col][m + 1], v);  // X[m][col], m >= col
            Tm[u * 256 + threadIdx.x] = v;
          }
          __syncthreads();
          for (int u = 0; u < nblkd; u++) {
            const int bi = dlev + u, bj = u;
            const int row = kSB * bi + ea, col = kSB * bj + eb;
            double v = 0.0;
            for (int m = 0; m <= ea; m++) v = fma(Ls[kSB * bi + m][row + 1], Tm[u * 256 + m * 16 + eb], v);  // X[row][16 bi + m]
            Ls[col][row + 1] = -v;
          }
          __syncthreads();
        }
        double *out = invD + ((size_t)mat * nblk + k) * kNB * kNB;  // column-major, ld 64: out[c 64 + r] = X[r][c]
        for (int c = q; c < kNB; c += 4) out[(size_t)c * kNB + j] = (j < nb && c < nb && j >= c) ? Ls[c][j + 1] : 0.0;
        CH_T(5);
      } else {
        // ---- panel tile: L[tile, k] = P X^T.  P goes through shared memory as the k-major A operand
        // (Ps[c'][r]); B(n = c, k = c') = X[c][c'] = Ls[c'][c + 1] for c' <= c, zero above the diagonal
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
          for (int y = 0; y < 4; y++)
#pragma unroll
            for (int z = 0; z < 2; z++) Ps[wn + 8 * y + 2 * fk + z][wm + 8 * x + fr] = acc[x][y][z];
        __syncthreads();
        double out[2][4][2] = {};
#pragma unroll 4
        for (int kk = 0; kk < kNB; kk += 4) {
          const int kc = kk + fk;
          double a[2], b[4];
#pragma unroll
          for (int x = 0; x < 2; x++) a[x] = Ps[kc][wm + 8 * x + fr];
#pragma unroll
          for (int y = 0; y < 4; y++) {
            const int c = wn + 8 * y + fr;
            b[y] = (kc <= c && c < nb) ? Ls[kc][c + 1] : 0.0;
          }
#pragma unroll
          for (int x = 0; x < 2; x++)
#pragma unroll
            for (int y = 0; y < 4; y++)  // X is lower triangular: k blocks past the column block contribute zeros
              if ((live >> (4 * x + y) & 1) && kk < wn + 8 * y + 8) dmma_8x8x4(out[x][y][0], out[x][y][1], a[x], b[y]);
        }
#pragma unroll
        for (int x = 0; x < 2; x++)
#pragma unroll
          for (int y = 0; y < 4; y++)
#pragma unroll
            for (int z = 0; z < 2; z++) {
              const int r = wm + 8 * x + fr, c = wn + 8 * y + 2 * fk + z;
              if (r < mv && c < nb) L[(size_t)(k0 + c) * n + m0 + r] = out[x][y][z];
            }
        __syncthreads();  // Ps is reused by the next tile
        CH_T(6);
      }
    }
    __syncthreads();  // block column k is in global memory for the updates of k + 1
    CH_T(7);
  }
#undef CH_T
  if (threadIdx.x == 0 && fail_flag) atomicExch(bad, mat + 1);
}

// Solve L L^T x = b for ONE right-hand side per matrix with the factor of chol_batched AND its
// diagonal-block inverses: block forward / backward substitution, x_i = invD_ii (b_i - sum_j L_ij x_j),
// two barriers per 64-wide block instead of two per column (the column-by-column kernel it replaced
// spent 2 n barriers, each behind a dependent global load).  One CTA per matrix, four lanes per row.
__global__ void __launch_bounds__(256)
k_chol_solve_blocked(int n, const double *__restrict__ Lall, size_t stride, const double *__restrict__ invD,
                     int nblk, double *__restrict__ rhs) {
  extern __shared__ double xs[];  // [nblk * 64] solution in progress, then ys[64]
  double *ys = xs + (size_t)nblk * kNB;
  const double *L = Lall + (size_t)blockIdx.x * stride;
  const double *iD = invD + (size_t)blockIdx.x * nblk * kNB * kNB;
  double *b = rhs + (size_t)blockIdx.x * n;
  const int r = threadIdx.x >> 2, q = threadIdx.x & 3;
  for (int i = threadIdx.x; i < nblk * kNB; i += blockDim.x) xs[i] = i < n ? b[i] : 0.0;
  __syncthreads();
  for (int bi = 0; bi < nblk; bi++) {  // L y = b
    const int i0 = bi * kNB;
    const bool in = i0 + r < n;
    double acc = 0.0;
    if (in)
      for (int c = q; c < i0; c += 4) acc = fma(L[(size_t)c * n + i0 + r], xs[c], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0) ys[r] = in ? xs[i0 + r] - acc : 0.0;
    __syncthreads();
    double t = 0.0;
    for (int c = q; c <= r; c += 4) t = fma(iD[(size_t)bi * kNB * kNB + (size_t)c * kNB + r], ys[c], t);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    if (q == 0) xs[i0 + r] = t;
    __syncthreads();
  }
  for (int bi = nblk - 1; bi >= 0; bi--) {  // L^T x = y; here r indexes the COLUMN i0 + r of L
    const int i0 = bi * kNB, i1 = min(n, i0 + kNB);
    const bool in = i0 + r < n;
    double acc = 0.0;
    if (in)
      for (int k = i1 + q; k < n; k += 4) acc = fma(L[(size_t)(i0 + r) * n + k], xs[k], acc);
    acc += __shfl_xor_sync(0xffffffffu, acc, 1);
    acc += __shfl_xor_sync(0xffffffffu, acc, 2);
    if (q == 0) ys[r] = in ? xs[i0 + r] - acc : 0.0;
    __syncthreads();
    double t = 0.0;  // x[r] = sum_{c >= r} invD[c][r] y[c]
    for (int c = r + q; c < kNB; c += 4) t = fma(iD[(size_t)bi * kNB * kNB + (size_t)r * kNB + c], ys[c], t);
    t += __shfl_xor_sync(0xffffffffu, t, 1);
    t += __shfl_xor_sync(0xffffffffu, t, 2);
    if (q == 0) xs[i0 + r] = t;
    __syncthreads();
  }
  for (int i = threadIdx.x; i < n; i += blockDim.x) b[i] = xs[i];
}

// dst[mat][0:rows, 0:cols] = src[mat][0:rows, 0:cols]  (column-major, own ld / stride each)
__global__ void k_copy_rect(int rows, int cols, const double *__restrict__ src, int lds, size_t ss,
                            double *__restrict__ dst, int ldd, size_t sd) {
  int e = blockIdx.x * blockDim.x + threadIdx.x;
  if (e >= rows * cols) return;
  int col = e / rows, row = e - col * rows;
  dst[(size_t)blockIdx.y * sd + (size_t)col * ldd + row] =
      src[(size_t)blockIdx.y * ss + (size_t)col * lds + row];
}

// ---- approximate i-vector modes (IvExtractor --mode ubmWeight | eigenDecomposition) ------
// normTMatrix (:1600-1609): T[j, i] *= sqrt(invvar[i])
__global__ void k_norm_t(int R, size_t sv, const double *__restrict__ invvar, double *__restrict__ T) {
  size_t total = (size_t)R * sv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x)
    T[i] *= sqrt(invvar[i % sv]);
}
// normStatistics (:1225-1242): F = (F - mean N) sqrt(invvar)
__global__ void k_norm_stats(size_t U, int C, int D, const double *__restrict__ N,
                             const double *__restrict__ mean, const double *__restrict__ invvar,
                             double *__restrict__ F) {
  size_t sv = (size_t)C * D, total = U * sv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x) {
    size_t s = i / sv, k = i - s * sv;
    F[i] = (F[i] - mean[k] * N[s * C + k / D]) * sqrt(invvar[k]);
  }
}
// Tw = T o weight[component of the column]
__global__ void k_scale_cols_comp(int R, size_t sv, int D, const double *__restrict__ T,
                                  const double *__restrict__ weight, double *__restrict__ Tw) {
  size_t total = (size_t)R * sv;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (size_t)gridDim.x * blockDim.x)
    Tw[i] = T[i] * weight[(i % sv) / D];
}
// Dm[c, i] = sum_{k < D} G[c D + k, i]^2   (approximateTcTc :3131-3134)
__global__ void k_tctc_diag(int C, int D, int R, const double *__restrict__ G,
                            double *__restrict__ Dm) {
  size_t total = (size_t)C * R;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total;
       e += (size_t)gridDim.x * blockDim.x) {
    size_t c = e / R, i = e - c * R;
    double d = 0.0;
    for (int k = 0; k < D; k++) {
      double g = G[(c * D + k) * R + i];
      d += g * g;
    }
    Dm[e] = d;
  }
}
// den[s, i] = 1 + (sum_c N[s, c]) lambda_i   (estimateWUbmWeight :2367-2373, in the eigenbasis of W)
__global__ void k_ubm_weight_den(size_t U, int C, int R, const double *__restrict__ N,
                                 const double *__restrict__ lambda, double *__restrict__ den) {
  size_t s = blockIdx.x;
  __shared__ double nsum;
  __shared__ double red[8];
  double p = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) p += N[s * C + c];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) p += __shfl_xor_sync(0xffffffffu, p, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = p;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
    nsum = t;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < R; i += blockDim.x) den[s * R + i] = 1.0 + nsum * lambda[i];
}
// Z[s, i] /= den[s, i] (+ add_one: den holds N D without the leading 1, :2573-2578)
__global__ void k_div_rows(size_t n, const double *__restrict__ den, double add, double *__restrict__ Z) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n;
       i += (size_t)gridDim.x * blockDim.x)
    Z[i] /= (den[i] + add);
}
// reverse the column order of a column-major n x n matrix and of the eigenvalue vector
// (syevd returns ascending eigenvalues; the reference sorts descending), transposing into the
// reference's row-major eigvec[k][j]; sign convention: largest-magnitude component positive.
__global__ void k_eig_reorder(int n, int rank, const double *__restrict__ V, const double *__restrict__ lam,
                              double *__restrict__ eigvec, double *__restrict__ eigval) {
  int j = blockIdx.x;  // output column (j-th largest)
  if (j >= rank) return;
  const double *col = V + (size_t)(n - 1 - j) * n;
  __shared__ double best[256];
  __shared__ int besti[256];
  double b = -1.0;
  int bi = 0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) {
    double a = fabs(col[k]);
    if (a > b) {
      b = a;
      bi = k;
    }
  }
  best[threadIdx.x] = b;
  besti[threadIdx.x] = bi;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) {
      double ob = best[threadIdx.x + o];
      int oi = besti[threadIdx.x + o];
      if (ob > best[threadIdx.x] || (ob == best[threadIdx.x] && oi < besti[threadIdx.x])) {
        best[threadIdx.x] = ob;
        besti[threadIdx.x] = oi;
      }
    }
    __syncthreads();
  }
  const double sg = col[besti[0]] < 0.0 ? -1.0 : 1.0;
  for (int k = threadIdx.x; k < n; k += blockDim.x) eigvec[(size_t)k * rank + j] = sg * col[k];
  if (threadIdx.x == 0) eigval[j] = lam[n - 1 - j];
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
  cudaFree(tv->d_tettp);
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
  cudaFree(tv->d_fmax);
  cudaFree(tv->d_tett_planes);
  cudaFree(tv->d_ts_planes);
  cudaFree(tv->d_tett_scale);
  cudaFree(tv->d_ts_scale);
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

// In-place blocked Cholesky of `nb` column-major n x n matrices (lower triangle; the strict upper
// triangle is left with garbage).  invD receives the inverses of the diagonal-block factors;
// panel: [nb x n x kNB] scratch.
lr_status chol_batched(lr_tv *tv, double *Lb, int n, int nb, double *invD, const double *packed, double diag_add,
                       const char *what) {
  Engine &e = engine();
  const size_t rr = (size_t)n * n;
  const int nblk = (n + kNB - 1) / kNB;
  const long long sD = (long long)nblk * kNB * kNB;
  int *bad = tv->d_info;
  LR_CUDA(cudaMemsetAsync(bad, 0, sizeof(int), e.stream));
  bool &attr = e.attr_set[Engine::kAttrTvDiag];
  if (!attr) {
    LR_CUDA(cudaFuncSetAttribute(k_chol_fused, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kCholFusedSmem));
    attr = true;
  }
  static const bool kProf = getenv("LR_CHOL_PROF") != nullptr;  // phase clocks of the kernel, printed per launch
  DevBuf<unsigned long long> prof;
  if (kProf) {
    LR_CUDA(prof.alloc(8));
    LR_CUDA(cudaMemsetAsync(prof.p, 0, 8 * sizeof(unsigned long long), e.stream));
  }
  k_chol_fused<<<nb, 256, kCholFusedSmem, e.stream>>>(n, Lb, rr, packed, diag_add, invD, nblk, bad, prof.p);
  LR_CHECK_LAUNCH();
  if (kProf) {
    unsigned long long hp[8];
    LR_CUDA(cudaMemcpyAsync(hp, prof.p, sizeof(hp), cudaMemcpyDeviceToHost, e.stream));
    LR_CUDA(cudaStreamSynchronize(e.stream));
    static const char *name[8] = {"update MMA", "P assembly", "diag store", "diag 16-col factor+inverse (warp 0)",
                                  "diag panel+trailing", "diag X blocks + write", "panel solve", "column sync"};
    fprintf(stderr, "[chol_prof] n=%d matrices=%d, clk per matrix:", n, nb);
    for (int i = 0; i < 8; i++) fprintf(stderr, " %s=%.0f", name[i], (double)hp[i] / nb);
    fprintf(stderr, "\n");
  }
  int h = 0;
  LR_CUDA(cudaMemcpyAsync(&h, bad, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  if (h != 0)
    return fail(LR_ERR_NUMERIC, "%s: matrix %d of the batch is not positive definite", what, h - 1);
  return LR_OK;
}

// Posterior of a batch of utterances [u0, u0+nb): L = I + N TETt (in d_Lb), Cholesky, then
//   want_inverse == false: W[u] = L^-1 aux (substitution kernel)
//   want_inverse == true : d_Eb = L^-1 (blocked triangular inverse), W[u] = L^-1 aux
// (estimateW :2126-2168 / estimateAandC :1722-1760)
lr_status posterior_batch(lr_tv *tv, size_t u0, int nb, bool want_inverse) {
  Engine &e = engine();
  const int R = tv->R, C = tv->C;
  const size_t rr = (size_t)R * R;
  const int rp = (int)tv->Rp();
  const double one = 1.0, zero = 0.0;
  // packed lower triangles: Lp[nb x Rp] = N_b[nb x C] * TETtp[C x Rp]  (staged in d_Eb), then
  // L = I + unpack(Lp) with a zeroed upper triangle
  const int s = tv->planes;  // > 0: the INT8 digit GEMM (gemm_i8.cu); 0: cuBLAS fp64 (cross-check)
  if (s > 0) {
    DevBuf<unsigned char> pN;
    DevBuf<double> sN;
    LR_CUDA(pN.alloc(gemm_i8_panel_bytes(nb, C, s, kI8TileM)));
    LR_CUDA(sN.alloc(gemm_i8_scale_count(nb, kI8TileM)));
    lr_status st = gemm_i8_prepare(tv->d_N + u0 * C, (size_t)C, 1, nb, C, s, kI8TileM, pN.p, sN.p);
    if (st != LR_OK) return st;
    st = gemm_i8_run(pN.p, sN.p, nb, tv->d_tett_planes, tv->d_tett_scale, rp, C, s, 1.0, 0.0, tv->d_Eb, (size_t)rp);
    if (st != LR_OK) return st;
  } else {
    LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_N, CUBLAS_OP_N, rp, nb, C, &one, tv->d_tettp, rp,
                          tv->d_N + u0 * C, C, &zero, tv->d_Eb, rp));
    count_launch();
  }
  // aux[nb x R] = Fc_b[nb x sv] * Ts^T  -> written straight into W
  if (s > 0) {
    DevBuf<unsigned char> pF;
    DevBuf<double> sF;
    LR_CUDA(pF.alloc(gemm_i8_panel_bytes(nb, (long)tv->sv, s, kI8TileM)));
    LR_CUDA(sF.alloc(gemm_i8_scale_count(nb, kI8TileM)));
    lr_status st = gemm_i8_prepare(tv->d_F + u0 * tv->sv, tv->sv, 1, nb, (long)tv->sv, s, kI8TileM, pF.p, sF.p,
                                   tv->fmax_valid ? tv->d_fmax + u0 : nullptr);
    if (st != LR_OK) return st;
    st = gemm_i8_run(pF.p, sF.p, nb, tv->d_ts_planes, tv->d_ts_scale, R, (long)tv->sv, s, 1.0, 0.0,
                     tv->d_W + u0 * R, (size_t)R);
    if (st != LR_OK) return st;
  } else {
    LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_T, CUBLAS_OP_N, R, nb, (int)tv->sv, &one, tv->d_Ts,
                          (int)tv->sv, tv->d_F + u0 * tv->sv, (int)tv->sv, &zero, tv->d_W + u0 * R, R));
    count_launch();
  }
  // the Cholesky reads L = I + (packed product in Eb) directly and writes its factor to Lb
  lr_status st = chol_batched(tv, tv->d_Lb, R, nb, tv->d_invD, tv->d_Eb, 1.0,
                              "i-vector posterior precision L");
  if (st != LR_OK) return st;
  if (!want_inverse) {
    // w = L^-1 aux: one CTA per utterance, aux sits in W and is overwritten by the i-vector
    const int nblk_s = (R + kNB - 1) / kNB;
    k_chol_solve_blocked<<<nb, 256, (size_t)(nblk_s + 1) * kNB * sizeof(double), e.stream>>>(
        R, tv->d_Lb, rr, tv->d_invD, nblk_s, tv->d_W + u0 * R);
    LR_CHECK_LAUNCH();
    return LR_OK;
  }
  // Explicit inverse Linv = Y^T Y with Y = Lfac^-1, built block row by block row from the
  // inverses of the 64 x 64 diagonal blocks (kept by chol_batched) -- everything heavy is a
  // strided-batched DGEMM:
  //   Y[i, 0:i] = -invD_ii (Lfac[i, 0:i] Y[0:i, 0:i]),   Y[i, i] = invD_ii
  const int nblk = (R + kNB - 1) / kNB;
  const double mone = -1.0;
  LR_CUDA(cudaMemsetAsync(tv->d_Yb, 0, (size_t)nb * rr * sizeof(double), e.stream));
  const long long sD = (long long)nblk * kNB * kNB;
  // (cuBLAS fp64 for these three product families: the repo's DMMA kernel k_bgemm, which runs the Cholesky
  // panels and the M-step, reaches 16 TFLOP/s on them against cuBLAS's 36 -- measured, r2c20 -- so the
  // library keeps them until that kernel is tiled wider)
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
    k_copy_rect<<<dim3(ceil_div((long)nbi * nbi, 256), nb), 256, 0, e.stream>>>(
        nbi, nbi, tv->d_invD + (size_t)i * kNB * kNB, kNB, (size_t)sD,
        tv->d_Yb + (size_t)r0 * R + r0, R, rr);
    LR_CHECK_LAUNCH();
  }
  // Linv = Y^T Y -> Eb, LOWER TRIANGLE ONLY, block column by block column: Y is lower triangular, so
  // Linv[c0:R, c0:c0+nbi] = Y[c0:R, c0:R]^T Y[c0:R, c0:c0+nbi] (a third of the full product's work; the
  // strict upper triangle of Eb keeps whatever the inverse left there and is never read)
  for (int i = 0; i < nblk; i++) {
    const int c0 = i * kNB, nbi = std::min(kNB, R - c0), m = R - c0;
    const double *Ysub = tv->d_Yb + (size_t)c0 * R + c0;
    LR_CUBLAS(cublasDgemmStridedBatched(e.blas, CUBLAS_OP_T, CUBLAS_OP_N, m, nbi, m, &one, Ysub, R,
                                        (long long)rr, Ysub, R, (long long)rr, &zero,
                                        tv->d_Eb + (size_t)c0 * R + c0, R, (long long)rr, nb));
    count_launch();
  }
  // W_b = Linv_b aux_b = Y^T (Y aux): aux sits in W; t = Y aux goes through Lb's first nb*R doubles
  LR_CUBLAS(cublasDgemmStridedBatched(e.blas, CUBLAS_OP_N, CUBLAS_OP_N, R, 1, R, &one, tv->d_Yb, R,
                                      (long long)rr, tv->d_W + u0 * R, R, R, &zero, tv->d_Lb, R, R, nb));
  count_launch();
  LR_CUBLAS(cublasDgemmStridedBatched(e.blas, CUBLAS_OP_T, CUBLAS_OP_N, R, 1, R, &one, tv->d_Yb, R,
                                      (long long)rr, tv->d_Lb, R, R, &zero, tv->d_W + u0 * R, R, R, nb));
  count_launch();
  return LR_OK;
}

// Digit planes of the operands that change only with T (after estimateTETt): TETt^T as the [Rp x C]
// B operand of L = N TETt, and Ts = T o invvar as the [R x sv] B operand of aux = F Ts^T.
lr_status prepare_t_planes(lr_tv *tv) {
  Engine &e = engine();
  const int s = e.tv_gemm == 1 ? 0 : e.tv_planes;
  if (s != tv->planes) {
    cudaFree(tv->d_tett_planes);
    cudaFree(tv->d_ts_planes);
    cudaFree(tv->d_tett_scale);
    cudaFree(tv->d_ts_scale);
    tv->d_tett_planes = tv->d_ts_planes = nullptr;
    tv->d_tett_scale = tv->d_ts_scale = nullptr;
    tv->planes = 0;
    if (s > 0) {
      const long rp = (long)tv->Rp();
      LR_CUDA(cudaMalloc(&tv->d_tett_planes, gemm_i8_panel_bytes(rp, tv->C, s, kI8TileN)));
      LR_CUDA(cudaMalloc(&tv->d_tett_scale, gemm_i8_scale_count(rp, kI8TileN) * sizeof(double)));
      LR_CUDA(cudaMalloc(&tv->d_ts_planes, gemm_i8_panel_bytes(tv->R, (long)tv->sv, s, kI8TileN)));
      LR_CUDA(cudaMalloc(&tv->d_ts_scale, gemm_i8_scale_count(tv->R, kI8TileN) * sizeof(double)));
      tv->planes = s;
    }
  }
  if (s == 0) return LR_OK;
  const long rp = (long)tv->Rp();
  lr_status st = gemm_i8_prepare(tv->d_tettp, 1, (size_t)rp, rp, tv->C, s, kI8TileN, tv->d_tett_planes,
                                 tv->d_tett_scale);
  if (st != LR_OK) return st;
  return gemm_i8_prepare(tv->d_Ts, tv->sv, 1, tv->R, (long)tv->sv, s, kI8TileN, tv->d_ts_planes, tv->d_ts_scale);
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
  // utterances per posterior batch: at most 2.6 GB for the largest of the per-utterance buffers (L / E / Y
  // hold R x R doubles, invD ceil(R / 64) 64 x 64 blocks -- the larger one at small R), capped at 4096 so that
  // small ranks with many utterances do not over-allocate.  A batch is a whole number of Cholesky waves when
  // one fits (k_chol_fused: one CTA per matrix, 3 per SM), else whole 128-row tiles / K chunks of the digit
  // GEMM: the Cholesky is the largest single share of the i-vector solve and a 1.4-wave batch leaves a
  // quarter of it idle.
  const size_t per_utt = std::max(rr, (size_t)((R + kNB - 1) / kNB) * kNB * kNB) * sizeof(double);
  size_t fit = std::min<size_t>(4096, std::max<size_t>(32, (size_t)2600000000ull / per_utt));
  const size_t wave = 3 * (size_t)std::max(1, engine().sm_count);
  if (U <= fit) fit = U;  // everything in one batch
  else if (fit >= wave) fit -= fit % wave;
  else if (fit > 128) fit -= fit % 128;
  tv->batch = (int)fit;
  const int nbmax = tv->batch;
  auto A = [&](double **p, size_t n) { return cudaMalloc(p, n * sizeof(double)) == cudaSuccess; };
  bool ok = A(&tv->d_N, U * C) && A(&tv->d_F, U * tv->sv) && A(&tv->d_T, R * tv->sv) &&
            A(&tv->d_Ts, R * tv->sv) && A(&tv->d_W, U * R) && A(&tv->d_mean, tv->sv) &&
            A(&tv->d_invvar, tv->sv) && A(&tv->d_tett, (size_t)C * rr) &&
            A(&tv->d_tettp, (size_t)C * tv->Rp()) && A(&tv->d_acc, tv->acc_len()) &&
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
    const double gb = ((double)U * C + 2.0 * U * tv->sv / 2 + 2.0 * R * tv->sv + (double)U * R + (double)C * rr +
                       (double)C * tv->Rp() + (double)tv->acc_len() + 3.0 * nbmax * rr +
                       (double)nbmax * ((R + kNB - 1) / kNB) * kNB * kNB) * sizeof(double) / 1e9;
    fail(LR_ERR_CUDA, "lr_tv_create: cudaMalloc failed (%s) while allocating about %.1f GB for C=%d D=%d R=%d U=%zu, "
                      "batch %d", cudaGetErrorString(cudaGetLastError()), gb, C, D, R, U, nbmax);
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
  tv->fmax_valid = false;
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
double *lr_tv_dev_F(lr_tv *tv) {
  if (tv) tv->fmax_valid = false;  // the caller may write F: its row maxima are recomputed by the next substractM
  return tv ? tv->d_F : nullptr;
}

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
  if (A) {  // the packed lower triangles are expanded to the reference's full symmetric _A
    double *full = (double *)scratch_get(kSlotTmpA, (size_t)tv->C * rr * sizeof(double));
    if (!full) return LR_ERR_CUDA;
    k_unpack_lower<<<grid_for((size_t)tv->C * rr), 256, 0, engine().stream>>>(
        (size_t)tv->C, tv->R, tv->A(), full, 0.0, 1);
    LR_CHECK_LAUNCH();
    TV_COPY(A, full, (size_t)tv->C * rr, cudaMemcpyDeviceToHost);
  }
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
  if (!tv->d_fmax) LR_CUDA(cudaMalloc(&tv->d_fmax, tv->U * sizeof(unsigned long long)));
  LR_CUDA(cudaMemsetAsync(tv->d_fmax, 0, tv->U * sizeof(unsigned long long), engine().stream));
  dim3 grid((unsigned)ceil_div((long)tv->sv, kSubSeg), (unsigned)std::min<size_t>(tv->U, 65535));
  k_subtract_m<<<grid, 256, 0, engine().stream>>>(tv->U, tv->C, tv->D, tv->d_N, tv->d_mean, tv->d_F, tv->d_fmax);
  LR_CHECK_LAUNCH();
  tv->fmax_valid = true;
  return LR_OK;
}

lr_status lr_tv_estimate_tett(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_estimate_tett: null handle");
  Engine &e = engine();
  k_scale_cols<<<grid_for((size_t)tv->R * tv->sv), 256, 0, e.stream>>>(tv->R, tv->sv, tv->d_T,
                                                                       tv->d_invvar, tv->d_Ts);
  LR_CHECK_LAUNCH();
  // TETt_c = (T_c o invvar_c) T_c^T, lower triangles only, packed
  {
    const int nblk = (tv->R + kTtTile - 1) / kTtTile;
    dim3 grid((unsigned)(nblk * (nblk + 1) / 2), (unsigned)tv->C);
    k_tett_packed<<<grid, 256, 0, e.stream>>>(tv->R, tv->D, tv->sv, tv->d_Ts, tv->d_T, tv->d_tettp);
    LR_CHECK_LAUNCH();
  }
  return prepare_t_planes(tv);
}

lr_status lr_set_tv_gemm(int which, int planes) {
  LR_REQUIRE(which == 0 || which == 1, "lr_set_tv_gemm: kernel %d (0 = INT8 digit GEMM, 1 = cuBLAS fp64)", which);
  LR_REQUIRE(planes == 0 || (planes >= 3 && planes <= kI8MaxSlices), "lr_set_tv_gemm: %d digit planes outside [3, %d]",
             planes, kI8MaxSlices);
  engine().tv_gemm = which;
  if (planes) engine().tv_planes = planes;
  return LR_OK;
}

lr_status lr_gemm_digits(size_t M, size_t N, size_t K, const double *A, const double *B, double *Cm, double alpha,
                         double beta, int planes) {
  LR_READY();
  LR_REQUIRE(A && B && Cm && M && N && K, "lr_gemm_digits: null / empty argument");
  Engine &e = engine();
  const int s = planes ? planes : e.tv_planes;
  DevBuf<double> dA, dB, dC, sA, sB;
  DevBuf<unsigned char> pA, pB;
  LR_CUDA(dA.alloc(M * K));
  LR_CUDA(dB.alloc(N * K));
  LR_CUDA(dC.alloc(M * N));
  LR_CUDA(pA.alloc(gemm_i8_panel_bytes((long)M, (long)K, s, kI8TileM)));
  LR_CUDA(pB.alloc(gemm_i8_panel_bytes((long)N, (long)K, s, kI8TileN)));
  LR_CUDA(sA.alloc(gemm_i8_scale_count((long)M, kI8TileM)));
  LR_CUDA(sB.alloc(gemm_i8_scale_count((long)N, kI8TileN)));
  LR_CUDA(cudaMemcpyAsync(dA.p, A, M * K * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(dB.p, B, N * K * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(dC.p, Cm, M * N * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  lr_status st = gemm_i8_prepare(dA.p, K, 1, (long)M, (long)K, s, kI8TileM, pA.p, sA.p);
  if (st != LR_OK) return st;
  if ((st = gemm_i8_prepare(dB.p, K, 1, (long)N, (long)K, s, kI8TileN, pB.p, sB.p)) != LR_OK) return st;
  if ((st = gemm_i8_run(pA.p, sA.p, (long)M, pB.p, sB.p, (long)N, (long)K, s, alpha, beta, dC.p, N)) != LR_OK) return st;
  LR_CUDA(cudaMemcpyAsync(Cm, dC.p, M * N * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
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
  const int rp = (int)tv->Rp();
  LR_CUDA(cudaMemsetAsync(tv->A(), 0, (size_t)C * rp * sizeof(double), e.stream));
  LR_CUDA(cudaMemsetAsync(tv->Rm(), 0, (rr + 2 * (size_t)R) * sizeof(double), e.stream));
  DevBuf<double> Rmp;  // packed sum of the E_b
  LR_CUDA(Rmp.alloc((size_t)rp));
  LR_CUDA(cudaMemsetAsync(Rmp.p, 0, (size_t)rp * sizeof(double), e.stream));
  for (size_t u0 = 0; u0 < tv->U; u0 += tv->batch) {
    int nb = (int)std::min<size_t>(tv->batch, tv->U - u0);
    lr_status st = posterior_batch(tv, u0, nb, true);
    if (st != LR_OK) return st;
    const double *Wb = tv->d_W + u0 * R;
    // r += sum_b w_b ; sumW likewise (:1762, :1772)
    LR_CUBLAS(cublasDgemv(e.blas, CUBLAS_OP_N, R, nb, &one, Wb, R, tv->d_ones, 1, &one, tv->r(), 1));
    count_launch();
    // E_b = Linv_b + w_b w_b^T (:1766-1768), formed straight into PACKED lower triangles (E_b is symmetric;
    // the reference's loops run over the full R x R)
    k_pack_lower_rank1<<<grid_for((size_t)nb * rp), 256, 0, e.stream>>>((size_t)nb, R, tv->d_Eb, Wb, tv->d_Lb);
    LR_CHECK_LAUNCH();
    // Rm += sum_b E_b (:1770), on the packed form (unpacked + mirrored once at the end)
    LR_CUBLAS(cublasDgemv(e.blas, CUBLAS_OP_N, rp, nb, &one, tv->d_Lb, rp, tv->d_ones, 1, &one, Rmp.p, 1));
    count_launch();
    // A[C x Rp] += N_b^T pack(E_b) (:1775-1782)
    if (tv->planes > 0) {
      const int s = tv->planes;
      {  // rows = components, K = the batch's utterances
        DevBuf<unsigned char> pNt, pE;
        DevBuf<double> sNt, sE;
        LR_CUDA(pNt.alloc(gemm_i8_panel_bytes(C, nb, s, kI8TileM)));
        LR_CUDA(sNt.alloc(gemm_i8_scale_count(C, kI8TileM)));
        LR_CUDA(pE.alloc(gemm_i8_panel_bytes(rp, nb, s, kI8TileN)));
        LR_CUDA(sE.alloc(gemm_i8_scale_count(rp, kI8TileN)));
        if ((st = gemm_i8_prepare(tv->d_N + u0 * C, 1, (size_t)C, C, nb, s, kI8TileM, pNt.p, sNt.p)) != LR_OK) return st;
        if ((st = gemm_i8_prepare(tv->d_Lb, 1, (size_t)rp, rp, nb, s, kI8TileN, pE.p, sE.p)) != LR_OK) return st;
        if ((st = gemm_i8_run(pNt.p, sNt.p, C, pE.p, sE.p, rp, nb, s, 1.0, 1.0, tv->A(), (size_t)rp)) != LR_OK) return st;
      }
      {  // Cmx[R x sv] += W_b^T Fc_b (:1784-1788)
        DevBuf<unsigned char> pWt, pFt;
        DevBuf<double> sWt, sFt;
        LR_CUDA(pWt.alloc(gemm_i8_panel_bytes(R, nb, s, kI8TileM)));
        LR_CUDA(sWt.alloc(gemm_i8_scale_count(R, kI8TileM)));
        LR_CUDA(pFt.alloc(gemm_i8_panel_bytes((long)tv->sv, nb, s, kI8TileN)));
        LR_CUDA(sFt.alloc(gemm_i8_scale_count((long)tv->sv, kI8TileN)));
        if ((st = gemm_i8_prepare(Wb, 1, (size_t)R, R, nb, s, kI8TileM, pWt.p, sWt.p)) != LR_OK) return st;
        if ((st = gemm_i8_prepare(tv->d_F + u0 * tv->sv, 1, tv->sv, (long)tv->sv, nb, s, kI8TileN, pFt.p, sFt.p)) != LR_OK) return st;
        if ((st = gemm_i8_run(pWt.p, sWt.p, R, pFt.p, sFt.p, (long)tv->sv, nb, s, 1.0, 1.0, tv->Cmx(), tv->sv)) != LR_OK) return st;
      }
    } else {
      LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_N, CUBLAS_OP_T, rp, C, nb, &one, tv->d_Lb, rp,
                            tv->d_N + u0 * C, C, &one, tv->A(), rp));
      count_launch();
      // Cmx[R x sv] += W_b^T Fc_b (:1784-1788)
      LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_N, CUBLAS_OP_T, (int)tv->sv, R, nb, &one,
                            tv->d_F + u0 * tv->sv, (int)tv->sv, Wb, R, &one, tv->Cmx(), (int)tv->sv));
      count_launch();
    }
  }
  // Rm = the symmetric matrix of the packed sum
  k_unpack_lower<<<grid_for(rr), 256, 0, e.stream>>>((size_t)1, R, Rmp.p, tv->Rm(), 0.0, 1);
  LR_CHECK_LAUNCH();
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

// M-step of the components [c0, c1): T_c = A_c^-1 Cmx_c (updateTestimate :981-1000 is independent
// per component, so a multi-GPU run gives every rank C / world of them).
lr_status lr_tv_update_t_range(lr_tv *tv, int c0, int c1) {
  LR_READY();
  LR_REQUIRE(tv && c0 >= 0 && c0 < c1 && c1 <= tv->C, "lr_tv_update_t_range: bad component range");
  Engine &e = engine();
  const int R = tv->R, D = tv->D, nc = c1 - c0;
  const size_t rr = (size_t)R * R;
  const double one = 1.0;
  // the factors of the (packed) A_c go to the TETt buffer (TETt is re-estimated from the new T anyway)
  // T_c = A_c^-1 Cmx_c.  Held as T_c^T: the D x R column-major view (ld sv) of the row-major slice
  // T[:, cD:(c+1)D]; with A_c = L L^T the system is Z L L^T = B (B = Cmx_c^T, copied in first), solved block
  // column by block column with the factor's diagonal-block inverses -- the same two panel products as the
  // Cholesky itself (own DMMA kernel; the reference inverts A_c, :985-990):
  //   forward   Y_j = (B_j - sum_{k<j} Y_k L_jk^T) invD_jj^T
  //   backward  Z_j = (Y_j - sum_{k>j} Z_k L_kj)   invD_jj
  LR_CUDA(cudaMemcpy2DAsync(tv->d_T + (size_t)c0 * D, tv->sv * sizeof(double), tv->Cmx() + (size_t)c0 * D,
                            tv->sv * sizeof(double), (size_t)nc * D * sizeof(double), R,
                            cudaMemcpyDeviceToDevice, e.stream));
  const int nblk = (R + kNB - 1) / kNB, ldt = (int)tv->sv;
  const size_t sD = (size_t)nblk * kNB * kNB;
  for (int b0 = 0; b0 < nc; b0 += tv->batch) {
    const int nbm = std::min(tv->batch, nc - b0);
    double *Ab = tv->d_tett + (size_t)(c0 + b0) * rr;
    double *Tc = tv->d_T + (size_t)(c0 + b0) * D;
    lr_status st = chol_batched(tv, Ab, R, nbm, tv->d_invD, tv->A() + (size_t)(c0 + b0) * tv->Rp(), 0.0,
                                "M-step accumulator A_c");
    if (st != LR_OK) return st;
    for (int j = 0; j < nblk; j++) {
      const int j0 = j * kNB, nbj = std::min(kNB, R - j0);
      double *P = Tc + (size_t)j0 * ldt;
      if (j0 > 0 && (st = bgemm(false, false, D, nbj, j0, -1.0, Tc, ldt, (size_t)D, Ab + j0, R, rr, 1.0, P, ldt,
                                (size_t)D, nbm)) != LR_OK)
        return st;
      if ((st = bgemm(false, false, D, nbj, nbj, 1.0, P, ldt, (size_t)D, tv->d_invD + (size_t)j * kNB * kNB, kNB, sD,
                      0.0, P, ldt, (size_t)D, nbm)) != LR_OK)
        return st;
    }
    for (int j = nblk - 1; j >= 0; j--) {
      const int j0 = j * kNB, nbj = std::min(kNB, R - j0), j1 = j0 + nbj;
      double *P = Tc + (size_t)j0 * ldt;
      if (j1 < R && (st = bgemm(false, true, D, nbj, R - j1, -1.0, Tc + (size_t)j1 * ldt, ldt, (size_t)D,
                                Ab + (size_t)j0 * R + j1, R, rr, 1.0, P, ldt, (size_t)D, nbm)) != LR_OK)
        return st;
      if ((st = bgemm(false, true, D, nbj, nbj, 1.0, P, ldt, (size_t)D, tv->d_invD + (size_t)j * kNB * kNB, kNB, sD,
                      0.0, P, ldt, (size_t)D, nbm)) != LR_OK)
        return st;
    }
  }
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_tv_update_t(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_update_t: null handle");
  return lr_tv_update_t_range(tv, 0, tv->C);
}

// The columns of T that belong to the components [c0, c1) as one contiguous device block
// [R x (c1 - c0) D] (and back): the all-gather payload of the component-sharded M-step.
lr_status lr_tv_pack_t(lr_tv *tv, int c0, int c1, double *d_dst) {
  LR_READY();
  LR_REQUIRE(tv && d_dst && c0 >= 0 && c0 < c1 && c1 <= tv->C, "lr_tv_pack_t: bad arguments");
  const size_t w = (size_t)(c1 - c0) * tv->D * sizeof(double);
  LR_CUDA(cudaMemcpy2DAsync(d_dst, w, tv->d_T + (size_t)c0 * tv->D, tv->sv * sizeof(double), w, tv->R,
                            cudaMemcpyDeviceToDevice, engine().stream));
  return LR_OK;
}
lr_status lr_tv_unpack_t(lr_tv *tv, int c0, int c1, const double *d_src) {
  LR_READY();
  LR_REQUIRE(tv && d_src && c0 >= 0 && c0 < c1 && c1 <= tv->C, "lr_tv_unpack_t: bad arguments");
  const size_t w = (size_t)(c1 - c0) * tv->D * sizeof(double);
  LR_CUDA(cudaMemcpy2DAsync(tv->d_T + (size_t)c0 * tv->D, tv->sv * sizeof(double), d_src, w, w, tv->R,
                            cudaMemcpyDeviceToDevice, engine().stream));
  return LR_OK;
}
// length (doubles) of the packed A_c block of ONE component inside lr_tv_dev_acc (components are
// contiguous: a reduce-scatter by component works on the block directly)
lr_status lr_tv_dims(const lr_tv *tv, int *C, int *D, int *R) {
  LR_REQUIRE(tv, "lr_tv_dims: null handle");
  if (C) *C = tv->C;
  if (D) *D = tv->D;
  if (R) *R = tv->R;
  return LR_OK;
}

size_t lr_tv_acc_a_stride(const lr_tv *tv) { return tv ? tv->Rp() : 0; }

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

// ---- approximate i-vector modes ----------------------------------------------------------
lr_status lr_tv_norm_t(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_norm_t: null handle");
  k_norm_t<<<grid_for((size_t)tv->R * tv->sv), 256, 0, engine().stream>>>(tv->R, tv->sv, tv->d_invvar,
                                                                        tv->d_T);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

lr_status lr_tv_norm_statistics(lr_tv *tv) {
  LR_READY();
  LR_REQUIRE(tv, "lr_tv_norm_statistics: null handle");
  tv->fmax_valid = false;
  k_norm_stats<<<grid_for(tv->U * tv->sv), 256, 0, engine().stream>>>(tv->U, tv->C, tv->D, tv->d_N,
                                                                      tv->d_mean, tv->d_invvar, tv->d_F);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

lr_status lr_tv_weighted_cov(lr_tv *tv, const double *weight, double *W) {
  LR_READY();
  LR_REQUIRE(tv && weight && W, "lr_tv_weighted_cov: null argument");
  Engine &e = engine();
  const int R = tv->R;
  const double one = 1.0, zero = 0.0;
  double *Tw = (double *)scratch_get(kSlotTmpB, ((size_t)R * tv->sv + tv->C) * sizeof(double));
  if (!Tw) return LR_ERR_CUDA;
  double *d_w = Tw + (size_t)R * tv->sv;
  LR_CUDA(cudaMemcpyAsync(d_w, weight, tv->C * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  k_scale_cols_comp<<<grid_for((size_t)R * tv->sv), 256, 0, e.stream>>>(R, tv->sv, tv->D, tv->d_T,
                                                                        d_w, Tw);
  LR_CHECK_LAUNCH();
  // W[R x R] = Tw T^T (symmetric): column-major (sv x R) views of the row-major matrices
  LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_T, CUBLAS_OP_N, R, R, (int)tv->sv, &one, tv->d_T,
                        (int)tv->sv, Tw, (int)tv->sv, &zero, tv->d_Lb, R));
  count_launch();
  LR_CUDA(cudaMemcpyAsync(W, tv->d_Lb, (size_t)R * R * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

namespace lr {
namespace {
// symmetric eigen-problem on the device: d_V (column-major n x n, overwritten with the
// eigenvectors), d_lam ascending.  cusolverDnDsyevd.
lr_status eig_sym_device(cusolverDnHandle_t solver, int n, double *d_V, double *d_lam, int *d_info) {
  Engine &e = engine();
  int lwork = 0;
  LR_CUSOLVER(cusolverDnDsyevd_bufferSize(solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n,
                                          d_V, n, d_lam, &lwork));
  double *work = (double *)scratch_get(kSlotTmpA, (size_t)std::max(lwork, 1) * sizeof(double));
  if (!work) return LR_ERR_CUDA;
  LR_CUSOLVER(cusolverDnDsyevd(solver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, n, d_V, n,
                               d_lam, work, lwork, d_info));
  count_launch();
  int h = 0;
  LR_CUDA(cudaMemcpyAsync(&h, d_info, sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  if (h != 0) return fail(LR_ERR_NUMERIC, "eigen-decomposition did not converge (info %d)", h);
  return LR_OK;
}

// W_b (op)= Q diag(1 / den_b) Q^T aux_b for every utterance; aux = Fn Tn^T.
// Qrm: device row-major Q[R x R] (reference layout: Q(i, k) = component i of eigenvector k).
// den: device [U x R]; den_add is added to it.  accumulate: W += (the reference's
// eigenDecomposition path never resets _W) or W = (ubmWeight path, :2354).
lr_status approx_solve(lr_tv *tv, const double *Qrm, const double *den, double den_add,
                       bool accumulate) {
  Engine &e = engine();
  const int R = tv->R;
  const double one = 1.0, zero = 0.0, beta = accumulate ? 1.0 : 0.0;
  double *aux = tv->d_Eb, *Z = tv->d_Yb;  // [batch x R] each
  for (size_t u0 = 0; u0 < tv->U; u0 += tv->batch) {
    const int nb = (int)std::min<size_t>(tv->batch, tv->U - u0);
    // aux[nb x R] = Fn_b[nb x sv] Tn^T   (:2380-2384, :2589-2593)
    LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_T, CUBLAS_OP_N, R, nb, (int)tv->sv, &one, tv->d_T,
                          (int)tv->sv, tv->d_F + u0 * tv->sv, (int)tv->sv, &zero, aux, R));
    // Z = aux Q : column-major Z^T[R x nb] = (Q^T)[R x R] aux^T; row-major Q seen column-major is Q^T
    LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_N, CUBLAS_OP_N, R, nb, R, &one, Qrm, R, aux, R, &zero, Z, R));
    k_div_rows<<<grid_for((size_t)nb * R), 256, 0, e.stream>>>((size_t)nb * R, den + u0 * R, den_add, Z);
    LR_CHECK_LAUNCH();
    // W_b (+)= Z Q^T : column-major W^T[R x nb] = Q[R x R] Z^T = op_T(Qrm seen column-major) Z^T
    LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_T, CUBLAS_OP_N, R, nb, R, &one, Qrm, R, Z, R, &beta,
                          tv->d_W + u0 * R, R));
    count_launch(3);
  }
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}
}  // namespace
}  // namespace lr

lr_status lr_eigen_problem(int n, const double *EP, int rank, double *eigvec, double *eigval) {
  LR_READY();
  LR_REQUIRE(n >= 1 && rank >= 1 && rank <= n && EP && eigvec && eigval,
             "lr_eigen_problem: bad arguments (n=%d rank=%d)", n, rank);
  Engine &e = engine();
  cusolverDnHandle_t solver = nullptr;
  LR_CUSOLVER(cusolverDnCreate(&solver));
  struct Guard {
    cusolverDnHandle_t h;
    ~Guard() { cusolverDnDestroy(h); }
  } guard{solver};
  LR_CUSOLVER(cusolverDnSetStream(solver, e.stream));
  const size_t nn = (size_t)n * n;
  double *buf = (double *)scratch_get(kSlotTmpB, (nn + n + nn + n + 2) * sizeof(double));
  if (!buf) return LR_ERR_CUDA;
  double *d_V = buf, *d_lam = buf + nn, *d_out = d_lam + n, *d_val = d_out + nn;
  int *d_info = (int *)(d_val + n);
  LR_CUDA(cudaMemcpyAsync(d_V, EP, nn * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  lr_status st = eig_sym_device(solver, n, d_V, d_lam, d_info);
  if (st != LR_OK) return st;
  k_eig_reorder<<<rank, 256, 0, e.stream>>>(n, rank, d_V, d_lam, d_out, d_val);
  LR_CHECK_LAUNCH();
  LR_CUDA(cudaMemcpyAsync(eigvec, d_out, (size_t)n * rank * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaMemcpyAsync(eigval, d_val, rank * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_tv_approximate_tctc(lr_tv *tv, const double *Q, double *Dm) {
  LR_READY();
  LR_REQUIRE(tv && Q && Dm, "lr_tv_approximate_tctc: null argument");
  Engine &e = engine();
  const int R = tv->R;
  const double one = 1.0, zero = 0.0;
  double *G = (double *)scratch_get(kSlotTmpB, (tv->sv * R + (size_t)R * R + (size_t)tv->C * R) * sizeof(double));
  if (!G) return LR_ERR_CUDA;
  double *d_Q = G + tv->sv * R, *d_D = d_Q + (size_t)R * R;
  LR_CUDA(cudaMemcpyAsync(d_Q, Q, (size_t)R * R * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  // G[sv x R] (row-major) = T^T Q : column-major G^T[R x sv] = Q^T[R x R] T[R x sv]
  //   = (row-major Q seen column-major) (row-major T seen column-major (sv x R), transposed)
  LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_N, CUBLAS_OP_T, R, (int)tv->sv, R, &one, d_Q, R, tv->d_T,
                        (int)tv->sv, &zero, G, R));
  count_launch();
  k_tctc_diag<<<grid_for((size_t)tv->C * R), 256, 0, e.stream>>>(tv->C, tv->D, R, G, d_D);
  LR_CHECK_LAUNCH();
  LR_CUDA(cudaMemcpyAsync(Dm, d_D, (size_t)tv->C * R * sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_tv_estimate_w_ubm_weight(lr_tv *tv, const double *W) {
  LR_READY();
  LR_REQUIRE(tv && W, "lr_tv_estimate_w_ubm_weight: null argument");
  Engine &e = engine();
  const int R = tv->R;
  const size_t rr = (size_t)R * R;
  // L_s = I + n_s W shares the eigenvectors of W: L_s^-1 = Q diag(1 / (1 + n_s lambda)) Q^T, so the
  // reference's per-utterance R x R inversion (:2376-2378) becomes one eigen-decomposition + GEMMs.
  double *buf = (double *)scratch_get(kSlotTmpB, (2 * rr + R + tv->U * R) * sizeof(double));
  if (!buf) return LR_ERR_CUDA;
  double *d_V = buf, *d_Qrm = buf + rr, *d_lam = d_Qrm + rr, *d_den = d_lam + R;
  LR_CUDA(cudaMemcpyAsync(d_V, W, rr * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  lr_status st = eig_sym_device(tv->solver, R, d_V, d_lam, tv->d_info);
  if (st != LR_OK) return st;
  // column-major V (eigenvector k in column k) -> row-major Q(i, k): a transpose
  const double one = 1.0, zero = 0.0;
  LR_CUBLAS(cublasDgeam(e.blas, CUBLAS_OP_T, CUBLAS_OP_N, R, R, &one, d_V, R, &zero, d_V, R, d_Qrm, R));
  count_launch();
  k_ubm_weight_den<<<(unsigned)tv->U, 256, 0, e.stream>>>(tv->U, tv->C, R, tv->d_N, d_lam, d_den);
  LR_CHECK_LAUNCH();
  return approx_solve(tv, d_Qrm, d_den, 0.0, false);
}

lr_status lr_tv_estimate_w_eigen_decomposition(lr_tv *tv, const double *Dm, const double *Q) {
  LR_READY();
  LR_REQUIRE(tv && Dm && Q, "lr_tv_estimate_w_eigen_decomposition: null argument");
  Engine &e = engine();
  const int R = tv->R, C = tv->C;
  const size_t rr = (size_t)R * R;
  const double one = 1.0, zero = 0.0;
  double *buf = (double *)scratch_get(kSlotTmpB, (rr + (size_t)C * R + tv->U * R) * sizeof(double));
  if (!buf) return LR_ERR_CUDA;
  double *d_Qrm = buf, *d_D = buf + rr, *d_den = d_D + (size_t)C * R;
  LR_CUDA(cudaMemcpyAsync(d_Qrm, Q, rr * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(d_D, Dm, (size_t)C * R * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  // den[U x R] = N[U x C] Dm[C x R]  (:2573-2578; the leading 1 is added in k_div_rows)
  LR_CUBLAS(cublasDgemm(e.blas, CUBLAS_OP_N, CUBLAS_OP_N, R, (int)tv->U, C, &one, d_D, R, tv->d_N, C,
                        &zero, d_den, R));
  count_launch();
  return approx_solve(tv, d_Qrm, d_den, 1.0, true);
}

double *lr_tv_dev_acc(lr_tv *tv) { return tv ? tv->d_acc : nullptr; }
size_t lr_tv_acc_len(const lr_tv *tv) { return tv ? tv->acc_len() : 0; }

}  // extern "C"
