// gemm_split.cu -- tcgen05 split-precision "skinny K" GEMM with a rank-1 style epilogue:
//
//     C[m, n] = alpha * sum_k A[m, k] B[n, k]  +  row_term[m]  +  col_term[n]        (K <= 256)
//
// the shape of every trial-matrix scorer of the i-vector back end (PLDA native scoring:
// PldaTools.cpp:4186-4271 -- models x segments over the rank-r speaker space).  fp64 operands are
// scaled by a power of two, split into fp16 hi + lo (22 significand bits) and contracted as three
// fp16 UMMAs with fp32 accumulation in TMEM:  A_hi B_hi + A_lo B_hi + A_hi B_lo  (the scheme of
// gmm_tc.cu).  The output is written ONCE, in fp32 or fp64: the kernel is bound by that write (4 B
// per trial at fp32) once the operand traffic stays in L2.
//
// Decomposition.  A tile of 128 rows of A (all of K, hi and lo: 256 TMEM columns) is the A operand
// of every UMMA and lives in TMEM for the whole work item (TS mode: only B is fetched from shared
// memory); the item's column tiles of B (128 rows x 64-wide K chunks, hi | lo = 32 KB per chunk)
// stream through a 6-stage bulk-copy ring; two 128-column fp32 accumulators alternate, so the
// epilogue of tile i (TMEM -> registers -> + row / column terms -> global) overlaps the UMMAs of
// tile i + 1.  Warp 0 bulk-copy producer, warp 1 UMMA issuer, warp 2 TMEM allocator, warps 4-7
// epilogue (one per TMEM lane quarter) -- they also move the A tile from the ring into TMEM.
#include <algorithm>
#include <cmath>
#include <type_traits>

#include "gemm_split.cuh"
#include "tc_ptx.cuh"

namespace lr {
namespace {
using namespace tcptx;

constexpr int kGThreads = 384;  // warps 0-3: producer, UMMA issuer, TMEM allocator, (idle); 4-11: epilogue
constexpr int kGStages = 6;
constexpr int kChunkBytes = 2 * 128 * 128;  // [hi | lo] x 128 rows x 64 fp16
constexpr size_t kGSmem = 1024 + kGStages * (size_t)kChunkBytes + 256;

// rows [r0, r0 + 128) of X (row-major, leading dimension ld) -> per K chunk [hi | lo] swizzled panels
__global__ void __launch_bounds__(256)
k_split_panels(const double *__restrict__ X, size_t ld, long rows, int K, int nchunk, double scale,
               unsigned char *__restrict__ out) {
  // one thread per (row, 16-byte group of 8 columns)
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long row = gid / (nchunk * 8);
  const int grp = (int)(gid % (nchunk * 8));
  const long rows_pad = (rows + 127) / 128 * 128;
  if (row >= rows_pad) return;
  const int kc = grp >> 3, j = grp & 7;
  __align__(16) __half hi[8], lo[8];
#pragma unroll
  for (int e = 0; e < 8; e++) {
    const int k = kc * 64 + j * 8 + e;
    float v = 0.f, rest = 0.f;
    if (row < rows && k < K) {
      const double x = X[(size_t)row * ld + k] * scale;
      v = (float)x;
      const __half h = __float2half_rn(v);
      rest = (float)(x - (double)__half2float(h));
      hi[e] = h;
    } else {
      hi[e] = __float2half_rn(0.f);
    }
    lo[e] = __float2half_rn(rest);
  }
  const long tile = row / 128;
  const int r = (int)(row % 128);
  unsigned char *base = out + ((size_t)tile * nchunk + kc) * kChunkBytes;
  const uint32_t off = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
  *reinterpret_cast<uint4 *>(base + off) = *reinterpret_cast<const uint4 *>(hi);
  *reinterpret_cast<uint4 *>(base + 128 * 128 + off) = *reinterpret_cast<const uint4 *>(lo);
}

// max |x| over a strided matrix (for the power-of-two operand scale)
__global__ void k_absmax(const double *__restrict__ X, size_t ld, long rows, int K, double *__restrict__ out) {
  double m = 0.0;
  const size_t total = (size_t)rows * K;
  for (size_t e = (size_t)blockIdx.x * blockDim.x + threadIdx.x; e < total; e += (size_t)gridDim.x * blockDim.x) {
    const size_t r = e / K, k = e - r * K;
    m = fmax(m, fabs(X[r * ld + k]));
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0 && m > 0.0)
    atomicMax(reinterpret_cast<unsigned long long *>(out), (unsigned long long)__double_as_longlong(m));
}

struct GBars {
  uint32_t base, bar;
  __device__ __forceinline__ uint32_t stage(int i) const { return base + i * kChunkBytes; }
  __device__ __forceinline__ uint32_t full(int i) const { return bar + 8 * i; }
  __device__ __forceinline__ uint32_t empty(int i) const { return bar + 48 + 8 * i; }
  __device__ __forceinline__ uint32_t acc_full(int i) const { return bar + 96 + 8 * i; }
  __device__ __forceinline__ uint32_t acc_empty(int i) const { return bar + 112 + 8 * i; }
  __device__ __forceinline__ uint32_t a_tmem() const { return bar + 128; }
  __device__ __forceinline__ uint32_t tmem_slot() const { return bar + 136; }
};

// work item = (row tile mt, column tiles [n0, n1)); items are dealt round-robin to the CTAs
struct Items {
  int mt_count, nt_count, seg, n_items;  // seg = column-tile segments per row tile
  __host__ __device__ void get(int it, int &mt, int &n0, int &n1) const {
    mt = it / seg;
    const int s = it - mt * seg;
    n0 = (int)((long)nt_count * s / seg);
    n1 = (int)((long)nt_count * (s + 1) / seg);
  }
};

__global__ void k_to_float(long n, const double *__restrict__ x, float *__restrict__ y) {
  long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) y[i] = (float)x[i];
}

// OutT = float: the epilogue runs in fp32 on fp32 copies of the row / column terms; OutT = double: fp64
template <typename OutT>
__global__ void __launch_bounds__(kGThreads, 1)
k_gemm_split(Items items, int nchunk, int k16, const unsigned char *__restrict__ Ap,
             const unsigned char *__restrict__ Bp, OutT *__restrict__ Cout, size_t ldc, long M, long N,
             OutT alpha, const OutT *__restrict__ row_term, const OutT *__restrict__ col_term) {
  extern __shared__ unsigned char smem_raw[];
  GBars sm;
  sm.base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  sm.bar = sm.base + kGStages * kChunkBytes;
  unsigned char *base_ptr = smem_raw + (sm.base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kGStages; i++) {
      mbar_init(sm.full(i), 1);
      mbar_init(sm.empty(i), 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(sm.acc_full(i), 1);
      mbar_init(sm.acc_empty(i), 8);
    }
    mbar_init(sm.a_tmem(), 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(sm.tmem_slot(), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_slot()));
  constexpr int kColAhi = 0, kColAlo = 128, kColAcc = 256;
  constexpr uint32_t idesc = make_idesc(128, 128, 0, 0);

  if (warp == 0) {
    // ---- producer: per item the A tile's chunks, then the B tiles' chunks, all through one ring
    const bool leader = elect_one();
    long e = 0;  // ring entry counter
    for (int it = blockIdx.x; it < items.n_items; it += gridDim.x) {
      int mt, n0, n1;
      items.get(it, mt, n0, n1);
      const int n_entries = nchunk * (1 + (n1 - n0));
      for (int i = 0; i < n_entries; i++, e++) {
        const int st = (int)(e % kGStages);
        mbar_wait(sm.empty(st), (uint32_t)(((e / kGStages) & 1) ^ 1));
        if (leader) {
          const unsigned char *src = i < nchunk
                                         ? Ap + ((size_t)mt * nchunk + i) * kChunkBytes
                                         : Bp + ((size_t)(n0 + (i - nchunk) / nchunk) * nchunk + (i - nchunk) % nchunk) * kChunkBytes;
          mbar_expect_tx(sm.full(st), kChunkBytes);
          bulk_g2s(sm.stage(st), src, kChunkBytes / 2, sm.full(st));
          bulk_g2s(sm.stage(st) + kChunkBytes / 2, src + kChunkBytes / 2, kChunkBytes / 2, sm.full(st));
        }
        __syncwarp();
      }
    }
  } else if (warp == 1) {
    // ---- UMMA issuer
    const bool leader = elect_one();
    long e = 0;
    int n_item = 0;
    long tile_seq = 0;  // accumulator buffer = tile_seq & 1
    for (int it = blockIdx.x; it < items.n_items; it += gridDim.x, n_item++) {
      int mt, n0, n1;
      items.get(it, mt, n0, n1);
      e += nchunk;                               // the A entries are consumed by the epilogue warps
      mbar_wait(sm.a_tmem(), (uint32_t)(n_item & 1));  // the A tile of this item is in TMEM
      for (int nt = n0; nt < n1; nt++, tile_seq++) {
        const int buf = (int)(tile_seq & 1);
        if (tile_seq >= 2) mbar_wait(sm.acc_empty(buf), (uint32_t)(((tile_seq >> 1) - 1) & 1));
        uint32_t acc = 0;
        for (int kc = 0; kc < nchunk; kc++, e++) {
          const int st = (int)(e % kGStages);
          mbar_wait(sm.full(st), (uint32_t)((e / kGStages) & 1));
          tc_fence_after();
          if (leader) {
            const uint64_t bhi = make_desc(sm.stage(st), 16, 1024);
            const uint64_t blo = make_desc(sm.stage(st) + 128 * 128, 16, 1024);
            const uint32_t d = tmem_base + kColAcc + buf * 128;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              if (kc * 4 + kk < k16) {
                const uint32_t a_hi = tmem_base + kColAhi + kc * 32 + kk * 8;
                const uint32_t a_lo = tmem_base + kColAlo + kc * 32 + kk * 8;
                umma_ts(d, a_hi, desc_add(bhi, kk * 32), idesc, acc);
                umma_ts(d, a_lo, desc_add(bhi, kk * 32), idesc, 1u);
                umma_ts(d, a_hi, desc_add(blo, kk * 32), idesc, 1u);
                acc = 1;
              }
            }
            umma_commit(sm.empty(st));
            if (kc == nchunk - 1) umma_commit(sm.acc_full(buf));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp >= 4) {
    // ---- eight epilogue warps: TMEM lane quarter q, column half `half` of every tile.  They also move
    // the A tile from the ring into TMEM (hi panels by half 0, lo panels by half 1).
    const int q = warp & 3, half = (warp - 4) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    long e = 0;
    long tile_seq = 0;
    for (int it = blockIdx.x; it < items.n_items; it += gridDim.x) {
      int mt, n0, n1;
      items.get(it, mt, n0, n1);
      // (every UMMA of the previous item has completed: its last accumulator was waited for below)
      for (int kc = 0; kc < nchunk; kc++, e++) {
        const int st = (int)(e % kGStages);
        mbar_wait(sm.full(st), (uint32_t)((e / kGStages) & 1));
        const int r = q * 32 + lane;
#pragma unroll
        for (int half16 = 0; half16 < 2; half16++) {
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int chunk = half16 * 4 + j;
            const uint32_t a = sm.stage(st) + half * (128 * 128) + r * 128 + ((chunk ^ (r & 7)) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[4 * j]), "=r"(v[4 * j + 1]), "=r"(v[4 * j + 2]), "=r"(v[4 * j + 3])
                         : "r"(a));
          }
          tmem_st16(tmem_base + lane_addr + (half ? kColAlo : kColAhi) + kc * 32 + half16 * 16, v);
        }
        tmem_wait_st();
        named_bar_sync(1, 256);  // all eight warps have read the stage
        if (threadIdx.x == 128) mbar_arrive(sm.empty(st));
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.a_tmem());
      e += (long)nchunk * (n1 - n0);  // the B entries are consumed by the UMMA issuer
      const long m = (long)mt * 128 + q * 32 + lane;
      const OutT rt = (row_term && m < M) ? row_term[m] : (OutT)0;
      const bool vec_ok = (ldc * sizeof(OutT)) % 16 == 0 && (reinterpret_cast<uintptr_t>(Cout) & 15) == 0;
      for (int nt = n0; nt < n1; nt++, tile_seq++) {
        const int buf = (int)(tile_seq & 1);
        mbar_wait(sm.acc_full(buf), (uint32_t)((tile_seq >> 1) & 1));
        tc_fence_after();
#pragma unroll 1
        for (int c32 = 0; c32 < 2; c32++) {
          uint32_t r32[32];
          tmem_ld32(tmem_base + lane_addr + kColAcc + buf * 128 + half * 64 + c32 * 32, r32);
          tmem_wait_ld();
          if (c32 == 1) {  // the accumulator is in registers: the issuer may reuse the buffer
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(sm.acc_empty(buf));
          }
          const long n_base = (long)nt * 128 + half * 64 + c32 * 32;
          if (m < M) {
            OutT *dst = Cout + (size_t)m * ldc + n_base;
            OutT o[32];
            const bool full = n_base + 32 <= N;
#pragma unroll
            for (int j = 0; j < 32; j++) {
              // (the same address in every lane: one broadcast transaction)
              const OutT ct = (col_term && (full || n_base + j < N)) ? __ldg(col_term + n_base + j) : (OutT)0;
              o[j] = alpha * (OutT)__uint_as_float(r32[j]) + (rt + ct);
            }
            if (vec_ok && full) {
              constexpr int kPer = 16 / (int)sizeof(OutT);  // elements per 16-byte store
#pragma unroll
              for (int j = 0; j < 32; j += kPer)
                *reinterpret_cast<uint4 *>(dst + j) = *reinterpret_cast<const uint4 *>(&o[j]);
            } else {
#pragma unroll
              for (int j = 0; j < 32; j++)
                if (n_base + j < N) dst[j] = o[j];
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

}  // namespace

size_t gemm_split_panel_bytes(long rows, int K) {
  const long tiles = (rows + 127) / 128;
  const int nchunk = (K + 63) / 64;
  return (size_t)tiles * nchunk * kChunkBytes;
}

// power-of-two scale that brings max |X| into [0.5, 1)
static lr_status operand_scale(const double *dX, size_t ld, long rows, int K, double *d_tmp, double *scale) {
  Engine &e = engine();
  LR_CUDA(cudaMemsetAsync(d_tmp, 0, sizeof(double), e.stream));
  const long blocks = std::min<long>(1024, ceil_div(rows * (long)K, 256));
  k_absmax<<<(unsigned)std::max<long>(1, blocks), 256, 0, e.stream>>>(dX, ld, rows, K, d_tmp);
  LR_CHECK_LAUNCH();
  double h = 0.0;
  LR_CUDA(cudaMemcpyAsync(&h, d_tmp, sizeof(double), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  if (!(h > 0.0) || !std::isfinite(h)) {
    *scale = 1.0;
    return LR_OK;
  }
  int ex = 0;
  std::frexp(h, &ex);  // h = f * 2^ex, f in [0.5, 1)
  *scale = std::ldexp(1.0, -ex);
  return LR_OK;
}

lr_status gemm_split_prepare(const double *dX, size_t ld, long rows, int K, unsigned char *d_panels,
                             double *d_tmp, double *scale_out) {
  Engine &e = engine();
  LR_REQUIRE(K >= 1 && K <= 256, "gemm_split: K = %d outside [1, 256]", K);
  lr_status st = operand_scale(dX, ld, rows, K, d_tmp, scale_out);
  if (st != LR_OK) return st;
  const int nchunk = (K + 63) / 64;
  const long rows_pad = (rows + 127) / 128 * 128;
  const long threads = rows_pad * nchunk * 8;
  k_split_panels<<<(unsigned)ceil_div(threads, 256), 256, 0, e.stream>>>(dX, ld, rows, K, nchunk, *scale_out,
                                                                        d_panels);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

template <typename OutT>
lr_status gemm_split_run(const unsigned char *dAp, double scaleA, long M, const unsigned char *dBp,
                         double scaleB, long N, int K, OutT *dC, size_t ldc, const double *d_row_term,
                         const double *d_col_term) {
  Engine &e = engine();
  if (M <= 0 || N <= 0) return LR_OK;
  bool &done = e.attr_set[Engine::kAttrTvGemm];
  if (!done) {
    LR_CUDA(cudaFuncSetAttribute(k_gemm_split<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGSmem));
    LR_CUDA(cudaFuncSetAttribute(k_gemm_split<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGSmem));
    done = true;
  }
  Items items;
  items.mt_count = (int)((M + 127) / 128);
  items.nt_count = (int)((N + 127) / 128);
  // column-tile segments per row tile: enough items to occupy every SM, at least 8 column tiles each
  items.seg = std::max(1, std::min(items.nt_count / 8 > 0 ? items.nt_count / 8 : 1,
                                   ceil_div(2 * e.sm_count, items.mt_count)));
  items.n_items = items.mt_count * items.seg;
  const int nchunk = (K + 63) / 64, k16 = (K + 15) / 16;
  const int grid = std::min(e.sm_count, items.n_items);
  const OutT *rowp = nullptr, *colp = nullptr;
  DevBuf<float> rowf, colf;
  if constexpr (std::is_same<OutT, float>::value) {
    if (d_row_term) {
      LR_CUDA(rowf.alloc((size_t)M));
      k_to_float<<<ceil_div(M, 256), 256, 0, e.stream>>>(M, d_row_term, rowf.p);
      LR_CHECK_LAUNCH();
      rowp = rowf.p;
    }
    if (d_col_term) {
      LR_CUDA(colf.alloc((size_t)N));
      k_to_float<<<ceil_div(N, 256), 256, 0, e.stream>>>(N, d_col_term, colf.p);
      LR_CHECK_LAUNCH();
      colp = colf.p;
    }
  } else {
    rowp = d_row_term;
    colp = d_col_term;
  }
  k_gemm_split<OutT><<<grid, kGThreads, kGSmem, e.stream>>>(items, nchunk, k16, dAp, dBp, dC, ldc, M, N,
                                                            (OutT)(1.0 / (scaleA * scaleB)), rowp, colp);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

template lr_status gemm_split_run<float>(const unsigned char *, double, long, const unsigned char *, double, long,
                                         int, float *, size_t, const double *, const double *);
template lr_status gemm_split_run<double>(const unsigned char *, double, long, const unsigned char *, double, long,
                                          int, double *, size_t, const double *, const double *);

}  // namespace lr
