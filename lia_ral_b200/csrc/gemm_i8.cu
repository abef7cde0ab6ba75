// gemm_i8.cu -- fp64-grade GEMM on the INT8 tensor pipe (tcgen05.mma kind::i8, int32 accumulators in
// TMEM), the contraction behind the i-vector rows of the hot path:
//
//     C[m, n] = beta C[m, n] + alpha * sum_k A[m, k] B[n, k]
//
//   L_s  = I + sum_c N[s, c] TETt_c            (AccumulateTVStat.cpp:2126-2137, :1722-1733)
//   aux  = F_s (T Sigma^-1)^T                  (:2146-2153, :1741-1748)
//   A_c += sum_s N[s, c] (L_s^-1 + w_s w_s^T)  (:1775-1782)
//   Cmx += w_s F_s^T                           (:1784-1788)
//
// These feed a Cholesky solve whose condition number reaches 1e4, so the 22-bit split of gmm_tc.cu /
// gemm_split.cu is not enough (4e-7 in L costs 4e-3 in the i-vector).  Here every fp64 operand row
// is scaled by a power of two and cut into `s` signed 7-bit digits,
//     x = sigma * sum_i d_i 2^-(6 + 7 i) + O(sigma 2^-7s),     d_i in [-64, 64],
// each digit plane is an int8 matrix, and the product is the sum over digit pairs with i + j < s of
// EXACT int32 matrix products (|d d'| <= 2^12, so K (i + j + 1) 2^12 < 2^31 up to K = 32768 per
// work item).  Pairs with equal i + j share one accumulator ("class"); the s classes are recombined
// in fp64 in the epilogue.  s = 6 leaves 2^-42 relative to (row scale x column scale), the dropped
// classes contribute about the same: fp64-grade for these rows at 21 int8 products instead of one
// fp64 product -- on a pipe that is ~100 x faster than the fp64 one.
//
// Decomposition.  One CTA per SM, persistent over work items (row tile of 128, column tile of 64,
// K range).  All s class accumulators of the tile live in TMEM (s x 64 columns <= 512): that is
// what sets the tile width -- with K as the outer loop every digit plane is fetched ONCE per tile
// (class-outer order would stream the operands s (s + 1) / 2 times).  Per 128-deep K chunk the
// producer warp bulk-copies the s B planes (8 KB each, one stage of a 2-stage ring) and the s A
// planes (16 KB each, 5-slot ring, heavy and light planes interleaved); the issuer warp runs the
// (s - i) x 4 UMMAs (M128 N64 K32, SS) of plane A_i against B_0 .. B_{s-1-i}; eight epilogue warps
// fold the classes: int32 -> fp64 by the 2^52 trick, Horner over the classes, row / column scales,
// then one read-modify-write (or fp64 RED adds when the K range is split between CTAs).
#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "gemm_i8.cuh"
#include "tc_ptx.cuh"

namespace lr {
namespace {
using namespace tcptx;

constexpr int kThreads = 512;  // warp 0 producer, 1 UMMA issuer, 2 TMEM allocator, 4-7 movers, 8-15 epilogue
constexpr int kAStages = 4;    // shared-memory slots of A planes (the TMEM slots behind them add up to 4 more)
constexpr int kABytes = kI8TileM * 128;  // one digit plane of the A tile, one K chunk
constexpr int kBBytes = kI8TileN * 128;
constexpr int kChunkK = 128;

static size_t smem_bytes(int s) { return 1024 + (size_t)kAStages * kABytes + (size_t)(s <= 6 ? 3 : 2) * s * kBBytes + 256; }

// ------------------------------------------------------------------ operand preparation
// max |x| per row, as the bit pattern of the (non-negative) double
template <bool kKContig>
__global__ void __launch_bounds__(256)
k_i8_rowmax(const double *__restrict__ X, size_t stride_row, size_t stride_k, long rows, long K,
            unsigned long long *__restrict__ maxbits) {
  if (kKContig) {
    // one warp per row
    const long row = (long)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (row >= rows) return;
    const double *p = X + (size_t)row * stride_row;
    double m = 0.0;
    for (long k = threadIdx.x & 31; k < K; k += 32) m = fmax(m, fabs(p[k]));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0) maxbits[row] = (unsigned long long)__double_as_longlong(m);
  } else {
    // one thread per row and 64-deep K segment (lanes = adjacent rows: coalesced)
    const long row = (long)blockIdx.x * 256 + threadIdx.x;
    if (row >= rows) return;
    const long k0 = (long)blockIdx.y * 64, k1 = min(K, k0 + 64);
    double m = 0.0;
    for (long k = k0; k < k1; k++) m = fmax(m, fabs(X[(size_t)row * stride_row + (size_t)k * stride_k]));
    if (m > 0.0) atomicMax(maxbits + row, (unsigned long long)__double_as_longlong(m));
  }
}

// digit planes.  Panel layout: [row tile][K chunk][plane i][tile_rows x 128 B], every plane in the
// canonical 128-byte-swizzle K-major layout (8-row atoms of 1 KB).  One thread per (row, 16 k).
template <bool kKContig>
__global__ void __launch_bounds__(256)
k_i8_planes(const double *__restrict__ X, size_t stride_row, size_t stride_k, long rows, long K,
            long rows_pad, int nchunk, int s, int tile_rows, const unsigned long long *__restrict__ maxbits,
            double *__restrict__ scale_out, unsigned char *__restrict__ out) {
  const long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long groups = (long)nchunk * 8;
  long row, grp;
  if (kKContig) {
    row = gid / groups;
    grp = gid - row * groups;
  } else {
    grp = gid / rows_pad;
    row = gid - grp * rows_pad;
  }
  if (row >= rows_pad || grp >= groups) return;
  double sigma = 0.0, inv = 0.0;
  if (row < rows) {
    const double m = __longlong_as_double((long long)maxbits[row]);
    if (m > 0.0 && m < 1.0e300) {
      int ex;
      frexp(m, &ex);  // m = f 2^ex, f in [0.5, 1)
      sigma = ldexp(1.0, ex);
      inv = ldexp(1.0, -ex);
    }
  }
  if (grp == 0) scale_out[row] = sigma;
  __align__(16) signed char dig[kI8MaxSlices][16];
#pragma unroll
  for (int e = 0; e < 16; e++) {
    const long k = grp * 16 + e;
    double t = 0.0;
    if (row < rows && k < K) t = X[(size_t)row * stride_row + (size_t)k * stride_k] * inv * 64.0;
#pragma unroll
    for (int i = 0; i < kI8MaxSlices; i++) {
      if (i < s) {
        const double d = rint(t);
        dig[i][e] = (signed char)(int)d;
        t = (t - d) * 128.0;
      }
    }
  }
  const long tile = row / tile_rows;
  const int r = (int)(row - tile * tile_rows);
  const int kc = (int)(grp >> 3), j = (int)(grp & 7);
  const size_t plane = (size_t)tile_rows * 128;
  unsigned char *base = out + ((size_t)tile * nchunk + kc) * s * plane;
  const uint32_t off = (uint32_t)r * 128u + (uint32_t)((j ^ (r & 7)) << 4);
#pragma unroll
  for (int i = 0; i < kI8MaxSlices; i++)
    if (i < s) *reinterpret_cast<uint4 *>(base + i * plane + off) = *reinterpret_cast<const uint4 *>(dig[i]);
}

// ------------------------------------------------------------------ the GEMM kernel
// D[tmem] (+)= A[tmem] B[smem], int8 x int8 -> int32
__device__ __forceinline__ void umma_ts_i8(uint32_t d_tmem, uint32_t a_tmem, uint64_t bdesc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "r"(a_tmem), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// instruction descriptor: s8 x s8 -> s32, M x N, both operands K-major
__host__ __device__ constexpr uint32_t make_idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// non-blocking look at an mbarrier phase
__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
      "selp.b32 %0, 1, 0, p;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

// TMEM: class accumulators in columns [0, 64 s), A planes in slots of 32 columns (128 int8 per lane) above
__host__ __device__ constexpr int tmem_slots(int s) { return (512 - 64 * s) / 32 < 4 ? (512 - 64 * s) / 32 : 4; }
__host__ __device__ constexpr int b_stages(int s) { return s <= 6 ? 3 : 2; }

struct Bars {
  uint32_t base, bar;
  int s;
  __device__ __forceinline__ uint32_t a_stage(int i) const { return base + i * kABytes; }
  __device__ __forceinline__ uint32_t b_stage(int i) const { return base + kAStages * kABytes + i * s * kBBytes; }
  __device__ __forceinline__ uint32_t a_full(int i) const { return bar + 8 * i; }        // kAStages <= 4
  __device__ __forceinline__ uint32_t a_empty(int i) const { return bar + 32 + 8 * i; }
  __device__ __forceinline__ uint32_t b_full(int i) const { return bar + 64 + 8 * i; }   // <= 3
  __device__ __forceinline__ uint32_t b_empty(int i) const { return bar + 88 + 8 * i; }
  __device__ __forceinline__ uint32_t at_full(int i) const { return bar + 112 + 8 * i; }  // <= 4
  __device__ __forceinline__ uint32_t at_empty(int i) const { return bar + 144 + 8 * i; }
  __device__ __forceinline__ uint32_t acc_full() const { return bar + 176; }
  __device__ __forceinline__ uint32_t acc_empty() const { return bar + 184; }
  __device__ __forceinline__ uint32_t tmem_slot() const { return bar + 192; }
};

// work item -> (row tile, column tile, K chunk range); row tiles fastest so that the CTAs running at
// one moment share a few B panels and all of A through L2
struct Sched {
  int mt_count, nt_count, ksplit, nchunk, n_items;
  __host__ __device__ void get(int it, int &mt, int &nt, int &c0, int &c1) const {
    mt = it % mt_count;
    const int rest = it / mt_count;
    const int ks = rest % ksplit;
    nt = rest / ksplit;
    c0 = (int)((long)nchunk * ks / ksplit);
    c1 = (int)((long)nchunk * (ks + 1) / ksplit);
  }
};

// order in which the A planes of a chunk are consumed: 0, s-1, 1, s-2, ... (s - i products each:
// heavy and light planes alternate, so the rings drain at an even rate)
__host__ __device__ constexpr int plane_order(int ii, int s) { return (ii & 1) ? s - 1 - (ii >> 1) : (ii >> 1); }

__device__ __forceinline__ double i32_to_f64(uint32_t bits) {
  // exact: 2^52 + 2^31 + x as a double whose low word is x + 2^31, minus the constant
  return __hiloint2double(0x43300000, (int)(bits ^ 0x80000000u)) - 4503601774854144.0;
}

// Warp roles (512 threads): 0 bulk-copy producer, 1 UMMA issuer, 2 TMEM allocator, 4-7 movers (one per
// TMEM lane quarter: A plane shared memory -> registers -> TMEM), 8-15 epilogue (lane quarter x column half).
// The A operand of every UMMA comes from TMEM (TS mode): the SS form of M128 N64 K32 is bound by the
// 128 B/clk shared-memory operand read (48 clk per UMMA measured, scripts/umma_i8_probe.cu), the TS form
// runs at the tensor pipe's 32 clk.
template <int S>  // digit planes per operand (compile time: the issue loop is fully unrolled)
__global__ void __launch_bounds__(kThreads, 1)
k_gemm_i8(Sched sched, const unsigned char *__restrict__ Ap, const unsigned char *__restrict__ Bp,
          const double *__restrict__ scaleA, const double *__restrict__ scaleB, double *__restrict__ Cout,
          size_t ldc, long M, long N, double alpha, double beta) {
  extern __shared__ unsigned char smem_raw[];
  constexpr int s = S;
  constexpr int kBSt = b_stages(S), kTSlots = tmem_slots(S);
  constexpr uint32_t kAcol = 64 * S;  // first TMEM column of the A plane slots
  Bars sm;
  sm.s = s;
  sm.base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  sm.bar = sm.base + kAStages * kABytes + kBSt * s * kBBytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kAStages; i++) {
      mbar_init(sm.a_full(i), 1);
      mbar_init(sm.a_empty(i), 4);  // the four mover warps have read the plane
    }
    for (int i = 0; i < kBSt; i++) {
      mbar_init(sm.b_full(i), 1);
      mbar_init(sm.b_empty(i), 1);
    }
    for (int i = 0; i < kTSlots; i++) {
      mbar_init(sm.at_full(i), 4);   // the four mover warps have written their lane quarter
      mbar_init(sm.at_empty(i), 1);  // the UMMAs reading the slot have completed
    }
    mbar_init(sm.acc_full(), 1);
    mbar_init(sm.acc_empty(), 8);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(sm.tmem_slot(), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_slot()));
  constexpr uint32_t idesc = make_idesc_i8(kI8TileM, kI8TileN);

  if (warp == 0) {
    // ---- producer: per chunk the s B planes (one stage), then the A planes in consumption order
    const bool leader = elect_one();
    long aseq = 0, bseq = 0;
    for (int it = blockIdx.x; it < sched.n_items; it += gridDim.x) {
      int mt, nt, c0, c1;
      sched.get(it, mt, nt, c0, c1);
      for (int kc = c0; kc < c1; kc++) {
        const int bst = (int)(bseq % kBSt);
        mbar_wait(sm.b_empty(bst), (uint32_t)(((bseq / kBSt) & 1) ^ 1));
        if (leader) {
          const unsigned char *src = Bp + ((size_t)nt * sched.nchunk + kc) * s * kBBytes;
          mbar_expect_tx(sm.b_full(bst), (uint32_t)(s * kBBytes));
#pragma unroll
          for (int j = 0; j < s; j++)
            bulk_g2s(sm.b_stage(bst) + j * kBBytes, src + (size_t)j * kBBytes, kBBytes, sm.b_full(bst));
        }
        __syncwarp();
        bseq++;
#pragma unroll
        for (int ii = 0; ii < s; ii++, aseq++) {
          const int i = plane_order(ii, s);
          const int ast = (int)(aseq % kAStages);
          mbar_wait(sm.a_empty(ast), (uint32_t)(((aseq / kAStages) & 1) ^ 1));
          if (leader) {
            const unsigned char *src = Ap + (((size_t)mt * sched.nchunk + kc) * s + i) * kABytes;
            mbar_expect_tx(sm.a_full(ast), kABytes);
            bulk_g2s(sm.a_stage(ast), src, kABytes, sm.a_full(ast));
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == 1) {
    // ---- UMMA issuer.  Everything that can be is a compile-time constant, and the barrier of the NEXT
    // plane is looked at (non-blocking) before the current plane's UMMAs are issued: the single issuing
    // thread has 32 clk per UMMA, a blocking wait between two planes idles the pipe.
    const bool leader = elect_one();
    long aseq = 0, bseq = 0, tile_seq = 0;
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t b_lo0 = ((sm.b_stage(0) >> 4) & 0x3FFFu) | (1u << 16);
    for (int it = blockIdx.x; it < sched.n_items; it += gridDim.x, tile_seq++) {
      int mt, nt, c0, c1;
      sched.get(it, mt, nt, c0, c1);
      if (tile_seq > 0) mbar_wait(sm.acc_empty(), (uint32_t)((tile_seq - 1) & 1));
      tc_fence_after();
      bool have = false;  // the barrier of plane `aseq` has already been seen complete
      for (int kc = c0; kc < c1; kc++, bseq++) {
        const int bst = (int)(bseq % kBSt);
        mbar_wait(sm.b_full(bst), (uint32_t)((bseq / kBSt) & 1));
        const uint32_t b_lo = b_lo0 + (uint32_t)bst * (uint32_t)((s * kBBytes) >> 4);
        const uint32_t later = kc > c0 ? 1u : 0u;  // plane 0 comes first and opens every class of the item
#pragma unroll
        for (int ii = 0; ii < s; ii++, aseq++) {
          const int i = plane_order(ii, s);
          const int ts = (int)(aseq % kTSlots);
          if (!have) mbar_wait(sm.at_full(ts), (uint32_t)((aseq / kTSlots) & 1));
          tc_fence_after();
          // look ahead (the next plane of this item, if any)
          const bool more = (ii + 1 < s) || (kc + 1 < c1);
          have = more && __all_sync(0xffffffffu, mbar_test(sm.at_full((int)((aseq + 1) % kTSlots)),
                                                           (uint32_t)(((aseq + 1) / kTSlots) & 1)));
          if (leader) {
            const uint32_t a_t = tmem_base + kAcol + (uint32_t)ts * 32u;
#pragma unroll
            for (int j = 0; j < s - i; j++) {
#pragma unroll
              for (int kk = 0; kk < 4; kk++) {
                const uint64_t bdesc = ((uint64_t)desc_hi << 32) | (uint64_t)(b_lo + j * (kBBytes >> 4) + kk * 2);
                umma_ts_i8(tmem_base + (i + j) * kI8TileN, a_t + kk * 8, bdesc, idesc,
                           (i == 0 && kk == 0) ? later : 1u);
              }
            }
            umma_commit(sm.at_empty(ts));
          }
          __syncwarp();
        }
        if (leader) umma_commit(sm.b_empty(bst));
        __syncwarp();
      }
      if (leader) umma_commit(sm.acc_full());
      __syncwarp();
    }
  } else if (warp >= 4 && warp < 8) {
    // ---- movers: lane quarter q of every A plane, shared memory -> registers -> TMEM slot
    const int q = warp & 3;
    const int r = q * 32 + lane;  // row of the tile = TMEM lane
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    long aseq = 0;
    for (int it = blockIdx.x; it < sched.n_items; it += gridDim.x) {
      int mt, nt, c0, c1;
      sched.get(it, mt, nt, c0, c1);
      for (long n = (long)(c1 - c0) * s; n > 0; n--, aseq++) {
        const int ast = (int)(aseq % kAStages), ts = (int)(aseq % kTSlots);
        mbar_wait(sm.a_full(ast), (uint32_t)((aseq / kAStages) & 1));
        uint32_t v[32];
#pragma unroll
        for (int c = 0; c < 8; c++) {
          const uint32_t a = sm.a_stage(ast) + (uint32_t)r * 128u + (uint32_t)((c ^ (r & 7)) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v[4 * c]), "=r"(v[4 * c + 1]), "=r"(v[4 * c + 2]), "=r"(v[4 * c + 3])
                       : "r"(a));
        }
        // (the loads above have returned: they feed the stores below) -- the shared-memory slot is free
        mbar_wait(sm.at_empty(ts), (uint32_t)(((aseq / kTSlots) & 1) ^ 1));
        tc_fence_after();
        {
          uint32_t lo16[16], hi16[16];
#pragma unroll
          for (int j = 0; j < 16; j++) {
            lo16[j] = v[j];
            hi16[j] = v[16 + j];
          }
          tmem_st16(tmem_base + lane_addr + kAcol + (uint32_t)ts * 32u, lo16);
          tmem_st16(tmem_base + lane_addr + kAcol + (uint32_t)ts * 32u + 16u, hi16);
        }
        tmem_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) {
          mbar_arrive(sm.a_empty(ast));
          mbar_arrive(sm.at_full(ts));
        }
      }
    }
  } else if (warp >= 8) {
    // ---- epilogue: TMEM lane quarter q (rows), column half h of the 64-wide tile, 16 columns at a time
    const int q = warp & 3, h = (warp - 8) >> 2;
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    long tile_seq = 0;
    const bool split = sched.ksplit > 1;
    for (int it = blockIdx.x; it < sched.n_items; it += gridDim.x, tile_seq++) {
      int mt, nt, c0, c1;
      sched.get(it, mt, nt, c0, c1);
      const long m = (long)mt * kI8TileM + q * 32 + lane;
      const double sa = m < M ? scaleA[m] * alpha * (1.0 / 4096.0) : 0.0;  // digit weights 2^-6 x 2^-6
      mbar_wait(sm.acc_full(), (uint32_t)(tile_seq & 1));
      tc_fence_after();
      double v[2][16];
#pragma unroll
      for (int g = 0; g < 2; g++) {
#pragma unroll
        for (int j = 0; j < 16; j++) v[g][j] = 0.0;
#pragma unroll
        for (int d = s - 1; d >= 0; d--) {  // Horner over the classes: v = v 2^-7 + acc_d
          uint32_t r16[16];
          tmem_ld16(tmem_base + lane_addr + d * kI8TileN + h * 32 + g * 16, r16);
          tmem_wait_ld();
#pragma unroll
          for (int j = 0; j < 16; j++) v[g][j] = fma(v[g][j], 0.0078125, i32_to_f64(r16[j]));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.acc_empty());
#pragma unroll
      for (int g = 0; g < 2; g++) {
        const long n_base = (long)nt * kI8TileN + h * 32 + g * 16;
        if (m < M && n_base < N) {
          double *dst = Cout + (size_t)m * ldc + n_base;
          const bool full = n_base + 16 <= N;
          const bool vec_ok = full && ((ldc * sizeof(double)) % 16 == 0) && ((reinterpret_cast<uintptr_t>(Cout) & 15) == 0);
#pragma unroll
          for (int j = 0; j < 16; j++) v[g][j] *= sa * __ldg(scaleB + n_base + j);  // scaleB is padded to the tile
          if (split) {
#pragma unroll
            for (int j = 0; j < 16; j++)
              if (full || n_base + j < N) atomicAdd(dst + j, v[g][j]);
          } else if (vec_ok) {
            if (beta != 0.0) {
#pragma unroll
              for (int j = 0; j < 16; j += 2) {
                const double2 c = *reinterpret_cast<const double2 *>(dst + j);
                v[g][j] = fma(beta, c.x, v[g][j]);
                v[g][j + 1] = fma(beta, c.y, v[g][j + 1]);
              }
            }
#pragma unroll
            for (int j = 0; j < 16; j += 2) *reinterpret_cast<double2 *>(dst + j) = make_double2(v[g][j], v[g][j + 1]);
          } else {
#pragma unroll
            for (int j = 0; j < 16; j++)
              if (n_base + j < N) dst[j] = (beta != 0.0 ? beta * dst[j] : 0.0) + v[g][j];
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}

__global__ void k_scale_rows(long M, long N, double beta, double *__restrict__ C, size_t ldc) {
  const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= M * N) return;
  const long m = i / N, n = i - m * N;
  C[(size_t)m * ldc + n] = beta == 0.0 ? 0.0 : beta * C[(size_t)m * ldc + n];
}

}  // namespace

size_t gemm_i8_panel_bytes(long rows, long K, int s, int tile_rows) {
  const long tiles = (rows + tile_rows - 1) / tile_rows;
  const long nchunk = (K + kChunkK - 1) / kChunkK;
  return (size_t)tiles * nchunk * s * tile_rows * 128;
}

size_t gemm_i8_scale_count(long rows, int tile_rows) {
  return (size_t)((rows + tile_rows - 1) / tile_rows) * tile_rows;
}

lr_status gemm_i8_prepare(const double *dX, size_t stride_row, size_t stride_k, long rows, long K, int s,
                          int tile_rows, unsigned char *d_panels, double *d_scale,
                          const unsigned long long *d_rowmax) {
  Engine &e = engine();
  LR_REQUIRE(s >= 1 && s <= kI8MaxSlices, "gemm_i8: %d digit planes outside [1, %d]", s, kI8MaxSlices);
  LR_REQUIRE(tile_rows == kI8TileM || tile_rows == kI8TileN, "gemm_i8: tile_rows %d", tile_rows);
  LR_REQUIRE(stride_row == 1 || stride_k == 1, "gemm_i8: one operand stride must be 1");
  LR_REQUIRE(rows >= 1 && K >= 1, "gemm_i8: empty operand");
  const long rows_pad = (long)gemm_i8_scale_count(rows, tile_rows);
  const int nchunk = (int)((K + kChunkK - 1) / kChunkK);
  DevBuf<unsigned long long> maxbuf;
  const bool k_contig = stride_k == 1;
  const unsigned long long *maxp = d_rowmax;  // the caller may already know the row maxima
  if (!maxp) {
    LR_CUDA(maxbuf.alloc((size_t)rows));
    if (k_contig) {
      k_i8_rowmax<true><<<(unsigned)ceil_div(rows, 8), 256, 0, e.stream>>>(dX, stride_row, stride_k, rows, K, maxbuf.p);
    } else {
      LR_CUDA(cudaMemsetAsync(maxbuf.p, 0, (size_t)rows * sizeof(unsigned long long), e.stream));
      dim3 grid((unsigned)ceil_div(rows, 256), (unsigned)ceil_div(K, 64));
      k_i8_rowmax<false><<<grid, 256, 0, e.stream>>>(dX, stride_row, stride_k, rows, K, maxbuf.p);
    }
    LR_CHECK_LAUNCH();
    maxp = maxbuf.p;
  }
  const long threads = rows_pad * nchunk * 8;
  const unsigned blocks = (unsigned)((threads + 255) / 256);
  if (k_contig)
    k_i8_planes<true><<<blocks, 256, 0, e.stream>>>(dX, stride_row, stride_k, rows, K, rows_pad, nchunk, s,
                                                     tile_rows, maxp, d_scale, d_panels);
  else
    k_i8_planes<false><<<blocks, 256, 0, e.stream>>>(dX, stride_row, stride_k, rows, K, rows_pad, nchunk, s,
                                                      tile_rows, maxp, d_scale, d_panels);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

lr_status gemm_i8_run(const unsigned char *dAp, const double *dAscale, long M, const unsigned char *dBp,
                      const double *dBscale, long N, long K, int s, double alpha, double beta, double *dC,
                      size_t ldc) {
  Engine &e = engine();
  if (M <= 0 || N <= 0) return LR_OK;
  LR_REQUIRE(s >= 1 && s <= kI8MaxSlices, "gemm_i8: %d digit planes outside [1, %d]", s, kI8MaxSlices);
  LR_REQUIRE(K >= 1, "gemm_i8: K = %ld", K);
  using KernelT = void (*)(Sched, const unsigned char *, const unsigned char *, const double *, const double *,
                           double *, size_t, long, long, double, double);
  static const KernelT kernels[kI8MaxSlices + 1] = {nullptr,      k_gemm_i8<1>, k_gemm_i8<2>, k_gemm_i8<3>,
                                                    k_gemm_i8<4>, k_gemm_i8<5>, k_gemm_i8<6>, k_gemm_i8<7>};
  const KernelT kern = kernels[s];
  bool &done = e.attr_set[Engine::kAttrGemmI8];
  if (!done) {
    for (int i = 1; i <= kI8MaxSlices; i++)
      LR_CUDA(cudaFuncSetAttribute(kernels[i], cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes(i)));
    done = true;
  }
  Sched sc;
  sc.mt_count = (int)((M + kI8TileM - 1) / kI8TileM);
  sc.nt_count = (int)((N + kI8TileN - 1) / kI8TileN);
  sc.nchunk = (int)((K + kChunkK - 1) / kChunkK);
  // int32 exactness: s K 2^12 < 2^31 per item -> at most 256 chunks (K = 32768) per item; and enough
  // items to fill the machine twice when the tile count alone does not
  const long tiles = (long)sc.mt_count * sc.nt_count;
  int ksplit = ceil_div(sc.nchunk, 256);
  if (tiles < 2L * e.sm_count) ksplit = std::max<int>(ksplit, std::min<long>(sc.nchunk / 4, ceil_div(2L * e.sm_count, tiles)));
  sc.ksplit = std::max(1, std::min(ksplit, sc.nchunk));
  sc.n_items = (int)(tiles * sc.ksplit);
  if (sc.ksplit > 1) {
    LR_REQUIRE(beta == 0.0 || beta == 1.0, "gemm_i8: beta must be 0 or 1 when the K range is split");
    if (beta == 0.0) {
      k_scale_rows<<<(unsigned)((M * N + 255) / 256), 256, 0, e.stream>>>(M, N, 0.0, dC, ldc);
      LR_CHECK_LAUNCH();
    }
  }
  const int grid = std::min(e.sm_count, sc.n_items);
  kern<<<grid, kThreads, smem_bytes(s), e.stream>>>(sc, dAp, dBp, dAscale, dBscale, dC, ldc, M, N, alpha, beta);
  LR_CUDA(cudaGetLastError());
  count_launch();
  return LR_OK;
}

}  // namespace lr
