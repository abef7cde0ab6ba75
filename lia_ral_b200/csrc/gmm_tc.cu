// gmm_tc.cu -- tcgen05 / TMEM implementation of the frames x components passes (sm_100a).
//
// Formulation.  In the globally normalised space xh = (x - g) / s (g, s = mixture mean / std per
// dimension) the log2 joint likelihood is an inner product over K = 128 columns
//     S2[t,c] = sum_k A[t,k] W[c,k],   A[t,:] = [xh (60) | 1 1 1 | 0 || xh^2 (60) | 0000]
//     W[c,:] = [-2 alpha beta | K_c split in 3 | 0 || -alpha^2 | 0],  alpha = s sa, beta = g sa + nm
// Both operands are split hi + lo into two fp16 values (22 significand bits) and the contraction
// is three fp16 UMMAs with fp32 accumulation in TMEM:  A_hi W_hi + A_lo W_hi + A_hi W_lo.
//
// Data layout.  k_tc_convert writes the frame operand once per call as 64 KB tiles of 128
// frames: four 16 KB panels [xh_hi,1 | xh^2_hi | xh_lo | xh^2_lo], each 128 rows x 64 fp16 in the
// canonical 128-byte-swizzle layout, so one cp.async.bulk per panel lands it ready for UMMA --
// as the K-major operand of the likelihood GEMM and, read MN-major, as the B operand of the
// statistics GEMM.  Each CTA owns a slice of 128 components whose weights (64 KB) stay resident
// in shared memory for the whole kernel; the grid is (slices x groups), every group streaming a
// contiguous range of frame tiles through an mbarrier ring (pass 1: two 64 KB tiles, pass 2: five
// 32 KB half tiles).
//
// Pass 1 (k_tc_lse):   D[t, c] = A W^T  (frames on TMEM lanes), per-frame online (max, sum) over
//                      the slice's 128 columns -> partial log-sum-exp per (slice, frame).
// Pass 2 (k_tc_acc):   D[c, t] = W A^T  (components on lanes), g' = 2^(S - lse + 14) packed to
//                      fp16 back into TMEM, then the TS-UMMA  F[c, d] += g'[c, t] A[t, d]  with the
//                      frame tile as MN-major B operand: N, sum g xh (hi, lo), sum g xh^2 (hi, lo)
//                      accumulate in TMEM across the tiles of a run and are flushed in fp64.
#include <cuda_fp16.h>

#include "gmm.cuh"
#include "tc_ptx.cuh"

namespace lr {

namespace {

constexpr int kTile = 128;                     // frames per tile
constexpr int kSlice = 128;                    // components per CTA
constexpr int kPanelBytes = 128 * 128;         // 128 rows x 64 fp16
constexpr int kTileBytes = 4 * kPanelBytes;    // 64 KB
constexpr int kStages = 2;
constexpr unsigned kPadIndex = 0xFFFFFFFFu;
constexpr int kOneCol = 60;                    // columns 60..62 of panel a carry 1.0
constexpr float kGammaShift = 14.f;            // posteriors are stored as fp16(2^14 g)
constexpr float kXClamp = 240.f;               // |xh| clamp (xh^2 must stay below fp16 max)
// fp32 TMEM partial sums are flushed to fp64 at least this often (8 k frames).  The tensor core
// TRUNCATES the fp32 accumulator at every UMMA (8 per tile), a systematic -2^-25.8 relative bias per
// accumulation step (measured on B200: 128-tile runs sat 1.4e-5 below 27-tile runs), so the run
// length bounds the bias: 64 tiles -> <= 9e-6 relative.  Statistics-pass time per 1 M frames at
// 32 / 64 / 128 tiles: 1.80 / 1.68 / 1.63 ms.
constexpr int kMaxRunTiles = 64;
constexpr int kTcThreads = 384;                 // warps: 0 bulk-copy producer, 1 MMA issuer, 2 TMEM allocator, 4-11 epilogue
constexpr int kEpiWarps = 8;
constexpr int kStageFloats = 32 * 32;             // per-warp flush staging: 32 comps x 32 dims

using namespace tcptx;

struct TileInfo {
  int row;    // statistics row the run is flushed into
  int flags;  // bit 0: first tile of a run (accumulator starts from zero); bit 1: flush after
};

// split v into fp16 hi + lo
__device__ __forceinline__ void split2(float v, __half &hi, __half &lo) {
  hi = __float2half_rn(v);
  lo = __float2half_rn(v - __half2float(hi));
}

// ------------------------------------------------------------------ operand preparation
// g, s per dimension: mixture mean / standard deviation (fp64)
__global__ void k_tc_norm(int C, int D, const double *__restrict__ w,
                          const double *__restrict__ mean, const double *__restrict__ cov,
                          double *__restrict__ g, double *__restrict__ s) {
  int i = blockIdx.x;
  if (i >= D) return;
  __shared__ double sh[3][256];
  double a = 0.0, b = 0.0, ws = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) {
    double m = mean[(size_t)c * D + i];
    a += w[c] * m;
    b += w[c] * (cov[(size_t)c * D + i] + m * m);
    ws += w[c];
  }
  sh[0][threadIdx.x] = a;
  sh[1][threadIdx.x] = b;
  sh[2][threadIdx.x] = ws;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o)
      for (int k = 0; k < 3; k++) sh[k][threadIdx.x] += sh[k][threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) {
    double wsum = sh[2][0] > 0.0 ? sh[2][0] : 1.0;
    double gi = sh[0][0] / wsum;
    double var = sh[1][0] / wsum - gi * gi;
    double si = var > 1e-300 ? sqrt(var) : 1.0;
    g[i] = gi;  // candidates: k_tc_norm_adopt decides
    s[i] = si;
  }
}

// weights of one component -> [slice][hi a | hi b | lo a | lo b] panels.  One thread per comp.
__global__ void k_tc_weights(int C, int D, int Cp, const double *__restrict__ w,
                             const double *__restrict__ mean, const double *__restrict__ covinv,
                             const double *__restrict__ cst, const double *__restrict__ g,
                             const double *__restrict__ s, unsigned char *__restrict__ Wp,
                             int *__restrict__ flag) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  unsigned char *base = Wp + (size_t)(c / kSlice) * (4 * kPanelBytes);
  const int row = c % kSlice;
  auto put = [&](int panel, int col, __half v) {
    *reinterpret_cast<__half *>(base + (size_t)panel * kPanelBytes + panel_off(row, col)) = v;
  };
  const __half z = __float2half_rn(0.f);
  for (int p = 0; p < 4; p++)
    for (int col = 0; col < 64; col++) put(p, col, z);
  if (c >= C) {
    put(0, kOneCol, __float2half_rn(-60000.f));  // padding component: S2 = -60000 -> 2^S2 = 0
    return;
  }
  const double kHalfLog2e = 0.72134752044448170368;
  double kc = log2(w[c]) + log2(cst[c]);
  bool bad = !(w[c] > 0.0) || !isfinite(kc);
  double maxabs = 0.0;
  for (int i = 0; i < D; i++) {
    double sa = sqrt(kHalfLog2e * covinv[(size_t)c * D + i]);
    double alpha = s[i] * sa;
    double beta = (g[i] - mean[(size_t)c * D + i]) * sa;
    double w1 = -2.0 * alpha * beta, w2 = -alpha * alpha;
    kc -= beta * beta;
    maxabs = fmax(maxabs, fmax(fabs(w1), fabs(w2)));
    __half h, l;
    split2((float)w1, h, l);
    // second-order correction of the float cast: lo also absorbs (w1 - float(w1))
    l = __float2half_rn((float)(w1 - (double)__half2float(h)));
    put(0, i, h);
    put(2, i, l);
    h = __float2half_rn((float)w2);
    l = __float2half_rn((float)(w2 - (double)__half2float(h)));
    put(1, i, h);
    put(3, i, l);
  }
  if (bad) {
    put(0, kOneCol, __float2half_rn(-60000.f));
    return;
  }
  if (fabs(kc) > 30000.0 || maxabs > 30000.0) atomicExch(flag, 1);  // outside the fp16 range
  __half k0 = __float2half_rn((float)kc);
  double r1 = kc - (double)__half2float(k0);
  __half k1 = __float2half_rn((float)r1);
  double r2 = r1 - (double)__half2float(k1);
  __half k2 = __float2half_rn((float)r2);
  put(0, kOneCol, k0);
  put(0, kOneCol + 1, k1);
  put(0, kOneCol + 2, k2);
}

// frames -> swizzled fp16 hi/lo tiles.  One thread per (row, 16-byte chunk).
__global__ void __launch_bounds__(256)
k_tc_convert(int D, const float *__restrict__ X, size_t ldx, const unsigned *__restrict__ index,
             long P, long P_pad, const float *__restrict__ gf, const float *__restrict__ rsf,
             unsigned char *__restrict__ Xh) {
  long gid = (long)blockIdx.x * blockDim.x + threadIdx.x;
  long p = gid >> 3;
  int j = (int)(gid & 7);
  if (p >= P_pad) return;
  long tile = p / kTile;
  int row = (int)(p - tile * kTile);
  bool valid = p < P;
  size_t fr = 0;
  if (valid) {
    if (index) {
      unsigned ix = index[p];
      valid = ix != kPadIndex;
      fr = ix;
    } else {
      fr = (size_t)p;
    }
  }
  __align__(16) __half ha[8], la[8], hb[8], lb[8];
#pragma unroll
  for (int e = 0; e < 8; e++) {
    int k = j * 8 + e;
    float xa = 0.f, xb = 0.f;
    if (valid) {
      if (k < D) {
        float v = (X[fr * ldx + k] - gf[k]) * rsf[k];
        v = fminf(fmaxf(v, -kXClamp), kXClamp);
        xa = v;
        xb = v * v;
      } else if (k >= kOneCol && k < kOneCol + 3) {
        xa = 1.f;
      }
    }
    split2(xa, ha[e], la[e]);
    split2(xb, hb[e], lb[e]);
    if (k >= kOneCol) la[e] = __float2half_rn(0.f);
  }
  unsigned char *t = Xh + (size_t)tile * kTileBytes;
  uint32_t off = (uint32_t)row * 128u + (uint32_t)((j ^ (row & 7)) << 4);
  *reinterpret_cast<uint4 *>(t + 0 * kPanelBytes + off) = *reinterpret_cast<const uint4 *>(ha);
  *reinterpret_cast<uint4 *>(t + 1 * kPanelBytes + off) = *reinterpret_cast<const uint4 *>(hb);
  *reinterpret_cast<uint4 *>(t + 2 * kPanelBytes + off) = *reinterpret_cast<const uint4 *>(la);
  *reinterpret_cast<uint4 *>(t + 3 * kPanelBytes + off) = *reinterpret_cast<const uint4 *>(lb);
}

// ------------------------------------------------------------------ shared kernel scaffolding
constexpr int kHStages = 5;                       // pass 2: ring of 64-frame half tiles
constexpr int kHalfPanel = kPanelBytes / 2;      // 64 rows x 128 B
constexpr int kHalfBytes = 4 * kHalfPanel;       // 32 KB
// 1 KB alignment slack + weights 64 KB + 5 half-tile stages (pass 1: 2 tile stages) + barriers.
// Pass 2's flush staging (8 warps x 4 KB) aliases the HI weights, which live in TMEM by then.
constexpr size_t kTcSmem = 1024 + 64 * 1024 + kHStages * (size_t)kHalfBytes + 256;

struct Smem {
  uint32_t w;           // weights slice: hi a | hi b | lo a | lo b
  uint32_t stage[kStages];     // pass 1: 2 x 64 KB
  uint32_t hstage[kHStages];   // pass 2: 5 x 32 KB (same memory)
  uint32_t full[kHStages], empty[kHStages];
  uint32_t s_full[3], s_empty[2], p_full[3], s_free[3];
  uint32_t f_full, f_empty, w_full;
  uint32_t tmem_slot;
};

__device__ __forceinline__ uint32_t carve_base(unsigned char *raw) {
  return (smem_u32(raw) + 1023u) & ~1023u;
}

__device__ __forceinline__ Smem carve(unsigned char *raw) {
  uint32_t base = (smem_u32(raw) + 1023u) & ~1023u;
  Smem s;
  s.w = base;
  for (int i = 0; i < kStages; i++) s.stage[i] = base + 64 * 1024 + i * kTileBytes;
  for (int i = 0; i < kHStages; i++) s.hstage[i] = base + 64 * 1024 + i * kHalfBytes;
  uint32_t b = base + 64 * 1024 + kHStages * kHalfBytes;
  static_assert(kHStages <= 5, "barrier block layout");
  for (int i = 0; i < kHStages; i++) {
    s.full[i] = b + 8 * i;
    s.empty[i] = b + 40 + 8 * i;
  }
  for (int i = 0; i < 3; i++) {
    s.s_full[i] = b + 80 + 8 * i;
    s.p_full[i] = b + 120 + 8 * i;
  }
  for (int i = 0; i < 2; i++) s.s_empty[i] = b + 104 + 8 * i;
  s.f_full = b + 144;
  s.f_empty = b + 152;
  s.w_full = b + 160;
  s.tmem_slot = b + 168;
  for (int i = 0; i < 3; i++) s.s_free[i] = b + 176 + 8 * i;
  return s;
}

// the likelihood GEMM of one tile: 3 products x 2 panels x 4 K-steps of 16
//   w_is_a: true  -> D[c, t] (A = weights, B = frames)   (pass 2)
//           false -> D[t, c] (A = frames,  B = weights)  (pass 1)
template <int NFRAMES>
__device__ __forceinline__ void issue_g1(uint32_t d_tmem, uint64_t w_desc0, uint64_t x_desc0,
                                         bool w_is_a, int nq = 6) {
  // frames panel = NFRAMES rows x 128 B; with the weights as A the frames are the N dimension.
  // w_desc0 / x_desc0: K-major SW128 descriptors of the first weights / frames panel.
  constexpr uint32_t idesc = make_idesc(128, NFRAMES, 0, 0);
  constexpr int kXPanel = NFRAMES * 128;
  // (weights panel, frames panel): hi_a P1a, hi_b P1b, hi_a P2a, hi_b P2b, lo_a P1a, lo_b P1b
  constexpr int wp[6] = {0, 1, 0, 1, 2, 3};
  constexpr int xp[6] = {0, 1, 2, 3, 0, 1};
  uint32_t acc = 0;
#pragma unroll
  for (int q = 0; q < 6; q++) {
    if (q >= nq) break;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      uint64_t wd = desc_add(w_desc0, wp[q] * kPanelBytes + kk * 32);
      uint64_t xd = desc_add(x_desc0, xp[q] * kXPanel + kk * 32);
      if (w_is_a)
        umma_ss(d_tmem, wd, xd, idesc, acc);
      else
        umma_ss(d_tmem, xd, wd, idesc, acc);
      acc = 1;
    }
  }
}

// The same contraction with the weights as the A operand read from TMEM (tcgen05.mma "TS"):
// only the frame half panels (B) are fetched from shared memory, 64 B/clk instead of the
// 192 B/clk an SS issue of M=128 x N=64 needs (the SM delivers 128 B/clk).
// w_tmem: 128 columns, panel p (hi a, hi b, lo a, lo b) at column 32 p, fp16 pairs per column.
// w_tmem: 64 columns holding the HI weights (panel a at +0, panel b at +32, fp16 pairs per
// column); the LO weights stay in shared memory (w_smem: lo a at +32 KB, lo b at +48 KB) and
// their product is issued SS -- 16 of the 24 UMMAs read only the frame operand from smem.
__device__ __forceinline__ void issue_g1_ts(uint32_t d_tmem, uint32_t w_tmem, uint64_t wlo_desc0,
                                            uint64_t x_desc0, int nq) {
  // wlo_desc0: descriptor of the LO a weights panel (LO b follows at +16 KB)
  constexpr uint32_t idesc = make_idesc(128, 64, 0, 0);
  constexpr int kXPanel = 64 * 128;
  constexpr int wp[6] = {0, 1, 0, 1, 0, 1};
  constexpr int xp[6] = {0, 1, 2, 3, 0, 1};
  uint32_t acc = 0;
#pragma unroll
  for (int q = 0; q < 6; q++) {
    if (q >= nq) break;
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
      uint64_t xd = desc_add(x_desc0, xp[q] * kXPanel + kk * 32);
      if (q < 4) {
        umma_ts(d_tmem, w_tmem + wp[q] * 32 + kk * 8, xd, idesc, acc);
      } else {
        umma_ss(d_tmem, desc_add(wlo_desc0, wp[q] * kPanelBytes + kk * 32), xd, idesc, acc);
      }
      acc = 1;
    }
  }
}

// ------------------------------------------------------------------ pass 1
// grid = slices * groups.  Partial (max, sum) of 2^S over the slice's components per frame.
__global__ void __launch_bounds__(kTcThreads, 1)
k_tc_lse(int n_slices, int csize, const unsigned char *__restrict__ Wp,
         const unsigned char *__restrict__ Xh, const int *__restrict__ group_tiles /*[groups + 1]*/,
         long P_pad, float2 *__restrict__ part /*[slices][P_pad]*/, int dbg,
         float *__restrict__ S_out /*[P][Cp] log2 joint likelihoods, or null*/, long P, int Cp) {
  extern __shared__ unsigned char smem_raw[];
  const Smem sm = carve(smem_raw);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % n_slices, group = blockIdx.x / n_slices;
  const int t_begin = group_tiles[group], t_end = group_tiles[group + 1];
  const int n_tiles = t_end - t_begin;

  const uint32_t crank = csize > 1 ? cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << csize) - 1);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kStages; i++) {
      mbar_init(sm.full[i], 1);
      mbar_init(sm.empty[i], csize);  // one commit from every CTA sharing the multicast tile
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(sm.s_full[i], 1);
      mbar_init(sm.s_empty[i], 4);  // the four warps of the team that owns the buffer
    }
    mbar_init(sm.w_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(sm.tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();  // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_slot));

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(sm.w_full, 4 * kPanelBytes);
      for (int p = 0; p < 4; p++)
        bulk_g2s(sm.w + p * kPanelBytes, Wp + (size_t)slice * 4 * kPanelBytes + (size_t)p * kPanelBytes,
                 kPanelBytes, sm.w_full);
    }
    for (int i = 0; i < n_tiles; i++) {
      int st = i % kStages;
      mbar_wait(sm.empty[st], ((i / kStages) & 1) ^ 1);
      if (leader) {
        mbar_expect_tx(sm.full[st], kTileBytes);
        const unsigned char *src = Xh + (size_t)(t_begin + i) * kTileBytes;
        if (csize > 1) {
          // this CTA fetches whole panels (4 / csize of them) for the entire cluster: large
          // copies (>= 8 KB) are what the copy engine needs to run at full rate
          for (int p = crank; p < 4; p += csize)
            bulk_g2s_mc(sm.stage[st] + p * kPanelBytes, src + (size_t)p * kPanelBytes, kPanelBytes,
                        sm.full[st], cmask);
        } else {
          for (int p = 0; p < 4; p++)
            bulk_g2s(sm.stage[st] + p * kPanelBytes, src + (size_t)p * kPanelBytes, kPanelBytes,
                     sm.full[st]);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // single issuer: alternating two issuer warps by tile was measured SLOWER here (1.08 -> 1.34 ms
    // per 1 M frames): this pass is bound by shared-memory operand bandwidth, not by issue latency
    const bool leader = elect_one();
    const uint64_t w_desc0 = make_desc(sm.w, 16, 1024);
    mbar_wait(sm.w_full, 0);
    for (int i = 0; i < n_tiles; i++) {
      const int st = i % kStages, buf = i & 1;
      mbar_wait(sm.full[st], (i / kStages) & 1);
      mbar_wait(sm.s_empty[buf], ((i >> 1) & 1) ^ 1);
      tc_fence_after();
      if (leader) {
        issue_g1<128>(tmem_base + buf * 128, w_desc0, make_desc(sm.stage[st], 16, 1024), false,
                      (dbg & 8) ? 1 : 6);
        if (csize > 1)
          umma_commit_mc(sm.empty[st], cmask);
        else
          umma_commit(sm.empty[st]);
        umma_commit(sm.s_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp >= 4) {
    // two teams of four warps (one per TMEM lane quarter) alternate tiles: team t owns S[t]
    const int q = warp & 3;
    const int team = (warp - 4) >> 2;
    for (int i = team; i < n_tiles; i += 2) {
      const int buf = team;
      mbar_wait(sm.s_full[buf], (i >> 1) & 1);
      tc_fence_after();
      float m = -3.0e38f, s = 0.f;
#pragma unroll 1
      for (int ch = 0; ch < 4; ch++) {
        uint32_t r[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + buf * 128 + ch * 32, r);
        tmem_wait_ld();
        if (S_out) {  // the top-K path nominates its candidates on these scores (gmm_topk.cu)
          const long pf = (long)(t_begin + i) * kTile + q * 32 + lane;
          if (pf < P) {
            uint4 *dst = reinterpret_cast<uint4 *>(S_out + (size_t)pf * Cp + (size_t)slice * kSlice + ch * 32);
#pragma unroll
            for (int e = 0; e < 8; e++) dst[e] = make_uint4(r[4 * e], r[4 * e + 1], r[4 * e + 2], r[4 * e + 3]);
          }
        }
        float c8[8];
#pragma unroll
        for (int e = 0; e < 8; e++)
          c8[e] = fmaxf(fmaxf(__uint_as_float(r[e]), __uint_as_float(r[e + 8])),
                        fmaxf(__uint_as_float(r[e + 16]), __uint_as_float(r[e + 24])));
        float cm = fmaxf(fmaxf(fmaxf(c8[0], c8[1]), fmaxf(c8[2], c8[3])),
                         fmaxf(fmaxf(c8[4], c8[5]), fmaxf(c8[6], c8[7])));
        float mn = fmaxf(m, cm);
        float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int e = 0; e < 32; e++) a4[e & 3] += ex2f(__uint_as_float(r[e]) - mn);
        s = s * ex2f(m - mn) + ((a4[0] + a4[1]) + (a4[2] + a4[3]));
        m = mn;
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.s_empty[buf]);
      long p = (long)(t_begin + i) * kTile + q * 32 + lane;
      part[(size_t)slice * P_pad + p] = make_float2(m, s);
    }
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();  // no CTA leaves while a peer may still signal / write to it
  if (warp == 2) tmem_dealloc(tmem_base, 256);
}

// combine the per-slice partials: lse2[p] = log2 sum_c 2^S ; llk_sum += ln2 * sum over valid p
__global__ void k_tc_combine(int n_slices, long P, long P_pad, const unsigned *__restrict__ index,
                             const float2 *__restrict__ part, float *__restrict__ lse2,
                             double *__restrict__ llk_sum) {
  long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  double mine = 0.0;
  if (p < P_pad) {
    float m = -3.0e38f;
    for (int sl = 0; sl < n_slices; sl++) m = fmaxf(m, part[(size_t)sl * P_pad + p].x);
    float s = 0.f;
    for (int sl = 0; sl < n_slices; sl++) {
      float2 v = part[(size_t)sl * P_pad + p];
      s += v.y * ex2f(v.x - m);
    }
    float l = m + log2f(s);
    lse2[p] = l;
    bool valid = p < P && (!index || index[p] != kPadIndex);
    if (valid) mine = (double)l;
  }
  if (llk_sum) {
    __shared__ double red[8];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) mine += __shfl_xor_sync(0xffffffffu, mine, o);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += red[w];
      if (t != 0.0) atomicAdd(llk_sum, t * 0.69314718055994530942);
    }
  }
}

// ------------------------------------------------------------------ pass 2
// Pipeline unit = half tile (64 frames, 4 x 8 KB half panels) through a 5-stage ring, so the
// bulk loads run up to four half tiles ahead of the tensor pipe.
// TMEM columns: three S buffers of 64 columns at 0, 64, 128 (the fp16 posteriors overwrite the
// first 16 columns of each 32-column half in place), the HI weights at 192..255 and the
// statistics accumulator at 256..511.
template <bool EM>
__global__ void __launch_bounds__(kTcThreads, 1)
k_tc_acc(int C, int D, int n_slices, int csize, const unsigned char *__restrict__ Wp,
         const unsigned char *__restrict__ Xh, const int *__restrict__ group_tiles,
         const TileInfo *__restrict__ tinfo, const float *__restrict__ lse2,
         const double *__restrict__ g, const double *__restrict__ s, double fw,
         double *__restrict__ out_N, double *__restrict__ out_F, double *__restrict__ out_S2,
         int dbg) {
  constexpr int N2 = EM ? 256 : 128;
  // B of the statistics GEMM: MN-major, 64-wide chunks = half panels
  constexpr uint32_t kLbo2 = EM ? (uint32_t)kHalfPanel : 2u * kHalfPanel;
  constexpr uint32_t idesc2 = make_idesc(128, N2, 0, 1);
  constexpr int kHF = 64;    // frames per half tile
  constexpr int kWCol = 192;  // TMEM columns [192, 256): HI weights of the slice (A operand of G1)
  constexpr int kNS = 3;      // S buffers: the likelihood GEMM runs two half tiles ahead
  extern __shared__ unsigned char smem_raw[];
  const Smem sm = carve(smem_raw);
  __shared__ float nl_s[4][kHF];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % n_slices, group = blockIdx.x / n_slices;
  const int t_begin = group_tiles[group], t_end = group_tiles[group + 1];
  const int n_half = 2 * (t_end - t_begin);

  const uint32_t crank = csize > 1 ? cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << csize) - 1);
  if (threadIdx.x == 0) {
    for (int i = 0; i < kHStages; i++) {
      mbar_init(sm.full[i], 1);
      mbar_init(sm.empty[i], csize);  // one commit from every CTA sharing the multicast tile
    }
    for (int i = 0; i < 3; i++) {
      mbar_init(sm.s_full[i], 1);
      mbar_init(sm.p_full[i], 4);  // the four warps of one epilogue team
      mbar_init(sm.s_free[i], 1);  // statistics GEMM done with the S / posterior buffer
    }
    mbar_init(sm.f_full, 1);
    mbar_init(sm.f_empty, kEpiWarps);
    mbar_init(sm.w_full, 1);
    mbar_init(sm.s_empty[0], kEpiWarps);  // "weights are in TMEM" (pass 2 has no s_empty use)
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(sm.tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();  // peers' barriers are initialised before anyone signals them
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_slot));
  const uint32_t tmem_f = tmem_base + 256;

  if (warp == 0) {
    const bool leader = elect_one();
    if (leader) {
      mbar_expect_tx(sm.w_full, 4 * kPanelBytes);
      for (int p = 0; p < 4; p++)
        bulk_g2s(sm.w + p * kPanelBytes, Wp + (size_t)slice * 4 * kPanelBytes + (size_t)p * kPanelBytes,
                 kPanelBytes, sm.w_full);
    }
    for (int h = 0; h < n_half; h++) {
      const int st = h % kHStages;
      mbar_wait(sm.empty[st], ((h / kHStages) & 1) ^ 1);
      if (leader && (dbg & 128) && h >= kHStages) {
        mbar_arrive(sm.full[st]);
      } else if (leader) {
        mbar_expect_tx(sm.full[st], kHalfBytes);
        const unsigned char *src =
            Xh + (size_t)(t_begin + (h >> 1)) * kTileBytes + (size_t)(h & 1) * kHalfPanel;
        if (csize > 1) {
          for (int p = crank; p < 4; p += csize)
            bulk_g2s_mc(sm.hstage[st] + p * kHalfPanel, src + (size_t)p * kPanelBytes, kHalfPanel,
                        sm.full[st], cmask);
        } else {
          for (int p = 0; p < 4; p++)
            bulk_g2s(sm.hstage[st] + p * kHalfPanel, src + (size_t)p * kPanelBytes, kHalfPanel,
                     sm.full[st]);
        }
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    // Likelihood-GEMM issuer.  Two issuer warps (this one and warp 3) keep the tensor pipe fed:
    // the tcgen05 queue is shallow, so one warp's barrier round trips would drain it
    // (measured: UMMA time was purely additive to the ~830 clk/half-tile wait latency).
    const bool leader = elect_one();
    const int nq = (dbg & 8) ? 1 : 6;
    const uint64_t wlo_desc0 = make_desc(sm.w + 2 * kPanelBytes, 16, 1024);
    const uint64_t x_desc0 = make_desc(sm.hstage[0], 16, 1024);  // K-major view of the frames
    mbar_wait(sm.w_full, 0);      // LO weights landed in shared memory
    mbar_wait(sm.s_empty[0], 0);  // epilogue warps copied the HI weights into TMEM
    for (int h = 0; h < n_half; h++) {
      const int st = h % kHStages, buf = h % kNS;
      mbar_wait(sm.full[st], (h / kHStages) & 1);
      if (h >= kNS) mbar_wait(sm.s_free[buf], (h / kNS - 1) & 1);  // G2(h - 3) consumed the buffer
      tc_fence_after();
      if (leader) {
        issue_g1_ts(tmem_base + buf * kHF, tmem_base + kWCol, wlo_desc0,
                    desc_add(x_desc0, st * kHalfBytes), nq);
        umma_commit(sm.s_full[buf]);
      }
      __syncwarp();
    }
  } else if (warp == 3) {
    // Statistics-GEMM issuer: F[c, d] (+)= P[c, t] A[t, d], 4 K-steps of 16 frames per half
    // tile; P is the fp16 posterior block the epilogue wrote over the S columns.
    const bool leader = elect_one();
    const int n_k2 = (dbg & 4) ? 1 : kHF / 16;
    const uint64_t b_desc0 = make_desc(sm.hstage[0], kLbo2, 1024);  // MN-major view of the frames
    int n_flush = 0;  // flushes requested so far
    TileInfo ti = n_half > 0 ? tinfo[t_begin] : TileInfo{0, 0};
    TileInfo ti_next = ti;
    for (int h = 0; h < n_half; h++) {
      const int st = h % kHStages, buf = h % kNS;
      if (!(h & 1)) {  // tile info of the NEXT tile is fetched a whole tile ahead of its use
        ti = ti_next;
        if (h + 2 < n_half) ti_next = tinfo[t_begin + (h >> 1) + 1];
      }
      const bool first = (ti.flags & 1) && !(h & 1), last = (ti.flags & 2) && (h & 1);
      mbar_wait(sm.full[st], (h / kHStages) & 1);  // (already complete: warp 1 consumed it first)
      mbar_wait(sm.p_full[buf], (h / kNS) & 1);
      if (first && n_flush > 0) mbar_wait(sm.f_empty, (n_flush - 1) & 1);
      tc_fence_after();
      if (leader) {
        uint32_t acc = first ? 0u : 1u;
        const uint64_t bd0 = desc_add(b_desc0, st * kHalfBytes);
#pragma unroll
        for (int kk = 0; kk < kHF / 16; kk++) {
          if (kk < n_k2) {
            umma_ts(tmem_f, tmem_base + buf * kHF + (kk >> 1) * 32 + (kk & 1) * 8,
                    desc_add(bd0, kk * 2048), idesc2, acc);
            acc = 1;
          }
        }
        if (csize > 1)
          umma_commit_mc(sm.empty[st], cmask);
        else
          umma_commit(sm.empty[st]);
        umma_commit(sm.s_free[buf]);
        if (last) umma_commit(sm.f_full);
      }
      if (last) n_flush++;
      __syncwarp();
    }
  } else if (warp >= 4) {
    // 8 epilogue warps = two teams of four (one warp per TMEM lane quarter q).  Team t converts
    // the half tiles h = t, t+2, ... (all 64 frame columns), so two half tiles are in flight;
    // in the flush, warp (q, t) owns the statistics columns 32t.. of components 32q..
    const int q = warp & 3;
    const int ch = (warp - 4) >> 2;  // team
    const int et = threadIdx.x - 128 - ch * 128;  // 0..127 inside the team
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    // flush staging: the HI weights' 32 KB of shared memory (free once they are in TMEM)
    float *stg = reinterpret_cast<float *>(smem_raw + (carve_base(smem_raw) - smem_u32(smem_raw))) +
                 (warp - 4) * kStageFloats;
    int n_flush = 0;
    {
      // HI weights: shared memory (swizzled panels) -> TMEM rows; warp (q, ch) copies the
      // rows 32q.. of panel ch (hi a / hi b)
      mbar_wait(sm.w_full, 0);
      const int r = q * 32 + lane;
      for (int p = ch; p < ch + 1; p++) {
#pragma unroll
        for (int half16 = 0; half16 < 2; half16++) {
          uint32_t v[16];
#pragma unroll
          for (int j = 0; j < 4; j++) {
            const int chunk = half16 * 4 + j;
            const uint32_t a = sm.w + p * kPanelBytes + r * 128 + ((chunk ^ (r & 7)) << 4);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                         : "=r"(v[4 * j]), "=r"(v[4 * j + 1]), "=r"(v[4 * j + 2]), "=r"(v[4 * j + 3])
                         : "r"(a));
          }
          tmem_st16(tmem_base + lane_addr + kWCol + p * 32 + half16 * 16, v);
        }
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.s_empty[0]);
    }
    // 14 - lse2 of the half tile's frames (the TMEM columns), prefetched one team-step ahead
    float nl_next = 0.f;
    if (et < kHF && ch < n_half) nl_next = kGammaShift - lse2[(size_t)t_begin * kTile + (size_t)ch * kHF + et];
    TileInfo eti_next = n_half > 0 ? tinfo[t_begin] : TileInfo{0, 0};
    for (int h = ch; h < n_half; h += 2) {
      const int buf = h % kNS;
      const TileInfo eti = eti_next;
      if (h + 2 < n_half) eti_next = tinfo[t_begin + ((h + 2) >> 1)];
      if (et < kHF) nl_s[h & 3][et] = nl_next;
      named_bar_sync(1 + ch, 128);
      if (et < kHF && h + 2 < n_half)
        nl_next = kGammaShift - lse2[(size_t)t_begin * kTile + (size_t)(h + 2) * kHF + et];
      mbar_wait(sm.s_full[buf], (h / kNS) & 1);
      tc_fence_after();
      if (!(dbg & 16)) {
        uint32_t r0[32], r1[32];
        tmem_ld32(tmem_base + lane_addr + buf * kHF, r0);
        tmem_ld32(tmem_base + lane_addr + buf * kHF + 32, r1);
        tmem_wait_ld();
        const float *nl = nl_s[h & 3];
        uint32_t pk[16];
#pragma unroll
        for (int e = 0; e < 16; e++) {
          float a = __uint_as_float(r0[2 * e]) + nl[2 * e];
          float b = __uint_as_float(r0[2 * e + 1]) + nl[2 * e + 1];
          if (!(dbg & 1)) {
            a = ex2f(a);
            b = ex2f(b);
          }
          __half2 hh = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));
          pk[e] = *reinterpret_cast<uint32_t *>(&hh);
        }
        tmem_st16(tmem_base + lane_addr + buf * kHF, pk);
#pragma unroll
        for (int e = 0; e < 16; e++) {
          float a = __uint_as_float(r1[2 * e]) + nl[32 + 2 * e];
          float b = __uint_as_float(r1[2 * e + 1]) + nl[32 + 2 * e + 1];
          if (!(dbg & 1)) {
            a = ex2f(a);
            b = ex2f(b);
          }
          __half2 hh = __floats2half2_rn(fminf(a, 65504.f), fminf(b, 65504.f));
          pk[e] = *reinterpret_cast<uint32_t *>(&hh);
        }
        tmem_st16(tmem_base + lane_addr + buf * kHF + 32, pk);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.p_full[buf]);

      const TileInfo ti = eti;
      if (ti.flags & 2) {
        // Flush the accumulator of this run.  TMEM: lane = component, columns = statistics,
        //   EM : [xh hi,1 | xh^2 hi | xh lo | xh^2 lo] (4 x 64);  BW: [xh hi,1 | xh lo] (2 x 64).
        // The increments are staged through shared memory (fp32) so that the fp64
        // read-modify-write of the row's [128 comps x D] block is coalesced (lane = dimension).
        mbar_wait(sm.f_full, n_flush & 1);
        n_flush++;
        tc_fence_after();
        const int comp0 = slice * kSlice + q * 32;
        double *oN = out_N, *oF = out_F, *oS = out_S2;
        const double sc = 1.0 / 16384.0;  // undo the 2^14 posterior scale
        constexpr int kLoCol = EM ? 128 : 64;
        const double n = (double)__uint_as_float(tmem_ld1(tmem_f + lane_addr + kOneCol)) * sc;
        const size_t rc0 = (size_t)ti.row * C + comp0;
        const int k0 = ch * 32;
        uint32_t a_hi[32], a_lo[32];
        tmem_ld32(tmem_f + lane_addr + k0, a_hi);
        tmem_ld32(tmem_f + lane_addr + kLoCol + k0, a_lo);
        tmem_wait_ld();
        float f[32];
#pragma unroll
        for (int e = 0; e < 32; e++)
          f[e] = (__uint_as_float(a_hi[e]) + __uint_as_float(a_lo[e])) * (float)sc;
        if (EM) {
          tmem_ld32(tmem_f + lane_addr + 64 + k0, a_hi);
          tmem_ld32(tmem_f + lane_addr + 192 + k0, a_lo);
          tmem_wait_ld();
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(sm.f_empty);  // accumulator columns are free again
        if (!(dbg & 2)) {
          if (ch == 1 && oN && comp0 + lane < C) atomicAdd(&oN[rc0 + lane], fw * n);
          // first moments: F += fw (s_k f + g_k n)
          __syncwarp();
#pragma unroll
          for (int e = 0; e < 32; e++) {
            const int k = k0 + e;
            float inc = 0.f;
            if (k < D) inc = (float)(fw * (s[k] * (double)f[e] + g[k] * n));
            stg[lane * 32 + (e ^ lane)] = inc;
          }
          __syncwarp();
          if (oF && k0 + lane < D) {
#pragma unroll 8
            for (int c = 0; c < 32; c++) {
              if (comp0 + c < C)
                atomicAdd(&oF[(rc0 + c) * D + k0 + lane], (double)stg[c * 32 + (lane ^ c)]);
            }
          }
          if (EM && oS) {
            __syncwarp();
#pragma unroll
            for (int e = 0; e < 32; e++) {
              const int k = k0 + e;
              float inc = 0.f;
              if (k < D) {
                double q2 = ((double)__uint_as_float(a_hi[e]) + (double)__uint_as_float(a_lo[e])) * sc;
                double sk = s[k], gk = g[k];
                inc = (float)(fw * (sk * sk * q2 + 2.0 * sk * gk * (double)f[e] + gk * gk * n));
              }
              stg[lane * 32 + (e ^ lane)] = inc;
            }
            __syncwarp();
            if (k0 + lane < D) {
#pragma unroll 4
              for (int c = 0; c < 32; c++) {
                if (comp0 + c < C)
                  atomicAdd(&oS[(rc0 + c) * D + k0 + lane], (double)stg[c * 32 + (lane ^ c)]);
              }
            }
          }
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (csize > 1) cluster_sync_all();  // no CTA leaves while a peer may still signal / write to it
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}


// ------------------------------------------------------------------ one-pass kernel
// Likelihood GEMM, per-frame log-sum-exp and statistics GEMM in ONE sweep over the frames: the
// per-frame normaliser is not precomputed by a first pass, it is exchanged between the slice-CTAs
// of a frame group while the posteriors wait in TMEM.
//
// Per CTA (slice of 128 components, one frame group) and half tile h (64 frames):
//   G1(h) : S[c, t] = W A^T, ALL weights (hi and lo) resident in TMEM -> 24 TS UMMAs that read only
//           the frame half panels from shared memory                                     (warp 1)
//   E1(h) : 16 warps = 2 parities x 2 column halves x 4 TMEM lane quarters; the warps of parity
//           h & 1 take half tile h, 32 of its 64 frame columns each.  S is read with the 16x256b TMEM
//           shape (a thread holds 4 component rows x 8 frame columns), so the per-frame max / sum over
//           a warp's 32 components is 3 in-thread steps + a 3-step shuffle butterfly.  Posteriors are
//           formed RELATIVE TO THE WARP'S MAXIMUM, fp16(2^14 2^(S - max_warp)), and parked in a TMEM
//           slot (holding them in registers over the exchange made the hot loop spill); the warp's
//           (max, sum) go to a shared-memory ring.
//   X(h)  : (warp 2) combines the four lane quarters' (max, sum) per frame and posts the slice's pair
//           to the group's exchange ring in global memory as ONE 64-bit word whose top bit is the
//           ring-lap tag (data = flag: no fence, no counter)
//   L(h)  : four warps (parity x column half), lane = frame: fetch the n_slices words of the frame with
//           weak no-allocate loads (served by L2), repeated until every lap tag matches, combine -> lse
//   E1'(h): the same E1 warps, one iteration (two half tiles) later: read the parked posteriors back,
//           rescale them by 2^(max_warp - lse) (fp16 mantissa x exact power of two), store them again
//   G2(h) : F[c, :] += P[c, t] A[t, :] as TS UMMAs, the hi and the lo frame panels accumulated into
//           the SAME columns (EM: [xh,1 | xh^2] = 128 columns, BW: [xh,1] = 64)           (warp 3)
// TMEM: accumulator [0,128) | weights hi a, hi b, lo a, lo b [128,256) | S 2 x 64 [256,384) |
//       posterior slots 4 x 32 [384,512).
// Shared memory: six 32 KB half-tile stages (a stage lives from its bulk copy until G2 has read it,
// about five half-tile periods); the weights pass through the last two stages on their way to TMEM.
// No block-level barrier in the main loop: every hand-over is an mbarrier.  All CTAs of a group
// advance in lock step (each needs every slice's partials), so the grid must be co-resident:
// launched cooperatively with grid <= SM count.
constexpr int kXRing = 16;    // exchange ring slots per group (see tc_run_stats)
constexpr int kOStages = 6;   // half-tile stages
constexpr int kNS = 2;        // S buffers
constexpr int kNP = 4;        // posterior slots in TMEM
constexpr int kNR = 4;        // depth of the (max, sum) and lse rings in shared memory
constexpr int kOneThreads = 768;
constexpr int kOColAcc = 0, kOColW = 128, kOColS = 256, kOColP = 384;
constexpr int kOStgOff = kOStages * kHalfBytes;        // flush staging: 16 warps x 512 B (32 x 4 floats)
constexpr int kORingOff = kOStgOff + 16 * 512;         // lane-quarter (max, sum) ring: 2 x [kNR][4][64] floats
constexpr int kOLseOff = kORingOff + 2 * kNR * 4 * 64 * 4;  // lse ring [kNR][64]
constexpr int kOBarOff = kOLseOff + kNR * 64 * 4;
constexpr size_t kOneSmem = 1024 + kOBarOff + 384;
static_assert(kOneSmem <= 227 * 1024, "one-pass kernel: shared memory budget");

// mbarrier addresses are computed (base + offset + 8 i), not kept in arrays: dynamically indexed
// arrays of a struct end up in local memory
struct SmemOne {
  uint32_t base, stage0, bar;
  __device__ __forceinline__ uint32_t full(int i) const { return bar + 8 * i; }
  __device__ __forceinline__ uint32_t empty(int i) const { return bar + 48 + 8 * i; }
  __device__ __forceinline__ uint32_t s_full(int i) const { return bar + 96 + 8 * i; }
  __device__ __forceinline__ uint32_t s_free(int i) const { return bar + 120 + 8 * i; }
  __device__ __forceinline__ uint32_t ms_written(int i) const { return bar + 144 + 8 * i; }
  __device__ __forceinline__ uint32_t lse_ready(int i) const { return bar + 176 + 8 * i; }
  __device__ __forceinline__ uint32_t p_ready(int i) const { return bar + 208 + 8 * i; }
  __device__ __forceinline__ uint32_t p_free(int i) const { return bar + 240 + 8 * i; }
  __device__ __forceinline__ uint32_t f_full() const { return bar + 272; }
  __device__ __forceinline__ uint32_t f_empty() const { return bar + 280; }
  __device__ __forceinline__ uint32_t w_full() const { return bar + 288; }
  __device__ __forceinline__ uint32_t w_tmem() const { return bar + 296; }
  __device__ __forceinline__ uint32_t tmem_slot() const { return bar + 304; }
};
static_assert(kNS <= 3 && kNR == 4 && kNP == 4 && kOStages == 6, "barrier block layout");

__device__ __forceinline__ SmemOne carve_one(unsigned char *raw) {
  SmemOne s;
  s.base = carve_base(raw);
  s.stage0 = s.base;
  s.bar = s.base + kOBarOff;
  return s;
}

// 16 lanes x 256 bits, N repeats along the columns: thread t of the warp receives the lanes
// (t / 4) and (t / 4 + 8) of the 16-lane group addressed, columns 8 j + 2 (t % 4) + e, in register
// 4 j + 2 rs + e  (j repeat, rs = 0/1 lane select, e = 0/1).
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, uint32_t (&r)[16]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr)
      : "memory");
}
// 16 lanes x 128 bits, N repeats: lanes (t / 4), (t / 4 + 8), column 4 j + (t % 4) in register 2 j + rs
__device__ __forceinline__ void tmem_ld_16x128b_x4(uint32_t taddr, uint32_t (&r)[8]) {
  asm volatile("tcgen05.ld.sync.aligned.16x128b.x4.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr)
               : "memory");
}
__device__ __forceinline__ void tmem_st_16x128b_x4(uint32_t taddr, const uint32_t (&r)[8]) {
  asm volatile(
      "tcgen05.st.sync.aligned.16x128b.x4.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
      : "memory");
}
__device__ __forceinline__ void st_relaxed_v2u64(unsigned long long *p, unsigned long long a,
                                                 unsigned long long b) {
  asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};" ::"l"(p), "l"(a), "l"(b) : "memory");
}
// weak load that does not allocate in L1: the exchange ring is touched by nothing else on the SM, so
// every such load is served by L2 (the coherence point) and, unlike ld.relaxed.gpu / .cg (both
// LDG.STRONG.GPU, ~50 issue clocks each, measured), a batch of them pipelines
__device__ __forceinline__ unsigned long long ld_na_u64(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.global.L1::no_allocate.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_relaxed_u64(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ float lg2f(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
template <int N>
__device__ __forceinline__ void reg_alloc() {
  asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N));
}
template <int N>
__device__ __forceinline__ void reg_dealloc() {
  asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N));
}

// One step of the lane reduce-scatter: the lanes whose bit `xr` is set keep the upper half of
// in[0 .. 2n), the others the lower half; the halves not kept are combined into the partner.
template <int N, bool MAX>
__device__ __forceinline__ void lane_halve(const float (&in)[2 * N], float (&out)[N], bool up,
                                           int xr) {
#pragma unroll
  for (int i = 0; i < N; i++) {
    const float mine = up ? in[N + i] : in[i];
    const float send = up ? in[i] : in[N + i];
    const float got = __shfl_xor_sync(0xFFFFFFFFu, send, xr);
    out[i] = MAX ? fmaxf(mine, got) : mine + got;
  }
}

// PROF: clock64() phase accounting of one warp per role, summed over the grid into
// prof[role * 16 + phase] (LR_TC_PROF=1; never the product path).
#define LR_PT(i)                        \
  do {                                  \
    if (PROF) {                         \
      const long long now_ = clock64(); \
      pacc[i] += now_ - tlast;          \
      tlast = now_;                     \
    }                                   \
  } while (0)
#define LR_PDUMP(role)                                                                          \
  do {                                                                                          \
    if (PROF && lane == 0)                                                                      \
      for (int i_ = 0; i_ < 8; i_++)                                                            \
        atomicAdd((unsigned long long *)prof + (role) * 16 + i_, (unsigned long long)pacc[i_]); \
  } while (0)

template <bool EM, bool PROF>
__global__ void __launch_bounds__(kOneThreads, 1)
k_tc_one(int C, int D, int n_slices, const unsigned char *__restrict__ Wp,
         const unsigned char *__restrict__ Xh, const int *__restrict__ group_tiles,
         const TileInfo *__restrict__ tinfo, unsigned long long *xch,
         float *__restrict__ lse_out, const unsigned *__restrict__ index, long P,
         double *llk_sum, const double *__restrict__ g, const double *__restrict__ s, double fw,
         double *__restrict__ out_N, double *__restrict__ out_F, double *__restrict__ out_S2,
         int dbg, long long *prof) {
  constexpr int N2 = EM ? 128 : 64;  // statistics columns: [xh, 1 | xh^2] or [xh, 1]
  constexpr uint32_t idesc1 = make_idesc(128, 64, 0, 0);
  constexpr uint32_t idesc2 = make_idesc(128, N2, 0, 1);
  constexpr int kHF = 64;
  long long pacc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  long long tlast = PROF ? clock64() : 0;
  extern __shared__ unsigned char smem_raw[];
  const SmemOne sm = carve_one(smem_raw);
  unsigned char *base_ptr = smem_raw + (sm.base - smem_u32(smem_raw));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int slice = blockIdx.x % n_slices, group = blockIdx.x / n_slices;
  const int t_begin = group_tiles[group], t_end = group_tiles[group + 1];
  const int n_half = 2 * (t_end - t_begin);
  // the lane quarters' (max, sum) per frame: [h % kNR][4 quarters][64 frames].  Depth 4: the slot of
  // half tile h is rewritten for h + 4, i.e. after this CTA's E1 warps finished h + 2 > h, which
  // needed every peer's (hence also this CTA's) publication of h -- so it has been read.
  float *wmax = reinterpret_cast<float *>(base_ptr + kORingOff);
  float *wsum = wmax + kNR * 4 * kHF;
  float *lse_s = reinterpret_cast<float *>(base_ptr + kOLseOff);  // [kNR][64]
  const uint32_t w_smem = sm.stage0 + 4 * kHalfBytes;             // weights staging = stages 4, 5

  if (threadIdx.x == 0) {
    for (int i = 0; i < kOStages; i++) {
      mbar_init(sm.full(i), 1);
      mbar_init(sm.empty(i), 1);
    }
    for (int i = 0; i < kNS; i++) {
      mbar_init(sm.s_full(i), 1);
      mbar_init(sm.s_free(i), 8);  // the eight E1 warps of the half tile's parity have S in registers
    }
    for (int i = 0; i < kNR; i++) {
      mbar_init(sm.ms_written(i), 8);  // E1 warps: (max, sum) of the half tile are in the ring
      mbar_init(sm.lse_ready(i), 2);   // L warps: lse of the half tile is in shared memory
    }
    for (int i = 0; i < kNP; i++) {
      mbar_init(sm.p_ready(i), 8);  // E1 warps: rescaled posteriors are in the TMEM slot
      mbar_init(sm.p_free(i), 1);   // G2 done with the slot
    }
    mbar_init(sm.f_full(), 1);
    mbar_init(sm.f_empty(), 16);
    mbar_init(sm.w_full(), 1);
    mbar_init(sm.w_tmem(), 16);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(sm.tmem_slot(), 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_slot()));
  const uint32_t tmem_f = tmem_base + kOColAcc;
  const long frame0 = (long)t_begin * kTile;
  // exchange ring of the group: [kXRing][2 column halves][n_slices][32 frames] words
  unsigned long long *ring = xch + (size_t)group * kXRing * n_slices * kHF;

  if (warp < 4) {
    reg_dealloc<40>();
    if (warp == 0) {
      // ---- bulk-copy producer
      const bool leader = elect_one();
      if (leader) {
        mbar_expect_tx(sm.w_full(), 4 * kPanelBytes);
        for (int p = 0; p < 4; p++)
          bulk_g2s(w_smem + p * kPanelBytes, Wp + (size_t)slice * 4 * kPanelBytes + (size_t)p * kPanelBytes,
                   kPanelBytes, sm.w_full());
      }
      for (int h = 0; h < n_half; h++) {
        const int st = h % kOStages;
        if (h == 4) mbar_wait(sm.w_tmem(), 0);  // stages 4, 5 held the weights until they were in TMEM
        mbar_wait(sm.empty(st), ((h / kOStages) & 1) ^ 1);
        if (leader) {
          const int np = (dbg & 512) ? 2 : 4;  // product level 2 never reads the lo panels
          mbar_expect_tx(sm.full(st), np * kHalfPanel);
          const unsigned char *src =
              Xh + (size_t)(t_begin + (h >> 1)) * kTileBytes + (size_t)(h & 1) * kHalfPanel;
          for (int p = 0; p < np; p++)
            bulk_g2s(sm.stage0 + st * kHalfBytes + p * kHalfPanel, src + (size_t)p * kPanelBytes,
                     kHalfPanel, sm.full(st));
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      // ---- likelihood-GEMM issuer: hi_a P1a, hi_b P1b, hi_a P2a, hi_b P2b, lo_a P1a, lo_b P1b
      const bool leader = elect_one();
      const uint64_t x_desc0 = make_desc(sm.stage0, 16, 1024);  // K-major view of the frames
      constexpr int wp[6] = {0, 1, 0, 1, 2, 3};
      constexpr int xp[6] = {0, 1, 2, 3, 0, 1};
      mbar_wait(sm.w_tmem(), 0);  // the weights are in TMEM
      LR_PT(0);
      for (int h = 0; h < n_half; h++) {
        const int st = h % kOStages, sb = h % kNS;
        mbar_wait(sm.full(st), (h / kOStages) & 1);
        LR_PT(1);
        if (h >= kNS) mbar_wait(sm.s_free(sb), ((h / kNS) - 1) & 1);
        LR_PT(2);
        tc_fence_after();
        if (leader) {
          const uint32_t d_tmem = tmem_base + kOColS + sb * kHF;
          const uint64_t xd0 = desc_add(x_desc0, st * kHalfBytes);
          uint32_t acc = 0;
#pragma unroll
          for (int q = 0; q < 6; q++) {
            if ((q == 2 || q == 3) && (dbg & 512)) continue;  // product level 2: no W_hi X_lo
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              umma_ts(d_tmem, tmem_base + kOColW + wp[q] * 32 + kk * 8,
                      desc_add(xd0, xp[q] * kHalfPanel + kk * 32), idesc1, acc);
              acc = 1;
            }
          }
          umma_commit(sm.s_full(sb));
        }
        __syncwarp();
        LR_PT(3);
      }
      LR_PDUMP(3);
    } else if (warp == 2) {
      // ---- publisher: slice (max, sum) of every frame of the half tile -> exchange ring
      for (int h = 0; h < n_half; h++) {
        const int rs = h % kNR;
        mbar_wait(sm.ms_written(rs), (h / kNR) & 1);
        const float *pm = wmax + rs * 4 * kHF + 2 * lane, *pz = wsum + rs * 4 * kHF + 2 * lane;
        const float2 m0 = *reinterpret_cast<const float2 *>(pm), m1 = *reinterpret_cast<const float2 *>(pm + kHF);
        const float2 m2 = *reinterpret_cast<const float2 *>(pm + 2 * kHF), m3 = *reinterpret_cast<const float2 *>(pm + 3 * kHF);
        const float2 z0 = *reinterpret_cast<const float2 *>(pz), z1 = *reinterpret_cast<const float2 *>(pz + kHF);
        const float2 z2 = *reinterpret_cast<const float2 *>(pz + 2 * kHF), z3 = *reinterpret_cast<const float2 *>(pz + 3 * kHF);
        const float ma = fmaxf(fmaxf(m0.x, m1.x), fmaxf(m2.x, m3.x));
        const float mb = fmaxf(fmaxf(m0.y, m1.y), fmaxf(m2.y, m3.y));
        const float za = ((z0.x * ex2f(m0.x - ma) + z1.x * ex2f(m1.x - ma)) +
                          (z2.x * ex2f(m2.x - ma) + z3.x * ex2f(m3.x - ma))) * (1.f / 16384.f);
        const float zb = ((z0.y * ex2f(m0.y - mb) + z1.y * ex2f(m1.y - mb)) +
                          (z2.y * ex2f(m2.y - mb) + z3.y * ex2f(m3.y - mb))) * (1.f / 16384.f);
        const unsigned tag = (((unsigned)h / kXRing) & 1u) ^ 1u;
        const unsigned long long wa =
            ((unsigned long long)(__float_as_uint(za) | (tag << 31)) << 32) | __float_as_uint(ma);
        const unsigned long long wb =
            ((unsigned long long)(__float_as_uint(zb) | (tag << 31)) << 32) | __float_as_uint(mb);
        // frames 2 lane, 2 lane + 1 -> column half lane / 16, position (2 lane) % 32
        st_relaxed_v2u64(ring + (((size_t)(h % kXRing) * 2 + (lane >> 4)) * n_slices + slice) * 32 + (2 * lane & 31),
                         wa, wb);
      }
    } else {
      // ---- statistics-GEMM issuer: F[c, :] (+)= P[c, t] A[t, :]; B = hi panels, then lo panels,
      // read MN-major (64-wide chunks = half panels, kHalfPanel apart)
      const bool leader = elect_one();
      const uint64_t b_desc0 = make_desc(sm.stage0, (uint32_t)kHalfPanel, 1024);
      int n_flush = 0;
      TileInfo ti = n_half > 0 ? tinfo[t_begin] : TileInfo{0, 0};
      TileInfo ti_next = ti;
      for (int h = 0; h < n_half; h++) {
        const int st = h % kOStages, ps = h % kNP;
        if (!(h & 1)) {
          ti = ti_next;
          if (h + 2 < n_half) ti_next = tinfo[t_begin + (h >> 1) + 1];
        }
        const bool first = (ti.flags & 1) && !(h & 1), last = (ti.flags & 2) && (h & 1);
        mbar_wait(sm.full(st), (h / kOStages) & 1);
        LR_PT(1);
        mbar_wait(sm.p_ready(ps), (h / kNP) & 1);
        LR_PT(2);
        if (first && n_flush > 0) mbar_wait(sm.f_empty(), (n_flush - 1) & 1);
        LR_PT(3);
        tc_fence_after();
        if (leader) {
          uint32_t acc = first ? 0u : 1u;
          const uint64_t bd0 = desc_add(b_desc0, st * kHalfBytes);
#pragma unroll
          for (int part = 0; part < 2; part++) {
            if (part == 1 && (dbg & 256)) break;  // product level >= 1: hi frame panels only
#pragma unroll
            for (int kk = 0; kk < kHF / 16; kk++) {
              umma_ts(tmem_f, tmem_base + kOColP + ps * 32 + kk * 8,
                      desc_add(bd0, part * 2 * kHalfPanel + kk * 2048), idesc2, acc);
              acc = 1;
            }
          }
          umma_commit(sm.empty(st));
          umma_commit(sm.p_free(ps));
          if (last) umma_commit(sm.f_full());
        }
        if (last) n_flush++;
        __syncwarp();
        LR_PT(4);
      }
      LR_PDUMP(4);
    }
  } else if (warp < 20) {
    // ---- E1: 16 warps = parity (2) x column half (2) x TMEM lane quarter q (4)
    // 4 x 40 + 16 x 96 + 4 x 56 = 24 x 80: the CTA's register allocation is redistributed, not grown
    reg_alloc<96>();
    const int q = warp & 3, ch = ((warp - 4) >> 2) & 1, par = (warp - 4) >> 3;
    const int t4 = 2 * par + ch;  // 0..3: the statistics columns 32 (t4 & 1).. of F (t4 < 2) or S2 in the flush
    const uint32_t lane_addr = (uint32_t)(q * 32) << 16;
    const uint32_t lane_addr16 = (uint32_t)(q * 32 + 16) << 16;
    float *stg = reinterpret_cast<float *>(base_ptr + kOStgOff) + (warp - 4) * 128;  // 32 x 4 floats
    {
      // weights: shared memory (swizzled panels) -> TMEM rows; warp (t4, q) copies the rows 32 q.. of panel t4
      mbar_wait(sm.w_full(), 0);
      const int r = q * 32 + lane;
#pragma unroll
      for (int half16 = 0; half16 < 2; half16++) {
        uint32_t v[16];
#pragma unroll
        for (int j = 0; j < 4; j++) {
          const int chunk = half16 * 4 + j;
          const uint32_t a = w_smem + t4 * kPanelBytes + r * 128 + ((chunk ^ (r & 7)) << 4);
          asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];"
                       : "=r"(v[4 * j]), "=r"(v[4 * j + 1]), "=r"(v[4 * j + 2]), "=r"(v[4 * j + 3])
                       : "r"(a));
        }
        tmem_st16(tmem_base + lane_addr + kOColW + t4 * 32 + half16 * 16, v);
      }
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.w_tmem());
    }
    const bool b4 = lane & 16, b3 = lane & 8;
    const int jo = (lane >> 3) & 3;  // the column group (of 8) whose warp results this lane ends up holding
    const int mycol = ch * 32 + 8 * jo + 2 * (lane & 3);  // first of the two frame columns this lane owns
    int n_flush = 0;
    float mw0 = 0.f, mw1 = 0.f;  // warp maxima of the frames mycol, mycol + 1 of the parked half tile

    // ---- second half of a half tile's life: rescale the held posteriors, store them, flush at run ends
    auto finish = [&](int h, const TileInfo ti, float mw0, float mw1) {
      const int rs = h % kNR;
      LR_PT(4);
      mbar_wait(sm.lse_ready(rs), (h / kNR) & 1);
      LR_PT(5);
      // rescale factors 2^d, d = max_warp - lse <= 0 (up to rounding), for the frames mycol, mycol + 1,
      // applied as two fp16 factors: a mantissa in (0.5, 1] times 2^ka (ka >= -13: a normal fp16) and
      // the exact power of two 2^(k - ka) >= 2^-24; below 2^-38 nothing of the warp's components
      // survives in fp16
      uint32_t fm2, fp2;
      {
        const float2 l2 = *reinterpret_cast<const float2 *>(lse_s + rs * kHF + mycol);
        const float d0 = mw0 - l2.x, d1 = mw1 - l2.y;
        const float k0 = ceilf(d0), k1 = ceilf(d1);
        const float ka0 = fmaxf(k0, -13.f), ka1 = fmaxf(k1, -13.f);
        const bool dead0 = !(d0 > -38.f), dead1 = !(d1 > -38.f);
        __half2 a = __floats2half2_rn(dead0 ? 0.f : ex2f((d0 - k0) + ka0), dead1 ? 0.f : ex2f((d1 - k1) + ka1));
        __half2 b = __floats2half2_rn(dead0 ? 0.f : ex2f(fmaxf(k0 - ka0, -24.f)),
                                      dead1 ? 0.f : ex2f(fmaxf(k1 - ka1, -24.f)));
        fm2 = *reinterpret_cast<uint32_t *>(&a);
        fp2 = *reinterpret_cast<uint32_t *>(&b);
      }
      const int ps = h % kNP;
      uint32_t pk0[8], pk1[8];
      tmem_ld_16x128b_x4(tmem_base + lane_addr + kOColP + ps * 32 + ch * 16, pk0);
      tmem_ld_16x128b_x4(tmem_base + lane_addr16 + kOColP + ps * 32 + ch * 16, pk1);
      tmem_wait_ld();
#pragma unroll
      for (int j = 0; j < 4; j++) {
        // packed column 4 j + lane % 4 = frames 8 j + 2 (lane % 4), +1: factors held by the lane 8 j + lane % 4
        const uint32_t ua = __shfl_sync(0xFFFFFFFFu, fm2, 8 * j + (lane & 3));
        const uint32_t ub = __shfl_sync(0xFFFFFFFFu, fp2, 8 * j + (lane & 3));
        const __half2 a = *reinterpret_cast<const __half2 *>(&ua), b = *reinterpret_cast<const __half2 *>(&ub);
#pragma unroll
        for (int r2 = 0; r2 < 2; r2++) {
          __half2 x = *reinterpret_cast<__half2 *>(&pk0[2 * j + r2]);
          x = __hmul2(__hmul2(x, a), b);
          pk0[2 * j + r2] = *reinterpret_cast<uint32_t *>(&x);
          __half2 y = *reinterpret_cast<__half2 *>(&pk1[2 * j + r2]);
          y = __hmul2(__hmul2(y, a), b);
          pk1[2 * j + r2] = *reinterpret_cast<uint32_t *>(&y);
        }
      }
      LR_PT(6);
      tmem_st_16x128b_x4(tmem_base + lane_addr + kOColP + ps * 32 + ch * 16, pk0);
      tmem_st_16x128b_x4(tmem_base + lane_addr16 + kOColP + ps * 32 + ch * 16, pk1);
      tmem_wait_st();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.p_ready(ps));
      LR_PT(7);
      if (!(ti.flags & 2)) return;
      // ---- flush of the run that ends with this tile.  TMEM: lane = component, columns
      // [xh, 1 (64) | xh^2 (64)]; warp (t4, q) owns the components 32 q.. and the columns 32 (t4 & 1)..
      // of F (t4 < 2) or of S2 (t4 >= 2).  The increments are staged through shared memory (fp32),
      // 4 columns at a time, so that the fp64 atomics of a row's block are coalesced.
      mbar_wait(sm.f_full(), n_flush & 1);
      n_flush++;
      tc_fence_after();
      const int comp0 = slice * kSlice + q * 32;
      const double sc = 1.0 / 16384.0;  // undo the 2^14 posterior scale
      const double n = (double)__uint_as_float(tmem_ld1(tmem_f + lane_addr + kOneCol)) * sc;
      const size_t rc0 = (size_t)ti.row * C + comp0;
      const int k0 = 32 * (t4 & 1);
      uint32_t a1[32], a2[32];
      tmem_ld32(tmem_f + lane_addr + k0, a1);
      if (EM && t4 >= 2) tmem_ld32(tmem_f + lane_addr + 64 + k0, a2);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.f_empty());  // accumulator columns are free again
      if (dbg & 2) return;
      if (t4 == 1 && out_N && comp0 + lane < C) atomicAdd(&out_N[rc0 + lane], fw * n);
      double *dst = t4 < 2 ? out_F : out_S2;
      if ((!EM && t4 >= 2) || !dst) return;
#pragma unroll
      for (int c4 = 0; c4 < 8; c4++) {  // 4 statistics columns per round
        __syncwarp();
#pragma unroll
        for (int e = 0; e < 4; e++) {
          const int k = k0 + 4 * c4 + e;
          float inc = 0.f;
          if (k < D) {
            const double f1 = (double)__uint_as_float(a1[4 * c4 + e]) * sc;
            if (t4 < 2) {
              inc = (float)(fw * (s[k] * f1 + g[k] * n));
            } else {
              const double q2 = (double)__uint_as_float(a2[4 * c4 + e]) * sc;
              const double sk = s[k], gk = g[k];
              inc = (float)(fw * (sk * sk * q2 + 2.0 * sk * gk * f1 + gk * gk * n));
            }
          }
          stg[lane * 4 + (e ^ (lane & 3))] = inc;
        }
        __syncwarp();
        // lane -> (component lane / 4 + 8 i, column lane % 4): 4 consecutive doubles per component
        const int kk = k0 + 4 * c4 + (lane & 3);
#pragma unroll
        for (int i = 0; i < 4; i++) {
          const int c = (lane >> 2) + 8 * i;
          if (kk < D && comp0 + c < C)
            atomicAdd(&dst[(rc0 + c) * D + kk], (double)stg[c * 4 + ((lane & 3) ^ (c & 3))]);
        }
      }
      __syncwarp();
    };

    int prev = -1;
    TileInfo ti_prev{0, 0};
    for (int h = par; h < n_half; h += 2) {
      const int sb = h % kNS, rs = h % kNR;
      const TileInfo ti_h = tinfo[t_begin + (h >> 1)];  // consumed one iteration later (run ends)
      LR_PT(0);
      mbar_wait(sm.s_full(sb), (h / kNS) & 1);
      LR_PT(1);
      tc_fence_after();
      uint32_t v0[16], v1[16];  // lanes 32q + {t/4, t/4+8} and 32q + 16 + {t/4, t/4+8}; reg 4 j + 2 rs + e
      tmem_ld_16x256b_x4(tmem_base + lane_addr + kOColS + sb * kHF + ch * 32, v0);
      tmem_ld_16x256b_x4(tmem_base + lane_addr16 + kOColS + sb * kHF + ch * 32, v1);
      tmem_wait_ld();
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.s_free(sb));  // S is in registers: a later G1 may overwrite it
      LR_PT(2);
      // ---- per-frame max over the warp's 32 components
      float a8[8], a4[4], a2[2];
#pragma unroll
      for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++)
          a8[2 * j + e] = fmaxf(fmaxf(__uint_as_float(v0[4 * j + e]), __uint_as_float(v0[4 * j + 2 + e])),
                                fmaxf(__uint_as_float(v1[4 * j + e]), __uint_as_float(v1[4 * j + 2 + e])));
      }
      lane_halve<4, true>(a8, a4, b4, 16);
      lane_halve<2, true>(a4, a2, b3, 8);
      a2[0] = fmaxf(a2[0], __shfl_xor_sync(0xFFFFFFFFu, a2[0], 4));
      a2[1] = fmaxf(a2[1], __shfl_xor_sync(0xFFFFFFFFu, a2[1], 4));
      // this lane holds the warp maxima of the columns 8 jo + 2 (lane % 4) + {0, 1}; the columns
      // 8 j + 2 (lane % 4) + e it works on are held by the lane 8 j + lane % 4
      const float nw0 = a2[0], nw1 = a2[1];
      float nm[8];  // 14 - warp maximum of this thread's 8 columns
#pragma unroll
      for (int j = 0; j < 4; j++) {
        nm[2 * j] = kGammaShift - __shfl_sync(0xFFFFFFFFu, nw0, 8 * j + (lane & 3));
        nm[2 * j + 1] = kGammaShift - __shfl_sync(0xFFFFFFFFu, nw1, 8 * j + (lane & 3));
      }
      // ---- posteriors relative to the warp maximum, scaled by 2^14
#pragma unroll
      for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int k = 0; k < 4; k++) {
          v0[4 * j + k] = __float_as_uint(ex2f(__uint_as_float(v0[4 * j + k]) + nm[2 * j + (k & 1)]));
          v1[4 * j + k] = __float_as_uint(ex2f(__uint_as_float(v1[4 * j + k]) + nm[2 * j + (k & 1)]));
        }
      }
#pragma unroll
      for (int j = 0; j < 4; j++) {
#pragma unroll
        for (int e = 0; e < 2; e++)
          a8[2 * j + e] = (__uint_as_float(v0[4 * j + e]) + __uint_as_float(v0[4 * j + 2 + e])) +
                          (__uint_as_float(v1[4 * j + e]) + __uint_as_float(v1[4 * j + 2 + e]));
      }
      lane_halve<4, false>(a8, a4, b4, 16);
      lane_halve<2, false>(a4, a2, b3, 8);
      a2[0] += __shfl_xor_sync(0xFFFFFFFFu, a2[0], 4);
      a2[1] += __shfl_xor_sync(0xFFFFFFFFu, a2[1], 4);
      if (!(lane & 4)) {  // lanes with bit 2 set hold duplicates
        *reinterpret_cast<float2 *>(wmax + (rs * 4 + q) * kHF + mycol) = make_float2(nw0, nw1);
        *reinterpret_cast<float2 *>(wsum + (rs * 4 + q) * kHF + mycol) = make_float2(a2[0], a2[1]);
      }
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.ms_written(rs));
      // ---- pack the frame pairs (2 pc, 2 pc + 1) into the fp16 column pc = 4 j + t % 4 and park them in
      // the TMEM slot of the half tile
      {
        uint32_t n0[8], n1[8];
#pragma unroll
        for (int j = 0; j < 4; j++) {
#pragma unroll
          for (int r2 = 0; r2 < 2; r2++) {
            __half2 h0 = __floats2half2_rn(__uint_as_float(v0[4 * j + 2 * r2]), __uint_as_float(v0[4 * j + 2 * r2 + 1]));
            __half2 h1 = __floats2half2_rn(__uint_as_float(v1[4 * j + 2 * r2]), __uint_as_float(v1[4 * j + 2 * r2 + 1]));
            n0[2 * j + r2] = *reinterpret_cast<uint32_t *>(&h0);
            n1[2 * j + r2] = *reinterpret_cast<uint32_t *>(&h1);
          }
        }
        const int ps = h % kNP;
        if (h >= kNP) {
          mbar_wait(sm.p_free(ps), ((h / kNP) - 1) & 1);  // G2(h - kNP) is done with the slot
          tc_fence_after();
        }
        tmem_st_16x128b_x4(tmem_base + lane_addr + kOColP + ps * 32 + ch * 16, n0);
        tmem_st_16x128b_x4(tmem_base + lane_addr16 + kOColP + ps * 32 + ch * 16, n1);
      }
      LR_PT(3);
      // the previous half tile of this warp: its normaliser has had this whole iteration to arrive
      if (prev >= 0) finish(prev, ti_prev, mw0, mw1);
      tmem_wait_st();  // (the parked posteriors of h are in TMEM before the next iteration may read them back)
      mw0 = nw0;
      mw1 = nw1;
      prev = h;
      ti_prev = ti_h;
    }
    tmem_wait_st();
    if (prev >= 0) finish(prev, ti_prev, mw0, mw1);
    if (q == 0 && ch == 0) LR_PDUMP(par);
  } else {
    // ---- L: four warps (parity x column half), lane = frame: log-sum-exp over all slices
    reg_dealloc<56>();
    const int ch = (warp - 20) & 1, par = (warp - 20) >> 1;
    const int lw = warp - 20;
    double llk_acc = 0.0;
    for (int h = par; h < n_half; h += 2) {
      const int rs = h % kNR;
      LR_PT(0);
      mbar_wait(sm.ms_written(rs), (h / kNR) & 1);  // the E1 warps' (max, sum) are in shared memory
      LR_PT(1);
      const unsigned long long tag = ((((unsigned)h / kXRing) & 1u) ^ 1u);
      const unsigned long long *slot = ring + ((size_t)(h % kXRing) * 2 + ch) * n_slices * 32 + lane;
      // this frame's word from each slice, 16 slices per round, fetched with weak no-allocate loads and
      // re-fetched until every lap tag matches
      float m = -3.0e38f, z = 0.f;
      for (int s0 = 0; s0 < n_slices; s0 += 16) {
        const int ns = min(16, n_slices - s0);
        unsigned long long u[16];
        for (int attempt = 0;; attempt++) {
#pragma unroll
          for (int i = 0; i < 16; i++) u[i] = (i < ns) ? ld_na_u64(slot + (size_t)(s0 + i) * 32) : (tag << 63);
          bool ready = true;
#pragma unroll
          for (int i = 0; i < 16; i++) ready = ready && ((u[i] >> 63) == tag);
          if (__all_sync(0xFFFFFFFFu, ready) || (dbg & 1)) break;
          if (PROF) pacc[4] += 1000;  // 1000 per repeated fetch
          if (attempt > 2) __nanosleep(100);
        }
        float mn = m;
#pragma unroll
        for (int i = 0; i < 16; i++)
          if (i < ns) mn = fmaxf(mn, __uint_as_float((unsigned)u[i]));
        float zb = 0.f;
#pragma unroll
        for (int i = 0; i < 16; i++)
          if (i < ns)
            zb += __uint_as_float((unsigned)(u[i] >> 32) & 0x7FFFFFFFu) * ex2f(__uint_as_float((unsigned)u[i]) - mn);
        z = z * ex2f(m - mn) + zb;
        m = mn;
      }
      const float lse = m + lg2f(z);
      lse_s[rs * kHF + ch * 32 + lane] = lse;
      __syncwarp();
      if (lane == 0) mbar_arrive(sm.lse_ready(rs));
      LR_PT(3);
      if (slice == 0) {
        const long f = frame0 + (long)h * kHF + ch * 32 + lane;
        if (lse_out) lse_out[f] = lse;
        if (f < P && (!index || index[f] != kPadIndex)) llk_acc += (double)lse;
      }
    }
    if (slice == 0 && llk_sum) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) llk_acc += __shfl_xor_sync(0xFFFFFFFFu, llk_acc, o);
      if (lane == 0 && llk_acc != 0.0) atomicAdd(llk_sum, llk_acc * 0.69314718055994530942);
    }
    if (lw == 0) LR_PDUMP(2);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, 512);
}
#undef LR_PT
#undef LR_PDUMP

}  // namespace

// ------------------------------------------------------------------ host side
struct TcState {
  unsigned char *d_W = nullptr;  // [slices][64 KB]
  int *d_flag = nullptr;         // [0] weights out of the fp16 range, [1] normalisation re-derived
  double *d_cand = nullptr;      // candidate normalisation: g[64] | s[64]
  bool ok = false;
  unsigned long long norm_id = 0;
};

unsigned long long tc_norm_id(const lr_gmm *g) {
  const TcState *st = g ? (const TcState *)g->d_tc_w : nullptr;
  return st ? st->norm_id : 0;
}

// The normalised space is a free choice (any shift / scale gives the same mathematics; it only
// conditions the fp16 split), so the current one is KEPT while the mixture's global mean stays within a
// quarter deviation and its deviation within 25 % of it: EM preserves the first two moments of the data,
// so after the first iteration the frames' converted operand can be reused (lr_feats cache).
__global__ void k_tc_norm_adopt(int D, int have, const double *__restrict__ cand, double *__restrict__ g,
                                double *__restrict__ s, float *__restrict__ gf, float *__restrict__ rsf,
                                int *__restrict__ flag) {
  __shared__ int change;
  if (threadIdx.x == 0) change = have ? 0 : 1;
  __syncthreads();
  const int i = threadIdx.x;
  if (i < D && have) {
    const double dg = fabs(cand[i] - g[i]), ratio = cand[64 + i] / s[i];
    if (!(dg <= 0.25 * s[i]) || !(ratio >= 0.8 && ratio <= 1.25)) change = 1;
  }
  __syncthreads();
  if (change && i < D) {
    const float gfl = (float)cand[i], rsfl = (float)(1.0 / cand[64 + i]);
    gf[i] = gfl;
    rsf[i] = rsfl;
    g[i] = (double)gfl;  // the converter normalises with the fp32 values; weights and flush use exactly those
    s[i] = 1.0 / (double)rsfl;
  }
  if (i == 0) flag[1] = change;
}

bool tc_supported(const lr_gmm *g) {
  if (!g || g->D > kOneCol) return false;
  const TcState *st = (const TcState *)g->d_tc_w;
  return st && st->ok;
}

// Auto mode (lr_set_gmm_kernel(0)) takes the tensor-core path when the model supports it and
// has at least two slices of components; the SIMT path serves everything else.
bool tc_selected(const lr_gmm *g, lr_status *err) {
  Engine &e = engine();
  if (err) *err = LR_OK;
  if (e.gmm_kernel == 1) return false;
  if (e.gmm_kernel >= 2) {
    if (!tc_supported(g)) {
      if (err)
        *err = fail(LR_ERR_ARG,
                    "tcgen05 GMM kernel forced but this model (C=%d, D=%d) is outside its range",
                    g->C, g->D);
      return false;
    }
    return true;
  }
  return tc_supported(g) && g->C >= 256;
}

lr_status tc_derive(lr_gmm *g) {
  Engine &e = engine();
  if (g->D > kOneCol) return LR_OK;
  TcState *st = (TcState *)g->d_tc_w;
  if (!st) {
    st = new TcState();
    g->d_tc_w = st;
    size_t wbytes = (size_t)(g->Cp / kSlice) * 4 * kPanelBytes;
    LR_CUDA(cudaMalloc(&st->d_W, wbytes));
    LR_CUDA(cudaMalloc(&st->d_flag, 2 * sizeof(int)));
    LR_CUDA(cudaMalloc(&st->d_cand, 128 * sizeof(double)));
    LR_CUDA(cudaMalloc(&g->d_g, g->D * sizeof(double)));
    LR_CUDA(cudaMalloc(&g->d_s, g->D * sizeof(double)));
    LR_CUDA(cudaMalloc(&g->d_gf, 64 * sizeof(float)));
    LR_CUDA(cudaMalloc(&g->d_rsf, 64 * sizeof(float)));
  }
  LR_CUDA(cudaMemsetAsync(st->d_flag, 0, 2 * sizeof(int), e.stream));
  k_tc_norm<<<g->D, 256, 0, e.stream>>>(g->C, g->D, g->d_w, g->d_mean, g->d_cov, st->d_cand, st->d_cand + 64);
  LR_CHECK_LAUNCH();
  k_tc_norm_adopt<<<1, 64, 0, e.stream>>>(g->D, st->norm_id != 0 ? 1 : 0, st->d_cand, g->d_g, g->d_s, g->d_gf,
                                         g->d_rsf, st->d_flag);
  LR_CHECK_LAUNCH();
  k_tc_weights<<<ceil_div(g->Cp, 128), 128, 0, e.stream>>>(g->C, g->D, g->Cp, g->d_w, g->d_mean,
                                                           g->d_covinv, g->d_cst, g->d_g, g->d_s,
                                                           st->d_W, st->d_flag);
  LR_CHECK_LAUNCH();
  int flag[2] = {0, 0};
  LR_CUDA(cudaMemcpyAsync(flag, st->d_flag, 2 * sizeof(int), cudaMemcpyDeviceToHost, e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  st->ok = flag[0] == 0;  // weights outside the fp16 range -> this model stays on the SIMT path
  if (flag[1]) st->norm_id = ++e.norm_seq;  // a new normalised space: cached frame operands are stale
  return LR_OK;
}

void tc_free(lr_gmm *g) {
  TcState *st = (TcState *)g->d_tc_w;
  if (!st) return;
  cudaFree(st->d_W);
  cudaFree(st->d_flag);
  cudaFree(st->d_cand);
  delete st;
  g->d_tc_w = nullptr;
}

// Split [0, n_tiles) into `groups` contiguous ranges; cut points snap to run starts when a run
// boundary is close so that most rows stay inside one group.
static void split_groups(int n_tiles, int groups, std::vector<int> &cuts) {
  cuts.resize(groups + 1);
  for (int gidx = 0; gidx <= groups; gidx++) cuts[gidx] = (int)((long)n_tiles * gidx / groups);
}

// cluster size: the slices of one group that share a multicast tile (must divide n_slices)
static int tc_cluster_size(int n_slices) {
  // Measured on B200 (profiles/r01_tc_notes.md): multicast clusters cut L2->SM traffic 4x but the
  // kernels are bound by copy LATENCY (prefetch depth), not bandwidth, and 4-CTA clusters leave 20
  // SMs idle -- slower overall.  Kept behind debug bit 6 for experiments.
  if (!(engine().tc_debug & 64)) return 1;
  if (n_slices % 4 == 0) return 4;
  if (n_slices % 2 == 0) return 2;
  return 1;
}

template <typename... Args>
static lr_status tc_launch(void (*kern)(Args...), int grid, int csize, Args... args) {
  Engine &e = engine();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kTcThreads);
  cfg.dynamicSmemBytes = kTcSmem;
  cfg.stream = e.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)csize;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LR_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  count_launch();
  return LR_OK;
}

// number of groups: co-resident CTAs (1 per SM, whole clusters) / slices
template <typename K>
static int tc_groups(K kern, int n_slices, int csize, int n_tiles) {
  Engine &e = engine();
  int max_ctas = e.sm_count;
  if (csize > 1) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)(csize * 64));
    cfg.blockDim = dim3(kTcThreads);
    cfg.dynamicSmemBytes = kTcSmem;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = (unsigned)csize;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    int n_clusters = 0;
    if (cudaOccupancyMaxActiveClusters(&n_clusters, kern, &cfg) == cudaSuccess && n_clusters > 0)
      max_ctas = n_clusters * csize;
    else
      cudaGetLastError();
  }
  return std::max(1, std::min(n_tiles, max_ctas / n_slices));
}

static lr_status tc_set_attrs() {
  bool &done = engine().attr_set[Engine::kAttrTc];
  if (done) return LR_OK;
  LR_CUDA(cudaFuncSetAttribute(k_tc_lse, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  LR_CUDA(cudaFuncSetAttribute(k_tc_acc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  LR_CUDA(cudaFuncSetAttribute(k_tc_acc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTcSmem));
  LR_CUDA(cudaFuncSetAttribute(k_tc_one<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOneSmem));
  LR_CUDA(cudaFuncSetAttribute(k_tc_one<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOneSmem));
  LR_CUDA(cudaFuncSetAttribute(k_tc_one<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOneSmem));
  LR_CUDA(cudaFuncSetAttribute(k_tc_one<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kOneSmem));
  done = true;
  return LR_OK;
}

// Cooperative launch of the one-pass kernel: every CTA of a frame group waits for its peers'
// partial log-sum-exps, so the whole grid must be co-resident (the launch fails otherwise).
template <typename... Args>
static lr_status tc_launch_coop(void (*kern)(Args...), int grid, Args... args) {
  Engine &e = engine();
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)grid);
  cfg.blockDim = dim3(kOneThreads);
  cfg.dynamicSmemBytes = kOneSmem;
  cfg.stream = e.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeCooperative;
  attr[0].val.cooperative = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  LR_CUDA(cudaLaunchKernelEx(&cfg, kern, args...));
  count_launch();
  return LR_OK;
}

// The tensor-core path works on a PADDED frame list (P_pad = multiple of 128; padding entries
// carry index 0xFFFFFFFF): see tc_pad_plan in gmm_api.cu.  fl.P is the padded length here.
lr_status tc_pass_lse(lr_gmm *g, const FrameList &fl, float *d_lse2, double *d_llk_sum, float *d_S) {
  Engine &e = engine();
  TcState *st = (TcState *)g->d_tc_w;
  lr_status rc = tc_set_attrs();
  if (rc != LR_OK) return rc;
  const long P = fl.P;
  const long P_pad = (P + kTile - 1) / kTile * kTile;
  const int n_tiles = (int)(P_pad / kTile);
  const int n_slices = g->Cp / kSlice;
  const int csize = tc_cluster_size(n_slices);
  const int groups = tc_groups(k_tc_lse, n_slices, csize, n_tiles);
  unsigned char *Xh = (unsigned char *)scratch_get(kSlotTmpA, (size_t)n_tiles * kTileBytes);
  float2 *part = (float2 *)scratch_get(kSlotTmpB, (size_t)n_slices * P_pad * sizeof(float2));
  int *d_cuts = (int *)scratch_get(kSlotRest, (groups + 1) * sizeof(int));
  if (!Xh || !part || !d_cuts) return LR_ERR_CUDA;
  std::vector<int> cuts;
  split_groups(n_tiles, groups, cuts);
  LR_CUDA(cudaMemcpyAsync(d_cuts, cuts.data(), (groups + 1) * sizeof(int), cudaMemcpyHostToDevice,
                          e.stream));
  k_tc_convert<<<(unsigned)((P_pad * 8 + 255) / 256), 256, 0, e.stream>>>(
      g->D, fl.dX, fl.ldx, fl.d_index, P, P_pad, g->d_gf, g->d_rsf, Xh);
  LR_CHECK_LAUNCH();
  {
    ProfileScope prof(0);
    rc = tc_launch(k_tc_lse, n_slices * groups, csize, n_slices, csize,
                   (const unsigned char *)st->d_W, (const unsigned char *)Xh, (const int *)d_cuts,
                   P_pad, part, e.tc_debug, d_S, P, g->Cp);
    if (rc != LR_OK) return rc;
  }
  k_tc_combine<<<(unsigned)((P_pad + 255) / 256), 256, 0, e.stream>>>(n_slices, P, P_pad, fl.d_index,
                                                                     part, d_lse2, d_llk_sum);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

// Likelihood + statistics over a frame list whose row runs are padded to whole tiles
// (build_plan(..., pad = true)): chunk.pos % 128 == 0, padding entries carry kPadIndex.
lr_status tc_run_stats(lr_gmm *g, const FrameList &fl, const std::vector<LrChunk> &chunks,
                       double fw, double *out_N, double *out_F, double *out_S2,
                       double *d_llk_sum, unsigned char *conv, bool conv_valid) {
  Engine &e = engine();
  TcState *st = (TcState *)g->d_tc_w;
  lr_status rc = tc_set_attrs();
  if (rc != LR_OK) return rc;
  const long P_pad = (fl.P + kTile - 1) / kTile * kTile;
  const int n_tiles = (int)(P_pad / kTile);
  if (n_tiles == 0) return LR_OK;
  const int n_slices = g->Cp / kSlice;
  const int csize = tc_cluster_size(n_slices);
  // both passes must cut the tiles identically (pass 2 reuses pass 1's group table)
  const int groups = tc_groups(k_tc_lse, n_slices, csize, n_tiles);
  std::vector<int> cuts;
  split_groups(n_tiles, groups, cuts);

  // per-tile run info.  A run = consecutive tiles of one row inside one group, at most
  // kMaxRunTiles long; runs are flushed with fp64 atomics (RED), so a row may be shared between
  // groups (the EM case: one row, every group).
  std::vector<int> tile_row(n_tiles, -1);
  for (const LrChunk &c : chunks) {
    if (c.pos % kTile) return fail(LR_ERR_ARG, "tc_run_stats: chunk not tile aligned");
    for (long t = c.pos / kTile; t < (c.pos + c.len + kTile - 1) / kTile; t++) tile_row[t] = c.row;
  }
  std::vector<int> tile_group(n_tiles);
  for (int gi = 0; gi < groups; gi++)
    for (int t = cuts[gi]; t < cuts[gi + 1]; t++) tile_group[t] = gi;
  std::vector<TileInfo> tinfo(n_tiles);
  // profiling experiments: bits 16..23 of lr_debug_flags override the flush interval
  const int run_tiles = ((e.tc_debug >> 16) & 0xFF) ? ((e.tc_debug >> 16) & 0xFF) : kMaxRunTiles;
  for (int t = 0; t < n_tiles;) {
    const int row = tile_row[t], gi = tile_group[t];
    int t2 = t;
    while (t2 < n_tiles && tile_row[t2] == row && tile_group[t2] == gi) t2++;
    if (row < 0) return fail(LR_ERR_ARG, "tc_run_stats: tile %d is covered by no chunk", t);
    // Runs of at most run_tiles tiles.  A long stretch of one row (the EM case: one row, every group)
    // starts with a SHORTER first run that differs from group to group, so that the groups do not
    // flush their accumulators into the same C x D addresses at the same moment (144 CTAs x 15 k fp64
    // atomics in one burst otherwise).
    const int first_run = (t2 - t > 2 * run_tiles) ? std::max(1, run_tiles * (gi % groups + 1) / groups) : run_tiles;
    for (int k = t; k < t2; k++) {
      const int off = k - t;
      const int in_run = off < first_run ? off : (off - first_run) % run_tiles;
      const int run_len = off < first_run ? first_run : run_tiles;
      int flags = 0;
      if (in_run == 0) flags |= 1;
      if (in_run == run_len - 1 || k == t2 - 1) flags |= 2;
      tinfo[k] = {row, flags};
    }
    t = t2;
  }

  const bool want_stats = out_N || out_F || out_S2;
  if (want_stats && e.gmm_kernel != 3 && n_slices <= e.sm_count) {
    // ---- one pass: likelihood GEMM, log-sum-exp exchange and statistics GEMM in one kernel
    unsigned char *Xh1 = conv ? conv : (unsigned char *)scratch_get(kSlotTmpA, (size_t)n_tiles * kTileBytes);
    int *d_cuts1 = (int *)scratch_get(kSlotRest, (groups + 1) * sizeof(int));
    TileInfo *d_tinfo1 = (TileInfo *)scratch_get(kSlotChunks, (size_t)n_tiles * sizeof(TileInfo));
    // exchange ring: kXRing half tiles x n_slices x 64 frames of (max, sum) words per group.  A CTA
    // overwrites a ring slot 16 half tiles after it was written; by then every CTA of the group has
    // read it (a CTA reaches half tile h only after all peers posted h - 4, i.e. finished h - 8).
    const size_t xbytes = (size_t)groups * kXRing * n_slices * 64 * sizeof(unsigned long long);
    unsigned long long *d_xch = (unsigned long long *)scratch_get(kSlotXchg, xbytes);
    if (!Xh1 || !d_cuts1 || !d_tinfo1 || !d_xch) return LR_ERR_CUDA;
    LR_CUDA(cudaMemcpyAsync(d_cuts1, cuts.data(), (groups + 1) * sizeof(int), cudaMemcpyHostToDevice,
                            e.stream));
    LR_CUDA(cudaMemcpyAsync(d_tinfo1, tinfo.data(), (size_t)n_tiles * sizeof(TileInfo),
                            cudaMemcpyHostToDevice, e.stream));
    LR_CUDA(cudaMemsetAsync(d_xch, 0, xbytes, e.stream));  // lap tags start at 1
    if (!(conv && conv_valid)) {
      k_tc_convert<<<(unsigned)((P_pad * 8 + 255) / 256), 256, 0, e.stream>>>(
          g->D, fl.dX, fl.ldx, fl.d_index, fl.P, P_pad, g->d_gf, g->d_rsf, Xh1);
      LR_CHECK_LAUNCH();
    }
#ifdef LR_DEBUG_BUILD  // phase profiler / experiment switches: `make DEBUG=1` builds only
    static const bool kProf = getenv("LR_TC_PROF") != nullptr;
    static const int kEnvDbg = getenv("LR_TC_DEBUG") ? atoi(getenv("LR_TC_DEBUG")) : 0;
#else
    constexpr bool kProf = false;
    constexpr int kEnvDbg = 0;
#endif
    long long *d_prof = nullptr;
    if (kProf) {
      d_prof = (long long *)scratch_get(kSlotLse, 96 * sizeof(long long));
      if (!d_prof) return LR_ERR_CUDA;
      LR_CUDA(cudaMemsetAsync(d_prof, 0, 96 * sizeof(long long), e.stream));
    }
    auto kern = out_S2 ? (kProf ? k_tc_one<true, true> : k_tc_one<true, false>)
                       : (kProf ? k_tc_one<false, true> : k_tc_one<false, false>);
    {
      ProfileScope prof(1);
      rc = tc_launch_coop(kern, n_slices * groups, g->C, g->D, n_slices, (const unsigned char *)st->d_W,
                          (const unsigned char *)Xh1, (const int *)d_cuts1, (const TileInfo *)d_tinfo1,
                          d_xch, (float *)nullptr, fl.d_index, fl.P, d_llk_sum, (const double *)g->d_g,
                          (const double *)g->d_s, fw, out_N, out_F, out_S2,
                          e.tc_debug | kEnvDbg | (e.gmm_products >= 1 ? 256 : 0) | (e.gmm_products >= 2 ? 512 : 0), d_prof);
    }
    if (rc == LR_OK && kProf) {
      long long hp[96];
      LR_CUDA(cudaMemcpyAsync(hp, d_prof, sizeof(hp), cudaMemcpyDeviceToHost, e.stream));
      LR_CUDA(cudaStreamSynchronize(e.stream));
      const double per = 1.0 / ((double)n_slices * groups) / std::max(1, 2 * (n_tiles / groups));
      static const char *role[5] = {"E1 parity0", "E1 parity1", "L (publish + lse)", "G1 issuer", "G2 issuer"};
      for (int r = 0; r < 5; r++) {
        fprintf(stderr, "[tc_prof] %s clk per %s:", role[r], r < 3 ? "iteration (2 half tiles)" : "half tile");
        for (int i = 0; i < 8; i++) fprintf(stderr, " %d:%.0f", i, hp[r * 16 + i] * per * (r < 3 ? 2.0 : 1.0));
        fprintf(stderr, "\n");
      }
    }
    return rc;
  }
  float *d_lse = (float *)scratch_get(kSlotLse, (size_t)P_pad * sizeof(float));
  if (!d_lse) return LR_ERR_CUDA;
  rc = tc_pass_lse(g, fl, d_lse, d_llk_sum);  // also leaves the converted tiles in kSlotTmpA
  if (rc != LR_OK) return rc;
  if (!want_stats) return LR_OK;
  unsigned char *Xh = (unsigned char *)scratch_get(kSlotTmpA, (size_t)n_tiles * kTileBytes);
  int *d_cuts = (int *)scratch_get(kSlotRest, (groups + 1) * sizeof(int));
  TileInfo *d_tinfo = (TileInfo *)scratch_get(kSlotChunks, (size_t)n_tiles * sizeof(TileInfo));
  if (!Xh || !d_cuts || !d_tinfo) return LR_ERR_CUDA;
  LR_CUDA(cudaMemcpyAsync(d_tinfo, tinfo.data(), (size_t)n_tiles * sizeof(TileInfo),
                          cudaMemcpyHostToDevice, e.stream));
  {
    ProfileScope prof(1);
    if (out_S2)
      rc = tc_launch(k_tc_acc<true>, n_slices * groups, csize, g->C, g->D, n_slices, csize,
                     (const unsigned char *)st->d_W, (const unsigned char *)Xh, (const int *)d_cuts,
                     (const TileInfo *)d_tinfo, (const float *)d_lse, (const double *)g->d_g,
                     (const double *)g->d_s, fw, out_N, out_F, out_S2, e.tc_debug);
    else
      rc = tc_launch(k_tc_acc<false>, n_slices * groups, csize, g->C, g->D, n_slices, csize,
                     (const unsigned char *)st->d_W, (const unsigned char *)Xh, (const int *)d_cuts,
                     (const TileInfo *)d_tinfo, (const float *)d_lse, (const double *)g->d_g,
                     (const double *)g->d_s, fw, out_N, out_F, (double *)nullptr, e.tc_debug);
    if (rc != LR_OK) return rc;
  }
  return LR_OK;
}

}  // namespace lr
