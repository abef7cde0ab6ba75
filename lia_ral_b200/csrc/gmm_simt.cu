// gmm_simt.cu -- fp32 SIMT implementation of the frames x components passes.
//
// This is the precise general-shape path (any C, D <= 63): the Mahalanobis term is evaluated in
// the "direct" form sum_i (x_i sa_ci - mu_ci sa_ci)^2 (two FFMA per term, no cancellation), the
// per-frame log-sum-exp runs in the log2 domain, and the Baum-Welch / EM statistics are
// accumulated CENTRED on each component mean, sum_t g (x - mu_c) and sum_t g (x - mu_c)^2, so
// that F - mu N (substractM) and m2/occ - mu^2 (getEM) lose no digits; partial sums leave the
// SM as fp64 atomics every kChunkFrames frames.
//
// Replaces, per frame: DistribGD::computeLK, MixtureGDStat::computeAndAccumulate{Occ,EM,LLK}
// [alize-core] and the loops AccumulateTVStat.cpp:332-349, AccumulateStat.cpp:103-128.
#include "gmm.cuh"

namespace lr {

namespace {

constexpr int kThreads = 256;
constexpr int kCT = 64;        // components per tile
constexpr int kF1 = 128;       // frames per tile, pass 1
constexpr int kF1Pitch = 132;  // xs pitch pass 1 (16B-aligned rows)
constexpr int kF2 = 64;        // frames per tile, pass 2
constexpr int kF2Pitch = 68;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Load the [D x CT] slices of sa / nm for component tile ct into shared memory.
__device__ __forceinline__ void load_model_tile(const float *__restrict__ sa,
                                                const float *__restrict__ nm, int D, int Cp,
                                                int ct, float *sas, float *nms) {
  for (int idx = threadIdx.x; idx < D * kCT; idx += kThreads) {
    int i = idx >> 6, j = idx & 63;
    sas[idx] = sa[(size_t)i * Cp + ct * kCT + j];
    nms[idx] = nm[(size_t)i * Cp + ct * kCT + j];
  }
}

// ------------------------------------------------------------------ pass 1: log-sum-exp
// One CTA per 128 positions; loops over all component tiles keeping a per-thread online
// (max, sum) for its 4 component columns, merged across the 16 column-threads at the end.
__global__ void __launch_bounds__(kThreads, 2)
k_lse(int D, int Cp, const float *__restrict__ X, size_t ldx, const unsigned *__restrict__ index,
      long P, const float *__restrict__ sa, const float *__restrict__ nm,
      const float *__restrict__ const2, float *__restrict__ lse2, float *__restrict__ S_out,
      double *__restrict__ llk_sum) {
  extern __shared__ __align__(16) float smem[];
  float *xs = smem;                   // [D][kF1Pitch]
  float *sas = xs + D * kF1Pitch;     // [D][64]
  float *nms = sas + D * kCT;         // [D][64]
  __shared__ double red[kThreads / 32];

  const int tid = threadIdx.x;
  const int tg = tid >> 4, cg = tid & 15;
  const long p0 = (long)blockIdx.x * kF1;

  for (int idx = tid; idx < kF1 * D; idx += kThreads) {
    int t = idx / D, i = idx - t * D;
    long p = p0 + t;
    float v = 0.f;
    if (p < P) {
      size_t fr = index ? (size_t)index[p] : (size_t)p;
      v = X[fr * ldx + i];
    }
    xs[i * kF1Pitch + t] = v;
  }

  float mrun[8], srun[8];
#pragma unroll
  for (int f = 0; f < 8; f++) {
    mrun[f] = -3.0e38f;
    srun[f] = 0.f;
  }

  const int nct = Cp / kCT;
  for (int ct = 0; ct < nct; ct++) {
    __syncthreads();  // previous tile's readers done (and xs visible on the first trip)
    load_model_tile(sa, nm, D, Cp, ct, sas, nms);
    __syncthreads();
    float acc[8][4];
#pragma unroll
    for (int f = 0; f < 8; f++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[f][c] = 0.f;
#pragma unroll 4
    for (int i = 0; i < D; i++) {
      float4 xa = *reinterpret_cast<const float4 *>(&xs[i * kF1Pitch + tg * 8]);
      float4 xb = *reinterpret_cast<const float4 *>(&xs[i * kF1Pitch + tg * 8 + 4]);
      float4 s4 = *reinterpret_cast<const float4 *>(&sas[i * kCT + cg * 4]);
      float4 n4 = *reinterpret_cast<const float4 *>(&nms[i * kCT + cg * 4]);
      float x[8] = {xa.x, xa.y, xa.z, xa.w, xb.x, xb.y, xb.z, xb.w};
      float s[4] = {s4.x, s4.y, s4.z, s4.w};
      float n[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
      for (int f = 0; f < 8; f++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          float e = fmaf(x[f], s[c], n[c]);
          acc[f][c] = fmaf(e, e, acc[f][c]);
        }
    }
    float4 k4 = *reinterpret_cast<const float4 *>(&const2[ct * kCT + cg * 4]);
    float k[4] = {k4.x, k4.y, k4.z, k4.w};
#pragma unroll
    for (int f = 0; f < 8; f++) {
      float v[4];
#pragma unroll
      for (int c = 0; c < 4; c++) v[c] = k[c] - acc[f][c];
      if (S_out) {
        long p = p0 + tg * 8 + f;
        if (p < P)
          *reinterpret_cast<float4 *>(&S_out[(size_t)p * Cp + ct * kCT + cg * 4]) =
              make_float4(v[0], v[1], v[2], v[3]);
      }
      float m = fmaxf(fmaxf(v[0], v[1]), fmaxf(v[2], v[3]));
      float mn = fmaxf(mrun[f], m);
      srun[f] = srun[f] * ex2(mrun[f] - mn) + ex2(v[0] - mn) + ex2(v[1] - mn) + ex2(v[2] - mn) +
                ex2(v[3] - mn);
      mrun[f] = mn;
    }
  }

  // merge the 16 column-threads of each frame (lanes 0-15 / 16-31 of a warp)
  double part = 0.0;
#pragma unroll
  for (int f = 0; f < 8; f++) {
    float m = mrun[f], s = srun[f];
#pragma unroll
    for (int o = 8; o > 0; o >>= 1) {
      float mo = __shfl_xor_sync(0xffffffffu, m, o);
      float so = __shfl_xor_sync(0xffffffffu, s, o);
      float mn = fmaxf(m, mo);
      s = s * ex2(m - mn) + so * ex2(mo - mn);
      m = mn;
    }
    long p = p0 + tg * 8 + f;
    if (cg == 0 && p < P) {
      float l = m + log2f(s);
      lse2[p] = l;
      part += (double)l;
    }
  }
  if (llk_sum) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    if (tid == 0) {
      double s = 0.0;
      for (int w = 0; w < kThreads / 32; w++) s += red[w];
      atomicAdd(llk_sum, s * 0.69314718055994530942);
    }
  }
}

// ------------------------------------------------------------------ pass 2: statistics
// One CTA per (chunk, component tile): recompute S for its 64 components, gamma =
// exp2(S - lse2), then acc[c][d] += gamma (x_d - mu_cd) (and (x_d - mu_cd)^2 for EM) with a
// 4 x 4 register tile per thread; column D of the value tile is 1 so the same loop yields N.
template <bool EM>
__global__ void __launch_bounds__(kThreads, 2)
k_acc(int C, int D, int Cp, const float *__restrict__ X, size_t ldx,
      const unsigned *__restrict__ index, const float *__restrict__ lse2,
      const LrChunk *__restrict__ chunks, const float *__restrict__ sa,
      const float *__restrict__ nm, const float *__restrict__ const2,
      const float *__restrict__ mean_f, const double *__restrict__ mean_d, double fw,
      double *__restrict__ out_N, double *__restrict__ out_F, double *__restrict__ out_S2) {
  extern __shared__ __align__(16) float smem[];
  float *xs = smem;                 // [D][kF2Pitch]  transposed frames (S computation)
  float *sas = xs + D * kF2Pitch;   // [D][64]
  float *nms = sas + D * kCT;       // [D][64]
  float *xr = nms + D * kCT;        // [kF2][64]      row-major [x | 1 | 0]
  float *gs = xr + kF2 * 64;        // [kF2][64]      posteriors
  float *ls = gs + kF2 * 64;        // [kF2]          lse2 of the tile (or +inf for padding rows)
  __shared__ double nsh[kCT];

  const int tid = threadIdx.x;
  const int nct = Cp / kCT;
  const int ct = blockIdx.x % nct;
  const LrChunk ch = chunks[blockIdx.x / nct];
  const int tg = tid >> 4, cg = tid & 15;  // S phase: frames tg*4.., comps cg*4..
  const int cgrp = tid >> 4, dgrp = tid & 15;  // accumulate phase: comps cgrp*4.., dims dgrp*4..

  load_model_tile(sa, nm, D, Cp, ct, sas, nms);
  float mu[4][4];
#pragma unroll
  for (int c = 0; c < 4; c++) {
    float4 m4 = *reinterpret_cast<const float4 *>(
        &mean_f[(size_t)(ct * kCT + cgrp * 4 + c) * 64 + dgrp * 4]);
    mu[c][0] = m4.x;
    mu[c][1] = m4.y;
    mu[c][2] = m4.z;
    mu[c][3] = m4.w;
  }
  float4 k4 = *reinterpret_cast<const float4 *>(&const2[ct * kCT + cg * 4]);
  const float kc[4] = {k4.x, k4.y, k4.z, k4.w};

  float a1[4][4], a2[4][4];
#pragma unroll
  for (int c = 0; c < 4; c++)
#pragma unroll
    for (int d = 0; d < 4; d++) {
      a1[c][d] = 0.f;
      a2[c][d] = 0.f;
    }

  for (int t0 = 0; t0 < ch.len; t0 += kF2) {
    __syncthreads();  // previous tile fully consumed
    for (int idx = tid; idx < kF2 * 64; idx += kThreads) {
      int t = idx >> 6, i = idx & 63;
      bool valid = (t0 + t) < ch.len;
      float v = 0.f;
      if (i < D) {
        if (valid) {
          long p = ch.pos + t0 + t;
          size_t fr = index ? (size_t)index[p] : (size_t)p;
          v = X[fr * ldx + i];
        }
        xs[i * kF2Pitch + t] = v;
      } else if (i == D) {
        v = valid ? 1.f : 0.f;
      }
      xr[idx] = v;
    }
    if (tid < kF2) ls[tid] = (t0 + tid) < ch.len ? lse2[ch.pos + t0 + tid] : 3.0e38f;
    __syncthreads();

    float acc[4][4];
#pragma unroll
    for (int f = 0; f < 4; f++)
#pragma unroll
      for (int c = 0; c < 4; c++) acc[f][c] = 0.f;
#pragma unroll 4
    for (int i = 0; i < D; i++) {
      float4 x4 = *reinterpret_cast<const float4 *>(&xs[i * kF2Pitch + tg * 4]);
      float4 s4 = *reinterpret_cast<const float4 *>(&sas[i * kCT + cg * 4]);
      float4 n4 = *reinterpret_cast<const float4 *>(&nms[i * kCT + cg * 4]);
      float x[4] = {x4.x, x4.y, x4.z, x4.w};
      float s[4] = {s4.x, s4.y, s4.z, s4.w};
      float n[4] = {n4.x, n4.y, n4.z, n4.w};
#pragma unroll
      for (int f = 0; f < 4; f++)
#pragma unroll
        for (int c = 0; c < 4; c++) {
          float e = fmaf(x[f], s[c], n[c]);
          acc[f][c] = fmaf(e, e, acc[f][c]);
        }
    }
#pragma unroll
    for (int f = 0; f < 4; f++) {
      float l = ls[tg * 4 + f];
      float4 g4;
      g4.x = ex2(kc[0] - acc[f][0] - l);
      g4.y = ex2(kc[1] - acc[f][1] - l);
      g4.z = ex2(kc[2] - acc[f][2] - l);
      g4.w = ex2(kc[3] - acc[f][3] - l);
      *reinterpret_cast<float4 *>(&gs[(tg * 4 + f) * 64 + cg * 4]) = g4;
    }
    __syncthreads();

#pragma unroll 4
    for (int t = 0; t < kF2; t++) {
      float4 g4 = *reinterpret_cast<const float4 *>(&gs[t * 64 + cgrp * 4]);
      float4 x4 = *reinterpret_cast<const float4 *>(&xr[t * 64 + dgrp * 4]);
      float g[4] = {g4.x, g4.y, g4.z, g4.w};
      float x[4] = {x4.x, x4.y, x4.z, x4.w};
#pragma unroll
      for (int c = 0; c < 4; c++)
#pragma unroll
        for (int d = 0; d < 4; d++) {
          float dd = x[d] - mu[c][d];
          a1[c][d] = fmaf(g[c], dd, a1[c][d]);
          if (EM) a2[c][d] = fmaf(g[c] * dd, dd, a2[c][d]);
        }
    }
  }

  // flush: N from the ones column, then F = Fc + mu N, S2 = S2c + 2 mu Fc + mu^2 N (fp64)
  if (dgrp == (D >> 2)) {
#pragma unroll
    for (int c = 0; c < 4; c++) {
      float v = a1[c][0];
      if ((D & 3) == 1) v = a1[c][1];
      if ((D & 3) == 2) v = a1[c][2];
      if ((D & 3) == 3) v = a1[c][3];
      nsh[cgrp * 4 + c] = (double)v;
    }
  }
  __syncthreads();
#pragma unroll
  for (int c = 0; c < 4; c++) {
    int comp = ct * kCT + cgrp * 4 + c;
    if (comp >= C) continue;
    double n = nsh[cgrp * 4 + c];
    size_t rc = (size_t)ch.row * C + comp;
    if (dgrp == 0 && out_N) atomicAdd(&out_N[rc], fw * n);
#pragma unroll
    for (int d = 0; d < 4; d++) {
      int dim = dgrp * 4 + d;
      if (dim >= D) continue;
      double m = mean_d[(size_t)comp * D + dim];
      double fc = (double)a1[c][d];
      if (out_F) atomicAdd(&out_F[rc * D + dim], fw * (fc + m * n));
      if (EM && out_S2)
        atomicAdd(&out_S2[rc * D + dim], fw * ((double)a2[c][d] + 2.0 * m * fc + m * m * n));
    }
  }
}

}  // namespace

static size_t lse_smem(int D) { return (size_t)(D * kF1Pitch + 2 * D * kCT) * sizeof(float); }
static size_t acc_smem(int D) {
  return (size_t)(D * kF2Pitch + 2 * D * kCT + 2 * kF2 * 64 + kF2) * sizeof(float);
}

lr_status gmm_pass_lse(lr_gmm *g, const FrameList &fl, float *d_lse2, float *d_S,
                       double *d_llk_sum) {
  if (fl.P <= 0) return LR_OK;
  Engine &e = engine();
  size_t sm = lse_smem(g->D);
  bool &attr_set = engine().attr_set[Engine::kAttrSimtLse];
  if (!attr_set) {
    LR_CUDA(cudaFuncSetAttribute(k_lse, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    attr_set = true;
  }
  long grid = (fl.P + kF1 - 1) / kF1;
  ProfileScope prof(0);
  k_lse<<<(unsigned)grid, kThreads, sm, e.stream>>>(g->D, g->Cp, fl.dX, fl.ldx, fl.d_index, fl.P,
                                                    g->d_sa, g->d_nm, g->d_const2, d_lse2, d_S,
                                                    d_llk_sum);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

lr_status gmm_pass_acc(lr_gmm *g, const FrameList &fl, const float *d_lse2,
                       const LrChunk *d_chunks, int n_chunks, double fw, double *out_N,
                       double *out_F, double *out_S2) {
  if (n_chunks <= 0) return LR_OK;
  Engine &e = engine();
  size_t sm = acc_smem(g->D);
  bool &attr_set = engine().attr_set[Engine::kAttrSimtAcc];
  if (!attr_set) {
    LR_CUDA(cudaFuncSetAttribute(k_acc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 100 * 1024));
    LR_CUDA(cudaFuncSetAttribute(k_acc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 100 * 1024));
    attr_set = true;
  }
  long grid = (long)n_chunks * (g->Cp / kCT);
  ProfileScope prof(1);
  if (out_S2)
    k_acc<true><<<(unsigned)grid, kThreads, sm, e.stream>>>(
        g->C, g->D, g->Cp, fl.dX, fl.ldx, fl.d_index, d_lse2, d_chunks, g->d_sa, g->d_nm,
        g->d_const2, g->d_mean_f, g->d_mean, fw, out_N, out_F, out_S2);
  else
    k_acc<false><<<(unsigned)grid, kThreads, sm, e.stream>>>(
        g->C, g->D, g->Cp, fl.dX, fl.ldx, fl.d_index, d_lse2, d_chunks, g->d_sa, g->d_nm,
        g->d_const2, g->d_mean_f, g->d_mean, fw, out_N, out_F, nullptr);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

}  // namespace lr
