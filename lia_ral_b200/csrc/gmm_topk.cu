// gmm_topk.cu -- top-distribution selection and top-distribution rescoring
// (MixtureGDStat::computeAndAccumulateLLK with DETERMINE_TOP_DISTRIBS / USE_TOP_DISTRIBS
// [alize-core]; call sites ComputeTest.cpp:162-167, TopGauss.cpp:166-192).
//
// Index selection must equal the fp64 reference ordering bit for bit, so the fp32 scores of
// pass 1 only nominate candidates: every component within kTopEps (log2 units, far above the
// fp32 error) of the K-th best is re-evaluated in fp64 with the reference's operation order,
// and the final ranking uses those fp64 likelihoods (ties -> lowest index).
#include "gmm_topk.cuh"

namespace lr {

namespace {

constexpr int kWarps = 4;
constexpr float kTopEps = 0.05f;

// p_c(x) = w_c cst_c exp(-0.5 sum_i (x_i - mu_ci)^2 covinv_ci) with the oracle's rounding
// sequence (no FMA contraction): d = x - mu; q += (d*d)*covinv.
__device__ __forceinline__ double comp_lk(int D, const float *__restrict__ x,
                                          const double *__restrict__ mean,
                                          const double *__restrict__ covinv, double w, double cst) {
  double q = 0.0;
  for (int i = 0; i < D; i++) {
    double d = __dsub_rn((double)x[i], mean[i]);
    q = __dadd_rn(q, __dmul_rn(__dmul_rn(d, d), covinv[i]));
  }
  double lk = __dmul_rn(cst, exp(__dmul_rn(-0.5, q)));
  if (isnan(lk)) lk = 1e-200;  // EPS_LK, TopGauss.cpp:67
  return __dmul_rn(w, lk);
}

__device__ __forceinline__ double clamp_llk(double lk, double lo, double hi) {
  double l = log(lk);
  if (isnan(l)) return lo;
  if (l <= lo) return lo;
  if (l >= hi) return hi;
  return l;
}

// One warp per frame position.  S row staged in shared memory.
__global__ void __launch_bounds__(kWarps * 32)
k_topk(int C, int D, int Cp, const float *__restrict__ X, size_t ldx,
       const unsigned *__restrict__ index, long P, const float *__restrict__ S, int K,
       int complete, double min_llk, double max_llk, const double *__restrict__ w,
       const double *__restrict__ mean, const double *__restrict__ covinv,
       const double *__restrict__ cst, double *__restrict__ llk, unsigned *__restrict__ idx_out,
       double *__restrict__ top_lk, double *__restrict__ rest_lk, double *__restrict__ rest_w) {
  extern __shared__ __align__(16) float smem_f[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *row = smem_f + (size_t)warp * Cp;
  __shared__ int cand_idx[kWarps][kMaxCand];
  __shared__ double cand_p[kWarps][kMaxCand];
  __shared__ int sorted_c[kWarps][kMaxCand];
  __shared__ double sorted_p[kWarps][kMaxCand];

  long p = (long)blockIdx.x * kWarps + warp;
  if (p >= P) return;
  const float *srow = S + (size_t)p * Cp;
  for (int c = lane; c < Cp; c += 32) row[c] = c < C ? srow[c] : -3.0e38f;
  __syncwarp();

  // nominate candidates in descending fp32 order
  int ncand = 0;
  float kth = -3.0e38f;
  while (ncand < kMaxCand) {
    float best = -3.0e38f;
    int bi = 0x7fffffff;
    for (int c = lane; c < Cp; c += 32) {
      float v = row[c];
      if (v > best) {
        best = v;
        bi = c;
      }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      float ov = __shfl_xor_sync(0xffffffffu, best, o);
      int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ov > best || (ov == best && oi < bi)) {
        best = ov;
        bi = oi;
      }
    }
    if (bi >= C) break;  // nothing left
    if (ncand >= K && best < kth - kTopEps) break;
    if (lane == 0) {
      cand_idx[warp][ncand] = bi;
      row[bi] = -3.0e38f;
    }
    ncand++;
    if (ncand == K) kth = best;
    __syncwarp();
  }
  __syncwarp();

  // fp64 re-evaluation
  size_t fr = index ? (size_t)index[p] : (size_t)p;
  const float *x = X + fr * ldx;
  for (int j = lane; j < ncand; j += 32) {
    int c = cand_idx[warp][j];
    cand_p[warp][j] = comp_lk(D, x, mean + (size_t)c * D, covinv + (size_t)c * D, w[c], cst[c]);
  }
  __syncwarp();

  // rank by (p desc, index asc); ranks are unique, so this is a scatter into sorted order
  double lsum_rest = 0.0;
  for (int j = lane; j < ncand; j += 32) {
    double pj = cand_p[warp][j];
    int cj = cand_idx[warp][j];
    int rank = 0;
    for (int q = 0; q < ncand; q++) {
      double pq = cand_p[warp][q];
      int cq = cand_idx[warp][q];
      if (pq > pj || (pq == pj && cq < cj)) rank++;
    }
    sorted_p[warp][rank] = pj;
    sorted_c[warp][rank] = cj;
    if (rank < K) {
      idx_out[(size_t)p * K + rank] = (unsigned)cj;
      if (top_lk) top_lk[(size_t)p * K + rank] = pj;
    } else {
      lsum_rest += pj;
    }
  }
  __syncwarp();
  // remaining (never nominated) components, fp32 log domain
  float mr = -3.0e38f;
  for (int c = lane; c < C; c += 32) mr = fmaxf(mr, row[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mr = fmaxf(mr, __shfl_xor_sync(0xffffffffu, mr, o));
  float rs = 0.f;
  if (mr > -1.0e37f)
    for (int c = lane; c < C; c += 32) {
      float v = row[c];
      if (v > -1.0e37f) rs += exp2f(v - mr);
    }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    rs += __shfl_xor_sync(0xffffffffu, rs, o);
    lsum_rest += __shfl_xor_sync(0xffffffffu, lsum_rest, o);
  }
  if (lane == 0) {
    double rest = lsum_rest;
    if (mr > -1.0e37f) rest += (double)rs * exp2((double)mr);
    double top_sum = 0.0, top_w = 0.0;
    int kk = K < ncand ? K : ncand;
    // summed in descending order like the reference's sorted LKVector
    for (int r = 0; r < kk; r++) {
      top_sum += sorted_p[warp][r];
      top_w += w[sorted_c[warp][r]];
    }
    if (rest_lk) rest_lk[p] = rest;
    if (rest_w) rest_w[p] = 1.0 - top_w;
    if (llk) llk[p] = clamp_llk(complete ? top_sum + rest : top_sum, min_llk, max_llk);
  }
}

// USE_TOP_DISTRIBS on a client model: one warp per frame, lanes over the K stored indices.
__global__ void __launch_bounds__(128)
k_use_topk(int D, const float *__restrict__ X, size_t ldx, const unsigned *__restrict__ index,
           long P, int K, const unsigned *__restrict__ idx, const double *__restrict__ rest_lk,
           int complete, double min_llk, double max_llk, const double *__restrict__ w,
           const double *__restrict__ mean, const double *__restrict__ covinv,
           const double *__restrict__ cst, double *__restrict__ llk) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  long p = (long)blockIdx.x * 4 + warp;
  if (p >= P) return;
  size_t fr = index ? (size_t)index[p] : (size_t)p;
  const float *x = X + fr * ldx;
  double lk = 0.0;
  for (int k = lane; k < K; k += 32) {
    unsigned c = idx[(size_t)p * K + k];
    lk += comp_lk(D, x, mean + (size_t)c * D, covinv + (size_t)c * D, w[c], cst[c]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) lk += __shfl_xor_sync(0xffffffffu, lk, o);
  if (lane == 0) {
    if (complete && rest_lk) lk += rest_lk[p];
    llk[p] = clamp_llk(lk, min_llk, max_llk);
  }
}

// TOP_DISTRIBS_NO_ACTION: llk = ln2 * lse2, clamped
__global__ void k_llk_from_lse(long P, const float *__restrict__ lse2, double min_llk,
                               double max_llk, double *__restrict__ llk) {
  long p = (long)blockIdx.x * blockDim.x + threadIdx.x;
  if (p >= P) return;
  double l = (double)lse2[p] * 0.69314718055994530942;
  if (isnan(l) || l <= min_llk) l = min_llk;
  if (l >= max_llk) l = max_llk;
  llk[p] = l;
}

}  // namespace

lr_status gmm_topk(lr_gmm *g, const FrameList &fl, const float *d_S, int K, int complete,
                   double min_llk, double max_llk, double *d_llk, unsigned *d_idx,
                   double *d_top_lk, double *d_rest_lk, double *d_rest_w) {
  if (fl.P <= 0) return LR_OK;
  Engine &e = engine();
  size_t sm = (size_t)kWarps * g->Cp * sizeof(float);
  if (sm > 200 * 1024) return fail(LR_ERR_ARG, "top-K selection supports at most %d components",
                                   200 * 1024 / (kWarps * 4));
  bool &attr_set = engine().attr_set[Engine::kAttrTopk];
  if (!attr_set) {
    LR_CUDA(cudaFuncSetAttribute(k_topk, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set = true;
  }
  long grid = (fl.P + kWarps - 1) / kWarps;
  k_topk<<<(unsigned)grid, kWarps * 32, sm, e.stream>>>(
      g->C, g->D, g->Cp, fl.dX, fl.ldx, fl.d_index, fl.P, d_S, K, complete, min_llk, max_llk,
      g->d_w, g->d_mean, g->d_covinv, g->d_cst, d_llk, d_idx, d_top_lk, d_rest_lk, d_rest_w);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

lr_status gmm_use_topk(lr_gmm *g, const FrameList &fl, int K, const unsigned *d_idx,
                       const double *d_rest_lk, int complete, double min_llk, double max_llk,
                       double *d_llk) {
  if (fl.P <= 0) return LR_OK;
  Engine &e = engine();
  long grid = (fl.P + 3) / 4;
  k_use_topk<<<(unsigned)grid, 128, 0, e.stream>>>(g->D, fl.dX, fl.ldx, fl.d_index, fl.P, K, d_idx,
                                                   d_rest_lk, complete, min_llk, max_llk, g->d_w,
                                                   g->d_mean, g->d_covinv, g->d_cst, d_llk);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

lr_status gmm_llk_from_lse(long P, const float *d_lse2, double min_llk, double max_llk,
                           double *d_llk) {
  if (P <= 0) return LR_OK;
  Engine &e = engine();
  k_llk_from_lse<<<(unsigned)((P + 255) / 256), 256, 0, e.stream>>>(P, d_lse2, min_llk, max_llk,
                                                                     d_llk);
  LR_CHECK_LAUNCH();
  return LR_OK;
}

}  // namespace lr
