// gmm.cu -- the device-resident MixtureGD twin: parameters, DistribGD::computeAll, the EM
// re-estimation (MixtureGDStat::getEM + varianceControl) and the FeatureServer frame block.
#include "gmm.cuh"

#include <cmath>

namespace lr {

// ---- DistribGD::computeAll [alize-core; constants probed on TrainWorld/test/wld.validate]:
// covInv = 1/cov, det = prod cov, cst = 1/((2 pi)^(D/2) sqrt(det)); then the fp32 operands of
// the frames x components passes.  One thread per component.
__global__ void k_gmm_derive(int C, int D, int Cp, const double *__restrict__ w,
                             const double *__restrict__ mean, const double *__restrict__ cov,
                             double *__restrict__ covinv, double *__restrict__ det,
                             double *__restrict__ cst, int cst_override, float *__restrict__ sa,
                             float *__restrict__ nm, float *__restrict__ const2,
                             float *__restrict__ mean_f) {
  int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= Cp) return;
  if (c >= C) {  // padding component: contributes exp2(-1e30) = 0 everywhere
    for (int i = 0; i < D; i++) {
      sa[(size_t)i * Cp + c] = 0.f;
      nm[(size_t)i * Cp + c] = 0.f;
    }
    for (int i = 0; i < 64; i++) mean_f[(size_t)c * 64 + i] = 0.f;
    const2[c] = -1e30f;
    return;
  }
  const double kHalfLog2e = 0.72134752044448170368;  // 0.5 * log2(e)
  double dt = 1.0;
  for (int i = 0; i < D; i++) {
    double v = cov[(size_t)c * D + i];
    dt *= v;
    double ci = 1.0 / v;
    covinv[(size_t)c * D + i] = ci;
    double s = sqrt(kHalfLog2e * ci);
    sa[(size_t)i * Cp + c] = (float)s;
    nm[(size_t)i * Cp + c] = (float)(-mean[(size_t)c * D + i] * s);
    mean_f[(size_t)c * 64 + i] = (float)mean[(size_t)c * D + i];
  }
  for (int i = D; i < 64; i++) mean_f[(size_t)c * 64 + i] = 0.f;
  det[c] = dt;
  double k;
  if (cst_override) {
    k = cst[c];
  } else {
    k = 1.0 / (pow(2.0 * 3.14159265358979323846, 0.5 * D) * sqrt(dt));
    cst[c] = k;
  }
  double l2 = log2(w[c]) + log2(k);
  const2[c] = (w[c] > 0.0 && k > 0.0 && isfinite(l2)) ? (float)l2 : -1e30f;
}

lr_status gmm_derive(lr_gmm *g) {
  Engine &e = engine();
  int th = 128;
  k_gmm_derive<<<ceil_div(g->Cp, th), th, 0, e.stream>>>(
      g->C, g->D, g->Cp, g->d_w, g->d_mean, g->d_cov, g->d_covinv, g->d_det, g->d_cst,
      g->cst_override ? 1 : 0, g->d_sa, g->d_nm, g->d_const2, g->d_mean_f);
  LR_CHECK_LAUNCH();
  return tc_derive(g);  // no-op for shapes the tensor-core path does not take
}

// ---- MixtureGDStat::getEM [alize-core] + varianceControl (TrainTools.cpp:567-587):
// stats = [occ C | m1 C*D | m2 C*D | llk | n].
__global__ void k_em_total(int C, const double *__restrict__ occ, double *__restrict__ tot) {
  __shared__ double sh[256];
  double s = 0.0;
  for (int c = threadIdx.x; c < C; c += blockDim.x) s += occ[c];
  sh[threadIdx.x] = s;
  __syncthreads();
  for (int o = blockDim.x / 2; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) *tot = sh[0];
}

__global__ void k_em_update(int C, int D, const double *__restrict__ occ,
                            const double *__restrict__ m1, const double *__restrict__ m2,
                            const double *__restrict__ tot, double flooring, double ceiling,
                            const double *__restrict__ cov_signal, double *__restrict__ w,
                            double *__restrict__ mean, double *__restrict__ cov) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= C * D) return;
  int c = idx / D, i = idx - c * D;
  double o = occ[c];
  if (i == 0 && *tot > 0.0) w[c] = o / *tot;  // (no frame selected at all: the weights stay)
  double cv = cov[idx];
  if (o > 0.0) {  // components with no occupation keep their parameters
    double mu = m1[idx] / o;
    mean[idx] = mu;
    cv = m2[idx] / o - mu * mu;
  }
  if (cov_signal) {  // floor first, then ceiling; "<=" / ">=" like the reference
    double lo = flooring * cov_signal[i], hi = ceiling * cov_signal[i];
    if (cv <= lo) cv = lo;
    if (cv >= hi) cv = hi;
  }
  cov[idx] = cv;
}

// ---- FrameAccGD via computeMeanCov (TrainTools.cpp:593-602)
__global__ void k_mean_cov_acc(const float *__restrict__ X, long T, size_t ldx, int D,
                               double *__restrict__ acc /*[2*D]*/) {
  // blockDim = (64, 4): x = dimension, y = frame lane
  int i = threadIdx.x;
  double s = 0.0, s2 = 0.0;
  if (i < D) {
    for (long t = (long)blockIdx.x * blockDim.y + threadIdx.y; t < T;
         t += (long)gridDim.x * blockDim.y) {
      double v = (double)X[(size_t)t * ldx + i];
      s += v;
      s2 += v * v;
    }
  }
  __shared__ double sh[2][4][64];
  sh[0][threadIdx.y][i] = s;
  sh[1][threadIdx.y][i] = s2;
  __syncthreads();
  if (threadIdx.y == 0 && i < D) {
    for (int y = 1; y < 4; y++) {
      s += sh[0][y][i];
      s2 += sh[1][y][i];
    }
    atomicAdd(&acc[i], s);
    atomicAdd(&acc[D + i], s2);
  }
}

static void gmm_free(lr_gmm *g) {
  if (!g) return;
  cudaFree(g->d_w);
  cudaFree(g->d_mean);
  cudaFree(g->d_cov);
  cudaFree(g->d_covinv);
  cudaFree(g->d_cst);
  cudaFree(g->d_det);
  cudaFree(g->d_tot);
  cudaFree(g->d_sa);
  cudaFree(g->d_nm);
  cudaFree(g->d_const2);
  cudaFree(g->d_mean_f);
  tc_free(g);
  cudaFree(g->d_g);
  cudaFree(g->d_s);
  cudaFree(g->d_gf);
  cudaFree(g->d_rsf);
  delete g;
}

}  // namespace lr

using namespace lr;

extern "C" {

lr_gmm *lr_gmm_create(int C, int D, const double *w, const double *mean, const double *cov) {
  if (!ensure_ready()) return nullptr;
  if (C < 1 || D < 1 || !w || !mean || !cov) {
    fail(LR_ERR_ARG, "lr_gmm_create: bad arguments (C=%d D=%d)", C, D);
    return nullptr;
  }
  if (D > kMaxD) {
    fail(LR_ERR_ARG, "lr_gmm_create: vectSize %d > %d is not supported by this build", D, kMaxD);
    return nullptr;
  }
  lr_gmm *g = new lr_gmm();
  g->C = C;
  g->D = D;
  g->Cp = ((C + 127) / 128) * 128;
  size_t cd = (size_t)C * D;
  bool ok = cudaMalloc(&g->d_w, C * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&g->d_mean, cd * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&g->d_cov, cd * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&g->d_covinv, cd * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&g->d_cst, C * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&g->d_det, C * sizeof(double)) == cudaSuccess &&
            cudaMalloc(&g->d_tot, sizeof(double)) == cudaSuccess &&
            cudaMalloc(&g->d_sa, (size_t)D * g->Cp * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&g->d_nm, (size_t)D * g->Cp * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&g->d_const2, g->Cp * sizeof(float)) == cudaSuccess &&
            cudaMalloc(&g->d_mean_f, (size_t)g->Cp * 64 * sizeof(float)) == cudaSuccess;
  if (!ok) {
    fail(LR_ERR_CUDA, "lr_gmm_create: cudaMalloc failed: %s",
         cudaGetErrorString(cudaGetLastError()));
    gmm_free(g);
    return nullptr;
  }
  if (lr_gmm_set(g, w, mean, cov) != LR_OK) {
    gmm_free(g);
    return nullptr;
  }
  return g;
}

lr_status lr_gmm_set(lr_gmm *g, const double *w, const double *mean, const double *cov) {
  LR_READY();
  LR_REQUIRE(g && w && mean && cov, "lr_gmm_set: null argument");
  Engine &e = engine();
  size_t cd = (size_t)g->C * g->D;
  LR_CUDA(cudaMemcpyAsync(g->d_w, w, g->C * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(g->d_mean, mean, cd * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(g->d_cov, cov, cd * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  g->cst_override = false;
  lr_status st = gmm_derive(g);
  if (st != LR_OK) return st;
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_gmm_set_cst(lr_gmm *g, const double *cst) {
  LR_READY();
  LR_REQUIRE(g && cst, "lr_gmm_set_cst: null argument");
  Engine &e = engine();
  LR_CUDA(cudaMemcpyAsync(g->d_cst, cst, g->C * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  g->cst_override = true;
  lr_status st = gmm_derive(g);
  if (st != LR_OK) return st;
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

lr_status lr_gmm_get(lr_gmm *g, double *w, double *mean, double *cov, double *covinv, double *cst,
                     double *det) {
  LR_READY();
  LR_REQUIRE(g, "lr_gmm_get: null model");
  Engine &e = engine();
  size_t cd = (size_t)g->C * g->D;
  auto get = [&](double *dst, const double *src, size_t n) -> cudaError_t {
    if (!dst) return cudaSuccess;
    return cudaMemcpyAsync(dst, src, n * sizeof(double), cudaMemcpyDeviceToHost, e.stream);
  };
  LR_CUDA(get(w, g->d_w, g->C));
  LR_CUDA(get(mean, g->d_mean, cd));
  LR_CUDA(get(cov, g->d_cov, cd));
  LR_CUDA(get(covinv, g->d_covinv, cd));
  LR_CUDA(get(cst, g->d_cst, g->C));
  LR_CUDA(get(det, g->d_det, g->C));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  return LR_OK;
}

void lr_gmm_destroy(lr_gmm *g) { gmm_free(g); }

size_t lr_gmm_em_stats_len(const lr_gmm *g) {
  return g ? (size_t)g->C + 2 * (size_t)g->C * g->D + 2 : 0;
}

lr_status lr_gmm_em_update_dev(lr_gmm *g, const double *d_stats, double flooring, double ceiling,
                               const double *d_cov_signal) {
  LR_READY();
  LR_REQUIRE(g && d_stats, "lr_gmm_em_update_dev: null argument");
  Engine &e = engine();
  size_t cd = (size_t)g->C * g->D;
  const double *occ = d_stats, *m1 = d_stats + g->C, *m2 = d_stats + g->C + cd;
  k_em_total<<<1, 256, 0, e.stream>>>(g->C, occ, g->d_tot);
  LR_CHECK_LAUNCH();
  k_em_update<<<ceil_div((long)cd, 256), 256, 0, e.stream>>>(g->C, g->D, occ, m1, m2, g->d_tot,
                                                             flooring, ceiling, d_cov_signal,
                                                             g->d_w, g->d_mean, g->d_cov);
  LR_CHECK_LAUNCH();
  g->cst_override = false;
  return gmm_derive(g);
}

lr_status lr_gmm_em_update(lr_gmm *g, const double *occ, const double *m1, const double *m2,
                           double flooring, double ceiling, const double *cov_signal) {
  LR_READY();
  LR_REQUIRE(g && occ && m1 && m2, "lr_gmm_em_update: null argument");
  Engine &e = engine();
  size_t cd = (size_t)g->C * g->D, n = lr_gmm_em_stats_len(g);
  DevBuf<double> st, sig;
  LR_CUDA(st.alloc(n));
  LR_CUDA(cudaMemsetAsync(st.p, 0, n * sizeof(double), e.stream));
  LR_CUDA(cudaMemcpyAsync(st.p, occ, g->C * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(st.p + g->C, m1, cd * sizeof(double), cudaMemcpyHostToDevice, e.stream));
  LR_CUDA(cudaMemcpyAsync(st.p + g->C + cd, m2, cd * sizeof(double), cudaMemcpyHostToDevice,
                          e.stream));
  if (cov_signal) {
    LR_CUDA(sig.alloc(g->D));
    LR_CUDA(cudaMemcpyAsync(sig.p, cov_signal, g->D * sizeof(double), cudaMemcpyHostToDevice,
                            e.stream));
  }
  lr_status rc = lr_gmm_em_update_dev(g, st.p, flooring, ceiling, cov_signal ? sig.p : nullptr);
  cudaStreamSynchronize(e.stream);
  return rc;
}

lr_status lr_frames_mean_cov(const float *X, size_t T, size_t ldx, int D, double *mean,
                             double *cov) {
  LR_READY();
  LR_REQUIRE(X && mean && cov && D >= 1 && D <= 64 && T > 0 && ldx >= (size_t)D,
             "lr_frames_mean_cov: bad arguments");
  Engine &e = engine();
  DevBuf<float> dx;
  DevBuf<double> acc;
  const size_t blk = (size_t)1 << 20;  // frames per staged block
  LR_CUDA(dx.alloc(std::min(T, blk) * ldx));
  LR_CUDA(acc.alloc(2 * D));
  LR_CUDA(cudaMemsetAsync(acc.p, 0, 2 * D * sizeof(double), e.stream));
  for (size_t t0 = 0; t0 < T; t0 += blk) {
    size_t n = std::min(blk, T - t0);
    LR_CUDA(cudaMemcpyAsync(dx.p, X + t0 * ldx, n * ldx * sizeof(float), cudaMemcpyHostToDevice,
                            e.stream));
    int grid = std::max(1, std::min((int)((n + 3) / 4), engine().sm_count * 8));
    k_mean_cov_acc<<<grid, dim3(64, 4), 0, e.stream>>>(dx.p, (long)n, ldx, D, acc.p);
    LR_CHECK_LAUNCH();
  }
  std::vector<double> h(2 * D);
  LR_CUDA(cudaMemcpyAsync(h.data(), acc.p, 2 * D * sizeof(double), cudaMemcpyDeviceToHost,
                          e.stream));
  LR_CUDA(cudaStreamSynchronize(e.stream));
  for (int i = 0; i < D; i++) {
    mean[i] = h[i] / (double)T;
    cov[i] = h[D + i] / (double)T - mean[i] * mean[i];
  }
  return LR_OK;
}

lr_feats *lr_feats_upload(const float *X, size_t T, size_t ldx, int D) {
  if (!ensure_ready()) return nullptr;
  if (!X || T == 0 || D < 1 || ldx < (size_t)D) {
    fail(LR_ERR_ARG, "lr_feats_upload: bad arguments");
    return nullptr;
  }
  float *d = nullptr;
  if (cudaMalloc(&d, T * ldx * sizeof(float)) != cudaSuccess) {
    fail(LR_ERR_CUDA, "lr_feats_upload: cudaMalloc of %zu bytes failed", T * ldx * sizeof(float));
    cudaGetLastError();
    return nullptr;
  }
  Engine &e = engine();
  if (cudaMemcpyAsync(d, X, T * ldx * sizeof(float), cudaMemcpyHostToDevice, e.stream) !=
          cudaSuccess ||
      cudaStreamSynchronize(e.stream) != cudaSuccess) {
    fail(LR_ERR_CUDA, "lr_feats_upload: copy failed");
    cudaFree(d);
    return nullptr;
  }
  lr_feats *f = new lr_feats();
  f->d_x = d;
  f->T = T;
  f->ldx = ldx;
  f->D = D;
  f->owned = true;
  return f;
}

lr_feats *lr_feats_wrap_device(const float *dX, size_t T, size_t ldx, int D) {
  if (!ensure_ready()) return nullptr;
  if (!dX || T == 0 || D < 1 || ldx < (size_t)D) {
    fail(LR_ERR_ARG, "lr_feats_wrap_device: bad arguments");
    return nullptr;
  }
  lr_feats *f = new lr_feats();
  f->d_x = dX;
  f->T = T;
  f->ldx = ldx;
  f->D = D;
  f->owned = false;
  return f;
}

void lr_feats_destroy(lr_feats *f) {
  if (!f) return;
  if (f->owned) cudaFree((void *)f->d_x);
  cudaFree(f->d_conv);
  delete f;
}

lr_status lr_feats_invalidate(lr_feats *f) {
  LR_REQUIRE(f, "lr_feats_invalidate: null handle");
  f->conv_norm = 0;  // the frames changed under the handle: the next EM pass converts them again
  return LR_OK;
}

}  // extern "C"
