/*
 * lr_oracle.c -- fp64 CPU restatement of the LIA_RAL hot path (see lr_oracle.h).
 * TEST INFRASTRUCTURE ONLY: the checker and the timed CPU baseline, never the product.
 *
 * Built twice by oracle/Makefile:
 *   _build/liblr_oracle.so       -O2, no fast-math      -> the parity oracle
 *   _build/liblr_oracle_fast.so  -O3 -ffast-math (configure.ac:23) + pthreads
 *                                (configure.ac:38-47)   -> the timed "--enable-MT" baseline
 */
#include "lr_oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

/* ------------------------------------------------------------------ helpers */
typedef void (*range_fn)(size_t begin, size_t end, int tid, void *arg);
typedef struct {
  range_fn fn;
  size_t begin, end;
  int tid;
  void *arg;
} range_job;

static void *range_tramp(void *p) {
  range_job *j = (range_job *)p;
  j->fn(j->begin, j->end, j->tid, j->arg);
  return NULL;
}

/* Contiguous ranges per thread: the split the reference uses for NDX lines
 * (AccumulateTVStat.cpp:498-507), speakers (:1989-2023) and components (:881-907). */
static void parallel_ranges(size_t n, int threads, range_fn fn, void *arg) {
  if (threads <= 1 || n < 2) {
    fn(0, n, 0, arg);
    return;
  }
  if ((size_t)threads > n) threads = (int)n;
  pthread_t *th = (pthread_t *)malloc(sizeof(pthread_t) * threads);
  range_job *jobs = (range_job *)malloc(sizeof(range_job) * threads);
  size_t per = n / threads, rem = n % threads, pos = 0;
  for (int t = 0; t < threads; t++) {
    size_t len = per + ((size_t)t < rem ? 1 : 0);
    jobs[t].fn = fn;
    jobs[t].begin = pos;
    jobs[t].end = pos + len;
    jobs[t].tid = t;
    jobs[t].arg = arg;
    pos += len;
    pthread_create(&th[t], NULL, range_tramp, &jobs[t]);
  }
  for (int t = 0; t < threads; t++) pthread_join(th[t], NULL);
  free(th);
  free(jobs);
}

/* ------------------------------------------------------------------ A.1 */
void orc_gmm_compute_all(int C, int D, const double *cov, double *covinv, double *det,
                         double *cst) {
  for (int c = 0; c < C; c++) {
    double dt = 1.0;
    for (int i = 0; i < D; i++) {
      double v = cov[(size_t)c * D + i];
      dt *= v;
      covinv[(size_t)c * D + i] = 1.0 / v;
    }
    det[c] = dt;
    cst[c] = 1.0 / (pow(2.0 * M_PI, 0.5 * D) * sqrt(dt));
  }
}

/* ------------------------------------------------------------------ A.2 */
double orc_distrib_lk(int D, const double *x, const double *mean, const double *covinv,
                      double cst) {
  double q = 0.0;
  for (int i = 0; i < D; i++) {
    double d = x[i] - mean[i];
    q += d * d * covinv[i];
  }
  double lk = cst * exp(-0.5 * q);
  if (isnan(lk)) lk = ORC_EPS_LK;
  return lk;
}

/* ------------------------------------------------------------------ A.3 */
static double frame_lk_d(int C, int D, const double *w, const double *mean, const double *covinv,
                         const double *cst, const double *xd, double *p) {
  double s = 0.0;
  for (int c = 0; c < C; c++) {
    double v = w[c] * orc_distrib_lk(D, xd, mean + (size_t)c * D, covinv + (size_t)c * D, cst[c]);
    p[c] = v;
    s += v;
  }
  return s;
}

double orc_frame_likelihoods(int C, int D, const double *w, const double *mean,
                             const double *covinv, const double *cst, const float *x,
                             double *p) {
  double xd[1024];
  for (int i = 0; i < D; i++) xd[i] = (double)x[i];
  return frame_lk_d(C, D, w, mean, covinv, cst, xd, p);
}

/* ------------------------------------------------------------------ A.4 */
typedef struct {
  int C, D;
  const double *w, *mean, *covinv, *cst;
  const float *X;
  size_t T, ldx, U;
  const int32_t *frame2row;
  double *N, *F;
} bw_args;

/* rows [begin,end): every thread scans all frames and keeps those of its own rows, so
 * rows of N/F are disjoint between threads exactly like StatTVthread (:376-475). */
static void bw_range(size_t begin, size_t end, int tid, void *argp) {
  (void)tid;
  bw_args *a = (bw_args *)argp;
  int C = a->C, D = a->D;
  double *p = (double *)malloc(sizeof(double) * C);
  double xd[1024];
  for (size_t t = 0; t < a->T; t++) {
    int32_t row = a->frame2row ? a->frame2row[t] : 0;
    if (row < 0 || (size_t)row < begin || (size_t)row >= end) continue;
    const float *x = a->X + t * a->ldx;
    for (int i = 0; i < D; i++) xd[i] = (double)x[i];
    double s = frame_lk_d(C, D, a->w, a->mean, a->covinv, a->cst, xd, p);
    double *n = a->N + (size_t)row * C;
    double *f = a->F + (size_t)row * C * D;
    for (int k = 0; k < C; k++) {
      double g = p[k] / s; /* getOccVect: normalised occupation */
      n[k] += g;
      for (int i = 0; i < D; i++) f[(size_t)k * D + i] += g * xd[i];
    }
  }
  free(p);
}

void orc_bwstats(int C, int D, const double *w, const double *mean, const double *covinv,
                 const double *cst, const float *X, size_t T, size_t ldx,
                 const int32_t *frame2row, size_t U, double *N, double *F, int threads) {
  bw_args a = {C, D, w, mean, covinv, cst, X, T, ldx, U, frame2row, N, F};
  parallel_ranges(U, threads, bw_range, &a);
}

/* JFAAcc::normalizeFeatures (AccumulateJFAStat.cpp:4623-4680): for every frame of every selected segment, in
 * order, Prob[k] = w_k lk_k(f) / sum under the session model, then ff[i] -= Prob[k] ux[k*D + i] (k outer, i inner)
 * and the frame is written back -- here into the float32 buffer the engine's FeatureServer keeps. */
void orc_jfa_normalize_features(int C, int D, const double *w, const double *mean, const double *covinv,
                                const double *cst, const double *ux, float *X, size_t ldx,
                                const int64_t *seg_begin, const int64_t *seg_len, size_t n_segs) {
  double *p = (double *)malloc(sizeof(double) * C);
  double xd[1024];
  for (size_t s = 0; s < n_segs; s++)
    for (int64_t t = seg_begin[s]; t < seg_begin[s] + seg_len[s]; t++) {
      float *x = X + (size_t)t * ldx;
      for (int i = 0; i < D; i++) xd[i] = (double)x[i];
      double sum = frame_lk_d(C, D, w, mean, covinv, cst, xd, p);
      for (int k = 0; k < C; k++) {
        double g = p[k] / sum;
        for (int i = 0; i < D; i++) xd[i] -= g * ux[(size_t)k * D + i];
      }
      for (int i = 0; i < D; i++) x[i] = (float)xd[i];
    }
  free(p);
}

/* ------------------------------------------------------------------ A.5 */
typedef struct {
  int C, D;
  const double *w, *mean, *covinv, *cst;
  const float *X;
  size_t ldx;
  double fw;
  double *occ, *m1, *m2; /* per-thread accumulators: [threads][...] */
  double *llk;           /* [threads] */
  size_t acc_stride;
} em_args;

static void em_range(size_t begin, size_t end, int tid, void *argp) {
  em_args *a = (em_args *)argp;
  int C = a->C, D = a->D;
  double *occ = a->occ + (size_t)tid * C;
  double *m1 = a->m1 + (size_t)tid * a->acc_stride;
  double *m2 = a->m2 + (size_t)tid * a->acc_stride;
  double *p = (double *)malloc(sizeof(double) * C);
  double xd[1024], x2[1024];
  double llk = 0.0;
  for (size_t t = begin; t < end; t++) {
    const float *x = a->X + t * a->ldx;
    for (int i = 0; i < D; i++) {
      xd[i] = (double)x[i];
      x2[i] = xd[i] * xd[i];
    }
    double s = frame_lk_d(C, D, a->w, a->mean, a->covinv, a->cst, xd, p);
    llk += log(s);
    for (int k = 0; k < C; k++) {
      double g = p[k] / s * a->fw;
      occ[k] += g;
      double *a1 = m1 + (size_t)k * D, *a2 = m2 + (size_t)k * D;
      for (int i = 0; i < D; i++) {
        a1[i] += g * xd[i];
        a2[i] += g * x2[i];
      }
    }
  }
  a->llk[tid] = llk;
  free(p);
}

double orc_em_accumulate(int C, int D, const double *w, const double *mean,
                         const double *covinv, const double *cst, const float *X, size_t T,
                         size_t ldx, double frame_weight, double *occ, double *m1, double *m2,
                         double *nframes, int threads) {
  if (threads < 1) threads = 1;
  size_t cd = (size_t)C * D;
  /* per-thread MixtureStat merged at the end with addAccEM (AccumulateStat.cpp:286-292) */
  double *tocc = (double *)calloc((size_t)threads * C, sizeof(double));
  double *tm1 = (double *)calloc((size_t)threads * cd, sizeof(double));
  double *tm2 = (double *)calloc((size_t)threads * cd, sizeof(double));
  double *tllk = (double *)calloc(threads, sizeof(double));
  em_args a = {C, D, w, mean, covinv, cst, X, ldx, frame_weight, tocc, tm1, tm2, tllk, cd};
  parallel_ranges(T, threads, em_range, &a);
  double llk = 0.0;
  for (int t = 0; t < threads; t++) {
    llk += tllk[t];
    for (int k = 0; k < C; k++) occ[k] += tocc[(size_t)t * C + k];
    for (size_t i = 0; i < cd; i++) {
      m1[i] += tm1[(size_t)t * cd + i];
      m2[i] += tm2[(size_t)t * cd + i];
    }
  }
  if (nframes) *nframes += (double)T * frame_weight;
  free(tocc);
  free(tm1);
  free(tm2);
  free(tllk);
  return llk;
}

void orc_em_get(int C, int D, const double *occ, const double *m1, const double *m2, double *w,
                double *mean, double *cov) {
  double tot = 0.0;
  for (int c = 0; c < C; c++) tot += occ[c];
  for (int c = 0; c < C; c++) {
    w[c] = occ[c] / tot;
    if (occ[c] > 0.0) {
      for (int i = 0; i < D; i++) {
        double mu = m1[(size_t)c * D + i] / occ[c];
        mean[(size_t)c * D + i] = mu;
        cov[(size_t)c * D + i] = m2[(size_t)c * D + i] / occ[c] - mu * mu;
      }
    }
  }
}

void orc_variance_control(int C, int D, double *cov, double flooring, double ceiling,
                          const double *cov_signal, long *n_floor, long *n_ceil) {
  long nf = 0, nc = 0;
  for (int c = 0; c < C; c++)
    for (int v = 0; v < D; v++) {
      double x = cov[(size_t)c * D + v];
      if (x <= flooring * cov_signal[v]) {
        x = flooring * cov_signal[v];
        nf++;
      }
      if (x >= ceiling * cov_signal[v]) {
        x = ceiling * cov_signal[v];
        nc++;
      }
      cov[(size_t)c * D + v] = x;
    }
  if (n_floor) *n_floor = nf;
  if (n_ceil) *n_ceil = nc;
}

double orc_set_it_parameter(double begin, double end, int nb_it, int it) {
  if (nb_it < 2) return begin;
  double step = (begin - end) / ((double)nb_it - 1.0);
  return begin - step * it;
}

void orc_mean_cov(int D, const float *X, size_t T, size_t ldx, double *mean, double *cov) {
  /* FrameAccGD: sum x, sum x^2, count; mean = sx/n ; cov = sxx/n - mean^2 */
  for (int i = 0; i < D; i++) mean[i] = cov[i] = 0.0;
  for (size_t t = 0; t < T; t++)
    for (int i = 0; i < D; i++) {
      double v = (double)X[t * ldx + i];
      mean[i] += v;
      cov[i] += v * v;
    }
  for (int i = 0; i < D; i++) {
    mean[i] /= (double)T;
    cov[i] = cov[i] / (double)T - mean[i] * mean[i];
  }
}

/* ------------------------------------------------------------------ A.6 */
static double clamp_llk(double lk, double min_llk, double max_llk) {
  /* restated in TopGauss.cpp:255-260 */
  double l = log(lk);
  if (isnan(l)) return min_llk;
  if (l <= min_llk) return min_llk;
  if (l >= max_llk) return max_llk;
  return l;
}

void orc_llk_determine_top(int C, int D, const double *w, const double *mean,
                           const double *covinv, const double *cst, const float *X, size_t T,
                           size_t ldx, int K, int complete, double min_llk, double max_llk,
                           double *llk, uint32_t *idx, double *top_lk, double *rest_lk,
                           double *rest_w) {
  double *p = (double *)malloc(sizeof(double) * C);
  char *used = (char *)malloc(C);
  if (K > C) K = C;
  for (size_t t = 0; t < T; t++) {
    double s = orc_frame_likelihoods(C, D, w, mean, covinv, cst, X + t * ldx, p);
    memset(used, 0, C);
    double top_sum = 0.0, top_w = 0.0;
    /* descending sort, ties -> lowest index (selection keeps it O(C*K)) */
    for (int k = 0; k < K; k++) {
      int best = -1;
      for (int c = 0; c < C; c++)
        if (!used[c] && (best < 0 || p[c] > p[best])) best = c;
      used[best] = 1;
      idx[t * K + k] = (uint32_t)best;
      if (top_lk) top_lk[t * K + k] = p[best];
      top_sum += p[best];
      top_w += w[best];
    }
    double rest = 0.0;
    for (int c = 0; c < C; c++)
      if (!used[c]) rest += p[c];
    if (rest_lk) rest_lk[t] = rest;
    if (rest_w) rest_w[t] = 1.0 - top_w;
    (void)s;
    double lk = complete ? top_sum + rest : top_sum;
    if (llk) llk[t] = clamp_llk(lk, min_llk, max_llk);
  }
  free(p);
  free(used);
}

void orc_llk_use_top(int C, int D, const double *w, const double *mean, const double *covinv,
                     const double *cst, const float *X, size_t T, size_t ldx, int K,
                     const uint32_t *idx, const double *rest_lk, int complete, double min_llk,
                     double max_llk, double *llk) {
  (void)C;
  double xd[1024];
  for (size_t t = 0; t < T; t++) {
    const float *x = X + t * ldx;
    for (int i = 0; i < D; i++) xd[i] = (double)x[i];
    double lk = 0.0;
    for (int k = 0; k < K; k++) {
      uint32_t c = idx[t * K + k];
      lk += w[c] * orc_distrib_lk(D, xd, mean + (size_t)c * D, covinv + (size_t)c * D, cst[c]);
    }
    if (complete && rest_lk) lk += rest_lk[t];
    llk[t] = clamp_llk(lk, min_llk, max_llk);
  }
}

void orc_llk_all(int C, int D, const double *w, const double *mean, const double *covinv,
                 const double *cst, const float *X, size_t T, size_t ldx, double min_llk,
                 double max_llk, double *llk) {
  double *p = (double *)malloc(sizeof(double) * C);
  for (size_t t = 0; t < T; t++) {
    double s = orc_frame_likelihoods(C, D, w, mean, covinv, cst, X + t * ldx, p);
    llk[t] = clamp_llk(s, min_llk, max_llk);
  }
  free(p);
}

/* ------------------------------------------------------------------ dense helpers */
int orc_invert(int n, const double *a, double *inv) {
  /* exact dense inverse (Gauss-Jordan, partial pivoting); DoubleSquareMatrix::invert [ALIZE] */
  double *m = (double *)malloc(sizeof(double) * n * n);
  memcpy(m, a, sizeof(double) * n * n);
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) inv[(size_t)i * n + j] = (i == j) ? 1.0 : 0.0;
  for (int col = 0; col < n; col++) {
    int piv = col;
    double best = fabs(m[(size_t)col * n + col]);
    for (int r = col + 1; r < n; r++) {
      double v = fabs(m[(size_t)r * n + col]);
      if (v > best) {
        best = v;
        piv = r;
      }
    }
    if (best == 0.0) {
      free(m);
      return -1;
    }
    if (piv != col)
      for (int j = 0; j < n; j++) {
        double t = m[(size_t)col * n + j];
        m[(size_t)col * n + j] = m[(size_t)piv * n + j];
        m[(size_t)piv * n + j] = t;
        t = inv[(size_t)col * n + j];
        inv[(size_t)col * n + j] = inv[(size_t)piv * n + j];
        inv[(size_t)piv * n + j] = t;
      }
    double d = 1.0 / m[(size_t)col * n + col];
    for (int j = 0; j < n; j++) {
      m[(size_t)col * n + j] *= d;
      inv[(size_t)col * n + j] *= d;
    }
    for (int r = 0; r < n; r++) {
      if (r == col) continue;
      double f = m[(size_t)r * n + col];
      if (f == 0.0) continue;
      double *mr = m + (size_t)r * n, *mc = m + (size_t)col * n;
      double *ir = inv + (size_t)r * n, *ic = inv + (size_t)col * n;
      for (int j = 0; j < n; j++) {
        mr[j] -= f * mc[j];
        ir[j] -= f * ic[j];
      }
    }
  }
  free(m);
  return 0;
}

int orc_upper_cholesky(int n, const double *a, double *u) {
  /* a = u^T u with u upper triangular; DoubleSquareMatrix::upperCholesky [ALIZE] */
  memset(u, 0, sizeof(double) * n * n);
  for (int i = 0; i < n; i++) {
    for (int j = i; j < n; j++) {
      double s = a[(size_t)i * n + j];
      for (int k = 0; k < i; k++) s -= u[(size_t)k * n + i] * u[(size_t)k * n + j];
      if (i == j) {
        if (s <= 0.0) return -1;
        u[(size_t)i * n + i] = sqrt(s);
      } else {
        u[(size_t)i * n + j] = s / u[(size_t)i * n + i];
      }
    }
  }
  return 0;
}

/* ------------------------------------------------------------------ A.7 */
void orc_tv_subtract_m(size_t U, int C, int D, const double *N, const double *ubm_mean,
                       double *F) {
  size_t sv = (size_t)C * D;
  for (size_t s = 0; s < U; s++)
    for (int i = 0; i < C; i++)
      for (int j = 0; j < D; j++)
        F[s * sv + (size_t)i * D + j] -= ubm_mean[(size_t)i * D + j] * N[s * C + i];
}

typedef struct {
  int C, D, R;
  const double *T, *invvar;
  double *tett;
} tett_args;

static void tett_range(size_t begin, size_t end, int tid, void *argp) {
  (void)tid;
  tett_args *a = (tett_args *)argp;
  int D = a->D, R = a->R;
  size_t sv = (size_t)a->C * D;
  for (size_t d = begin; d < end; d++) {
    double *o = a->tett + d * (size_t)R * R;
    for (int i = 0; i < R; i++)
      for (int j = 0; j <= i; j++) {
        double s = 0.0;
        const double *ti = a->T + (size_t)i * sv + d * D;
        const double *tj = a->T + (size_t)j * sv + d * D;
        const double *e = a->invvar + d * D;
        for (int k = 0; k < D; k++) s += ti[k] * e[k] * tj[k];
        o[(size_t)i * R + j] = s;
      }
    for (int i = 0; i < R; i++)
      for (int j = i + 1; j < R; j++) o[(size_t)i * R + j] = o[(size_t)j * R + i];
  }
}

void orc_tv_tett(int C, int D, int R, const double *T, const double *invvar, double *tett,
                 int threads) {
  tett_args a = {C, D, R, T, invvar, tett};
  parallel_ranges((size_t)C, threads, tett_range, &a);
}

typedef struct {
  size_t U;
  int C, D, R, threads;
  const double *N, *F, *T, *invvar, *tett;
  double *W;
  /* E-step extras (NULL for plain i-vector extraction); per-thread copies */
  double *A, *Cmx, *Rm, *r, *meanW;
} iv_args;

/* Posterior of one utterance: L = I + sum_c N_c TETt_c, Linv = L^-1, aux, y = Linv aux */
static void iv_one(const iv_args *a, size_t spk, double *L, double *Linv, double *aux,
                   double *y) {
  int C = a->C, D = a->D, R = a->R;
  size_t sv = (size_t)C * D;
  memset(L, 0, sizeof(double) * R * R);
  for (int i = 0; i < R; i++) L[(size_t)i * R + i] = 1.0;
  for (int dis = 0; dis < C; dis++) {
    const double *te = a->tett + (size_t)dis * R * R;
    double n = a->N[spk * C + dis];
    for (int i = 0; i < R; i++)
      for (int j = 0; j <= i; j++) L[(size_t)i * R + j] += te[(size_t)i * R + j] * n;
  }
  for (int i = 0; i < R; i++)
    for (int j = i + 1; j < R; j++) L[(size_t)i * R + j] = L[(size_t)j * R + i];
  orc_invert(R, L, Linv);
  const double *f = a->F + spk * sv;
  for (int i = 0; i < R; i++) {
    double s = 0.0;
    const double *ti = a->T + (size_t)i * sv;
    for (size_t k = 0; k < sv; k++) s += f[k] * a->invvar[k] * ti[k];
    aux[i] = s;
  }
  for (int i = 0; i < R; i++) {
    double s = 0.0;
    for (int k = 0; k < R; k++) s += aux[k] * Linv[(size_t)i * R + k];
    y[i] = s;
  }
}

static void iv_range(size_t begin, size_t end, int tid, void *argp) {
  iv_args *a = (iv_args *)argp;
  int C = a->C, D = a->D, R = a->R;
  size_t sv = (size_t)C * D, rr = (size_t)R * R;
  double *L = (double *)malloc(sizeof(double) * rr);
  double *Linv = (double *)malloc(sizeof(double) * rr);
  double *aux = (double *)malloc(sizeof(double) * R);
  double *A = a->A ? a->A + (size_t)tid * C * rr : NULL;
  double *Cmx = a->A ? a->Cmx + (size_t)tid * R * sv : NULL;
  double *Rm = a->A ? a->Rm + (size_t)tid * rr : NULL;
  double *r = a->A ? a->r + (size_t)tid * R : NULL;
  double *meanW = a->A ? a->meanW + (size_t)tid * R : NULL;
  for (size_t spk = begin; spk < end; spk++) {
    double *y = a->W + spk * R;
    iv_one(a, spk, L, Linv, aux, y);
    if (!A) continue;
    /* E-step accumulators, AccumulateTVStat.cpp:1762-1788 */
    for (int k = 0; k < R; k++) meanW[k] += y[k];
    for (int i = 0; i < R; i++) {
      for (int j = 0; j < R; j++) {
        Linv[(size_t)i * R + j] += y[i] * y[j];
        Rm[(size_t)i * R + j] += Linv[(size_t)i * R + j];
      }
      r[i] += y[i];
    }
    for (int dis = 0; dis < C; dis++) {
      double n = a->N[spk * C + dis];
      double *Ad = A + (size_t)dis * rr;
      for (size_t e = 0; e < rr; e++) Ad[e] += Linv[e] * n;
    }
    const double *f = a->F + spk * sv;
    for (int i = 0; i < R; i++) {
      double yi = y[i];
      double *ci = Cmx + (size_t)i * sv;
      for (size_t j = 0; j < sv; j++) ci[j] += yi * f[j];
    }
  }
  free(L);
  free(Linv);
  free(aux);
}

void orc_tv_ivectors(size_t U, int C, int D, int R, const double *N, const double *F,
                     const double *T, const double *invvar, const double *tett, double *W,
                     int threads) {
  iv_args a = {U, C, D, R, threads, N, F, T, invvar, tett, W, NULL, NULL, NULL, NULL, NULL};
  memset(W, 0, sizeof(double) * U * R);
  parallel_ranges(U, threads, iv_range, &a);
}

void orc_tv_estep(size_t U, int C, int D, int R, const double *N, const double *F,
                  const double *T, const double *invvar, const double *tett, double *W,
                  double *A, double *Cmx, double *Rm, double *r, double *meanW, int threads) {
  if (threads < 1) threads = 1;
  if ((size_t)threads > U) threads = (int)(U ? U : 1);
  size_t sv = (size_t)C * D, rr = (size_t)R * R;
  /* the reference serialises the A / C updates behind two mutexes (:1920-1937); per-thread
   * partials summed at the end are arithmetically the same reduction */
  double *tA = (double *)calloc((size_t)threads * C * rr, sizeof(double));
  double *tC = (double *)calloc((size_t)threads * R * sv, sizeof(double));
  double *tR = (double *)calloc((size_t)threads * rr, sizeof(double));
  double *tr = (double *)calloc((size_t)threads * R, sizeof(double));
  double *tm = (double *)calloc((size_t)threads * R, sizeof(double));
  iv_args a = {U, C, D, R, threads, N, F, T, invvar, tett, W, tA, tC, tR, tr, tm};
  memset(W, 0, sizeof(double) * U * R);
  memset(A, 0, sizeof(double) * C * rr); /* _A.setAllValues(0.0) :1719 */
  memset(Rm, 0, sizeof(double) * rr);
  memset(r, 0, sizeof(double) * R);
  memset(meanW, 0, sizeof(double) * R);
  parallel_ranges(U, threads, iv_range, &a);
  for (int t = 0; t < threads; t++) {
    for (size_t e = 0; e < (size_t)C * rr; e++) A[e] += tA[(size_t)t * C * rr + e];
    for (size_t e = 0; e < (size_t)R * sv; e++) Cmx[e] += tC[(size_t)t * R * sv + e];
    for (size_t e = 0; e < rr; e++) Rm[e] += tR[(size_t)t * rr + e];
    for (int e = 0; e < R; e++) {
      r[e] += tr[(size_t)t * R + e];
      meanW[e] += tm[(size_t)t * R + e];
    }
  }
  for (int k = 0; k < R; k++) meanW[k] /= (double)U; /* :1792-1794, _n_speakers */
  free(tA);
  free(tC);
  free(tR);
  free(tr);
  free(tm);
}

void orc_tv_mstep(int C, int D, int R, const double *A, const double *Cmx, double *T) {
  size_t sv = (size_t)C * D, rr = (size_t)R * R;
  double *invA = (double *)malloc(sizeof(double) * rr);
  for (int d = 0; d < C; d++) {
    orc_invert(R, A + (size_t)d * rr, invA);
    for (int i = 0; i < R; i++)
      for (int j = 0; j < D; j++) {
        double s = 0.0;
        for (int k = 0; k < R; k++)
          s += invA[(size_t)i * R + k] * Cmx[(size_t)k * sv + (size_t)d * D + j];
        T[(size_t)i * sv + (size_t)d * D + j] = s;
      }
  }
  free(invA);
}

int orc_tv_mindiv(int C, int D, int R, double n_sessions, double *Rm, double *r,
                  const double *meanW, double *ubm_mean, double *T) {
  size_t sv = (size_t)C * D;
  for (int i = 0; i < R; i++) r[i] /= n_sessions;
  for (int i = 0; i < R; i++)
    for (int j = 0; j < R; j++)
      Rm[(size_t)i * R + j] = Rm[(size_t)i * R + j] / n_sessions - r[i] * r[j];
  double *Ch = (double *)malloc(sizeof(double) * R * R);
  if (orc_upper_cholesky(R, Rm, Ch) != 0) {
    free(Ch);
    return -1;
  }
  for (size_t j = 0; j < sv; j++) {
    double s = ubm_mean[j];
    for (int k = 0; k < R; k++) s += meanW[k] * T[(size_t)k * sv + j];
    ubm_mean[j] = s;
  }
  double *tmp = (double *)calloc((size_t)R * sv, sizeof(double));
  for (int i = 0; i < R; i++)
    for (int k = 0; k < R; k++) {
      double c = Ch[(size_t)i * R + k];
      if (c == 0.0) continue;
      const double *tk = T + (size_t)k * sv;
      double *o = tmp + (size_t)i * sv;
      for (size_t j = 0; j < sv; j++) o[j] += c * tk[j];
    }
  memcpy(T, tmp, sizeof(double) * R * sv);
  free(tmp);
  free(Ch);
  return 0;
}

void orc_tv_orthonormalize(int R, size_t sv, double *T) {
  /* classical Gram-Schmidt over rows, projections taken against the ORIGINAL row */
  double *Q = (double *)calloc((size_t)R * sv, sizeof(double));
  double *v = (double *)malloc(sizeof(double) * sv);
  for (int j = 0; j < R; j++) {
    const double *tj = T + (size_t)j * sv;
    memcpy(v, tj, sizeof(double) * sv);
    for (int i = 0; i < j; i++) {
      const double *qi = Q + (size_t)i * sv;
      double rij = 0.0;
      for (size_t k = 0; k < sv; k++) rij += qi[k] * tj[k];
      for (size_t k = 0; k < sv; k++) v[k] -= rij * qi[k];
    }
    double nv = 0.0;
    for (size_t k = 0; k < sv; k++) nv += v[k] * v[k];
    nv = sqrt(nv);
    double *qj = Q + (size_t)j * sv;
    if (nv == 0.0)
      for (size_t k = 0; k < sv; k++) qj[k] = 0.0;
    else
      for (size_t k = 0; k < sv; k++) qj[k] = v[k] / nv;
  }
  memcpy(T, Q, sizeof(double) * R * sv);
  free(Q);
  free(v);
}

/* ------------------------------------------------------------------ A.9 */
static void matmul(int m, int k, int n, const double *a, const double *b, double *c) {
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) c[(size_t)i * n + j] = 0.0;
  for (int i = 0; i < m; i++)
    for (int l = 0; l < k; l++) {
      double v = a[(size_t)i * k + l];
      for (int j = 0; j < n; j++) c[(size_t)i * n + j] += v * b[(size_t)l * n + j];
    }
}

static double logdet_via_chol(int n, const double *k) {
  /* alpha = 2 * sum log diag(chol(K)) (PldaTools.cpp:4511-4516) */
  double *u = (double *)malloc(sizeof(double) * n * n);
  double s = 0.0;
  if (orc_upper_cholesky(n, k, u) == 0)
    for (int i = 0; i < n; i++) s += log(u[(size_t)i * n + i]);
  else
    s = NAN;
  free(u);
  return 2.0 * s;
}

static void k_of(int r, double n, const double *phi, double *K) {
  double *tmp = (double *)malloc(sizeof(double) * r * r);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < r; j++) tmp[(size_t)i * r + j] = n * phi[(size_t)i * r + j] + (i == j);
  orc_invert(r, tmp, K);
  free(tmp);
}

static double quad_form(int r, const double *K, const double *v) {
  double s = 0.0;
  for (int i = 0; i < r; i++) {
    double t = 0.0;
    for (int j = 0; j < r; j++) t += K[(size_t)i * r + j] * v[j];
    s += v[i] * t;
  }
  return s;
}

int orc_plda_native_scoring(int d, int rF, int rG, const double *F, const double *G,
                            const double *Sigma, const double *models, size_t n_enrol,
                            const int32_t *model_of, size_t n_models, const double *segments,
                            size_t n_test, double *scores) {
  /* preComputation (PldaTools.cpp:2950-2972) */
  double *iS = (double *)malloc(sizeof(double) * d * d);
  if (orc_invert(d, Sigma, iS) != 0) {
    free(iS);
    return -1;
  }
  double *Ft = (double *)malloc(sizeof(double) * rF * d); /* F^T */
  for (int i = 0; i < d; i++)
    for (int j = 0; j < rF; j++) Ft[(size_t)j * d + i] = F[(size_t)i * rF + j];
  double *Ftw = (double *)malloc(sizeof(double) * rF * d);
  matmul(rF, d, d, Ft, iS, Ftw);
  double *FTJ = (double *)malloc(sizeof(double) * rF * d);
  memcpy(FTJ, Ftw, sizeof(double) * rF * d);
  if (rG > 0) {
    double *Gt = (double *)malloc(sizeof(double) * rG * d);
    for (int i = 0; i < d; i++)
      for (int j = 0; j < rG; j++) Gt[(size_t)j * d + i] = G[(size_t)i * rG + j];
    double *Gtw = (double *)malloc(sizeof(double) * rG * d);
    matmul(rG, d, d, Gt, iS, Gtw);
    double *GtwG = (double *)malloc(sizeof(double) * rG * rG);
    matmul(rG, d, rG, Gtw, G, GtwG);
    for (int i = 0; i < rG; i++) GtwG[(size_t)i * rG + i] += 1.0;
    double *iGG = (double *)malloc(sizeof(double) * rG * rG);
    orc_invert(rG, GtwG, iGG);
    double *FtwG = (double *)malloc(sizeof(double) * rF * rG);
    matmul(rF, d, rG, Ftw, G, FtwG);
    double *t1 = (double *)malloc(sizeof(double) * rF * rG);
    matmul(rF, rG, rG, FtwG, iGG, t1);
    double *t2 = (double *)malloc(sizeof(double) * rF * d);
    matmul(rF, rG, d, t1, Gtw, t2);
    for (size_t e = 0; e < (size_t)rF * d; e++) FTJ[e] -= t2[e];
    free(Gt);
    free(Gtw);
    free(GtwG);
    free(iGG);
    free(FtwG);
    free(t1);
    free(t2);
  }
  double *phi = (double *)malloc(sizeof(double) * rF * rF); /* FTJF */
  matmul(rF, d, rF, FTJ, F, phi);
  /* rotateLeft (:3770-3790): project models and segments, rF x n */
  double *pm = (double *)malloc(sizeof(double) * rF * n_enrol);
  double *ps = (double *)malloc(sizeof(double) * rF * n_test);
  matmul(rF, d, (int)n_enrol, FTJ, models, pm);
  matmul(rF, d, (int)n_test, FTJ, segments, ps);
  double *K1 = (double *)malloc(sizeof(double) * rF * rF);
  k_of(rF, 1.0, phi, K1);
  double alpha1 = logdet_via_chol(rF, K1);
  /* pldaScoringUnThreaded (:4186-4271) */
  double *KL = (double *)malloc(sizeof(double) * rF * rF);
  double *KL1 = (double *)malloc(sizeof(double) * rF * rF);
  double *m = (double *)malloc(sizeof(double) * rF);
  double *v = (double *)malloc(sizeof(double) * rF);
  double *s1 = (double *)malloc(sizeof(double) * n_test);
  for (size_t i = 0; i < n_test; i++) {
    for (int k = 0; k < rF; k++) v[k] = ps[(size_t)k * n_test + i];
    s1[i] = quad_form(rF, K1, v);
  }
  size_t sess = 0;
  long cur_nb = 0;
  double constant = 0.0;
  for (size_t mod = 0; mod < n_models; mod++) {
    for (int k = 0; k < rF; k++) m[k] = 0.0;
    long nb = 0;
    int32_t id = model_of[sess];
    while (sess < n_enrol && model_of[sess] == id) {
      for (int k = 0; k < rF; k++) m[k] += pm[(size_t)k * n_enrol + sess];
      nb++;
      sess++;
    }
    if (nb != cur_nb) {
      cur_nb = nb;
      k_of(rF, (double)nb, phi, KL);
      k_of(rF, (double)nb + 1.0, phi, KL1);
      double aL = logdet_via_chol(rF, KL), aL1 = logdet_via_chol(rF, KL1);
      constant = (aL1 - aL - alpha1) / 2.0;
    }
    double s2 = quad_form(rF, KL, m);
    for (size_t i = 0; i < n_test; i++) {
      for (int k = 0; k < rF; k++) v[k] = ps[(size_t)k * n_test + i] + m[k];
      double s3 = quad_form(rF, KL1, v);
      scores[mod * n_test + i] = (s3 - s2 - s1[i]) / 2.0 + constant;
    }
  }
  free(iS);
  free(Ft);
  free(Ftw);
  free(FTJ);
  free(phi);
  free(pm);
  free(ps);
  free(K1);
  free(KL);
  free(KL1);
  free(m);
  free(v);
  free(s1);
  return 0;
}

/* ====================================================================== approximate i-vectors
 *  IvExtractor --mode ubmWeight | eigenDecomposition and the
 * TotalVariability approximationMode outputs. */

/* normTMatrix, AccumulateTVStat.cpp:1600-1609 */
void orc_tv_norm_t(int R, size_t sv, const double *invvar, double *T) {
  for (size_t i = 0; i < sv; i++) {
    double sq = sqrt(invvar[i]);
    for (int j = 0; j < R; j++) T[(size_t)j * sv + i] = T[(size_t)j * sv + i] * sq;
  }
}

/* normStatisticsUnThreaded, AccumulateTVStat.cpp:1225-1242 */
void orc_tv_norm_statistics(size_t U, int C, int D, const double *N, const double *ubm_mean,
                            const double *invvar, double *F) {
  size_t sv = (size_t)C * D;
  for (int i = 0; i < C; i++)
    for (int j = 0; j < D; j++) {
      double sq = sqrt(invvar[(size_t)i * D + j]);
      for (size_t spk = 0; spk < U; spk++) {
        F[spk * sv + (size_t)i * D + j] -= ubm_mean[(size_t)i * D + j] * N[spk * C + i];
        F[spk * sv + (size_t)i * D + j] *= sq;
      }
    }
}

/* getWeightedCovUnThreaded, AccumulateTVStat.cpp:2837-2855.  W[R x R] */
void orc_tv_weighted_cov(int C, int D, int R, const double *T, const double *weight, double *W) {
  size_t sv = (size_t)C * D;
  for (int i = 0; i < R * R; i++) W[i] = 0.0;
  for (int cc = 0; cc < C; cc++)
    for (int i = 0; i < R; i++)
      for (int j = 0; j < i + 1; j++) {
        const double *ti = T + (size_t)i * sv + (size_t)cc * D, *tj = T + (size_t)j * sv + (size_t)cc * D;
        for (int k = 0; k < D; k++) W[i * R + j] += weight[cc] * ti[k] * tj[k];
      }
  for (int i = 0; i < R; i++)
    for (int j = i; j < R; j++) W[i * R + j] = W[j * R + i];
}

/* computeEigenProblem, AccumulateTVStat.cpp:2999-3052: LAPACKE_dgeev on the (symmetric) matrix,
 * eigenvalues sorted descending, eigvec[k][j] = component k of the j-th largest eigenvector
 * (unit Euclidean norm; the sign is LAPACK's and not defined by the reference).  LAPACK is not
 * available here: cyclic Jacobi, which converges to the same eigen-pairs for a symmetric input.
 * Our sign convention: the component of largest magnitude is positive. */
int orc_eigen_sym(int n, const double *EP, int rank, double *eigvec, double *eigval) {
  double *a = (double *)malloc(sizeof(double) * n * n), *v = (double *)calloc((size_t)n * n, sizeof(double));
  memcpy(a, EP, sizeof(double) * n * n);
  for (int i = 0; i < n; i++) v[i * n + i] = 1.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0, diag = 0.0;
    for (int i = 0; i < n; i++)
      for (int j = 0; j < n; j++) {
        if (i == j) diag += a[i * n + j] * a[i * n + j];
        else off += a[i * n + j] * a[i * n + j];
      }
    if (off <= 1e-30 * diag || off == 0.0) break;
    for (int p = 0; p < n - 1; p++)
      for (int q = p + 1; q < n; q++) {
        double apq = a[p * n + q];
        if (apq == 0.0) continue;
        double theta = (a[q * n + q] - a[p * n + p]) / (2.0 * apq);
        double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          double akp = a[k * n + p], akq = a[k * n + q];
          a[k * n + p] = c * akp - s * akq;
          a[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          double apk = a[p * n + k], aqk = a[q * n + k];
          a[p * n + k] = c * apk - s * aqk;
          a[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          double vkp = v[k * n + p], vkq = v[k * n + q];
          v[k * n + p] = c * vkp - s * vkq;
          v[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  int *order = (int *)malloc(sizeof(int) * n);
  for (int i = 0; i < n; i++) order[i] = i;
  for (int i = 0; i < n; i++)  /* descendingSort */
    for (int j = i + 1; j < n; j++)
      if (a[order[j] * n + order[j]] > a[order[i] * n + order[i]]) {
        int t = order[i];
        order[i] = order[j];
        order[j] = t;
      }
  for (int j = 0; j < rank; j++) {
    int col = order[j], big = 0;
    for (int k = 1; k < n; k++)
      if (fabs(v[k * n + col]) > fabs(v[big * n + col])) big = k;
    double sg = v[big * n + col] < 0 ? -1.0 : 1.0;
    for (int k = 0; k < n; k++) eigvec[(size_t)k * rank + j] = sg * v[k * n + col];
    eigval[j] = a[col * n + col];
  }
  free(a);
  free(v);
  free(order);
  return 0;
}

/* approximateTcTcUnThreaded, AccumulateTVStat.cpp:3116-3136.  Q[R x R], Dm[C x R] (overwritten;
 * the reference accumulates into a zeroed D) */
void orc_tv_approximate_tctc(int C, int D, int R, const double *T, const double *Q, double *Dm) {
  size_t sv = (size_t)C * D;
  double *A = (double *)malloc(sizeof(double) * D * R);
  for (int cc = 0; cc < C; cc++) {
    for (int i = 0; i < D * R; i++) A[i] = 0.0;
    for (int i = 0; i < D; i++)
      for (int j = 0; j < R; j++)
        for (int k = 0; k < R; k++) A[i * R + j] += T[(size_t)k * sv + (size_t)cc * D + i] * Q[k * R + j];
    for (int i = 0; i < R; i++) {
      double d = 0.0;
      for (int k = 0; k < D; k++) d += A[k * R + i] * A[k * R + i];
      Dm[(size_t)cc * R + i] = d;
    }
  }
  free(A);
}

/* estimateWUbmWeightUnThreaded, AccumulateTVStat.cpp:2348-2396.  T and F are the NORMALISED
 * matrix / statistics; Wcov[R x R]; W[U x R] overwritten. */
void orc_tv_ivectors_ubm_weight(size_t U, int C, int D, int R, const double *N, const double *F,
                                const double *T, const double *Wcov, double *W) {
  size_t sv = (size_t)C * D;
  double *L = (double *)malloc(sizeof(double) * R * R), *Linv = (double *)malloc(sizeof(double) * R * R);
  double *aux = (double *)malloc(sizeof(double) * R);
  for (size_t spk = 0; spk < U; spk++) {
    double n_sum = 0.0;
    for (int c = 0; c < C; c++) n_sum += N[spk * C + c];
    for (int i = 0; i < R; i++) {
      for (int j = 0; j < R; j++) L[i * R + j] = n_sum * Wcov[i * R + j];
      L[i * R + i] += 1.0;
    }
    orc_invert(R, L, Linv);
    for (int i = 0; i < R; i++) {
      aux[i] = 0.0;
      for (size_t k = 0; k < sv; k++) aux[i] += F[spk * sv + k] * T[(size_t)i * sv + k];
    }
    for (int i = 0; i < R; i++) {
      double y = 0.0;
      for (int k = 0; k < R; k++) y += aux[k] * Linv[i * R + k];
      W[spk * R + i] = y;
    }
  }
  free(L);
  free(Linv);
  free(aux);
}

/* estimateWEigenDecompositionUnThreaded, AccumulateTVStat.cpp:2566-2609.  Dm[C x R], Q[R x R];
 * W[U x R] is ACCUMULATED into (the reference never resets _W in this function). */
void orc_tv_ivectors_eigen(size_t U, int C, int D, int R, const double *N, const double *F,
                           const double *T, const double *Dm, const double *Q, double *W) {
  size_t sv = (size_t)C * D;
  double *invL = (double *)malloc(sizeof(double) * R), *aux = (double *)malloc(sizeof(double) * R);
  double *appL = (double *)malloc(sizeof(double) * R * R);
  for (size_t spk = 0; spk < U; spk++) {
    for (int i = 0; i < R; i++) {
      double tmp = 1.0;
      for (int cc = 0; cc < C; cc++) tmp += N[spk * C + cc] * Dm[(size_t)cc * R + i];
      invL[i] = 1 / tmp;
    }
    for (int i = 0; i < R; i++) {
      aux[i] = 0.0;
      for (size_t k = 0; k < sv; k++) aux[i] += F[spk * sv + k] * T[(size_t)i * sv + k];
    }
    for (int i = 0; i < R; i++)
      for (int j = 0; j < R; j++) {
        double s = 0.0;
        for (int k = 0; k < R; k++) s += Q[i * R + k] * invL[k] * Q[j * R + k];
        appL[i * R + j] = s;
      }
    for (int i = 0; i < R; i++)
      for (int k = 0; k < R; k++) W[spk * R + i] += aux[k] * appL[i * R + k];
  }
  free(invL);
  free(aux);
  free(appL);
}

/* ====================================================================== i-vector back-end
 *  PldaDev statistics / normalisation and the non-PLDA scorings of IvTest.
 * Vectors are COLUMNS: data[d x n] like the reference's _data / _models / _segments. */

/* PldaDev::computeAll, PldaTools.cpp:353-385.  class_of[n] = speaker of each session */
void orc_iv_compute_all(int d, size_t n, const double *data, const int32_t *class_of, size_t n_spk,
                        double *mean, double *spk_means) {
  size_t *cnt = (size_t *)calloc(n_spk, sizeof(size_t));
  for (int k = 0; k < d; k++) mean[k] = 0.0;
  for (size_t i = 0; i < (size_t)d * n_spk; i++) spk_means[i] = 0.0;
  for (size_t s = 0; s < n; s++) {
    cnt[class_of[s]]++;
    for (int k = 0; k < d; k++) {
      spk_means[(size_t)k * n_spk + class_of[s]] += data[(size_t)k * n + s];
      mean[k] += data[(size_t)k * n + s];
    }
  }
  for (int k = 0; k < d; k++) {
    mean[k] /= (double)n;
    for (size_t c = 0; c < n_spk; c++) spk_means[(size_t)k * n_spk + c] /= (double)cnt[c];
  }
  free(cnt);
}

/* PldaDev::computeCovMatUnThreaded, PldaTools.cpp:527-571: total, within and between covariance */
void orc_iv_cov_mat(int d, size_t n, const double *data, const int32_t *class_of, size_t n_spk,
                    const double *mean, const double *spk_means, double *Sigma, double *W, double *B) {
  size_t *cnt = (size_t *)calloc(n_spk, sizeof(size_t));
  for (size_t s = 0; s < n; s++) cnt[class_of[s]]++;
  for (int i = 0; i < d * d; i++) Sigma[i] = W[i] = B[i] = 0.0;
  for (int i = 0; i < d; i++)
    for (int j = i; j < d; j++) {
      double sg = 0.0, w = 0.0, b = 0.0;
      for (size_t s = 0; s < n; s++) {
        size_t c = (size_t)class_of[s];
        sg += (data[(size_t)i * n + s] - mean[i]) * (data[(size_t)j * n + s] - mean[j]);
        w += (data[(size_t)i * n + s] - spk_means[(size_t)i * n_spk + c]) *
             (data[(size_t)j * n + s] - spk_means[(size_t)j * n_spk + c]);
      }
      for (size_t c = 0; c < n_spk; c++)
        b += (double)cnt[c] * (spk_means[(size_t)i * n_spk + c] - mean[i]) *
             (spk_means[(size_t)j * n_spk + c] - mean[j]);
      Sigma[i * d + j] = Sigma[j * d + i] = sg / (double)n;
      W[i * d + j] = W[j * d + i] = w / (double)n;
      B[i * d + j] = B[j * d + i] = b / (double)n;
    }
  free(cnt);
}

/* PldaDev::computeWccnCholUnThreaded, PldaTools.cpp:1124-1175: W = mean over speakers of the
 * per-speaker covariance, WCCN = upperCholesky(W^-1) */
int orc_iv_wccn_chol(int d, size_t n, const double *data, const int32_t *class_of, size_t n_spk,
                     const double *spk_means, double *WCCN) {
  double *W = (double *)calloc((size_t)d * d, sizeof(double)), *cov = (double *)malloc(sizeof(double) * d * d);
  double *invW = (double *)malloc(sizeof(double) * d * d);
  size_t *cnt = (size_t *)calloc(n_spk, sizeof(size_t));
  for (size_t s = 0; s < n; s++) cnt[class_of[s]]++;
  size_t s = 0;
  while (s < n) {
    size_t spk = (size_t)class_of[s];
    for (int i = 0; i < d * d; i++) cov[i] = 0.0;
    while (s < n && (size_t)class_of[s] == spk) {
      for (int i = 0; i < d; i++)
        for (int j = i; j < d; j++)
          cov[i * d + j] += (data[(size_t)i * n + s] - spk_means[(size_t)i * n_spk + spk]) *
                            (data[(size_t)j * n + s] - spk_means[(size_t)j * n_spk + spk]);
      s++;
    }
    for (int i = 0; i < d; i++)
      for (int j = i; j < d; j++) W[i * d + j] += cov[i * d + j] / (double)cnt[spk];
  }
  for (int i = 0; i < d; i++)
    for (int j = i; j < d; j++) {
      W[i * d + j] /= (double)n_spk;
      W[j * d + i] = W[i * d + j];
    }
  int rc = orc_invert(d, W, invW);
  if (rc == 0) rc = orc_upper_cholesky(d, invW, WCCN);
  free(W);
  free(cov);
  free(invW);
  free(cnt);
  return rc;
}

/* lengthNorm, PldaTools.cpp:436-464 / 3706-3750 */
void orc_iv_length_norm(int d, size_t n, double *data) {
  for (size_t s = 0; s < n; s++) {
    double t = 0.0;
    for (int k = 0; k < d; k++) t += data[(size_t)k * n + s] * data[(size_t)k * n + s];
    t = sqrt(t);
    for (int k = 0; k < d; k++) data[(size_t)k * n + s] /= t;
  }
}
/* center, PldaTools.cpp:466-474 / 3754-3767 */
void orc_iv_center(int d, size_t n, const double *mu, double *data) {
  for (size_t s = 0; s < n; s++)
    for (int k = 0; k < d; k++) data[(size_t)k * n + s] -= mu[k];
}
/* rotateLeft, PldaTools.cpp:498-514 / 3770-3790: out[r x n] = M[r x d] data[d x n] */
void orc_iv_rotate_left(int r, int d, size_t n, const double *M, const double *data, double *out) {
  for (int i = 0; i < r; i++)
    for (size_t s = 0; s < n; s++) {
      double t = 0.0;
      for (int k = 0; k < d; k++) t += M[i * d + k] * data[(size_t)k * n + s];
      out[(size_t)i * n + s] = t;
    }
}
/* the normalisation matrix of one sphericalNuisanceNormalization iteration,
 * PldaTools.cpp:1853-1900: (eigenVect diag(1 / sqrt(eigenVal)))^T of Sigma (EFR) or W (sphNorm) */
int orc_iv_efr_matrix(int d, const double *cov, double *mat) {
  double *vec = (double *)malloc(sizeof(double) * d * d), *val = (double *)malloc(sizeof(double) * d);
  orc_eigen_sym(d, cov, d, vec, val);
  int rc = 0;
  for (int i = 0; i < d; i++) {   /* mat[i][k] = vec[k][i] / sqrt(val[i]) */
    if (!(val[i] > 0.0)) rc = 1;
    for (int k = 0; k < d; k++) mat[i * d + k] = vec[k * d + i] / sqrt(val[i]);
  }
  free(vec);
  free(val);
  return rc;
}

/* PldaDev::computeLDA, PldaTools.cpp:1381-1415: leading eigenvectors of W^-1 B (dgeev in the
 * reference; here through the symmetric form U^-T B U^-1 with W = U^T U), unit norm, rows of
 * ldaMat[rank x d]; sign: largest-magnitude component positive. */
int orc_iv_lda(int d, const double *W, const double *B, int rank, double *ldaMat) {
  double *U = (double *)malloc(sizeof(double) * d * d), *Ui = (double *)malloc(sizeof(double) * d * d);
  double *Cm = (double *)malloc(sizeof(double) * d * d), *tmp = (double *)malloc(sizeof(double) * d * d);
  double *vec = (double *)malloc(sizeof(double) * d * rank), *val = (double *)malloc(sizeof(double) * rank);
  int rc = orc_upper_cholesky(d, W, U);
  if (rc == 0) rc = orc_invert(d, U, Ui);
  if (rc == 0) {
    /* Cm = Ui^T B Ui */
    for (int i = 0; i < d; i++)
      for (int j = 0; j < d; j++) {
        double t = 0.0;
        for (int k = 0; k < d; k++) t += B[i * d + k] * Ui[k * d + j];
        tmp[i * d + j] = t;
      }
    for (int i = 0; i < d; i++)
      for (int j = 0; j < d; j++) {
        double t = 0.0;
        for (int k = 0; k < d; k++) t += Ui[k * d + i] * tmp[k * d + j];
        Cm[i * d + j] = t;
      }
    for (int i = 0; i < d; i++)
      for (int j = i + 1; j < d; j++) Cm[i * d + j] = Cm[j * d + i] = 0.5 * (Cm[i * d + j] + Cm[j * d + i]);
    orc_eigen_sym(d, Cm, rank, vec, val);
    for (int j = 0; j < rank; j++) {   /* v = Ui y, normalised */
      double nrm = 0.0;
      int big = 0;
      for (int i = 0; i < d; i++) {
        double t = 0.0;
        for (int k = 0; k < d; k++) t += Ui[i * d + k] * vec[(size_t)k * rank + j];
        ldaMat[(size_t)j * d + i] = t;
        nrm += t * t;
      }
      nrm = sqrt(nrm);
      for (int i = 1; i < d; i++)
        if (fabs(ldaMat[(size_t)j * d + i]) > fabs(ldaMat[(size_t)j * d + big])) big = i;
      double sg = ldaMat[(size_t)j * d + big] < 0 ? -1.0 : 1.0;
      for (int i = 0; i < d; i++) ldaMat[(size_t)j * d + i] *= sg / nrm;
    }
  }
  free(U);
  free(Ui);
  free(Cm);
  free(tmp);
  free(vec);
  free(val);
  return rc;
}

/* PldaTest::cosineDistance, PldaTools.cpp:3842-3880.  trials[nm x nt] (may be NULL = all);
 * untested pairs keep the score 0 (the reference's zero-initialised _scores). */
void orc_iv_cosine(int d, size_t nm, size_t nt, const double *models, const double *segments,
                   const uint8_t *trials, double *scores) {
  double *nM = (double *)calloc(nm, sizeof(double)), *nS = (double *)calloc(nt, sizeof(double));
  for (int k = 0; k < d; k++) {
    for (size_t m = 0; m < nm; m++) nM[m] += models[(size_t)k * nm + m] * models[(size_t)k * nm + m];
    for (size_t s = 0; s < nt; s++) nS[s] += segments[(size_t)k * nt + s] * segments[(size_t)k * nt + s];
  }
  for (size_t m = 0; m < nm; m++) nM[m] = sqrt(nM[m]);
  for (size_t s = 0; s < nt; s++) nS[s] = sqrt(nS[s]);
  for (size_t m = 0; m < nm; m++)
    for (size_t s = 0; s < nt; s++) {
      double sc = 0.0;
      if (!trials || trials[m * nt + s]) {
        for (int k = 0; k < d; k++) sc += models[(size_t)k * nm + m] * segments[(size_t)k * nt + s];
        sc /= (nM[m] * nS[s]);
      }
      scores[m * nt + s] = sc;
    }
  free(nM);
  free(nS);
}

/* PldaTest::mahalanobisDistance, PldaTools.cpp:3882-3910 */
void orc_iv_mahalanobis(int d, size_t nm, size_t nt, const double *models, const double *segments,
                        const double *Mah, const uint8_t *trials, double *scores) {
  double *tmp = (double *)malloc(sizeof(double) * d), *t = (double *)malloc(sizeof(double) * d);
  for (size_t m = 0; m < nm; m++)
    for (size_t s = 0; s < nt; s++) {
      double sc = 0.0;
      if (!trials || trials[m * nt + s]) {
        for (int k = 0; k < d; k++) tmp[k] = models[(size_t)k * nm + m] - segments[(size_t)k * nt + s];
        for (int k = 0; k < d; k++) {
          t[k] = 0.0;
          for (int i = 0; i < d; i++) t[k] += -0.5 * tmp[i] * Mah[i * d + k];
        }
        for (int i = 0; i < d; i++) sc += t[i] * tmp[i];
      }
      scores[m * nt + s] = sc;
    }
  free(tmp);
  free(t);
}

/* PldaTest::twoCovScoring + twoCovScoringMixPartUnThreaded, PldaTools.cpp:4083-4173, 3923-3950 */
int orc_iv_two_cov(int d, size_t nm, size_t nt, const double *models, const double *segments,
                   const double *W, const double *B, double *scores) {
  size_t dd = (size_t)d * d;
  double *invW = (double *)malloc(sizeof(double) * dd), *invB = (double *)malloc(sizeof(double) * dd);
  double *sumG = (double *)malloc(sizeof(double) * dd), *sumH = (double *)malloc(sizeof(double) * dd);
  double *tG = (double *)malloc(sizeof(double) * dd), *tH = (double *)malloc(sizeof(double) * dd);
  double *tG2 = (double *)calloc(dd, sizeof(double)), *tH2 = (double *)calloc(dd, sizeof(double));
  double *G = (double *)calloc(dd, sizeof(double)), *H = (double *)calloc(dd, sizeof(double));
  int rc = orc_invert(d, W, invW) | orc_invert(d, B, invB);
  for (size_t i = 0; i < dd; i++) {
    sumG[i] = invB[i] + 2 * invW[i];
    sumH[i] = invB[i] + invW[i];
  }
  rc |= orc_invert(d, sumG, tG) | orc_invert(d, sumH, tH);
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++)
      for (int k = 0; k < d; k++) {
        tH2[i * d + j] += invW[i * d + k] * tH[k * d + j];
        tG2[i * d + j] += invW[i * d + k] * tG[k * d + j];
      }
  for (int i = 0; i < d; i++)
    for (int j = 0; j < d; j++)
      for (int k = 0; k < d; k++) {
        G[i * d + j] += tG2[i * d + k] * invW[k * d + j];
        H[i * d + j] += tH2[i * d + k] * invW[k * d + j];
      }
  double *mds = (double *)calloc(nm, sizeof(double)), *sds = (double *)calloc(nt, sizeof(double));
  double *a = (double *)malloc(sizeof(double) * d);
  for (size_t m = 0; m < nm; m++) {
    for (int k = 0; k < d; k++) {
      a[k] = 0.0;
      for (int i = 0; i < d; i++) a[k] += models[(size_t)i * nm + m] * H[i * d + k];
    }
    for (int j = 0; j < d; j++) mds[m] += a[j] * models[(size_t)j * nm + m];
  }
  for (size_t s = 0; s < nt; s++) {
    for (int k = 0; k < d; k++) {
      a[k] = 0.0;
      for (int i = 0; i < d; i++) a[k] += segments[(size_t)i * nt + s] * H[i * d + k];
    }
    for (int j = 0; j < d; j++) sds[s] += a[j] * segments[(size_t)j * nt + s];
  }
  double *diff = (double *)malloc(sizeof(double) * d);
  for (size_t m = 0; m < nm; m++)
    for (size_t s = 0; s < nt; s++) {
      double sc = 0.0;
      for (int j = 0; j < d; j++) diff[j] = models[(size_t)j * nm + m] + segments[(size_t)j * nt + s];
      for (int k = 0; k < d; k++) {
        a[k] = 0.0;
        for (int i = 0; i < d; i++) a[k] += diff[i] * G[i * d + k];
      }
      for (int j = 0; j < d; j++) sc += a[j] * diff[j];
      scores[m * nt + s] = sc - (mds[m] + sds[s]);
    }
  free(invW); free(invB); free(sumG); free(sumH); free(tG); free(tH); free(tG2); free(tH2);
  free(G); free(H); free(mds); free(sds); free(a); free(diff);
  return rc;
}

/* ====================================================================== PLDA EM training
 * PldaModel::em_iteration, PldaTools.cpp:2329-2343: _Dev.center(_Delta), computeCovMatEigen
 * (:931-950, the un-normalised scatter), getExpectedValuesUnThreaded (:2359-2485) and mStep
 * (:2790-2813), restated without Eigen.  data[d x n] is centred in place; class_of is
 * non-decreasing (sessions of a speaker are adjacent, :2428-2434).  F[d x rF], G[d x rG],
 * Sigma[d x d], Delta[d] are updated in place.  The EigenSolver / dgeev call on the symmetric
 * matrix A (:2377-2400) is the cyclic Jacobi solver here. */
static void mm(int m, int n, int k, const double *A, int ta, const double *B, int tb, double *Cm) {
  /* Cm[m x n] = op(A) op(B); A is [m x k] (or [k x m] when ta), B is [k x n] (or [n x k] when tb) */
  for (int i = 0; i < m; i++)
    for (int j = 0; j < n; j++) {
      double t = 0.0;
      for (int l = 0; l < k; l++) t += (ta ? A[(size_t)l * m + i] : A[(size_t)i * k + l]) * (tb ? B[(size_t)j * k + l] : B[(size_t)l * n + j]);
      Cm[(size_t)i * n + j] = t;
    }
}

int orc_plda_em_iteration(int d, int rF, int rG, size_t n, double *data, const int32_t *class_of,
                          size_t n_spk, double *F, double *G, double *Sigma, double *Delta) {
  const int r = rF + rG;
  int rc = 0;
#define NEW(count) ((double *)calloc((size_t)(count) > 0 ? (size_t)(count) : 1, sizeof(double)))
  /* _Dev.center(_Delta) */
  for (size_t s = 0; s < n; s++)
    for (int k = 0; k < d; k++) data[(size_t)k * n + s] -= Delta[k];
  /* computeCovMatEigen: sigmaObs = X X^T */
  double *sigmaObs = NEW((size_t)d * d);
  for (int i = 0; i < d; i++)
    for (int j = i; j < d; j++) {
      double t = 0.0;
      for (size_t s = 0; s < n; s++) t += data[(size_t)i * n + s] * data[(size_t)j * n + s];
      sigmaObs[i * d + j] = sigmaObs[j * d + i] = t;
    }
  /* preComputation :2950-2972 */
  double *iS = NEW((size_t)d * d), *Ftw = NEW((size_t)rF * d), *Gtw = NEW((size_t)rG * d);
  double *GtwG = NEW((size_t)rG * rG), *FtwG = NEW((size_t)rF * rG), *iGG = NEW((size_t)rG * rG);
  rc |= orc_invert(d, Sigma, iS);
  mm(rF, d, d, F, 1, iS, 0, Ftw);
  mm(rG, d, d, G, 1, iS, 0, Gtw);
  mm(rG, rG, d, Gtw, 0, G, 0, GtwG);
  mm(rF, rG, d, Ftw, 0, G, 0, FtwG);
  for (int i = 0; i < rG; i++) GtwG[i * rG + i] += 1.0;
  if (rG > 0) rc |= orc_invert(rG, GtwG, iGG);
  /* :2370-2373 */
  double *FtwF = NEW((size_t)rF * rF), *S = NEW((size_t)rG * rF), *A = NEW((size_t)rF * rF), *tmp = NEW((size_t)rF * rG);
  mm(rF, rF, d, Ftw, 0, F, 0, FtwF);
  mm(rG, rF, rG, iGG, 0, FtwG, 1, S);
  mm(rF, rG, rG, FtwG, 0, iGG, 0, tmp);
  mm(rF, rF, rG, tmp, 0, FtwG, 1, A);
  for (int i = 0; i < rF * rF; i++) A[i] = FtwF[i] - A[i];
  for (int i = 0; i < rF; i++)
    for (int j = i + 1; j < rF; j++) A[i * rF + j] = A[j * rF + i] = 0.5 * (A[i * rF + j] + A[j * rF + i]);
  double *V = NEW((size_t)rF * rF), *Dv = NEW(rF);
  orc_eigen_sym(rF, A, rF, V, Dv);
  /* accumulators :2403-2405 */
  double *Ehh = NEW((size_t)r * r), *xh = NEW((size_t)d * r), *U = NEW(r);
  double *M = NEW((size_t)rF * rF), *MsT = NEW((size_t)rF * rG), *SMsT = NEW((size_t)rG * rG);
  size_t curNb = 0, sc = 0;
  for (size_t spk = 0; spk < n_spk && sc < n; spk++) {
    const size_t first = sc;
    const int32_t cls = class_of[sc];
    while (sc < n && class_of[sc] == cls) sc++;
    const size_t ns = sc - first;
    if (ns != curNb) { /* :2417-2427 */
      curNb = ns;
      for (int i = 0; i < rF; i++)
        for (int j = 0; j < rF; j++) {
          double t = 0.0;
          for (int k = 0; k < rF; k++) t += V[i * rF + k] * V[j * rF + k] / ((double)ns * Dv[k] + 1.0);
          M[i * rF + j] = t;
        }
      mm(rF, rG, rF, M, 0, S, 1, MsT);
      mm(rG, rG, rF, S, 0, MsT, 0, SMsT);
    }
    /* Sigma_x, fi, gi :2436-2451 */
    double *Sx = NEW((size_t)d * ns), *fi = NEW((size_t)rF * ns), *gi = NEW((size_t)rG * ns);
    double *f = NEW(rF), *g = NEW(rG), *eh = NEW(rF), *v = NEW(rF), *Eh = NEW((size_t)r * ns);
    for (int i = 0; i < d; i++)
      for (size_t j = 0; j < ns; j++) {
        double t = 0.0;
        for (int k = 0; k < d; k++) t += iS[i * d + k] * data[(size_t)k * n + first + j];
        Sx[(size_t)i * ns + j] = t;
      }
    mm(rF, (int)ns, d, F, 1, Sx, 0, fi);
    mm(rG, (int)ns, d, G, 1, Sx, 0, gi);
    for (int i = 0; i < rF; i++)
      for (size_t j = 0; j < ns; j++) f[i] += fi[(size_t)i * ns + j];
    for (int i = 0; i < rG; i++)
      for (size_t j = 0; j < ns; j++) g[i] += gi[(size_t)i * ns + j];
    /* thisEh = M (f - S^T g) :2453 */
    for (int i = 0; i < rF; i++) {
      double t = f[i];
      for (int k = 0; k < rG; k++) t -= S[k * rF + i] * g[k];
      v[i] = t;
    }
    for (int i = 0; i < rF; i++) {
      double t = 0.0;
      for (int k = 0; k < rF; k++) t += M[i * rF + k] * v[k];
      eh[i] = t;
    }
    /* Eh :2455-2464 */
    for (int i = 0; i < rF; i++)
      for (size_t j = 0; j < ns; j++) Eh[(size_t)i * ns + j] = eh[i];
    for (int i = 0; i < rG; i++) {
      double se = 0.0;
      for (int k = 0; k < rF; k++) se += S[i * rF + k] * eh[k];
      for (size_t j = 0; j < ns; j++) {
        double t = 0.0;
        for (int k = 0; k < rG; k++) t += iGG[i * rG + k] * gi[(size_t)k * ns + j];
        Eh[(size_t)(rF + i) * ns + j] = t - se;
      }
    }
    /* EhhSum += ns * tmpM + Eh Eh^T :2467-2473 */
    for (int i = 0; i < r; i++)
      for (int j = 0; j < r; j++) {
        double tm;
        if (i < rF && j < rF) tm = M[i * rF + j];
        else if (i < rF) tm = -MsT[i * rG + (j - rF)];
        else if (j < rF) tm = -MsT[j * rG + (i - rF)];
        else tm = iGG[(i - rF) * rG + (j - rF)] + SMsT[(i - rF) * rG + (j - rF)];
        double e2 = 0.0;
        for (size_t k = 0; k < ns; k++) e2 += Eh[(size_t)i * ns + k] * Eh[(size_t)j * ns + k];
        Ehh[i * r + j] += (double)ns * tm + e2;
      }
    /* xhSum :2476-2479, Umx :2482-2483 */
    for (int i = 0; i < d; i++)
      for (int j = 0; j < r; j++)
        for (size_t k = 0; k < ns; k++) xh[(size_t)i * r + j] += data[(size_t)i * n + first + k] * Eh[(size_t)j * ns + k];
    for (size_t k = 0; k < ns; k++)
      for (int j = 0; j < r; j++) U[j] += Eh[(size_t)j * ns + k];
    free(Sx); free(fi); free(gi); free(f); free(g); free(eh); free(v); free(Eh);
  }
  /* mStep :2790-2813 */
  double *iEhh = NEW((size_t)r * r), *FG = NEW((size_t)d * r), *SL = NEW((size_t)d * d);
  rc |= orc_invert(r, Ehh, iEhh);
  mm(d, r, r, xh, 0, iEhh, 0, FG);
  mm(d, d, r, FG, 0, xh, 1, SL);
  for (int i = 0; i < d * d; i++) Sigma[i] = (sigmaObs[i] - SL[i]) / (double)n;
  for (int j = 0; j < r; j++) U[j] /= (double)n;
  double *c = NEW((size_t)r * r), *cF = NEW((size_t)rF * rF), *cG = NEW((size_t)rG * rG);
  double *Rh = NEW((size_t)rF * rF), *Rw = NEW((size_t)rG * rG);
  for (int i = 0; i < r; i++)
    for (int j = 0; j < r; j++) c[i * r + j] = Ehh[i * r + j] / (double)n - U[i] * U[j];
  for (int i = 0; i < rF; i++)
    for (int j = 0; j < rF; j++) cF[i * rF + j] = c[i * r + j];
  for (int i = 0; i < rG; i++)
    for (int j = 0; j < rG; j++) cG[i * rG + j] = c[(rF + i) * r + rF + j];
  rc |= orc_upper_cholesky(rF, cF, Rh);
  if (rG > 0) rc |= orc_upper_cholesky(rG, cG, Rw);
  /* F = FGEst[:, :rF] Rh^T ; G = FGEst[:, rF:] Rw^T ; Delta += FGEst Umx */
  for (int i = 0; i < d; i++) {
    for (int j = 0; j < rF; j++) {
      double t = 0.0;
      for (int k = 0; k < rF; k++) t += FG[(size_t)i * r + k] * Rh[j * rF + k];
      F[(size_t)i * rF + j] = t;
    }
    for (int j = 0; j < rG; j++) {
      double t = 0.0;
      for (int k = 0; k < rG; k++) t += FG[(size_t)i * r + rF + k] * Rw[j * rG + k];
      G[(size_t)i * rG + j] = t;
    }
    double t = 0.0;
    for (int k = 0; k < r; k++) t += FG[(size_t)i * r + k] * U[k];
    Delta[i] += t;
  }
  free(sigmaObs); free(iS); free(Ftw); free(Gtw); free(GtwG); free(FtwG); free(iGG); free(FtwF); free(S);
  free(A); free(tmp); free(V); free(Dv); free(Ehh); free(xh); free(U); free(M); free(MsT); free(SMsT);
  free(iEhh); free(FG); free(SL); free(c); free(cF); free(cG); free(Rh); free(Rw);
#undef NEW
  return rc;
}
