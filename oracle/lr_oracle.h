/*
 * lr_oracle.h -- fp64 CPU restatement of the LIA_RAL GMM / i-vector / PLDA hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product: only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may load this library, and only as the checker / the timed CPU baseline.
 * The product (lia_ral_b200/) never links, imports or falls back to it.
 *
 * Provenance.  The reference (ALIZE-Speaker-Recognition/LIA_RAL, /root/reference)
 * cannot be compiled here: every per-frame arithmetic call lives in alize-core
 * (../alize-core, not vendored, no version pin, configure.ac:50-62), and autotools
 * are absent.  So this file restates
 *   (a) the LIA_RAL loops literally (file:line cited at each function), and
 *   (b) the alize-core semantics those loops call (computeAll, computeLK,
 *       computeAndAccumulate{Occ,EM,LLK}, getEM, invert, upperCholesky) from the
 *       library's published behaviour; these are marked [ALIZE] below.
 * Pinning: the restatement is checked against the reference's own intact fixtures
 *   - LIA_Utils/GmmTokenizer/test: arg-max Gaussian stream (test1.sym.ref) and the
 *     top-20 confusion matrix (mce_matrix.mat.ref)  -> strict KATs, both pass;
 *   - LIA_SpkDet/TrainWorld/test/wld.validate: computeAll identity (det/cst);
 *   - ComputeTest test1.validate.res property: client == world => LLR == 0.
 * For the i-vector / T-matrix EM / PLDA rows the reference ships no test at all:
 * PARITY UNPINNED there (oracle == literal restatement of the cited loops,
 * cross-checked against an independent numpy/scipy formulation).
 *
 * All matrices are row-major double unless stated.  Frames are float32 on input
 * (the on-disk type; Feature::getDataVector() widens them to double,
 * AccumulateTVStat.cpp:336).
 */
#ifndef LR_ORACLE_H
#define LR_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_EPS_LK 1e-200 /* TopGauss.cpp:67 */

/* ---- A.1 DistribGD::computeAll [ALIZE]; probed on wld.validate / RAW fixtures */
void orc_gmm_compute_all(int C, int D, const double *cov, double *covinv, double *det,
                         double *cst);

/* ---- A.2 DistribGD::computeLK [ALIZE]; in-repo restatements GeneralTools.cpp:816-826 */
double orc_distrib_lk(int D, const double *x, const double *mean, const double *covinv,
                      double cst);

/* ---- A.3 weighted component likelihoods p_c = w_c lk_c(x); returns sum_c p_c */
double orc_frame_likelihoods(int C, int D, const double *w, const double *mean,
                             const double *covinv, const double *cst, const float *x,
                             double *p /*[C]*/);

/* ---- A.4 Baum-Welch N/F per NDX line (AccumulateTVStat.cpp:332-349).
 * frame2row[t] = row (NDX line) of frame t, or <0 to skip the frame.
 * N[U*C], F[U*C*D] are accumulated into (+=).  threads>1 = contiguous row ranges
 * per thread (AccumulateTVStat.cpp:498-507). */
void orc_bwstats(int C, int D, const double *w, const double *mean, const double *covinv,
                 const double *cst, const float *X, size_t T, size_t ldx,
                 const int32_t *frame2row, size_t U, double *N, double *F, int threads);

/* JFAAcc::normalizeFeatures (AccumulateJFAStat.cpp:4623-4680): x_t -= sum_k P(k | x_t) ux[k*D + i] for the
 * frames of the segments, in order, in place, posteriors under the session model (w, mean, covinv, cst). */
void orc_jfa_normalize_features(int C, int D, const double *w, const double *mean, const double *covinv,
                                const double *cst, const double *ux, float *X, size_t ldx,
                                const int64_t *seg_begin, const int64_t *seg_len, size_t n_segs);

/* ---- A.5 EM accumulate (MixtureGDStat::computeAndAccumulateEM [ALIZE], driver
 * AccumulateStat.cpp:103-128, threaded :170-299).  Accumulates into occ[C],
 * m1[C*D], m2[C*D]; returns sum_t log(sum_c p_c) and adds T*weight to *nframes. */
double orc_em_accumulate(int C, int D, const double *w, const double *mean,
                         const double *covinv, const double *cst, const float *X, size_t T,
                         size_t ldx, double frame_weight, double *occ, double *m1, double *m2,
                         double *nframes, int threads);
/* getEM [ALIZE]: w=occ/sum occ, mean=m1/occ, cov=m2/occ-mean^2 (components with
 * occ==0 keep their previous parameters). */
void orc_em_get(int C, int D, const double *occ, const double *m1, const double *m2, double *w,
                double *mean, double *cov);
/* TrainTools.cpp:567-587 */
void orc_variance_control(int C, int D, double *cov, double flooring, double ceiling,
                          const double *cov_signal, long *n_floor, long *n_ceil);
/* TrainTools.cpp:560-564 */
double orc_set_it_parameter(double begin, double end, int nb_it, int it);
/* FrameAccGD [ALIZE] via computeMeanCov TrainTools.cpp:593-602 */
void orc_mean_cov(int D, const float *X, size_t T, size_t ldx, double *mean, double *cov);

/* ---- A.6 frame log-likelihood with top-K (computeAndAccumulateLLK [ALIZE];
 * call sites ComputeTest.cpp:162-167, TopGauss.cpp:166-192). */
/* DETERMINE_TOP_DISTRIBS on the world model.  Outputs per frame: llk[T] (clamped),
 * idx[T*K] (descending p_c, ties -> lowest index), top_lk[T*K] (p_c of the kept
 * components), rest_lk[T] = sum of p_c outside the top K, rest_w[T] likewise for weights. */
void orc_llk_determine_top(int C, int D, const double *w, const double *mean,
                           const double *covinv, const double *cst, const float *X, size_t T,
                           size_t ldx, int K, int complete, double min_llk, double max_llk,
                           double *llk, uint32_t *idx, double *top_lk, double *rest_lk,
                           double *rest_w);
/* USE_TOP_DISTRIBS on a client model. */
void orc_llk_use_top(int C, int D, const double *w, const double *mean, const double *covinv,
                     const double *cst, const float *X, size_t T, size_t ldx, int K,
                     const uint32_t *idx, const double *rest_lk, int complete, double min_llk,
                     double max_llk, double *llk);
/* TOP_DISTRIBS_NO_ACTION: llk over all components (accumulateStatLLK, AccumulateStat.cpp:69-94) */
void orc_llk_all(int C, int D, const double *w, const double *mean, const double *covinv,
                 const double *cst, const float *X, size_t T, size_t ldx, double min_llk,
                 double max_llk, double *llk);

/* ---- A.7 / A.8 Total Variability (AccumulateTVStat.cpp) */
/* substractM :1088-1105 */
void orc_tv_subtract_m(size_t U, int C, int D, const double *N, const double *ubm_mean,
                       double *F);
/* estimateTETtUnThreaded :777-805.  T[R x C*D], tett[C][R*R] */
void orc_tv_tett(int C, int D, int R, const double *T, const double *invvar, double *tett,
                 int threads);
/* estimateWUnThreaded :2114-2169.  W[U*R] is overwritten. */
void orc_tv_ivectors(size_t U, int C, int D, int R, const double *N, const double *F,
                     const double *T, const double *invvar, const double *tett, double *W,
                     int threads);
/* estimateAandCUnthreaded :1702-1795.  A[C x R*R], Cmx[R x C*D] (+= as the reference:
 * A is zeroed inside, Cmx is NOT), Rm[R*R], r[R], meanW[R], W[U*R] overwritten. */
void orc_tv_estep(size_t U, int C, int D, int R, const double *N, const double *F,
                  const double *T, const double *invvar, const double *tett, double *W,
                  double *A, double *Cmx, double *Rm, double *r, double *meanW, int threads);
/* updateTestimate :974-1005.  T <- A_c^-1 Cmx_c per component */
void orc_tv_mstep(int C, int D, int R, const double *A, const double *Cmx, double *T);
/* minDivergence :2056-2099 (Rm, r are modified in place like the reference) */
int orc_tv_mindiv(int C, int D, int R, double n_sessions, double *Rm, double *r,
                  const double *meanW, double *ubm_mean, double *T);
/* orthonormalizeT :1548-1596 */
void orc_tv_orthonormalize(int R, size_t sv, double *T);

/* ---- A.8b approximate i-vector modes */
/* normTMatrix :1600-1609 */
void orc_tv_norm_t(int R, size_t sv, const double *invvar, double *T);
/* normStatisticsUnThreaded :1225-1242 */
void orc_tv_norm_statistics(size_t U, int C, int D, const double *N, const double *ubm_mean,
                            const double *invvar, double *F);
/* getWeightedCovUnThreaded :2837-2855 */
void orc_tv_weighted_cov(int C, int D, int R, const double *T, const double *weight, double *W);
/* computeEigenProblem :2999-3052 (dgeev on a symmetric matrix; here cyclic Jacobi); eigvec[n x rank] */
int orc_eigen_sym(int n, const double *EP, int rank, double *eigvec, double *eigval);
/* approximateTcTcUnThreaded :3116-3136 */
void orc_tv_approximate_tctc(int C, int D, int R, const double *T, const double *Q, double *Dm);
/* estimateWUbmWeightUnThreaded :2348-2396 */
void orc_tv_ivectors_ubm_weight(size_t U, int C, int D, int R, const double *N, const double *F,
                                const double *T, const double *Wcov, double *W);
/* estimateWEigenDecompositionUnThreaded :2566-2609 (W += ...) */
void orc_tv_ivectors_eigen(size_t U, int C, int D, int R, const double *N, const double *F,
                           const double *T, const double *Dm, const double *Q, double *W);

/* ---- A.9 PLDA native scoring (PldaTools.cpp:2950-2972, 4489-4519, 4186-4271).
 * F[d x rF], G[d x rG] (rG may be 0), Sigma[d x d].  models[d x n_enrol] and
 * segments[d x n_test] are column-per-i-vector like the reference.  model_of[n_enrol]
 * gives, for each enrolment column, its model number (non-decreasing, consecutive
 * columns of a model are adjacent as in _modelIndexLine).  scores[n_models x n_test]. */
int orc_plda_native_scoring(int d, int rF, int rG, const double *F, const double *G,
                            const double *Sigma, const double *models, size_t n_enrol,
                            const int32_t *model_of, size_t n_models, const double *segments,
                            size_t n_test, double *scores);

/* ---- A.10 i-vector back-end: PldaDev statistics / normalisation and the
 * cosine / Mahalanobis / two-covariance scorings of IvTest.  Vectors are columns: data[d x n]. */
void orc_iv_compute_all(int d, size_t n, const double *data, const int32_t *class_of, size_t n_spk,
                        double *mean, double *spk_means);                 /* PldaTools.cpp:353-385 */
void orc_iv_cov_mat(int d, size_t n, const double *data, const int32_t *class_of, size_t n_spk,
                    const double *mean, const double *spk_means, double *Sigma, double *W,
                    double *B);                                            /* :527-571 */
int orc_iv_wccn_chol(int d, size_t n, const double *data, const int32_t *class_of, size_t n_spk,
                     const double *spk_means, double *WCCN);               /* :1124-1175 */
void orc_iv_length_norm(int d, size_t n, double *data);                    /* :436-464, 3706-3750 */
void orc_iv_center(int d, size_t n, const double *mu, double *data);       /* :466-474, 3754-3767 */
void orc_iv_rotate_left(int r, int d, size_t n, const double *M, const double *data,
                        double *out);                                      /* :498-514, 3770-3790 */
int orc_iv_efr_matrix(int d, const double *cov, double *mat);              /* :1853-1900 */
int orc_iv_lda(int d, const double *W, const double *B, int rank, double *ldaMat); /* :1381-1415 */
void orc_iv_cosine(int d, size_t nm, size_t nt, const double *models, const double *segments,
                   const uint8_t *trials, double *scores);                 /* :3842-3880 */
void orc_iv_mahalanobis(int d, size_t nm, size_t nt, const double *models, const double *segments,
                        const double *Mah, const uint8_t *trials, double *scores); /* :3882-3910 */
int orc_iv_two_cov(int d, size_t nm, size_t nt, const double *models, const double *segments,
                   const double *W, const double *B, double *scores);      /* :4083-4173 */

/* ---- A.11 PLDA EM training: one PldaModel::em_iteration (PldaTools.cpp:2329-2343, 2359-2485,
 * 2790-2813, 931-950).  data[d x n] is centred by Delta in place; F, G, Sigma, Delta updated. */
int orc_plda_em_iteration(int d, int rF, int rG, size_t n, double *data, const int32_t *class_of,
                          size_t n_spk, double *F, double *G, double *Sigma, double *Delta);

/* dense helpers ([ALIZE] DoubleSquareMatrix::invert / upperCholesky) */
int orc_invert(int n, const double *a, double *inv);
int orc_upper_cholesky(int n, const double *a, double *u); /* a = u^T u, u upper */

#ifdef __cplusplus
}
#endif
#endif
