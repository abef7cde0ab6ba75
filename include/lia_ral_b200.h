/*
 * lia_ral_b200.h -- C ABI of the B200-native engine for LIA_RAL's GMM / i-vector hot path.
 *
 * Drop-in boundary: the reference has no FFI; its hot path is ordinary C++
 * calls from libliatools (LIA_SpkTools) into alize-core, ONE FRAME AT A TIME.  This ABI is
 * inserted at the per-BATCH seam instead -- the LIA_SpkTools functions that own the frame /
 * utterance loops -- and every entry point cites the loop it replaces.
 *
 * Conventions
 *   - plain C types only; all host matrices row-major double exactly as the reference's
 *     Matrix<double>::getArray() / DoubleVector::getArray() hand them over; frames are the
 *     on-disk float32 (Feature::getDataVector() widens them, AccumulateTVStat.cpp:336).
 *   - statistics are ACCUMULATED INTO (+=) the caller's buffers like the reference
 *     (resetEM / resetAcc stay the caller's job: TrainTools.cpp:1057, AccumulateTVStat.cpp:613).
 *   - every function returns lr_status (0 = LR_OK) or a handle (NULL on failure);
 *     lr_last_error() returns the thread-local message.  The C++ host mirror rethrows it as
 *     an alize::Exception-compatible exception (reference convention: AccumulateTVStat.cpp:481).
 *   - NO CPU FALLBACK: without a CUDA device every compute entry point fails with LR_ERR_CUDA.
 *   - "_dev" variants take DEVICE pointers (e.g. torch tensors' data_ptr()) and enqueue on the
 *     engine stream without synchronising; host variants copy H2D/D2H inside the call.
 *   - handles are not thread-safe; one host thread per device (the reference's numThread is
 *     parsed by the host mirror and ignored by this backend).
 */
#ifndef LIA_RAL_B200_H
#define LIA_RAL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int lr_status;
enum {
  LR_OK = 0,
  LR_ERR_ARG = 1,   /* bad argument (reference: Exception thrown by the caller-side checks) */
  LR_ERR_CUDA = 2,  /* CUDA / cuBLAS failure, or no device */
  LR_ERR_NUMERIC = 3, /* singular / non-SPD matrix (reference: invert()/upperCholesky() failing) */
  LR_ERR_IO = 4      /* rendezvous file of lr_comm_init_file unreadable / unwritable */
};

const char *lr_last_error(void);
const char *lr_version(void);

/* Bind the calling process to one GPU (one process per GPU; rank r -> device r). */
lr_status lr_init(int device);
lr_status lr_shutdown(void);
lr_status lr_synchronize(void);
/* cudaStream_t of the engine as an integer, so torch can order against it. */
uint64_t lr_stream_handle(void);
int lr_sm_count(void);
/* Number of kernels this library has launched since the last reset (bench.py gpu_launches). */
uint64_t lr_launch_count(void);
void lr_reset_launch_count(void);
/* Per-kernel device timing for bench.py's roofline line: while enabled, every launch of the
 * two frames x components kernels is bracketed by CUDA events on the engine stream.
 * kind 0 = log-likelihood pass, 1 = statistics pass.  lr_profile_read synchronises. */
lr_status lr_profile(int enable);
lr_status lr_profile_read(int kind, double *total_ms, uint64_t *n_launches);
/* Kernel selection for the frames x components pass: 0 = auto, 1 = fp32 SIMT, 2 = tcgen05 (one-pass
 * likelihood + statistics kernel, two-pass for likelihoods only), 3 = tcgen05 with the two-pass
 * statistics kernels of round 1 (kept as the cross-check of the one-pass kernel). */
lr_status lr_set_gmm_kernel(int which);
int lr_get_gmm_kernel(void);
/* fp16 products the one-pass tcgen05 kernel issues per frame tile.  The frame operand is carried as
 * hi + lo fp16 panels, the model operand as hi + lo, the posteriors as ONE fp16 value:
 *   0: (default) likelihood  W_hi X_hi + W_hi X_lo + W_lo X_hi, statistics  P X_hi + P X_lo   (5 products)
 *   1: statistics on X_hi only (4 products): a zero-mean 2^-12 relative rounding per frame term.  Inside
 *      the contract wherever a component sees >= 100 frames; a component fed by a handful of frames with
 *      one-hot posteriors shows it directly (first-order statistics to 1e-4, variances to 2e-3)
 *   2: likelihood without W_hi X_lo as well (3 products): per-frame log-likelihoods to ~3e-5 relative,
 *      i-vectors of 20 000-frame utterances to 1.1e-4 -- outside the 1e-4 contract
 * Levels 1 and 2 are opt-in (+8 % / +19 % frames/s at 2048c/60d).
 * Measured errors and speeds per level: DESIGN.md 4.6 (scripts/products_probe.py). */
lr_status lr_set_gmm_products(int level);
int lr_get_gmm_products(void);

/* ------------------------------------------------------------------ GMM (MixtureGD) ------
 * Replaces MixtureGD + DistribGD::computeAll (alize-core; constants probed on
 * LIA_SpkDet/TrainWorld/test/wld.validate): covInv = 1/cov, det = prod cov,
 * cst = 1/((2pi)^(D/2) sqrt(det)).  w[C], mean[C*D], cov[C*D]. */
typedef struct lr_gmm lr_gmm;
lr_gmm *lr_gmm_create(int C, int D, const double *w, const double *mean, const double *cov);
lr_status lr_gmm_set(lr_gmm *g, const double *w, const double *mean, const double *cov);
/* any output may be NULL */
lr_status lr_gmm_get(lr_gmm *g, double *w, double *mean, double *cov, double *covinv,
                     double *cst, double *det);
/* override the stored cst (RAW model files carry their own cst/det records) */
lr_status lr_gmm_set_cst(lr_gmm *g, const double *cst);
void lr_gmm_destroy(lr_gmm *g);

/* ------------------------------------------------------------------ frames in HBM ---------
 * Replaces the FeatureServer buffer (featureServerBufferSize ALL_FEATURES): a [T x D] float32
 * block resident on the device so EM iterations do not re-cross PCIe. */
typedef struct lr_feats lr_feats;
lr_feats *lr_feats_upload(const float *X, size_t T, size_t ldx, int D);
lr_feats *lr_feats_wrap_device(const float *dX, size_t T, size_t ldx, int D);
void lr_feats_destroy(lr_feats *f);
/* A handle keeps, between lr_gmm_em_accumulate_dev calls over the same frame range, the frames' tensor-core
 * operand (512 B per frame; dropped above LR_CONV_CACHE_GB, default 24): EM iterations re-read unchanged
 * frames.  Frames of a WRAPPED buffer must therefore not change while the handle lives -- or call this. */
lr_status lr_feats_invalidate(lr_feats *f);

/* A run of selected frames and the statistics row it feeds: the reference's Seg
 * (sourceName/begin/length after fs.getFirstFeatureIndexOfASource) + the NDX line from
 * TVTranslate::locIndices (AccumulateTVStat.cpp:313-346).  A file listed on several NDX
 * lines is passed as several segments over the same frames. */
typedef struct {
  int64_t begin;  /* first frame (index into X) */
  int64_t length; /* number of frames */
  int32_t row;    /* statistics row (NDX line); ignored by the EM entry points */
  int32_t pad_;
} lr_seg;

/* ---- a4/a5: accumulateStatEM (AccumulateStat.cpp:103-140, threaded :170-299) over
 * MixtureGDStat::computeAndAccumulateEM.  occ[C] += g, m1[C*D] += g x, m2[C*D] += g x^2
 * (g = frame_weight * posterior), *sum_log_lk += sum_t log(sum_c w_c lk_c(x_t)),
 * *n_frames += frame_weight * #frames.  The log-likelihood sum is NOT weighted: it is the reference's
 * llkAcc (AccumulateStat.cpp:104-107 adds log(computeAndAccumulateEM(f)) per frame), n_frames its
 * getEMFeatureCount().  segs == NULL -> all T frames. */
lr_status lr_gmm_em_accumulate(lr_gmm *g, const float *X, size_t T, size_t ldx,
                               const lr_seg *segs, size_t n_segs, double frame_weight,
                               double *occ, double *m1, double *m2, double *sum_log_lk,
                               double *n_frames);
/* device-resident variant: d_stats = [occ C | m1 C*D | m2 C*D | sum_log_lk | n_frames] doubles
 * in device memory, accumulated into; frames [t0, t0+T) of f. */
lr_status lr_gmm_em_accumulate_dev(lr_gmm *g, const lr_feats *f, size_t t0, size_t T,
                                   double frame_weight, double *d_stats);
size_t lr_gmm_em_stats_len(const lr_gmm *g); /* C + 2*C*D + 2 */
/* MixtureGDStat::getEM + varianceControl (TrainTools.cpp:567-587,1076-1077) on the device:
 * w = occ/sum occ, mean = m1/occ, cov = m2/occ - mean^2, then clamp cov to
 * [flooring*cov_signal, ceiling*cov_signal] (floor first), then computeAll.  d_cov_signal may
 * be NULL (no variance control).  g is updated in place on the device; the call ends with ONE 8-byte
 * readback (the fp16 range guard of the tensor-core operands and the "normalised space re-derived" flag
 * that invalidates cached frame operands), i.e. it waits for the M-step kernels.
 * Weights are left unchanged when the total occupation is not positive. */
lr_status lr_gmm_em_update_dev(lr_gmm *g, const double *d_stats, double flooring, double ceiling,
                               const double *d_cov_signal);
/* host convenience: getEM + varianceControl from host statistics */
lr_status lr_gmm_em_update(lr_gmm *g, const double *occ, const double *m1, const double *m2,
                           double flooring, double ceiling, const double *cov_signal);
/* FrameAccGD via computeMeanCov (TrainTools.cpp:593-602): global mean / cov of the frames */
lr_status lr_frames_mean_cov(const float *X, size_t T, size_t ldx, int D, double *mean,
                             double *cov);

/* ---- a2/a3: TVAcc::computeAndAccumulateTVStat (AccumulateTVStat.cpp:268-351, threaded
 * :376-548): N[U x C] += posterior, F[U x C*D] += posterior * x for the frames of each
 * segment, into row seg.row. */
lr_status lr_gmm_bwstats(lr_gmm *g, const float *X, size_t T, size_t ldx, const lr_seg *segs,
                         size_t n_segs, size_t U, double *N, double *F);
/* device-resident variant: d_N / d_F device doubles, d_segs host array (small) */
lr_status lr_gmm_bwstats_dev(lr_gmm *g, const lr_feats *f, const lr_seg *segs, size_t n_segs,
                             size_t U, double *d_N, double *d_F);
/* JFAAcc::computeAndAccumulateJFAStat (AccumulateJFAStat.cpp:520-576; threaded :600-700): the same
 * posteriors accumulate per SESSION (N_h[n_sessions x C], F_h[n_sessions x C*D]) and per SPEAKER
 * (N[n_speakers x C], F[n_speakers x C*D]); segs[].row = session (JFATranslate::sessionNb),
 * speaker_of_session[session] = NDX line (locNb).  += like the reference's loop. */
lr_status lr_jfa_bwstats(lr_gmm *g, const float *X, size_t T, size_t ldx, const lr_seg *segs, size_t n_segs,
                         size_t n_sessions, const int32_t *speaker_of_session, size_t n_speakers, double *N_h,
                         double *F_h, double *N, double *F);
/* JFAAcc::normalizeFeatures (AccumulateJFAStat.cpp:4623-4680; called by substractUXfromFeatures :4689-4698 from
 * ComputeTestJFA, ComputeTest.cpp:455): every frame of the segments, IN PLACE in the host buffer X,
 *   x_t -= sum_k P(k | x_t) ux[k*D + i],   P under session_model (means M + U x, the world's weights / variances).
 * ux[C*D] = U x of the session.  Segments are visited in order; a frame covered by several segments is
 * compensated once per occurrence from its current value, like the reference's read-modify-write loop.
 * The reference's topGauss branch throws ("no topgauss yet"), so there is none here. */
lr_status lr_jfa_normalize_features(lr_gmm *session_model, const double *ux, float *X, size_t T, size_t ldx,
                                    const lr_seg *segs, size_t n_segs);

/* ---- a7: MixtureGDStat::computeAndAccumulateLLK (call sites ComputeTest.cpp:162-167,
 * TopGauss.cpp:166-192, AccumulateStat.cpp:77).
 * DETERMINE_TOP_DISTRIBS: idx[T*K] = the K most likely components in descending p_c order
 * (ties: lowest index), top_lk[T*K] their p_c (may be NULL), rest_lk[T] / rest_w[T] = sum of
 * p_c / weights outside the top K (TopGauss.cpp:181-192 snsl/snsw; may be NULL),
 * llk[T] = log(lk) clamped to [min_llk, max_llk] with lk = all components (complete != 0,
 * "COMPLETE") or the top K only ("PARTIAL").  Index selection is exact w.r.t. the fp64
 * likelihoods: fp32 candidates are re-evaluated in fp64 on the device. */
lr_status lr_gmm_llk_topk(lr_gmm *world, const float *X, size_t T, size_t ldx, int K,
                          int complete, double min_llk, double max_llk, double *llk,
                          uint32_t *idx, double *top_lk, double *rest_lk, double *rest_w);
/* USE_TOP_DISTRIBS on a client model: lk = sum_k w'_{idx} lk'_{idx}(x) (+ rest_lk if complete) */
lr_status lr_gmm_llk_use_topk(lr_gmm *client, const float *X, size_t T, size_t ldx, int K,
                              const uint32_t *idx, const double *rest_lk, int complete,
                              double min_llk, double max_llk, double *llk);
/* TOP_DISTRIBS_NO_ACTION: llk over all components (accumulateStatLLK) */
lr_status lr_gmm_llk(lr_gmm *g, const float *X, size_t T, size_t ldx, double min_llk,
                     double max_llk, double *llk);
/* The frame loop of ComputeTest() (ComputeTest.cpp:154-199) for one test file: world top-K
 * every frame (worldDecime 1; lr_compute_test_decime below for the general case), n_clients client models through USE_TOP_DISTRIBS;
 * mean_llk_world[n_segs_out], mean_llk_client[n_clients * n_segs_out]; per_segment != 0 =
 * segmentalMode (one mean per segment) else one mean over all segments (n_segs_out = 1). */
lr_status lr_compute_test(lr_gmm *world, lr_gmm *const *clients, int n_clients, const float *X,
                          size_t T, size_t ldx, const lr_seg *segs, size_t n_segs, int K,
                          int complete, double min_llk, double max_llk, int per_segment,
                          double *mean_llk_world, double *mean_llk_client);
/* The same loop with the reference's worldDecime (ComputeTest.cpp:111-113, 162-165): inside every segment
 * only frames idxFrame % world_decime == 0 run DETERMINE_TOP_DISTRIBS on the world; the others score the
 * world AND the clients through USE_TOP_DISTRIBS with the top list (and COMPLETE rest) of the last such frame. */
lr_status lr_compute_test_decime(lr_gmm *world, lr_gmm *const *clients, int n_clients, const float *X,
                                 size_t T, size_t ldx, const lr_seg *segs, size_t n_segs, int K,
                                 int complete, double min_llk, double max_llk, int per_segment,
                                 int world_decime, double *mean_llk_world, double *mean_llk_client);

/* ------------------------------------------------------------------ Total Variability -----
 * Device twin of the TVAcc object (AccumulateTVStat.h): owns _statN [U x C], _statF
 * [U x C*D], _T [R x C*D], _W [U x R], _TETt, _A [C x R*R], _Cmx [R x C*D], _R, _r, _meanW,
 * _ubm_means, _ubm_invvar in HBM (fp64). */
typedef struct lr_tv lr_tv;
lr_tv *lr_tv_create(int C, int D, int R, size_t U, const double *ubm_mean,
                    const double *ubm_invvar);
void lr_tv_destroy(lr_tv *tv);
lr_status lr_tv_set_stats(lr_tv *tv, const double *N, const double *F); /* loadN / loadF_X */
lr_status lr_tv_get_stats(lr_tv *tv, double *N, double *F);
/* device pointers to the statistics so lr_gmm_bwstats_dev can fill them in place */
double *lr_tv_dev_N(lr_tv *tv);
double *lr_tv_dev_F(lr_tv *tv);
lr_status lr_tv_set_T(lr_tv *tv, const double *T); /* loadT / initT result */
lr_status lr_tv_get_T(lr_tv *tv, double *T);
lr_status lr_tv_get_mean(lr_tv *tv, double *ubm_mean);
lr_status lr_tv_set_mean(lr_tv *tv, const double *ubm_mean); /* loadMeanEstimate :671 */
lr_status lr_tv_get_W(lr_tv *tv, double *W);
/* any output may be NULL; A[C x R*R], Cmx[R x C*D], Rm[R*R], r[R], meanW[R] */
lr_status lr_tv_get_acc(lr_tv *tv, double *A, double *Cmx, double *Rm, double *r, double *meanW);
lr_status lr_tv_reset_tmp_acc(lr_tv *tv);  /* resetTmpAcc: zero Cmx (A, R, r are zeroed by estep) */
lr_status lr_tv_subtract_m(lr_tv *tv);     /* substractM :1088-1105 */
lr_status lr_tv_estimate_tett(lr_tv *tv);  /* estimateTETt :766-805 */
lr_status lr_tv_estimate_w(lr_tv *tv);     /* estimateW :2103-2169 */
lr_status lr_tv_estimate_a_and_c(lr_tv *tv); /* estimateAandC :1691-1795 */
lr_status lr_tv_update_t(lr_tv *tv);       /* updateTestimate :974-1005 */
lr_status lr_tv_min_divergence(lr_tv *tv, double n_sessions); /* minDivergence :2056-2099 */
lr_status lr_tv_orthonormalize_t(lr_tv *tv); /* orthonormalizeT :1548-1596 */
/* ---- approximate i-vector extraction (IvExtractor --mode ubmWeight | eigenDecomposition,
 * IvExtractor.cpp:151-363; TotalVariability approximationMode outputs, TotalVariability.cpp:181-241) */
lr_status lr_tv_norm_t(lr_tv *tv);          /* normTMatrix :1600-1609: T[j,i] *= sqrt(invvar[i]) */
lr_status lr_tv_norm_statistics(lr_tv *tv); /* normStatistics :1215-1242: F = (F - mean N) sqrt(invvar) */
/* getWeightedCov :2826-2855: W[R x R] = sum_c weight[c] T_c T_c^T (T as currently held) */
lr_status lr_tv_weighted_cov(lr_tv *tv, const double *weight, double *W);
/* computeEigenProblem :2999-3052 (LAPACKE_dgeev on the symmetric W): eigenvalues sorted
 * descending, eigvec[n x rank] row-major with eigenvector j in COLUMN j, unit norm.  The reference
 * leaves the sign to LAPACK; here the component of largest magnitude is positive. */
lr_status lr_eigen_problem(int n, const double *EP, int rank, double *eigvec, double *eigval);
/* approximateTcTc :3106-3136: Dm[C x R], Dm[c,i] = diag(Q^T T_c T_c^T Q)_i.  Dm is overwritten
 * (the reference accumulates into a matrix its caller zeroed). */
lr_status lr_tv_approximate_tctc(lr_tv *tv, const double *Q, double *Dm);
/* estimateWUbmWeight :2337-2396: W_s = (I + (sum_c N[s,c]) Wcov)^-1 T F_s on the normalised T / F.
 * _W is overwritten. */
lr_status lr_tv_estimate_w_ubm_weight(lr_tv *tv, const double *Wcov);
/* estimateWEigenDecomposition :2556-2609: W_s += Q diag(1 / (1 + N_s Dm)) Q^T T F_s.  Like the
 * reference this ACCUMULATES into _W (zero after lr_tv_create). */
lr_status lr_tv_estimate_w_eigen_decomposition(lr_tv *tv, const double *Dm, const double *Q);

/* multi-GPU: device pointer + length (doubles) of the contiguous E-step accumulator block
 * [A | Cmx | Rm | r | sumW] that one NCCL all-reduce per EM iteration exchanges.  Inside the
 * block A is held as C packed lower triangles of R (R + 1) / 2 doubles (A_c is symmetric);
 * lr_tv_get_acc expands it to the reference's full [C x R*R]. */
double *lr_tv_dev_acc(lr_tv *tv);
size_t lr_tv_acc_len(const lr_tv *tv);
/* component-sharded M-step (updateTestimate is independent per component, :981-1000):
 * reduce-scatter the A part of the block by component (lr_tv_acc_a_stride doubles per component,
 * components contiguous), all-reduce the rest, lr_tv_update_t_range on the rank's components, then
 * all-gather the new columns of T through lr_tv_pack_t / lr_tv_unpack_t ([R x (c1 - c0) D] blocks). */
lr_status lr_tv_update_t_range(lr_tv *tv, int c0, int c1);
lr_status lr_tv_pack_t(lr_tv *tv, int c0, int c1, double *d_dst);
lr_status lr_tv_unpack_t(lr_tv *tv, int c0, int c1, const double *d_src);
size_t lr_tv_acc_a_stride(const lr_tv *tv);
/* after the all-reduce: meanW = sumW / n_speakers_total */
lr_status lr_tv_finish_estep(lr_tv *tv, double n_speakers_total);

/* ------------------------------------------------------------------ collectives (NCCL) ------
 * One process per GPU (lr_init(local rank) first).  The hot path has ONE exchange step per EM
 * iteration -- the sum of the ranks' sufficient statistics, the analogue of emAcc.addAccEM
 * (AccumulateStat.cpp:286-292) and of the mutex-guarded A / C updates (AccumulateTVStat.cpp:1920-1937);
 * everything else (BW statistics, ComputeTest, PLDA scoring) shards by NDX line / model row with no
 * collective.  NCCL is loaded at run time (dlopen), the library has no link-time dependency on it. */
#define LR_COMM_ID_BYTES 128
/* rank 0 creates the id (ncclGetUniqueId), distributes it by any means, every rank calls lr_comm_init */
lr_status lr_comm_unique_id(void *id128);
lr_status lr_comm_init(int rank, int world, const void *id128);
/* the same through a rendezvous file on a shared filesystem (rank 0 writes, the others wait) */
lr_status lr_comm_init_file(int rank, int world, const char *path);
int lr_comm_rank(void);
int lr_comm_world(void);
lr_status lr_comm_destroy(void);
/* in-place SUM all-reduce of n doubles, device resident, on the engine stream (no host sync) /
 * host buffer (H2D, all-reduce, D2H, sync) */
lr_status lr_allreduce_stats(double *d_buf, size_t n);
lr_status lr_allreduce_host(double *buf, size_t n);
/* dst[world * n] = the ranks' src[n] in rank order */
lr_status lr_allgather(const double *d_src, size_t n, double *d_dst);
lr_status lr_allgather_host(const double *src, size_t n, double *dst);
/* TotalVariability exchange + M-step after lr_tv_estimate_a_and_c on this rank's utterances:
 * reduce-scatter A by component, all-reduce [Cmx | R | r | sumW], meanW = sumW / total speakers,
 * updateTestimate on the rank's C / world components (:974-1005 is independent per component), all-gather
 * of the new T columns.  C % world != 0: one all-reduce + replicated M-step.  world == 1: finish + M-step. */
lr_status lr_tv_exchange_sharded(lr_tv *tv, double n_speakers_local, double *n_speakers_total);
lr_status lr_tv_dims(const lr_tv *tv, int *C, int *D, int *R);
/* The contraction kernel behind lr_tv_estimate_w / lr_tv_estimate_a_and_c (the scalar triple loops
 * AccumulateTVStat.cpp:2129-2137, :2146-2153, :1776-1782, :1784-1788), exposed for the parity tests:
 * C[M x N] = beta C + alpha A[M x K] B[N x K]^T, host buffers, row-major, fp64 in and out.  The product
 * runs on the INT8 tensor pipe: every operand row is scaled by a power of two and cut into `planes`
 * signed 7-bit digit planes (0 = the engine's setting, default 6); digit products accumulate exactly
 * in int32 and are recombined in fp64 (error ~ 2^-(7 planes) of row scale x column scale). */
lr_status lr_gemm_digits(size_t M, size_t N, size_t K, const double *A, const double *B, double *C,
                         double alpha, double beta, int planes);
/* Contraction kernel of the TV rows: which = 0 the INT8 digit GEMM (default), 1 cuBLAS fp64 (the
 * cross-check of the parity tests); planes = digit planes per operand (3..7, 0 = keep).  Takes effect
 * at the next lr_tv_estimate_tett. */
lr_status lr_set_tv_gemm(int which, int planes);

/* ------------------------------------------------------------------ PLDA scoring ----------
 * PldaTest::pldaNativeScoring + pldaScoring (PldaTools.cpp:4489-4519, 4175-4271) with
 * PldaModel::preComputation (:2950-2972) and rotateLeft (:3770-3790).  F[d x rF], G[d x rG]
 * (rG may be 0, G NULL), Sigma[d x d]; models[d x n_enrol], segments[d x n_test]: one
 * i-vector per COLUMN like the reference's _models/_segments; model_of[n_enrol] = model
 * number of each enrolment column (non-decreasing).  scores[n_models x n_test]. */
lr_status lr_plda_native_scoring(int d, int rF, int rG, const double *F, const double *G,
                                 const double *Sigma, const double *models, size_t n_enrol,
                                 const int32_t *model_of, size_t n_models,
                                 const double *segments, size_t n_test, double *scores);
/* The same scoring for trial matrices that do not fit / should not travel through host memory (1 M x 10 k
 * at configs[4]): d_models / d_segments are DEVICE pointers ([d x n] row-major as above), d_scores a
 * device fp32 [n_models x n_test] block with leading dimension ld_scores.  A rank scores its own shard of
 * the models (PldaTools.cpp:4302-4412 splits the models over threads the same way); no collective.
 * The trial matrix is computed in fp16 hi/lo split precision (22 bits) with fp32 accumulation: scores
 * agree with the fp64 loops to ~1e-6 of the largest score.  rF <= 256. */
lr_status lr_plda_native_scoring_dev(int d, int rF, int rG, const double *F, const double *G,
                                     const double *Sigma, const double *d_models, size_t n_enrol,
                                     const int32_t *model_of, size_t n_models, const double *d_segments,
                                     size_t n_test, float *d_scores, size_t ld_scores);

/* ------------------------------------------------------------------ i-vector back-end -----
 * Development-set statistics / normalisations of PldaDev and the non-PLDA scorings of PldaTest
 * (PldaTools.cpp; IvTest.cpp:112-391, IvNorm).  Vectors are COLUMNS of row-major [d x n] matrices
 * like the reference's _data / _models / _segments.  class_of[n] = speaker index of each session
 * (the reference's _class), n_spk speakers. */
/* PldaDev::computeAll :353-385 + computeCovMat :516-571: global / speaker means, total (Sigma),
 * within-class (W) and between-class (B) covariance, all divided by n.  Any output may be NULL. */
lr_status lr_iv_cov_mat(int d, size_t n, const double *data, const int32_t *class_of, size_t n_spk,
                        double *mean, double *spk_means, double *Sigma, double *W, double *B);
/* PldaDev::computeWccnChol :1113-1175: WCCN = upperCholesky((mean over speakers of cov_spk)^-1) */
lr_status lr_iv_wccn_chol(int d, size_t n, const double *data, const int32_t *class_of, size_t n_spk,
                          double *WCCN);
/* PldaDev::computeMahalanobis :1366-1378: M = W^-1 */
lr_status lr_iv_mahalanobis_matrix(int d, size_t n, const double *data, const int32_t *class_of,
                                   size_t n_spk, double *M);
/* one sphericalNuisanceNormalization iteration's matrix :1853-1900: mat = (V diag(1/sqrt(lambda)))^T
 * for cov = Sigma (mode EFR) or W (mode sphNorm); eigenvalues descending, sign: largest component > 0 */
lr_status lr_iv_efr_matrix(int d, const double *cov, double *mat);
/* PldaDev::computeLDA :1381-1415: rows of ldaMat[rank x d] = leading unit-norm eigenvectors of
 * W^-1 B (dgeev in the reference; the symmetric-definite pencil B v = lambda W v here) */
lr_status lr_iv_lda(int d, const double *W, const double *B, int rank, double *ldaMat);
/* center (mu, may be NULL) -> rotateLeft (M[r x d], may be NULL) -> lengthNorm (if length_norm):
 * PldaTools.cpp:466-474, 498-514, 436-464 (PldaTest: :3754-3790, :3706-3750) -- i.e. one
 * applySphericalNuisanceNormalization iteration (:1931-1975), an LDA / WCCN rotation, or any
 * prefix of it.  out[(M ? r : d) x n]. */
lr_status lr_iv_normalize(int d, size_t n, const double *data, const double *mu, const double *M, int r,
                          int length_norm, double *out);
/* PldaTest::cosineDistance :3842-3880, mahalanobisDistance :3882-3910, twoCovScoring :4083-4173.
 * models[d x n_models], segments[d x n_test], scores[n_models x n_test]; trials[n_models x n_test]
 * (bytes, may be NULL = all): pairs outside the mask score 0 like the reference's untouched _scores. */
lr_status lr_iv_cosine_scoring(int d, size_t n_models, size_t n_test, const double *models,
                               const double *segments, const uint8_t *trials, double *scores);
lr_status lr_iv_mahalanobis_scoring(int d, size_t n_models, size_t n_test, const double *models,
                                    const double *segments, const double *Mah, const uint8_t *trials,
                                    double *scores);
lr_status lr_iv_two_cov_scoring(int d, size_t n_models, size_t n_test, const double *models,
                                const double *segments, const double *W, const double *B,
                                double *scores);

/* PLDA training: one PldaModel::em_iteration (PldaTools.cpp:2329-2343) -- _Dev.center(_Delta),
 * computeCovMatEigen (:931-950), getExpectedValues (:2359-2485), mStep with minimum divergence
 * (:2790-2813).  data[d x n] (vectors in columns, sessions of a speaker adjacent, class_of
 * non-decreasing) is centred by Delta IN PLACE like the reference's _Dev; F[d x rF], G[d x rG]
 * (rG may be 0, G NULL), Sigma[d x d], Delta[d] are updated in place. */
lr_status lr_plda_em_iteration(int d, int rF, int rG, size_t n, double *data, const int32_t *class_of,
                               size_t n_spk, double *F, double *G, double *Sigma, double *Delta);

#ifdef __cplusplus
}
#endif
#endif
