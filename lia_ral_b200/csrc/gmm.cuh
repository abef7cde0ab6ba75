// gmm.cuh -- device-resident diagonal GMM (the MixtureGD twin) and the shared launch plumbing
// of the frames x components passes.
#pragma once
#include "common.cuh"

// Work chunk of the accumulate pass: a run of positions of the frame list feeding one
// statistics row.
struct LrChunk {
  long long pos;  // first position in the frame list
  int len;        // number of positions (<= kChunkFrames)
  int row;        // statistics row (NDX line); 0 for EM
};

struct lr_gmm {
  int C = 0, D = 0;
  int Cp = 0;  // C rounded up to a multiple of 128 (padding components carry const = -1e30)
  // fp64 model (authoritative copy lives on the device so the M-step never leaves HBM)
  double *d_w = nullptr, *d_mean = nullptr, *d_cov = nullptr;
  double *d_covinv = nullptr, *d_cst = nullptr, *d_det = nullptr;
  double *d_tot = nullptr;  // scratch scalar (sum of occupations)
  bool cst_override = false;
  // fp32 operands of the SIMT pass, log2 domain:
  //   S2[t,c] = const2[c] - sum_i (x_i * sa[i][c] + nm[i][c])^2
  //   sa = sqrt(0.5 log2(e) covinv), nm = -mean * sa, const2 = log2(w cst)
  float *d_sa = nullptr;     // [D][Cp]
  float *d_nm = nullptr;     // [D][Cp]
  float *d_const2 = nullptr; // [Cp]
  float *d_mean_f = nullptr; // [Cp][64] fp32 means, row per component, col 60.. = 0
  // operands of the tcgen05 pass (gmm_tc.cu); see tc_derive
  void *d_tc_w = nullptr;    // fp16 [2 (hi,lo)][2 (k panel)][Cp][64]
  double *d_g = nullptr, *d_s = nullptr;   // per-dimension shift / scale of the normalised space
  float *d_gf = nullptr, *d_rsf = nullptr; // fp32 shift, 1/scale
};

struct lr_feats {
  const float *d_x = nullptr;
  size_t T = 0, ldx = 0;
  int D = 0;
  bool owned = false;
  // tensor-core frame operand of the range [conv_t0, conv_t0 + conv_T) (k_tc_convert output, 512 B per
  // frame), kept between EM iterations over resident frames: it depends on the frames and on the
  // normalisation (conv_norm) only, and tc_derive keeps the normalisation while the mixture's global
  // mean / deviation stay close to it.  mutable: filled behind the const handle of the _dev entry points.
  mutable unsigned char *d_conv = nullptr;
  mutable size_t conv_cap = 0;
  mutable size_t conv_t0 = 0, conv_T = 0;
  mutable unsigned long long conv_norm = 0;  // 0 = nothing cached
};

namespace lr {

constexpr int kMaxD = 63;          // value tile of the accumulate pass holds [x | 1] in 64 columns
constexpr int kChunkFrames = 512;  // positions per accumulate chunk (fp32 partial sums, then fp64)

// (re)derive every kernel operand from d_w / d_mean / d_cov; enqueued on the engine stream.
lr_status gmm_derive(lr_gmm *g);

// A frame list: positions 0..P-1 map to frames of the block dX[.. x ldx]; d_index == nullptr
// means position p is frame p.
struct FrameList {
  const float *dX = nullptr;
  size_t ldx = 0;
  const unsigned *d_index = nullptr;
  long P = 0;
};

//   lse2[p]  = log2 sum_c w_c lk_c(x_p)                          (pass 1)
//   d_S      = optional [P x Cp] log2 joint likelihoods           (pass 1, top-K path)
//   llk_sum  = optional device double, += ln(2) * sum_p lse2[p]
lr_status gmm_pass_lse(lr_gmm *g, const FrameList &fl, float *d_lse2, float *d_S,
                       double *d_llk_sum);
//   accumulate pass: per chunk, out_N[row*C + c] += fw * sum gamma, out_F[(row*C+c)*D+i] +=
//   fw * sum gamma x_i (and out_S2 likewise with x_i^2 when non-null).
lr_status gmm_pass_acc(lr_gmm *g, const FrameList &fl, const float *d_lse2,
                       const LrChunk *d_chunks, int n_chunks, double fw, double *out_N,
                       double *out_F, double *out_S2);

// tcgen05 implementations of the same two passes (gmm_tc.cu)
bool tc_supported(const lr_gmm *g);
lr_status tc_derive(lr_gmm *g);
// d_S (optional): [P x Cp] fp32 log2 joint likelihoods, the scores the top-K path nominates on
lr_status tc_pass_lse(lr_gmm *g, const FrameList &fl, float *d_lse2, double *d_llk_sum, float *d_S = nullptr);
// likelihood + statistics over a tile-padded frame list (chunks tile aligned, host copy)
// conv != nullptr: the converted operand of this frame list lives there (n_tiles x 64 KB); it is
// (re)written unless conv_valid
lr_status tc_run_stats(lr_gmm *g, const FrameList &fl, const std::vector<LrChunk> &chunks,
                       double fw, double *out_N, double *out_F, double *out_S2,
                       double *d_llk_sum, unsigned char *conv = nullptr, bool conv_valid = false);
// identifier of the normalisation the model's tensor-core operands are expressed in (0: none)
unsigned long long tc_norm_id(const lr_gmm *g);
constexpr size_t kTcTileBytesPub = 65536;  // converted operand per 128 frames
void tc_free(lr_gmm *g);
// true when the tensor-core path serves this model under the current lr_set_gmm_kernel choice;
// sets *err when the choice is "tcgen05" but the model cannot be served
bool tc_selected(const lr_gmm *g, lr_status *err);
constexpr unsigned kPadFrame = 0xFFFFFFFFu;  // padding entry of a tile-padded frame list

}  // namespace lr
