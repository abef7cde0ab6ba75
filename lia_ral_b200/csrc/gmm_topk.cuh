// gmm_topk.cuh -- top-distribution selection / rescoring passes (see gmm_topk.cu)
#pragma once
#include "gmm.cuh"

namespace lr {

constexpr int kMaxCand = 128;  // nominated candidates per frame (K <= kMaxTopK)
constexpr int kMaxTopK = 96;

lr_status gmm_topk(lr_gmm *g, const FrameList &fl, const float *d_S, int K, int complete,
                   double min_llk, double max_llk, double *d_llk, unsigned *d_idx,
                   double *d_top_lk, double *d_rest_lk, double *d_rest_w);
lr_status gmm_use_topk(lr_gmm *g, const FrameList &fl, int K, const unsigned *d_idx,
                       const double *d_rest_lk, int complete, double min_llk, double max_llk,
                       double *d_llk);
lr_status gmm_llk_from_lse(long P, const float *d_lse2, double min_llk, double max_llk,
                           double *d_llk);

}  // namespace lr
