// gmm_tc.cu -- tcgen05 / TMEM implementation of the frames x components passes (sm_100a).
#include "gmm.cuh"

namespace lr {

bool tc_supported(const lr_gmm *g) {
  (void)g;
  return false;
}
lr_status tc_derive(lr_gmm *g) {
  (void)g;
  return LR_OK;
}
lr_status tc_pass_lse(lr_gmm *, const FrameList &, float *, double *) {
  return fail(LR_ERR_ARG, "tcgen05 GMM pass not built");
}
lr_status tc_pass_acc(lr_gmm *, const FrameList &, const float *, const LrChunk *, int, double,
                      double *, double *, double *) {
  return fail(LR_ERR_ARG, "tcgen05 GMM pass not built");
}

}  // namespace lr
