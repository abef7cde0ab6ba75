// gemm_split.cuh -- tcgen05 split-precision skinny-K GEMM (gemm_split.cu):
//   C[m, n] = sum_k A[m, k] B[n, k] + row_term[m] + col_term[n],  K <= 256, fp64 operands -> fp16 hi/lo
#pragma once
#include "common.cuh"

namespace lr {

// bytes of the operand panels of a [rows x K] matrix
size_t gemm_split_panel_bytes(long rows, int K);
// dX [rows x K] fp64 row-major (device, leading dimension ld) -> scaled fp16 hi/lo panels;
// d_tmp: one device double of scratch; *scale_out = the power-of-two scale applied
lr_status gemm_split_prepare(const double *dX, size_t ld, long rows, int K, unsigned char *d_panels,
                             double *d_tmp, double *scale_out);
// C (device, row-major, leading dimension ldc, float or double) from prepared panels; d_row_term[M],
// d_col_term[N] may be null
template <typename OutT>
lr_status gemm_split_run(const unsigned char *dAp, double scaleA, long M, const unsigned char *dBp,
                         double scaleB, long N, int K, OutT *dC, size_t ldc, const double *d_row_term,
                         const double *d_col_term);

}  // namespace lr
