// gemm_i8.cuh -- fp64-grade GEMM on the INT8 tensor pipe (gemm_i8.cu):
//   C[m, n] = beta C[m, n] + alpha * sum_k A[m, k] B[n, k]
// fp64 operands are cut row by row into `s` signed 7-bit digits (int8 panels); every digit product is
// an exact int32 UMMA accumulation in TMEM, the digit classes are recombined in fp64.
#pragma once
#include "common.cuh"

namespace lr {

constexpr int kI8MaxSlices = 7;  // 64 s accumulator columns + A plane slots <= 512 TMEM columns
constexpr int kI8TileM = 128, kI8TileN = 64;

// bytes of the digit panels of a [rows x K] operand cut in tiles of tile_rows (128: A side, 64: B side)
size_t gemm_i8_panel_bytes(long rows, long K, int s, int tile_rows);
// number of row scales (doubles) the operand needs
size_t gemm_i8_scale_count(long rows, int tile_rows);
// dX: element (r, k) at dX[r * stride_row + k * stride_k], one of the strides == 1.
// d_panels: gemm_i8_panel_bytes; d_scale: gemm_i8_scale_count doubles (power-of-two row scales)
// d_rowmax (optional): max |x| of every row as the bit pattern of the double, when the caller has it
lr_status gemm_i8_prepare(const double *dX, size_t stride_row, size_t stride_k, long rows, long K, int s,
                          int tile_rows, unsigned char *d_panels, double *d_scale,
                          const unsigned long long *d_rowmax = nullptr);
// C (device, row-major, leading dimension ldc); beta must be 0 or 1 when the K range is split
lr_status gemm_i8_run(const unsigned char *dAp, const double *dAscale, long M, const unsigned char *dBp,
                      const double *dBscale, long N, long K, int s, double alpha, double beta, double *dC,
                      size_t ldc);

}  // namespace lr
