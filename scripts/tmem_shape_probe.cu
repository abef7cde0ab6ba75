// tmem_shape_probe.cu -- prints which (lane, column) every register of the 16x256b.x8 / 16x128b.x8
// tcgen05.ld shapes receives, by writing lane * 1000 + column with the 32x32b shape first.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -o scripts/tmem_shape_probe scripts/tmem_shape_probe.cu
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

__global__ void __launch_bounds__(128, 1) k_probe(uint32_t *out256, uint32_t *out128, uint32_t *out_st) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&slot)), "r"(128u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = slot;
  const uint32_t lane_addr = (uint32_t)(warp * 32) << 16;
  // write lane * 1000 + col for 64 columns with 32x32b.x1 stores
  for (int c = 0; c < 64; c++) {
    uint32_t v = (uint32_t)((warp * 32 + lane) * 1000 + c);
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(base + lane_addr + c), "r"(v) : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  uint32_t r[32];
  for (int half = 0; half < 2; half++) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x256b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(base + lane_addr + ((uint32_t)(16 * half) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 32; i++) out256[((warp * 2 + half) * 32 + lane) * 32 + i] = r[i];
  }
  uint32_t s[16];
  for (int half = 0; half < 2; half++) {
    asm volatile(
        "tcgen05.ld.sync.aligned.16x128b.x8.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7]), "=r"(s[8]),
          "=r"(s[9]), "=r"(s[10]), "=r"(s[11]), "=r"(s[12]), "=r"(s[13]), "=r"(s[14]), "=r"(s[15])
        : "r"(base + lane_addr + ((uint32_t)(16 * half) << 16)));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    for (int i = 0; i < 16; i++) out128[((warp * 2 + half) * 32 + lane) * 16 + i] = s[i];
  }
  __syncthreads();
  // store with 16x128b.x8 into columns 64..95 the value (expected lane) * 1000 + (expected col), read back 32x32b
  for (int half = 0; half < 2; half++) {
    for (int j = 0; j < 8; j++)
      for (int rs = 0; rs < 2; rs++)
        s[2 * j + rs] = (uint32_t)((warp * 32 + 16 * half + 8 * rs + lane / 4) * 1000 + 4 * j + (lane & 3));
    asm volatile(
        "tcgen05.st.sync.aligned.16x128b.x8.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(base + lane_addr + ((uint32_t)(16 * half) << 16) + 64),
        "r"(s[0]), "r"(s[1]), "r"(s[2]), "r"(s[3]), "r"(s[4]), "r"(s[5]), "r"(s[6]), "r"(s[7]), "r"(s[8]), "r"(s[9]),
        "r"(s[10]), "r"(s[11]), "r"(s[12]), "r"(s[13]), "r"(s[14]), "r"(s[15])
        : "memory");
  }
  asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
  __syncthreads();
  for (int c = 0; c < 32; c++) {
    uint32_t v;
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v) : "r"(base + lane_addr + 64 + c) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
    out_st[(warp * 32 + lane) * 32 + c] = v;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(128u) : "memory");
}

int main() {
  uint32_t *d256, *d128, *dst;
  cudaMalloc(&d256, 8 * 32 * 32 * 4);
  cudaMalloc(&d128, 8 * 32 * 16 * 4);
  cudaMalloc(&dst, 128 * 32 * 4);
  k_probe<<<1, 128>>>(d256, d128, dst);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    std::printf("probe failed: %s\n", cudaGetErrorString(e));
    return 1;
  }
  static uint32_t h256[8 * 32 * 32], h128[8 * 32 * 16], hst[128 * 32];
  cudaMemcpy(h256, d256, sizeof(h256), cudaMemcpyDeviceToHost);
  cudaMemcpy(h128, d128, sizeof(h128), cudaMemcpyDeviceToHost);
  cudaMemcpy(hst, dst, sizeof(hst), cudaMemcpyDeviceToHost);
  int bad256 = 0, bad128 = 0, badst = 0;
  for (int w = 0; w < 4; w++)
    for (int half = 0; half < 2; half++)
      for (int t = 0; t < 32; t++) {
        for (int i = 0; i < 32; i++) {
          const int j = i / 4, rs = (i / 2) & 1, e2 = i & 1;
          const uint32_t want = (uint32_t)((w * 32 + 16 * half + 8 * rs + t / 4) * 1000 + 8 * j + 2 * (t & 3) + e2);
          if (h256[((w * 2 + half) * 32 + t) * 32 + i] != want) bad256++;
        }
        for (int i = 0; i < 16; i++) {
          const int j = i / 2, rs = i & 1;
          const uint32_t want = (uint32_t)((w * 32 + 16 * half + 8 * rs + t / 4) * 1000 + 4 * j + (t & 3));
          if (h128[((w * 2 + half) * 32 + t) * 16 + i] != want) bad128++;
        }
      }
  for (int l = 0; l < 128; l++)
    for (int c = 0; c < 32; c++)
      if (hst[l * 32 + c] != (uint32_t)(l * 1000 + c)) badst++;
  std::printf("16x256b.x8 ld mismatches %d, 16x128b.x8 ld mismatches %d, 16x128b.x8 st mismatches %d\n", bad256, bad128, badst);
  if (bad256 || bad128 || badst) {
    std::printf("warp 0 half 0, threads 0..7, 16x256b regs 0..7 (lane*1000+col):\n");
    for (int t = 0; t < 8; t++) {
      for (int i = 0; i < 8; i++) std::printf(" %6u", h256[t * 32 + i]);
      std::printf("\n");
    }
    std::printf("16x128b regs 0..3:\n");
    for (int t = 0; t < 8; t++) {
      for (int i = 0; i < 4; i++) std::printf(" %6u", h128[t * 16 + i]);
      std::printf("\n");
    }
    std::printf("st readback lane 0..3 cols 0..7:\n");
    for (int l = 0; l < 4; l++) {
      for (int c = 0; c < 8; c++) std::printf(" %6u", hst[l * 32 + c]);
      std::printf("\n");
    }
  }
  return 0;
}
