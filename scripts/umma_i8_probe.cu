// umma_i8_probe.cu -- clocks per tcgen05.mma for the shapes the digit GEMM could use (one CTA per SM,
// operands resident, nothing but UMMA issue + one commit per batch).  nvcc -arch=sm_100a.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../lia_ral_b200/csrc/tc_ptx.cuh"
using namespace lr::tcptx;

__device__ __forceinline__ void mma_i8_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], %1, %2, %3, p;\n}\n" ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_i8_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n.reg .pred p;\nsetp.ne.b32 p, %4, 0;\ntcgen05.mma.cta_group::1.kind::i8 [%0], [%1], %2, %3, p;\n}\n" ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_i8(int M, int N) {
  return (2u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// mode: 0 SS i8, 1 TS i8, 2 SS f16, 3 TS f16
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) k_probe(int iters, long long *out) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 96 * 1024, slot = bar + 8;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 96 * 1024 / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0x01010101u));
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  long long t0 = 0, t1 = 0;
  if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t id = (MODE < 2) ? idesc_i8(128, N) : make_idesc(128, N, 0, 0);
    const uint64_t adesc = make_desc(base, 16, 1024);             // 16 KB A plane
    const uint64_t bdesc = make_desc(base + 32 * 1024, 16, 1024);  // up to 32 KB B plane
    t0 = clock64();
    if (leader) {
      for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
          if (MODE == 0 || MODE == 2) {
            if (MODE == 0) mma_i8_ss(tmem_base, desc_add(adesc, kk * 32), desc_add(bdesc, kk * 32), id, 1u);
            else umma_ss(tmem_base, desc_add(adesc, kk * 32), desc_add(bdesc, kk * 32), id, 1u);
          } else {
            if (MODE == 1) mma_i8_ts(tmem_base, tmem_base + 256 + kk * 8, desc_add(bdesc, kk * 32), id, 1u);
            else umma_ts(tmem_base, tmem_base + 256 + kk * 8, desc_add(bdesc, kk * 32), id, 1u);
          }
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    t1 = clock64();
    if (leader) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

// the digit GEMM's issue pattern (6 planes, 21 products x 4 UMMAs per chunk) with nothing else going on:
// PAT 0 = as the kernel issues (A slot / B plane / class accumulator rotate), 1 = same accumulator for
// every UMMA, 2 = same A and B for every UMMA but rotating accumulators, 3 = TS (A from TMEM), rotating
template <int PAT>
__global__ void __launch_bounds__(128, 1) k_pattern(int chunks, long long *out) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t bar = base + 176 * 1024, slot = bar + 8;
  const int warp = threadIdx.x >> 5;
  for (uint32_t i = threadIdx.x; i < 176 * 1024 / 4; i += blockDim.x)
    asm volatile("st.shared.b32 [%0], %1;" ::"r"(base + i * 4), "r"(0x01010101u));
  if (threadIdx.x == 0) {
    mbar_init(bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(slot));
  if (warp == 1) {
    const bool leader = elect_one();
    constexpr uint32_t id = idesc_i8(128, 64);
    constexpr uint32_t desc_hi = (1024u >> 4) | (1u << 14) | (2u << 29);
    const uint32_t a_lo0 = ((base >> 4) & 0x3FFFu) | (1u << 16);
    const uint32_t b_lo0 = (((base + 80 * 1024) >> 4) & 0x3FFFu) | (1u << 16);
    long long t0 = clock64();
    if (leader) {
      long aseq = 0;
      for (int c = 0; c < chunks; c++) {
        const uint32_t b_lo = b_lo0 + (uint32_t)(c & 1) * (48 * 1024 >> 4);
#pragma unroll
        for (int ii = 0; ii < 6; ii++, aseq++) {
          const int i = (ii & 1) ? 5 - (ii >> 1) : (ii >> 1);
          const uint32_t a_lo = a_lo0 + (uint32_t)(aseq % 5) * (16384 >> 4);
#pragma unroll
          for (int j = 0; j < 6 - i; j++)
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
              const uint64_t ad = ((uint64_t)desc_hi << 32) | (uint64_t)((PAT == 2 ? a_lo0 : a_lo) + kk * 2);
              const uint64_t bd = ((uint64_t)desc_hi << 32) | (uint64_t)((PAT == 2 ? b_lo0 : b_lo + j * (8192 >> 4)) + kk * 2);
              const uint32_t d = tmem_base + (PAT == 1 ? 0 : (i + j) * 64);
              if (PAT == 3) mma_i8_ts(d, tmem_base + 384 + (uint32_t)(aseq & 3) * 32 + kk * 8, bd, id, 1u);
              else mma_i8_ss(d, ad, bd, id, 1u);
            }
        }
      }
      umma_commit(bar);
    }
    __syncwarp();
    mbar_wait(bar, 0);
    long long t1 = clock64();
    if (leader) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

template <int PAT>
void run_pattern(const char *name, long long *d_out) {
  const int chunks = 256, smem = 176 * 1024 + 2048;
  cudaFuncSetAttribute(k_pattern<PAT>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_pattern<PAT><<<148, 128, smem>>>(8, d_out);
  cudaDeviceSynchronize();
  k_pattern<PAT><<<148, 128, smem>>>(chunks, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; i++) avg += (double)h[i];
  avg /= 148.0 * chunks * 84;
  printf("pattern %-34s %7.1f clk/UMMA (%s)\n", name, avg, cudaGetErrorString(e));
}

template <int MODE, int N>
void run(const char *name, long long *d_out) {
  const int iters = 4096, smem = 96 * 1024 + 2048;
  cudaFuncSetAttribute(k_probe<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_probe<MODE, N><<<148, 128, smem>>>(64, d_out);
  cudaDeviceSynchronize();
  k_probe<MODE, N><<<148, 128, smem>>>(iters, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[148];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  double avg = 0;
  for (int i = 0; i < 148; i++) avg += (double)h[i];
  avg /= 148.0 * iters * 4;
  const double macs = 128.0 * N * ((MODE < 2) ? 32 : 16);
  printf("%-18s N=%3d  %7.1f clk/UMMA  %7.0f MAC/clk/SM  (%s)\n", name, N, avg, macs / avg, cudaGetErrorString(e));
}

int main() {
  long long *d_out;
  cudaMalloc(&d_out, 148 * sizeof(long long));
  run_pattern<0>("kernel order (SS)", d_out);
  run_pattern<1>("one accumulator (SS)", d_out);
  run_pattern<2>("one A, one B, rotating acc (SS)", d_out);
  run_pattern<3>("kernel order, A in TMEM (TS)", d_out);
  run<0, 64>("i8 SS", d_out);
  run<0, 128>("i8 SS", d_out);
  run<0, 256>("i8 SS", d_out);
  run<1, 64>("i8 TS", d_out);
  run<1, 128>("i8 TS", d_out);
  run<1, 256>("i8 TS", d_out);
  run<2, 64>("f16 SS", d_out);
  run<2, 128>("f16 SS", d_out);
  run<2, 256>("f16 SS", d_out);
  run<3, 64>("f16 TS", d_out);
  run<3, 128>("f16 TS", d_out);
  run<3, 256>("f16 TS", d_out);
  return 0;
}
