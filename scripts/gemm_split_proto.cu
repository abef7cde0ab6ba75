// gemm_split_proto.cu -- stand-alone prototype of the split-precision tcgen05 GEMM planned for the
// TV rows: C[M x N] = A[M x K] B[N x K]^T with both operands split into fp16
// hi + lo (22 significand bits), three UMMA products (hi hi + lo hi + hi lo), fp32 accumulation in
// TMEM -- the arithmetic of csrc/gmm_tc.cu's likelihood GEMM generalised to an arbitrary K, e.g.
// L[u, p] = sum_c N[u, c] TETt[c, p] with A = N and B = TETt^T.
//
// Correctness prototype: one 128 x 128 output tile per CTA, single-buffered 64-wide K panels
// (4 x 16 KB swizzled panels per step, cp.async.bulk), one elected thread issues the UMMAs, four
// warps read the accumulator back.  Self-checking against an fp64 host GEMM.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/gemm_split_proto scripts/gemm_split_proto.cu
//   ./scripts/gemm_split_proto [M N K]
//
// Compile-checked in round 1 (no GPU budget left to run it); not part of the product library.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kPanelBytes = 128 * 128;  // 128 rows x 64 fp16, 128-byte swizzle

// ---- PTX wrappers (the validated set of csrc/gmm_tc.cu) -------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "LAB_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra LAB_WAIT;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst), "r"(ncols) : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void umma_ss(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                        uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n"
      ".reg .b32 rx;\n"
      ".reg .pred px;\n"
      "elect.sync rx|px, 0xffffffff;\n"
      "selp.b32 %0, 1, 0, px;\n"
      "}\n"
      : "=r"(pred));
  return pred != 0;
}
// shared-memory matrix descriptor: K-major, 128-byte swizzle, version 1 (sm_100)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr >> 4) & 0x3FFF);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}
__device__ __forceinline__ uint64_t desc_add(uint64_t d, uint32_t bytes) { return d + (uint64_t)(bytes >> 4); }
__host__ __device__ constexpr uint32_t make_idesc(int M, int N) {  // fp16 x fp16 -> fp32, both operands K-major
  return (1u << 4) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__host__ __device__ __forceinline__ uint32_t panel_off(int row, int col) {
  return (uint32_t)row * 128u + (uint32_t)((((col >> 3) ^ (row & 7)) << 4) | ((col & 7) << 1));
}

// ---- operand packing: row-major fp64 [rows x K] (ld) -> hi / lo swizzled panels ----------------
// layout: panel (row block rb, k panel kp) at ((rb * n_kp + kp) * 2 + {0: hi, 1: lo}) * 16 KB
__global__ void k_pack(const double *__restrict__ src, int rows, int K, int ld, int n_kp, unsigned char *__restrict__ out) {
  long e = (long)blockIdx.x * blockDim.x + threadIdx.x;
  const long rows_pad = (long)((rows + 127) / 128) * 128, K_pad = (long)n_kp * 64;
  if (e >= rows_pad * K_pad) return;
  const int row = (int)(e / K_pad), k = (int)(e - (long)row * K_pad);
  const double v = (row < rows && k < K) ? src[(size_t)row * ld + k] : 0.0;
  const __half hi = __float2half_rn((float)v);
  const __half lo = __float2half_rn((float)(v - (double)__half2float(hi)));
  unsigned char *p = out + ((size_t)(row / 128) * n_kp + k / 64) * 2 * kPanelBytes;
  const uint32_t off = panel_off(row % 128, k % 64);
  *reinterpret_cast<__half *>(p + off) = hi;
  *reinterpret_cast<__half *>(p + kPanelBytes + off) = lo;
}

// ---- the GEMM: one CTA per 128 x 128 tile of C ---------------------------------------------------
constexpr size_t kSmem = 1024 + 4 * kPanelBytes + 64;
__global__ void __launch_bounds__(128, 1)
k_gemm_split(const unsigned char *__restrict__ Ap, const unsigned char *__restrict__ Bp, int n_kp, int M, int N,
             float *__restrict__ Cm /*[M x N] row-major*/) {
  extern __shared__ unsigned char smem_raw[];
  const uint32_t base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  const uint32_t s_a_hi = base, s_a_lo = base + kPanelBytes, s_b_hi = base + 2 * kPanelBytes, s_b_lo = base + 3 * kPanelBytes;
  const uint32_t bar_full = base + 4 * kPanelBytes, bar_mma = bar_full + 8, bar_done = bar_full + 16,
                 tmem_slot = bar_full + 24;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int mb = blockIdx.x, nb = blockIdx.y;
  if (threadIdx.x == 0) {
    mbar_init(bar_full, 1);
    mbar_init(bar_mma, 1);
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) tmem_alloc(tmem_slot, 128);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(tmem_base) : "r"(tmem_slot));

  if (warp == 0) {
    const bool leader = elect_one();
    constexpr uint32_t idesc = make_idesc(128, 128);
    const uint64_t a_hi = make_desc(s_a_hi, 16, 1024), a_lo = make_desc(s_a_lo, 16, 1024);
    const uint64_t b_hi = make_desc(s_b_hi, 16, 1024), b_lo = make_desc(s_b_lo, 16, 1024);
    for (int kp = 0; kp < n_kp; kp++) {
      if (kp > 0) mbar_wait(bar_mma, (kp - 1) & 1);  // the previous panels were consumed
      if (leader) {
        mbar_expect_tx(bar_full, 4 * kPanelBytes);
        const unsigned char *a = Ap + ((size_t)mb * n_kp + kp) * 2 * kPanelBytes;
        const unsigned char *b = Bp + ((size_t)nb * n_kp + kp) * 2 * kPanelBytes;
        bulk_g2s(s_a_hi, a, kPanelBytes, bar_full);
        bulk_g2s(s_a_lo, a + kPanelBytes, kPanelBytes, bar_full);
        bulk_g2s(s_b_hi, b, kPanelBytes, bar_full);
        bulk_g2s(s_b_lo, b + kPanelBytes, kPanelBytes, bar_full);
      }
      __syncwarp();
      mbar_wait(bar_full, kp & 1);
      tc_fence_after();
      if (leader) {
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {  // hi hi + lo hi + hi lo, K = 16 per UMMA
          const uint32_t o = kk * 32;
          umma_ss(tmem_base, desc_add(a_hi, o), desc_add(b_hi, o), idesc, (kp > 0 || kk > 0) ? 1u : 0u);
          umma_ss(tmem_base, desc_add(a_lo, o), desc_add(b_hi, o), idesc, 1u);
          umma_ss(tmem_base, desc_add(a_hi, o), desc_add(b_lo, o), idesc, 1u);
        }
        umma_commit(bar_mma);
        if (kp == n_kp - 1) umma_commit(bar_done);  // tracks every UMMA issued so far
      }
      __syncwarp();
    }
  }
  // every warp waits for the final commit (a barrier of its own: a parity wait on bar_mma would be
  // satisfied by the FIRST panel already), then reads its 32 lanes (rows of the tile)
  mbar_wait(bar_done, 0);
  tc_fence_after();
  const int row = mb * 128 + warp * 32 + lane;
#pragma unroll 1
  for (int ch = 0; ch < 4; ch++) {
    uint32_t r[32];
    tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + ch * 32, r);
    tmem_wait_ld();
    if (row < M) {
#pragma unroll
      for (int e = 0; e < 32; e++) {
        const int col = nb * 128 + ch * 32 + e;
        if (col < N) Cm[(size_t)row * N + col] = __uint_as_float(r[e]);
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 128);
}

#define CK(x)                                                                         \
  do {                                                                                \
    cudaError_t e__ = (x);                                                            \
    if (e__ != cudaSuccess) {                                                         \
      std::printf("%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e__)); \
      return 1;                                                                       \
    }                                                                                 \
  } while (0)

int main(int argc, char **argv) {
  const int M = argc > 3 ? std::atoi(argv[1]) : 300, N = argc > 3 ? std::atoi(argv[2]) : 1000,
            K = argc > 3 ? std::atoi(argv[3]) : 2048;
  const int n_kp = (K + 63) / 64, mbs = (M + 127) / 128, nbs = (N + 127) / 128;
  std::vector<double> A((size_t)M * K), B((size_t)N * K);
  uint64_t s = 88172645463325252ull;
  auto rnd = [&]() {
    s ^= s << 13;
    s ^= s >> 7;
    s ^= s << 17;
    return (double)(s >> 11) / 9007199254740992.0;
  };
  for (auto &v : A) v = rnd() < 0.9 ? 0.0 : rnd() * 300.0;  // occupancies: sparse, non-negative
  for (auto &v : B) v = (rnd() - 0.5) * 2.0;                // TETt entries scaled into [-1, 1]
  double *dA, *dB;
  unsigned char *pA, *pB;
  float *dC;
  CK(cudaMalloc(&dA, A.size() * 8));
  CK(cudaMalloc(&dB, B.size() * 8));
  CK(cudaMalloc(&pA, (size_t)mbs * n_kp * 2 * kPanelBytes));
  CK(cudaMalloc(&pB, (size_t)nbs * n_kp * 2 * kPanelBytes));
  CK(cudaMalloc(&dC, (size_t)M * N * 4));
  CK(cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice));
  const long ea = (long)mbs * 128 * n_kp * 64, eb = (long)nbs * 128 * n_kp * 64;
  k_pack<<<(unsigned)((ea + 255) / 256), 256>>>(dA, M, K, K, n_kp, pA);
  k_pack<<<(unsigned)((eb + 255) / 256), 256>>>(dB, N, K, K, n_kp, pB);
  CK(cudaGetLastError());
  CK(cudaFuncSetAttribute(k_gemm_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSmem));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaEventRecord(e0));
    k_gemm_split<<<dim3(mbs, nbs), 128, kSmem>>>(pA, pB, n_kp, M, N, dC);
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<float> Cm((size_t)M * N);
  CK(cudaMemcpy(Cm.data(), dC, Cm.size() * 4, cudaMemcpyDeviceToHost));
  double worst = 0.0, scale = 0.0;
  for (int i = 0; i < M; i += 7)
    for (int j = 0; j < N; j += 13) {
      double ref = 0.0;
      for (int k = 0; k < K; k++) ref += A[(size_t)i * K + k] * B[(size_t)j * K + k];
      worst = std::fmax(worst, std::fabs(ref - (double)Cm[(size_t)i * N + j]));
      scale = std::fmax(scale, std::fabs(ref));
    }
  std::printf("C[%d x %d] = A[.. x %d] B^T: max |err| %.3e of max |C| %.3e (rel %.2e; expected ~1e-6), %.3f ms, "
              "%.1f TFLOP/s algorithmic\n",
              M, N, K, worst, scale, worst / scale, ms, 2.0 * M * N * K / (ms * 1e-3) / 1e12);
  return worst / scale < 1e-5 ? 0 : 3;
}
