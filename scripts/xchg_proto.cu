// xchg_proto.cu -- stand-alone prototype of the two mechanisms the round-2 one-pass kernel needs
// (DESIGN.md section 4.6), so they can be validated and timed on a B200 before they go into gmm_tc.cu:
//
//  (1) per-frame max and sum ACROSS the 128 TMEM lanes (components) of a CTA without shuffles:
//      redux.sync.max.s32 on the order-preserving integer image of the fp32 score, redux.sync.add.u32
//      on round(2^24 * 2^(S - m)), the four warps combined through shared memory;
//  (2) the exchange of the 64 (max, sum) pairs of a half tile between the 16 slice-CTAs of one
//      (non-portable) 16-CTA cluster: st.shared::cluster into every peer's ring slot, one remote
//      mbarrier arrival per peer, acquire wait on the local barrier, combine -> log-sum-exp.
//
// Each CTA (cluster rank = slice) synthesises S[component, frame] from a hash, so the host can check
// the exchanged log-sum-exp exactly; clock64() brackets give the cost per half tile.
//
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/xchg_proto scripts/xchg_proto.cu
//   ./scripts/xchg_proto [half_tiles]        (prints max |error| and clocks per half tile)
//
// Compile-checked in round 1 (no GPU budget left to run it); not part of the product library.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int kSlices = 16;   // cluster size = slice-CTAs of one frame group
constexpr int kCols = 64;     // frames per half tile
constexpr int kThreads = 128; // one thread per component (TMEM lane) of the slice
constexpr int kRing = 4;      // exchange slots in flight

__host__ __device__ inline float synth_score(unsigned cluster, unsigned h, unsigned slice, unsigned lane,
                                             unsigned col) {
  unsigned x = cluster * 0x9E3779B1u ^ (h + 1) * 0x85EBCA77u ^ (slice + 1) * 0xC2B2AE3Du ^ (lane + 1) * 0x27D4EB2Fu ^
               (col + 1) * 0x165667B1u;
  x ^= x >> 15;
  x *= 0x2C1B3C6Du;
  x ^= x >> 12;
  x *= 0x297A2D39u;
  x ^= x >> 15;
  // log2 scores in [-700, -100]: most components negligible, a few competitive
  float u = (float)(x & 0xFFFFFF) / 16777216.f;
  return -100.f - 600.f * u * u;
}

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ unsigned cluster_rank() {
  unsigned r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ unsigned map_to_peer(unsigned local_addr, unsigned peer) {
  unsigned r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(peer));
  return r;
}
__device__ __forceinline__ void st_peer_f2(unsigned addr, float a, float b) {
  asm volatile("st.shared::cluster.v2.f32 [%0], {%1, %2};" ::"r"(addr), "f"(a), "f"(b) : "memory");
}
__device__ __forceinline__ void mbar_init(unsigned bar, unsigned count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_peer(unsigned remote_bar) {
  asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned bar, unsigned parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra WAIT_DONE;\n"
      "bra WAIT_LOOP;\n"
      "WAIT_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// order-preserving map fp32 -> s32 (and back)
__device__ __forceinline__ int f2ord(float f) {
  int i = __float_as_int(f);
  return i ^ ((i >> 31) & 0x7FFFFFFF);
}
__device__ __forceinline__ float ord2f(int i) { return __int_as_float(i ^ ((i >> 31) & 0x7FFFFFFF)); }

struct Shared {
  float wmax[kCols][4];               // per-warp column maxima, float4 per column
  unsigned wsum[kCols][4];            // per-warp fixed-point column sums
  float2 slot[kRing][kSlices][kCols]; // (max, sum) of every slice of the cluster, ring of half tiles
  unsigned long long bar[kRing];
};

__global__ void __launch_bounds__(kThreads, 1)
k_xchg(int n_half, float *__restrict__ lse_out /*[clusters][n_half][kCols]*/, long long *__restrict__ clocks) {
  __shared__ Shared sh;
  const unsigned rank = cluster_rank();
  const unsigned cluster = blockIdx.x / kSlices;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int i = 0; i < kRing; i++) mbar_init(smem_u32(&sh.bar[i]), kSlices);  // one arrival per peer CTA
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  cluster_sync_all();  // every peer's barriers exist before anyone signals them

  long long t_reduce = 0, t_xchg = 0;
  for (int h = 0; h < n_half; h++) {
    const int ring = h % kRing;
    float s[kCols];
#pragma unroll
    for (int j = 0; j < kCols; j++) s[j] = synth_score(cluster, h, rank, threadIdx.x, j);
    long long c0 = clock64();
    // ---- (1a) column max over the 128 lanes
#pragma unroll
    for (int j = 0; j < kCols; j++) {
      int m = __reduce_max_sync(0xFFFFFFFFu, f2ord(s[j]));
      if (lane == (j & 31)) sh.wmax[j][warp] = ord2f(m);
    }
    __syncthreads();
    float mloc[kCols];
#pragma unroll
    for (int j = 0; j < kCols; j++) {
      float4 v = *reinterpret_cast<const float4 *>(sh.wmax[j]);
      mloc[j] = fmaxf(fmaxf(v.x, v.y), fmaxf(v.z, v.w));
    }
    // ---- (1b) column sum of 2^(S - m) in 24-bit fixed point
#pragma unroll
    for (int j = 0; j < kCols; j++) {
      float e = exp2f(s[j] - mloc[j]);
      unsigned q = __float2uint_rn(e * 16777216.f);
      unsigned t = __reduce_add_sync(0xFFFFFFFFu, q);
      if (lane == (j & 31)) sh.wsum[j][warp] = t;
    }
    __syncthreads();
    long long c1 = clock64();
    // ---- (2) exchange: thread (peer group p, column j) writes this CTA's pair into peer slots
    // 128 threads x 8 stores = 16 peers x 64 columns
    {
      const int j = threadIdx.x & (kCols - 1);
      const uint4 w = *reinterpret_cast<const uint4 *>(sh.wsum[j]);
      const float z = (float)((double)(w.x + w.y) + (double)(w.z + w.w)) * (1.f / 16777216.f);
      const unsigned local = smem_u32(&sh.slot[ring][rank][j]);
#pragma unroll
      for (int k = 0; k < kSlices / 2; k++) {
        const unsigned peer = (unsigned)(2 * k + (threadIdx.x >> 6));
        st_peer_f2(map_to_peer(local, peer), mloc[j], z);
      }
    }
    __syncthreads();  // all of this CTA's remote stores are ordered before the arrivals below
    if (threadIdx.x < kSlices) {
      asm volatile("fence.acq_rel.cluster;" ::: "memory");
      mbar_arrive_peer(map_to_peer(smem_u32(&sh.bar[ring]), threadIdx.x));
    }
    mbar_wait_cluster(smem_u32(&sh.bar[ring]), (unsigned)((h / kRing) & 1));
    if (threadIdx.x < kCols) {
      const int j = threadIdx.x;
      float m = -3.0e38f;
#pragma unroll
      for (int r = 0; r < kSlices; r++) m = fmaxf(m, sh.slot[ring][r][j].x);
      float z = 0.f;
#pragma unroll
      for (int r = 0; r < kSlices; r++) z += sh.slot[ring][r][j].y * exp2f(sh.slot[ring][r][j].x - m);
      if (rank == 0) lse_out[((size_t)cluster * n_half + h) * kCols + j] = m + log2f(z);
    }
    long long c2 = clock64();
    t_reduce += c1 - c0;
    t_xchg += c2 - c1;
    __syncthreads();  // wmax / wsum are reused by the next half tile
  }
  if (threadIdx.x == 0) {
    clocks[2 * blockIdx.x] = t_reduce;
    clocks[2 * blockIdx.x + 1] = t_xchg;
  }
  cluster_sync_all();  // no CTA leaves while a peer may still store into it
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
  const int n_half = argc > 1 ? std::atoi(argv[1]) : 256;
  CK(cudaFuncSetAttribute(k_xchg, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
  cudaLaunchConfig_t cfg = {};
  cfg.blockDim = dim3(kThreads);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = kSlices;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cfg.gridDim = dim3(kSlices * 64);
  int max_clusters = 0;
  CK(cudaOccupancyMaxActiveClusters(&max_clusters, k_xchg, &cfg));
  std::printf("co-resident 16-CTA clusters: %d\n", max_clusters);
  if (max_clusters < 1) return 2;
  const int clusters = max_clusters;
  cfg.gridDim = dim3(kSlices * clusters);
  float *d_lse;
  long long *d_clk;
  CK(cudaMalloc(&d_lse, (size_t)clusters * n_half * kCols * sizeof(float)));
  CK(cudaMalloc(&d_clk, (size_t)clusters * kSlices * 2 * sizeof(long long)));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int rep = 0; rep < 2; rep++) {
    CK(cudaEventRecord(e0));
    CK(cudaLaunchKernelEx(&cfg, k_xchg, n_half, d_lse, d_clk));
    CK(cudaEventRecord(e1));
    CK(cudaDeviceSynchronize());
  }
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  std::vector<float> lse((size_t)clusters * n_half * kCols);
  std::vector<long long> clk((size_t)clusters * kSlices * 2);
  CK(cudaMemcpy(lse.data(), d_lse, lse.size() * sizeof(float), cudaMemcpyDeviceToHost));
  CK(cudaMemcpy(clk.data(), d_clk, clk.size() * sizeof(long long), cudaMemcpyDeviceToHost));
  double worst = 0.0;
  for (int c = 0; c < clusters; c++)
    for (int h = 0; h < n_half; h += 17)
      for (int j = 0; j < kCols; j += 5) {
        double m = -1e300, z = 0.0;
        for (int sl = 0; sl < kSlices; sl++)
          for (int l = 0; l < kThreads; l++) m = std::fmax(m, (double)synth_score(c, h, sl, l, j));
        for (int sl = 0; sl < kSlices; sl++)
          for (int l = 0; l < kThreads; l++) z += std::exp2((double)synth_score(c, h, sl, l, j) - m);
        worst = std::fmax(worst, std::fabs(m + std::log2(z) - (double)lse[((size_t)c * n_half + h) * kCols + j]));
      }
  double red = 0, xch = 0;
  for (size_t i = 0; i < clk.size(); i += 2) {
    red += (double)clk[i];
    xch += (double)clk[i + 1];
  }
  red /= (double)(clk.size() / 2) * n_half;
  xch /= (double)(clk.size() / 2) * n_half;
  std::printf("max |lse error| %.3e log2 units (contract: 1e-4)\n", worst);
  std::printf("per half tile and CTA: lane reductions %.0f clk, cluster exchange %.0f clk; kernel %.3f ms "
              "for %d half tiles x %d clusters (%.2f us per half tile)\n",
              red, xch, ms, n_half, clusters, 1e3 * ms / n_half);
  return worst < 1e-4 ? 0 : 3;
}
