// Micro-benchmark: global -> shared copy throughput per SM on B200 for
//   (a) cp.async.bulk (1-D bulk copy, UBLKCP) and (b) cp.async.bulk.tensor.2d (TMA, UTMALDG),
// as a function of the chunk size and the number of stages in flight.  No compute.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o bulk_bw bulk_bw.cu -lcuda
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                  \
  do {                                                                         \
    cudaError_t e = (x);                                                       \
    if (e != cudaSuccess) {                                                    \
      printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e));         \
      exit(1);                                                                 \
    }                                                                          \
  } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t c) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(c));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n.reg .pred P1;\nLAB_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n@P1 bra DONE;\nbra LAB_WAIT;\nDONE:\n}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tma_2d(uint32_t dst, const CUtensorMap *map, int c0, int c1, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(bar)
      : "memory");
}

// each CTA streams `n_tiles` tiles of `tile_bytes` (chunks of chunk_bytes) through `stages` slots
template <bool TMA>
__global__ void __launch_bounds__(128, 1)
k_stream(const unsigned char *src, const __grid_constant__ CUtensorMap map, int n_tiles, int tile_bytes,
         int chunk_bytes, int stages, int same_tiles, long long *cycles) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) uint64_t full[8];
  uint32_t base = (smem_u32(smem) + 1023u) & ~1023u;
  if (threadIdx.x == 0) {
    for (int i = 0; i < stages; i++) mbar_init(smem_u32(&full[i]), 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  long long t0 = clock64();
  if (threadIdx.x == 0) {
    // producer and consumer in one thread: keep `stages` tiles in flight
    auto issue = [&](int i) {
      int st = i % stages;
      uint32_t bar = smem_u32(&full[st]);
      mbar_expect_tx(bar, tile_bytes);
      size_t tile = same_tiles ? (size_t)i : (size_t)blockIdx.x * n_tiles + i;
      for (int off = 0; off < tile_bytes; off += chunk_bytes) {
        if (TMA) {
          // 2-D view: rows of 128 bytes; chunk = chunk_bytes/128 rows
          int row = (int)((tile * tile_bytes + off) / 128);
          tma_2d(base + st * tile_bytes + off, &map, 0, row, bar);
        } else {
          bulk_g2s(base + st * tile_bytes + off, src + tile * tile_bytes + off, chunk_bytes, bar);
        }
      }
    };
    for (int i = 0; i < stages && i < n_tiles; i++) issue(i);
    for (int i = 0; i < n_tiles; i++) {
      int st = i % stages;
      mbar_wait(smem_u32(&full[st]), (i / stages) & 1);
      if (i + stages < n_tiles) issue(i + stages);
    }
  }
  __syncthreads();
  long long t1 = clock64();
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

typedef CUresult (*EncodeFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                             const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                             CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main() {
  int dev = 0;
  CK(cudaSetDevice(dev));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, dev));
  const int n_sm = prop.multiProcessorCount;
  const size_t total = (size_t)2 << 30;  // 2 GB source
  unsigned char *src;
  CK(cudaMalloc(&src, total));
  CK(cudaMemset(src, 1, total));
  long long *cyc;
  CK(cudaMalloc(&cyc, n_sm * sizeof(long long)));
  void *fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  EncodeFn encode = (EncodeFn)fn;
  CK(cudaFuncSetAttribute(k_stream<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  CK(cudaFuncSetAttribute(k_stream<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  printf("SMs %d, clock %.0f MHz\n", n_sm, prop.clockRate / 1e3);
  printf("%6s %8s %8s %6s %5s | %9s %9s %9s\n", "mode", "tile", "chunk", "stages", "same", "ms", "GB/s/SM", "TB/s tot");
  for (int tma = 0; tma < 2; tma++)
    for (int same = 0; same < 2; same++)
      for (int tile_kb : {32, 64})
        for (int chunk_kb : {2, 8, 16, 32})
          for (int stages : {2, 3}) {
            int tile_bytes = tile_kb * 1024, chunk_bytes = chunk_kb * 1024;
            if (chunk_bytes > tile_bytes) continue;
            if (tma && chunk_bytes / 128 > 256) continue;  // box rows <= 256
            if (stages * tile_bytes > 190 * 1024) continue;
            int n_tiles = (int)(total / tile_bytes / n_sm);
            if (n_tiles > 800) n_tiles = 800;
            CUtensorMap map;
            memset(&map, 0, sizeof(map));
            if (tma) {
              cuuint64_t gdim[2] = {64, (cuuint64_t)(total / 128)};  // 64 fp16 x rows
              cuuint64_t gstride[1] = {128};
              cuuint32_t box[2] = {64, (cuuint32_t)(chunk_bytes / 128)};
              cuuint32_t estr[2] = {1, 1};
              CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, src, gdim, gstride, box, estr,
                                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                  CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
              if (r != CUDA_SUCCESS) {
                printf("encode failed %d\n", (int)r);
                continue;
              }
            }
            size_t smem = (size_t)stages * tile_bytes + 2048;
            for (int rep = 0; rep < 2; rep++) {
              CK(cudaEventRecord(e0));
              if (tma)
                k_stream<true><<<n_sm, 128, smem>>>(src, map, n_tiles, tile_bytes, chunk_bytes, stages, same, cyc);
              else
                k_stream<false><<<n_sm, 128, smem>>>(src, map, n_tiles, tile_bytes, chunk_bytes, stages, same, cyc);
              CK(cudaEventRecord(e1));
              CK(cudaEventSynchronize(e1));
            }
            CK(cudaGetLastError());
            float ms;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            double bytes_sm = (double)n_tiles * tile_bytes;
            printf("%6s %7dK %7dK %6d %5d | %9.3f %9.1f %9.2f\n", tma ? "tma2d" : "bulk", tile_kb, chunk_kb, stages, same,
                   ms, bytes_sm / (ms * 1e-3) / 1e9, bytes_sm * n_sm / (ms * 1e-3) / 1e12);
          }
  return 0;
}
