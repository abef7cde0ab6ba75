// Dependent-chain latencies of the instructions the 16 x 16 Cholesky recurrence is made of (one warp, B200):
// DFMA, DADD, rsqrt(double), 1/x (double), SHFL of a double, shared-memory store -> load, DMMA m8n8k4.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency_probe fp64_latency_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};"
               : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

__global__ void probe(double *out, long long *clk, double seed, int n) {
  __shared__ double sm[64];
  double x = seed + threadIdx.x * 1e-3, y = 1.0000001;
  long long t0, t1;
  t0 = clock64();
  for (int i = 0; i < n; i++) x = fma(x, y, 1e-9);
  t1 = clock64();
  if (threadIdx.x == 0) clk[0] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) x = x + y;
  t1 = clock64();
  if (threadIdx.x == 0) clk[1] = t1 - t0;
  x = fabs(x) + 1.0;
  t0 = clock64();
  for (int i = 0; i < n; i++) x = rsqrt(x) + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) clk[2] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) x = 1.0 / x + 1.5;
  t1 = clock64();
  if (threadIdx.x == 0) clk[3] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) x = __shfl_xor_sync(0xffffffffu, x, 1) + 0.0;
  t1 = clock64();
  if (threadIdx.x == 0) clk[4] = t1 - t0;
  t0 = clock64();
  for (int i = 0; i < n; i++) {
    sm[threadIdx.x] = x;
    __syncwarp();
    x = sm[threadIdx.x ^ 1];
    __syncwarp();
  }
  t1 = clock64();
  if (threadIdx.x == 0) clk[5] = t1 - t0;
  double c0 = x, c1 = y;
  t0 = clock64();
  for (int i = 0; i < n; i++) dmma(c0, c1, y, y);
  t1 = clock64();
  if (threadIdx.x == 0) clk[6] = t1 - t0;
  // throughput: 8 independent DFMA chains
  double z[8];
  for (int k = 0; k < 8; k++) z[k] = x + k;
  t0 = clock64();
  for (int i = 0; i < n; i++)
#pragma unroll
    for (int k = 0; k < 8; k++) z[k] = fma(z[k], y, 1e-9);
  t1 = clock64();
  if (threadIdx.x == 0) clk[7] = t1 - t0;
  for (int k = 0; k < 8; k++) x += z[k];
  out[threadIdx.x] = x + c0 + c1;
}

int main() {
  double *out;
  long long *clk, h[8];
  cudaMalloc(&out, 32 * 8);
  cudaMalloc(&clk, 8 * 8);
  const int n = 4096;
  for (int rep = 0; rep < 2; rep++) probe<<<1, 32>>>(out, clk, 1.0, n);
  cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
  const char *name[8] = {"DFMA", "DADD", "rsqrt(double)+DADD", "1/x(double)+DADD", "SHFL(double)+DADD", "STS->LDS round trip",
                         "DMMA m8n8k4 (dependent)", "8 independent DFMA (per 8)"};
  for (int i = 0; i < 8; i++) printf("%-30s %.1f clk\n", name[i], (double)h[i] / n);
  return cudaGetLastError() != cudaSuccess;
}
