// Dependent-chain latency of FFMA vs FFMA2 / FMUL2 / FADD2 on sm_100a (one warp, clock64).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_latency ffma2_latency.cu && ./ffma2_latency
#include <cuda_runtime.h>
#include <cstdio>
#define N 4096
__global__ void lat_scalar(float* out, long long* cyc, float a, float b) {
  float x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = fmaf(x, a, b);
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
__global__ void lat_packed(float* out, long long* cyc, float a, float b) {
  float2 x = make_float2(threadIdx.x, threadIdx.x + 1.f), A = make_float2(a, a), B = make_float2(b, b);
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) x = __ffma2_rn(x, A, B);
  long long t1 = clock64();
  out[threadIdx.x] = x.x + x.y;
  if (threadIdx.x == 0) cyc[1] = t1 - t0;
}
// packed op whose result is consumed by a scalar op on one half, then re-packed (mixed chains)
__global__ void lat_mixed(float* out, long long* cyc, float a, float b) {
  float2 x = make_float2(threadIdx.x, threadIdx.x + 1.f), A = make_float2(a, a), B = make_float2(b, b);
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) { x = __ffma2_rn(x, A, B); x.x = fmaxf(x.x, 0.5f); x.y = fmaxf(x.y, 0.5f); }
  long long t1 = clock64();
  out[threadIdx.x] = x.x + x.y;
  if (threadIdx.x == 0) cyc[2] = t1 - t0;
}
__global__ void lat_scalar_mixed(float* out, long long* cyc, float a, float b) {
  float x = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 64
  for (int i = 0; i < N; ++i) { x = fmaf(x, a, b); x = fmaxf(x, 0.5f); }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0) cyc[3] = t1 - t0;
}
int main() {
  float* out; long long* cyc; long long h[4];
  cudaMalloc(&out, 4096); cudaMalloc(&cyc, 32);
  for (int rep = 0; rep < 2; ++rep) {
    lat_scalar<<<1, 32>>>(out, cyc, 0.999f, 0.5f);
    lat_packed<<<1, 32>>>(out, cyc, 0.999f, 0.5f);
    lat_mixed<<<1, 32>>>(out, cyc, 0.999f, 0.5f);
    lat_scalar_mixed<<<1, 32>>>(out, cyc, 0.999f, 0.5f);
    cudaDeviceSynchronize();
  }
  cudaMemcpy(h, cyc, 32, cudaMemcpyDeviceToHost);
  printf("FFMA  dependent latency : %.2f cycles\n", (double)h[0] / N);
  printf("FFMA2 dependent latency : %.2f cycles\n", (double)h[1] / N);
  printf("FFMA2 + 2x FMNMX chain  : %.2f cycles per iteration\n", (double)h[2] / N);
  printf("FFMA  + FMNMX chain     : %.2f cycles per iteration\n", (double)h[3] / N);
  return 0;
}
