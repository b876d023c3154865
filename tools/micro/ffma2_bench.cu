// Microbenchmark: FP32 throughput of scalar FFMA vs packed FFMA2 (fma.rn.f32x2) on sm_100a.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2_bench ffma2_bench.cu && ./ffma2_bench
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b){ u64 r; asm("mov.b64 %0, {%1,%2};" : "=l"(r) : "f"(a), "f"(b)); return r; }
__device__ __forceinline__ float2 upk(u64 r){ float2 d; asm("mov.b64 {%0,%1}, %2;" : "=f"(d.x), "=f"(d.y) : "l"(r)); return d; }
__device__ __forceinline__ u64 fma2(u64 a, u64 b, u64 c){ u64 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

template <int MIX>  // MIX: extra integer ALU instructions per 8 fmas (0 or 4)
__global__ void k_scalar(float* out, float a, float b, int iters, int salt) {
  float acc[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) acc[i] = threadIdx.x * 0.001f + i;
  int z = salt + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) acc[i] = fmaf(acc[i], a, b);
    if (MIX) {
#pragma unroll
      for (int j = 0; j < MIX; ++j) z = (z ^ (z >> 3)) + it;
    }
  }
  float s = (float)z;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += acc[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MIX>
__global__ void k_packed(float* out, float a, float b, int iters, int salt) {
  u64 acc[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) acc[i] = pk(threadIdx.x * 0.001f + i, threadIdx.x * 0.002f + i);
  u64 A = pk(a, a), B = pk(b, b);
  int z = salt + threadIdx.x;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = fma2(acc[i], A, B);
    if (MIX) {
#pragma unroll
      for (int j = 0; j < MIX; ++j) z = (z ^ (z >> 3)) + it;
    }
  }
  float s = (float)z;
#pragma unroll
  for (int i = 0; i < 8; ++i) { float2 v = upk(acc[i]); s += v.x + v.y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <class F> float timeit(F f) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  f(); cudaDeviceSynchronize();
  cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  int sms; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  const int threads = 256, blocks = sms * 8, iters = 20000;
  float* out; cudaMalloc(&out, sizeof(float) * threads * blocks);
  const double fmas = 16.0 * iters * threads * blocks;  // FMAs per launch in every variant
  float t;
  t = timeit([&] { k_scalar<0><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters, 1); });
  printf("scalar FFMA            : %.3f ms  %.2f TFLOP/s\n", t, 2 * fmas / t / 1e9);
  t = timeit([&] { k_packed<0><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters, 1); });
  printf("packed FFMA2           : %.3f ms  %.2f TFLOP/s\n", t, 2 * fmas / t / 1e9);
  t = timeit([&] { k_scalar<4><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters, 1); });
  printf("scalar FFMA + 12 ALU/16: %.3f ms  %.2f TFLOP/s\n", t, 2 * fmas / t / 1e9);
  t = timeit([&] { k_packed<4><<<blocks, threads>>>(out, 1.0001f, 0.5f, iters, 1); });
  printf("packed FFMA2 + 12 ALU/16: %.3f ms  %.2f TFLOP/s\n", t, 2 * fmas / t / 1e9);
  return 0;
}
