// Micro-benchmark: fp32 FMA throughput per SM with scalar FFMA (three register operands) against packed FFMA2
// (fma.rn.f32x2, sm_100+).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ffma2 ffma2.cu
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kAcc = 16;  // independent accumulators per thread (scalar) / 8 pairs (packed)

__global__ void __launch_bounds__(256) scalar_kernel(float* out, float b, float c) {
  float a[kAcc];
#pragma unroll
  for (int i = 0; i < kAcc; ++i) a[i] = threadIdx.x * 1e-3f + i;
  float bb = b + threadIdx.x * 1e-9f, cc = c + threadIdx.x * 1e-9f;
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kAcc; ++i) a[i] = fmaf(a[i], bb, cc);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kAcc; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void __launch_bounds__(256) packed_kernel(float* out, float b, float c) {
  uint64_t a[kAcc / 2];
#pragma unroll
  for (int i = 0; i < kAcc / 2; ++i) {
    float lo = threadIdx.x * 1e-3f + 2 * i, hi = lo + 1.f;
    asm("mov.b64 %0, {%1, %2};" : "=l"(a[i]) : "f"(lo), "f"(hi));
  }
  uint64_t bb, cc;
  float b0 = b + threadIdx.x * 1e-9f, c0 = c + threadIdx.x * 1e-9f;
  asm("mov.b64 %0, {%1, %1};" : "=l"(bb) : "f"(b0));
  asm("mov.b64 %0, {%1, %1};" : "=l"(cc) : "f"(c0));
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kAcc / 2; ++i) asm volatile("fma.rn.f32x2 %0, %0, %1, %2;" : "+l"(a[i]) : "l"(bb), "l"(cc));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kAcc / 2; ++i) {
    float lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(a[i]));
    s += lo + hi;
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
  int sms = 0, khz = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
  float* out;
  const int grid = sms * 8, block = 256;
  cudaMalloc(&out, (size_t)grid * block * sizeof(float));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (int which = 0; which < 2; ++which) {
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
      cudaEventRecord(e0);
      if (which == 0) scalar_kernel<<<grid, block>>>(out, 0.999f, 0.001f);
      else packed_kernel<<<grid, block>>>(out, 0.999f, 0.001f);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      if (ms < best) best = ms;
    }
    const double fma = (double)grid * block * kIters * kAcc;
    printf("%s: %.3f ms  %.2f TFMA/s  = %.1f FMA/clk/SM at the nominal %d MHz (%d SMs)\n", which ? "FFMA2 (f32x2)" : "FFMA scalar ", best,
           fma / best / 1e9, fma / (best * 1e-3) / sms / (khz * 1e3), khz / 1000, sms);
  }
  cudaError_t err = cudaDeviceSynchronize();
  printf("status: %s\n", cudaGetErrorString(err));
  return err != cudaSuccess;
}
