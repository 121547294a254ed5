// Micro-benchmark: fp32 FMA issue rate on sm_100a -- packed fma.rn.f32x2 (FFMA2) vs scalar FFMA vs a 1:1 mix,
// at 1, 2 and 4 warps per scheduler.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_rate fma_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ uint64_t ffma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ float ffma(float a, float b, float c) {
  float d;
  asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(d) : "f"(a), "f"(b), "f"(c));
  return d;
}

__device__ __forceinline__ uint64_t pack2(float a, float b) {
  uint64_t r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(a), "f"(b));
  return r;
}

template <int MODE>   // 0: FFMA2 only, 1: FFMA only, 2: mixed, 3: FFMA2 whose multiplicand is a scalar broadcast pack2(f, f)
__global__ void k(float* out, int iters, float x) {
  uint64_t a2[8];
  float a1[16];
  const uint64_t b2 = ((uint64_t)__float_as_uint(x) << 32) | __float_as_uint(x);
#pragma unroll
  for (int i = 0; i < 8; ++i) a2[i] = (uint64_t)(threadIdx.x + i);
#pragma unroll
  for (int i = 0; i < 16; ++i) a1[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 8; ++i) a2[i] = ffma2(a2[i], b2, b2);
    } else if (MODE == 1) {
#pragma unroll
      for (int r = 0; r < 2; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) a1[i] = ffma(a1[i], x, x);
    } else if (MODE == 3) {
      const float f0 = a1[0], f1 = a1[1];
#pragma unroll
      for (int i = 0; i < 8; ++i) a2[i] = ffma2(pack2(f0, f0), a2[i], b2);
#pragma unroll
      for (int i = 0; i < 8; ++i) a2[i] = ffma2(pack2(f1, f1), a2[i], b2);
    } else {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        a2[i] = ffma2(a2[i], b2, b2);
        a1[2 * i] = ffma(a1[2 * i], x, x);
        a1[2 * i + 1] = ffma(a1[2 * i + 1], x, x);
      }
    }
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += __uint_as_float((uint32_t)a2[i]) + __uint_as_float((uint32_t)(a2[i] >> 32));
#pragma unroll
  for (int i = 0; i < 16; ++i) s += a1[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int threads, float fmas_per_iter_per_thread) {
  float* out;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  const int iters = 200000;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148, threads>>>(out, 1000, 0.999f);
  cudaEventRecord(e0);
  k<MODE><<<148, threads>>>(out, iters, 0.999f);
  cudaEventRecord(e1);
  cudaEventSynchronize(e1);
  float ms;
  cudaEventElapsedTime(&ms, e0, e1);
  const double fma = (double)148 * threads * iters * fmas_per_iter_per_thread;
  printf("%-6s %4d threads/SM: %.3f ms  %.1f TFLOP/s  %.1f FMA/clk/SM (at 1.965 GHz)\n", name, threads, ms,
         2 * fma / ms / 1e9, fma / (ms * 1e-3) / 148 / 1.965e9);
  cudaFree(out);
}

int main() {
  for (int threads : {128, 256, 512}) {
    run<0>("FFMA2", threads, 32.f);    // 16 FFMA2 = 32 FMA
    run<1>("FFMA", threads, 32.f);     // 32 FFMA
    run<2>("mixed", threads, 32.f);    // 8 FFMA2 + 16 FFMA
    run<3>("bcast", threads, 32.f);    // 16 FFMA2 with a pack2(f, f) operand
  }
  return 0;
}
