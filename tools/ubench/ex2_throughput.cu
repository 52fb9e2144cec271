// Microbenchmark: MUFU.EX2 throughput per SM for f32, f16x2 and bf16x2 operands, and an FMA-pipe
// polynomial exp2.  Decides how the attention softmax should compute its exponentials.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ex2_throughput ex2_throughput.cu
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

constexpr int ITERS = 4096;
constexpr int CHAINS = 8;

__device__ __forceinline__ float ex2_f32(float x) { float y; asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_f16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.f16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
__device__ __forceinline__ uint32_t ex2_bf16x2(uint32_t x) { uint32_t y; asm volatile("ex2.approx.ftz.bf16x2 %0, %1;" : "=r"(y) : "r"(x)); return y; }
// Cody-Waite + degree-3 minimax polynomial on [0,1): 2^x = 2^floor(x) * p(frac(x))
__device__ __forceinline__ float ex2_poly(float x) {
  x = fmaxf(x, -126.f);
  float fl = floorf(x);
  float f = x - fl;
  float p = fmaf(fmaf(fmaf(0.0555041086648216f, f, 0.2402264923172690f), f, 0.6931471805599453f), f, 1.0f);
  return __int_as_float(__float_as_int(p) + (((int)fl) << 23));
}

template <int MODE>
__global__ void bench(float* out, long long* cycles) {
  float a[CHAINS];
  uint32_t h[CHAINS];
  for (int i = 0; i < CHAINS; ++i) { a[i] = -0.001f * (threadIdx.x + i); h[i] = 0xb800b800u + threadIdx.x + i; }
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < ITERS; ++it) {
#pragma unroll
    for (int i = 0; i < CHAINS; ++i) {
      if (MODE == 0) a[i] = ex2_f32(a[i]) - 1.0009f;
      if (MODE == 1) h[i] = ex2_f16x2(h[i]) ^ 0x80008000u;
      if (MODE == 2) h[i] = ex2_bf16x2(h[i]) ^ 0x80008000u;
      if (MODE == 3) a[i] = ex2_poly(a[i]) - 1.0009f;
    }
  }
  long long t1 = clock64();
  float s = 0; uint32_t x = 0;
  for (int i = 0; i < CHAINS; ++i) { s += a[i]; x ^= h[i]; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)x;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

int main() {
  float* out; long long* cyc;
  cudaMalloc(&out, 148 * 1024 * sizeof(float));
  cudaMalloc(&cyc, 148 * sizeof(long long));
  const char* names[4] = {"ex2.approx.ftz.f32", "ex2.approx.f16x2 (2 results/op)", "ex2.approx.ftz.bf16x2 (2 results/op)", "poly3 exp2 on FMA pipe"};
  for (int mode = 0; mode < 4; ++mode) {
    for (int threads = 128; threads <= 1024; threads *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) bench<0><<<148, threads>>>(out, cyc);
        if (mode == 1) bench<1><<<148, threads>>>(out, cyc);
        if (mode == 2) bench<2><<<148, threads>>>(out, cyc);
        if (mode == 3) bench<3><<<148, threads>>>(out, cyc);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long c;
      cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
      double ops = (double)ITERS * CHAINS * threads;   // instructions (lanes) per SM
      double per = (mode == 1 || mode == 2) ? 2.0 : 1.0;
      printf("%-40s threads/SM=%4d  cycles=%lld  lane-ops/clk/SM=%.2f  results/clk/SM=%.2f\n", names[mode], threads, c,
             ops / c, per * ops / c);
    }
  }
  return 0;
}
