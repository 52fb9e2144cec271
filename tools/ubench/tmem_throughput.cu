// Microbenchmark: tcgen05.ld / tcgen05.st throughput per SM (32x32b.x32 shape), 4 or 8 warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I../../self_forcing_b200/csrc -o tmem_throughput tmem_throughput.cu
#include <stdio.h>
#include "common.cuh"
using namespace sfb;
namespace sfb { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }

template <int MODE>   // 0: ld 128 cols + wait per iter, 1: st 64 cols + wait, 2: ld 128 + st 64 (softmax-like traffic)
__global__ void bench(long long* cycles, float* sink, int iters) {
  __shared__ uint32_t slot;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = slot + ((uint32_t)((warp & 3) * 32) << 16) + (warp >> 2) * 256;
  float acc = 0.f;
  uint32_t v[128];
#pragma unroll
  for (int i = 0; i < 128; ++i) v[i] = threadIdx.x + i;
  __syncthreads();
  long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0 || MODE == 2) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(base + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&v[c * 32]));
      tmem_ld_wait();
#pragma unroll
      for (int i = 0; i < 128; i += 16) acc += __uint_as_float(v[i]);
    }
    if (MODE == 1 || MODE == 2) {
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_st16(base + c * 16, &v[c * 16]);
      tmem_st_wait();
    }
  }
  long long t1 = clock64();
  sink[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(slot, 512); }
}

int main() {
  long long* cyc; float* sink;
  cudaMalloc(&cyc, 148 * 8); cudaMalloc(&sink, 148 * 256 * 4);
  const int iters = 2000;
  const char* names[3] = {"ld 128 cols (16 KB/warp)", "st 64 cols (8 KB/warp)", "ld 128 + st 64"};
  for (int mode = 0; mode < 3; ++mode)
    for (int threads = 128; threads <= 256; threads *= 2) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) bench<0><<<148, threads>>>(cyc, sink, iters);
        if (mode == 1) bench<1><<<148, threads>>>(cyc, sink, iters);
        if (mode == 2) bench<2><<<148, threads>>>(cyc, sink, iters);
      }
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
      long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
      double bytes_ld = (mode == 1 ? 0.0 : 128.0 * 4 * threads), bytes_st = (mode == 0 ? 0.0 : 64.0 * 4 * threads);
      printf("%-28s warps=%d  clk/iter=%.1f  ld B/clk/SM=%.1f  st B/clk/SM=%.1f\n", names[mode], threads / 32,
             (double)c / iters, bytes_ld * iters / c, bytes_st * iters / c);
    }
  return 0;
}
