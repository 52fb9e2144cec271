// Microbenchmark: tcgen05.mma issue/execute rate per SM for the shapes the attention and GEMM kernels use.
//   SS: A and B from shared memory (128-byte swizzle, K-major);  TS: A from TMEM, B from shared memory (MN-major)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I../../self_forcing_b200/csrc -o mma_throughput mma_throughput.cu
#include <stdio.h>
#include "common.cuh"
using namespace sfb;
namespace sfb { void set_error(const char*, ...) {} int check_cuda(cudaError_t, const char*) { return 0; } }

// MODE 0: SS, N columns, distinct smem k-slices (like QK^T / GEMM main loop)   MODE 1: TS (A in TMEM), B MN-major N=128
template <int MODE, int N>
__global__ void __launch_bounds__(128, 1) bench(long long* cycles, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem);                 // A: 128 rows x 128 k (two 16 KB halves)
    const uint32_t b_addr = smem_u32(smem + 32768);         // B: up to 256 rows x 128 k
    constexpr uint32_t idesc_ss = umma_idesc_bf16(128, N, 0, 0);
    constexpr uint32_t idesc_ts = umma_idesc_bf16(128, 128, 0, 1);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (MODE == 0) {
          const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
          const uint32_t boff = (k >> 2) * (N * 128) + (k & 3) * 32;
          umma_ss(tm + (it & 1) * 256, umma_desc_sw128(a_addr + off, 16, 1024), umma_desc_sw128(b_addr + boff, 16, 1024),
                  idesc_ss, k != 0);
        } else {
          umma_ts(tm + 256, tm + k * 8, umma_desc_sw128(b_addr + (k & 3) * 2048, 8192, 1024), idesc_ts, k != 0);
        }
      }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

// MMA stream (QK-like SS N=128 then PV-like TS, alternating per iteration) in warp 0 while warps 4..11 (two
// softmax-like warpgroups) run NOISE: 0 nothing, 1 tcgen05.ld 64 cols + wait, 2 ld + st 32 cols, 3 MUFU.EX2 loop,
// 4 ld + 64 ex2 + st (softmax-like mix)
template <int NOISE>
__global__ void __launch_bounds__(384, 1) contend(long long* cycles, float* sink, int iters) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint32_t slot;
  __shared__ uint64_t bar;
  __shared__ volatile int stop;
  __shared__ uint64_t bar2[6];
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); for (int i = 0; i < 6; ++i) mbar_init(&bar2[i], 1); fence_barrier_init(); stop = 0; }
  if (warp == 0) tmem_alloc(&slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tm = slot;
  if (threadIdx.x == 0) {
    const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 32768);
    constexpr uint32_t idesc_ss = umma_idesc_bf16(128, 128, 0, 0);
    constexpr uint32_t idesc_ts = umma_idesc_bf16(128, 128, 0, 1);
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        const uint32_t off = (k >> 2) * 16384 + (k & 3) * 32;
        umma_ss(tm + (it & 1) * 128, umma_desc_sw128(a_addr + off, 16, 1024), umma_desc_sw128(b_addr + off, 16, 1024),
                idesc_ss, k != 0);
      }
      if (NOISE >= 5) umma_commit(&bar2[0]);
      if (NOISE == 6) tc_fence_after();
      if (NOISE == 7) { umma_commit(&bar2[1]); umma_commit(&bar2[2]); }
#pragma unroll
      for (int k = 0; k < 8; ++k)
        umma_ts(tm + 256 + (it & 1) * 128, tm + (it & 1) * 128 + k * 8, umma_desc_sw128(b_addr + k * 2048, 16384, 1024),
                idesc_ts, k != 0);
      if (NOISE >= 5) umma_commit(&bar2[3]);
      if (NOISE == 6) tc_fence_after();
      if (NOISE == 7) { umma_commit(&bar2[4]); umma_commit(&bar2[5]); }
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    cycles[blockIdx.x] = t1 - t0;
    stop = 1;
  } else if (warp >= 4 && NOISE > 0) {
    const uint32_t base = tm + ((uint32_t)((warp & 3) * 32) << 16) + ((warp - 4) >> 2) * 128;
    float acc = 0.f;
    uint32_t v[64];
#pragma unroll
    for (int i = 0; i < 64; ++i) v[i] = 0x3c003c00u;
    while (!stop) {
      if (NOISE == 1 || NOISE == 2 || NOISE == 4) {
        tmem_ld32(base, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(base + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
      }
      if (NOISE == 3 || NOISE == 4) {
#pragma unroll
        for (int i = 0; i < 64; ++i) {
          float y;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(__uint_as_float(v[i]) * 1e-9f));
          acc += y;
        }
      }
      if (NOISE == 2 || NOISE == 4) {
        uint32_t w[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) w[i] = 0;
        tmem_st16(base + 64, w);     // columns 64.. of the S region: never an accumulator of a pending MMA in this test
        tmem_st16(base + 80, w);
        tmem_st_wait();
      }
    }
    sink[blockIdx.x * 384 + threadIdx.x] = acc + __uint_as_float(v[3]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) { tc_fence_after(); tmem_dealloc(tm, 512); }
}

template <int NOISE>
void run_contend(const char* name, long long* cyc, float* sink) {
  const int iters = 2000, smem = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(contend<NOISE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) contend<NOISE><<<148, 384, smem>>>(cyc, sink, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); return; }
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  printf("%-52s clk per (8 QK + 8 PV) = %.0f  (floor 1024)\n", name, (double)c / iters);
}

template <int MODE, int N>
void run(const char* name, long long* cyc) {
  const int iters = 2000, smem = 32768 + 65536 + 1024;
  cudaFuncSetAttribute(bench<MODE, N>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int rep = 0; rep < 2; ++rep) bench<MODE, N><<<148, 128, smem>>>(cyc, iters);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("%s: error %s\n", name, cudaGetErrorString(e)); return; }
  long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
  const double per = (double)c / (iters * 8.0);
  const int n = MODE == 0 ? N : 128;
  printf("%-44s clk/MMA=%.1f  (math floor %d)  FLOP/clk/SM=%.0f\n", name, per, n / 2, 2.0 * 128 * n * 16 / per);
}

int main() {
  long long* cyc; cudaMalloc(&cyc, 148 * 8);
  run<0, 64>("SS  M128 N64  K16 (QK^T, 64-wide KV)", cyc);
  run<0, 128>("SS  M128 N128 K16 (QK^T, 128-wide KV)", cyc);
  run<0, 256>("SS  M128 N256 K16 (GEMM 128x256 tile)", cyc);
  run<1, 128>("TS  M128 N128 K16 (P.V, A in TMEM)", cyc);
  float* sink; cudaMalloc(&sink, 148 * 384 * 4);
  run_contend<0>("QK+PV stream, other warps idle", cyc, sink);
  run_contend<1>("QK+PV stream + 8 warps tcgen05.ld", cyc, sink);
  run_contend<2>("QK+PV stream + 8 warps tcgen05.ld + st", cyc, sink);
  run_contend<3>("QK+PV stream + 8 warps MUFU.EX2", cyc, sink);
  run_contend<4>("QK+PV stream + 8 warps ld + ex2 + st", cyc, sink);
  run_contend<5>("QK+PV stream, commit after every 8 MMAs", cyc, sink);
  run_contend<6>("QK+PV stream, commit + fence::after every 8 MMAs", cyc, sink);
  run_contend<7>("QK+PV stream, 3 commits after every 8 MMAs", cyc, sink);
  return 0;
}
