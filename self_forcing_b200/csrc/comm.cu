// Cross-GPU synchronisation for kernels that exchange data through peer-mapped memory (NVLink / NVSwitch).
//
// The Ulysses head-parallel path does its two all-to-alls per transformer block with plain global stores into
// peer memory from inside the producing kernels (sfb_qk_norm_rope_sp, sfb_attention_fwd_sp).  What is left of the
// "collective" is this barrier: every rank publishes an epoch number into a flag slot on every peer and spins until
// all peers' epochs have arrived.  One process per GPU, so the spinning kernels of different ranks always run
// concurrently (they are on different devices).
#include "common.cuh"

namespace sfb {

struct PeerFlags {
  int* flags[8];   // flags[p]: rank p's array of n ints, mapped into this process; slot [r] is written by rank r
};

// The epoch lives in device memory (slot [n] of this rank's own flag array) and is advanced by the kernel itself, so
// the launch has no per-call argument and can be replayed inside a CUDA graph.
__global__ void peer_barrier_kernel(PeerFlags pf, int rank, int n) {
  __shared__ int next_epoch;
  if (threadIdx.x == 0) next_epoch = ++pf.flags[rank][n];
  __syncthreads();
  const int epoch = next_epoch;
  const int p = threadIdx.x;
  if (p >= n) return;
  __threadfence_system();   // order this GPU's earlier peer stores before the flag
  asm volatile("st.release.sys.global.s32 [%0], %1;" ::"l"(pf.flags[p] + rank), "r"(epoch) : "memory");
  int v;
  do {
    asm volatile("ld.acquire.sys.global.s32 %0, [%1];" : "=r"(v) : "l"(pf.flags[rank] + p) : "memory");
  } while (v < epoch);
}

}  // namespace sfb

// flag_ptrs[p] = device pointer (as mapped in THIS process) of rank p's flag array of n + 1 int32, zero-initialised
// (slot [n] is the rank's own call counter).  Every rank must call it the same number of times.
extern "C" int sfb_peer_barrier(void* const* flag_ptrs, int rank, int n, void* stream) {
  using namespace sfb;
  if (n < 1 || n > 8 || rank < 0 || rank >= n) { set_error("sfb_peer_barrier: bad rank %d of %d", rank, n); return SFB_ERR_INVALID; }
  PeerFlags pf{};
  for (int i = 0; i < n; ++i) pf.flags[i] = static_cast<int*>(flag_ptrs[i]);
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(pf, rank, n);
  return check_cuda(cudaGetLastError(), "peer_barrier launch");
}
