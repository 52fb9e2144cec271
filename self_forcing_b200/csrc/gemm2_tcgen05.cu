// CTA-pair bf16 GEMM for sm_100a: one 256 x 256 output tile per cluster of two CTAs (cta_group::2).
//
// Why: with one CTA per tile the kernel in gemm_tcgen05.cu is bound by the L2 -> SM fill rate, not by the
// tensor pipe (ncu, profiles/r01a: 128x256 tiles move 48 KB per 64-wide k-block and sustain ~43 B/clk/SM
// of l1tex__m_xbar2l1tex_read_bytes, i.e. 44 % tensor-pipe utilisation).  A CTA pair shares the B tile:
// each CTA stages its own 128 rows of A and only HALF of B (128 of the 256 columns); the UMMA reads both
// halves.  That is 32 KB per k-block per SM for the same 512 tensor cycles -> 1.5x the arithmetic
// intensity per byte that crosses the crossbar.
//
// Roles per CTA (192 threads):  warp 0 TMA producer (own A rows, own B half; completion bytes land on the
// LEADER's mbarrier), warp 1 = MMA issuer (leader CTA only; tcgen05.mma.cta_group::2, M=256 N=256 K=16,
// commits multicast to both CTAs), warps 2-5 epilogue of this CTA's 128 accumulator rows (TMEM double
// buffered: 2 x 256 columns).
//
// Epilogue: the accumulator rows are staged in shared memory (four 128 x 64 bf16 sub-tiles, 128-byte
// swizzle) and written with TMA stores; the residual tile is prefetched into the same staging buffers with
// TMA loads while the main loop of the tile is still running, and updated in place.  (A per-thread
// row-strided epilogue -- 16-byte global loads/stores at a 3 KB row pitch -- measured 24 k .. 54 k cycles per
// tile on B200, longer than the 18 k-cycle main loop of a K=1536 tile, and became the bound.)
#include <stdlib.h>

#include "gemm_common.cuh"

namespace sfb {

constexpr int G2_ROWS = 128;           // accumulator rows per CTA (pair tile = 256 rows)
constexpr int G2_BN = 256;             // pair tile columns; each CTA stages 128 of them
constexpr int G2_BK = 64;
constexpr int G2_STAGES = 5;
constexpr int G2_SUB = 64;             // staging sub-tile width (one 128-byte swizzle row of bf16)
constexpr int G2_NSUB = G2_BN / G2_SUB;
constexpr int G2_SUB_BYTES = G2_ROWS * G2_SUB * 2;
constexpr int G2_A_BYTES = G2_ROWS * G2_BK * 2;
constexpr int G2_B_BYTES = (G2_BN / 2) * G2_BK * 2;
constexpr int G2_STAGE_BYTES = G2_A_BYTES + G2_B_BYTES;
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + G2_NSUB * G2_SUB_BYTES + 1024 + 256;
constexpr int G2_THREADS = 192;
constexpr int G2_MAX_PAIRS = 96;       // stream-K workspace slots

// This pair's contiguous range of (tile, k-block) work, walked tile by tile.  Without stream-K the range is a whole
// number of tiles; with it the range is W / pairs k-blocks and a tile can be cut once (host guarantees tiles >= pairs).
struct PairRange {
  long long cur, end;
  int kbs;
  __device__ PairRange(const GemmParams& p, int pair, int num_pairs) : kbs(p.num_k_blocks) {
    const long long tiles = (long long)p.num_m_blocks * p.num_n_blocks;
    if (p.streamk) {
      cur = tiles * kbs * pair / num_pairs;
      end = tiles * kbs * (pair + 1) / num_pairs;
    } else {
      cur = (tiles * pair / num_pairs) * kbs;
      end = (tiles * (pair + 1) / num_pairs) * kbs;
    }
  }
  __device__ bool next(int& tile, int& kb0, int& kb1) {
    if (cur >= end) return false;
    tile = (int)(cur / kbs);
    kb0 = (int)(cur - (long long)tile * kbs);
    kb1 = (end - cur) < (long long)(kbs - kb0) ? kb0 + (int)(end - cur) : kbs;
    cur += kb1 - kb0;
    return true;
  }
};

template <int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G2_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_out0, const __grid_constant__ CUtensorMap tma_out1,
                  const __grid_constant__ CUtensorMap tma_out2, const __grid_constant__ CUtensorMap tma_res,
                  const GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + G2_STAGES * G2_STAGE_BYTES;   // [G2_NSUB][128 rows][128 B], swizzled
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_out + G2_NSUB * G2_SUB_BYTES);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tmem_full = empty_bar + G2_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]  (only the leader's copy is waited on)
  uint64_t* res_full = tmem_empty + 2;           // [1]  residual tile landed in the staging buffers
  uint64_t* acc_init = res_full + 1;             // [1]  stream-K: the head tile's accumulator was preloaded (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(acc_init + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const bool leader = rank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);    // leader: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(&empty_bar[s], 1);   // one multicast commit per use
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 8);  // 4 epilogue warps x 2 CTAs arrive on the leader's barrier
    }
    mbar_init(res_full, 1);
    mbar_init(acc_init, 8);        // 4 epilogue warps x 2 CTAs
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int pair = blockIdx.x >> 1;
  const int num_pairs = gridDim.x >> 1;

  if (warp == 0) {
    // ------------------------------ TMA producer (both CTAs; warp-uniform, one lane issues) ----
    {
      int stage = 0;
      uint32_t phase = 0;
      PairRange range(p, pair, num_pairs);
      int tile, kb0, kb1;
      while (range.next(tile, kb0, kb1)) {
        const int m_blk = tile % p.num_m_blocks, n_blk = tile / p.num_m_blocks;
        const int a_row = m_blk * (2 * G2_ROWS) + rank * G2_ROWS;
        const int b_row = n_blk * G2_BN + rank * (G2_BN / 2);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* a_dst = smem + stage * G2_STAGE_BYTES;
            if (leader) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
            const uint32_t bar = mapa_cluster(smem_u32(&full_bar[stage]), 0);
            tma_load_2d_pair(a_dst, &tma_a, bar, kb * G2_BK, a_row);
            tma_load_2d_pair(a_dst + G2_A_BYTES, &tma_b, bar, kb * G2_BK, b_row);
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (leader only) ------------------
    if (leader) {   // warp-uniform; one lane issues
      constexpr uint32_t idesc = umma_idesc_bf16(2 * G2_ROWS, G2_BN, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      PairRange range(p, pair, num_pairs);
      int tile, kb0, kb1;
      for (; range.next(tile, kb0, kb1); ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        // stream-K head piece [0, kb1) of a tile whose tail another pair computed: the epilogue warps have preloaded
        // that fp32 partial into this accumulator, so every MMA accumulates
        const bool preloaded = p.streamk && kb0 == 0 && kb1 < p.num_k_blocks;
        if (preloaded) mbar_wait(acc_init, 0);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * G2_BN;
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + stage * G2_STAGE_BYTES);
            const uint64_t a_desc = umma_desc_sw128(a_addr, 16, 1024);
            const uint64_t b_desc = umma_desc_sw128(a_addr + G2_A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k)
              umma_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (preloaded || kb != kb0 || k != 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage], 3);   // both CTAs' smem slots are free once these retire
            if (kb == kb1 - 1) umma_commit_pair(&tmem_full[acc], 3);
          }
          __syncwarp();
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------ epilogue warps (both CTAs) ----------------
    constexpr bool HAS_RES = (EPI == EPI_RESIDUAL || EPI == EPI_GATE_RES);
    const int quarter = warp & 3;
    const int r_local = quarter * 32 + lane;                 // accumulator row (TMEM lane) of this thread
    const uint32_t leader_empty0 = mapa_cluster(smem_u32(&tmem_empty[0]), 0);
    uint8_t* my_row = stage_out + r_local * 128;
    const int sw = r_local & 7;                              // 128-byte swizzle: 16-byte chunk index ^ (row % 8)
    if (warp == 2 && elect_one()) {
      tma_prefetch_desc(&tma_out0);
      if (HAS_RES) tma_prefetch_desc(&tma_res);
    }
    uint32_t res_phase = 0;
    int it = 0;
    PairRange range(p, pair, num_pairs);
    const uint32_t leader_acc_init = mapa_cluster(smem_u32(acc_init), 0);
    int tile, kb0, kb1;
    for (; range.next(tile, kb0, kb1); ++it) {
      const int m_blk = tile % p.num_m_blocks, n_blk = tile / p.num_m_blocks;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row0 = m_blk * (2 * G2_ROWS) + rank * G2_ROWS;
      const int n0 = n_blk * G2_BN;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * G2_BN;
      const bool tail_part = kb0 > 0;                        // stream-K: this pair holds k-blocks [kb0, K) of the tile

      if (tail_part) {
        // park the fp32 partial (transposed: consecutive lanes = consecutive rows -> coalesced) for the pair that
        // owns the head of this tile (the previous pair), then publish it
        float* slot = p.sk_ws + ((long long)(pair - 1) * 2 + rank) * (G2_BN * G2_ROWS);
        mbar_wait(&tmem_full[acc], acc_phase);
        tc_fence_after();
#pragma unroll 1
        for (int c = 0; c < G2_BN / 32; ++c) {
          uint32_t v[32];
          tmem_ld32(t_row + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) slot[(c * 32 + i) * G2_ROWS + r_local] = __uint_as_float(v[i]);
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(leader_empty0 + acc * 8);
        __threadfence();
        named_barrier_sync(1, 128);
        if (warp == 2 && lane == 0)
          asm volatile("st.release.gpu.global.s32 [%0], %1;" ::"l"(p.sk_flags + (pair - 1) * 2 + rank), "r"(p.sk_epoch) : "memory");
      }

      // Look ahead: if the NEXT (= last) segment of this pair is the head piece of a split tile, preload the tail
      // partial (computed by the next pair at the very start of its range) into that segment's accumulator now,
      // while the MMAs of the current segment are still running.  The other accumulator buffer is free: its
      // previous user's epilogue finished in the previous iteration.  (After this pair's own tail partial has been
      // published, so that waiting for the neighbour's flag never delays a flag somebody else waits for.)
      if (p.streamk) {
        PairRange peek = range;
        int t2, a2, b2;
        if (peek.next(t2, a2, b2) && a2 == 0 && b2 < p.num_k_blocks) {
          const float* part = p.sk_ws + ((long long)pair * 2 + rank) * (G2_BN * G2_ROWS);
          const int* flag = p.sk_flags + pair * 2 + rank;
          int seen;
          do {
            asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
          } while (seen != p.sk_epoch);
          const uint32_t t_next = tmem_base + ((uint32_t)(quarter * 32) << 16) + ((it + 1) & 1) * G2_BN;
#pragma unroll 1
          for (int c = 0; c < G2_BN / 32; ++c) {
            uint32_t v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__ldcg(part + (c * 32 + i) * G2_ROWS + r_local));
            tmem_st32(t_next + c * 32, v);
          }
          tmem_st_wait();
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(leader_acc_init);
        }
      }

      if (tail_part) continue;

      const int seg = n0 / p.seg_cols;
      const int seg_col0 = n0 - seg * p.seg_cols;
      const CUtensorMap* omap = seg == 0 ? &tma_out0 : (seg == 1 ? &tma_out1 : &tma_out2);
      // staging buffers are free once the previous tile's TMA stores have read them
      if (warp == 2 && elect_one()) tma_store_wait_read<0>();
      named_barrier_sync(1, 128);
      if (HAS_RES && warp == 2 && elect_one()) {
        mbar_expect_tx(res_full, G2_NSUB * G2_SUB_BYTES);
#pragma unroll
        for (int sb = 0; sb < G2_NSUB; ++sb)
          tma_load_2d(stage_out + sb * G2_SUB_BYTES, &tma_res, res_full, n0 + sb * G2_SUB, row0);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (HAS_RES) {
        mbar_wait(res_full, res_phase);
        res_phase ^= 1;
      }
      const int row = row0 + r_local;
      const __nv_bfloat16* grow = nullptr;
      if (EPI == EPI_GATE_RES)
        grow = p.gate + (long long)(((row < p.M ? row : p.M - 1) + p.gate_row_offset) / p.rows_per_gate) * p.gate_stride + n0;
#pragma unroll 1
      for (int sb = 0; sb < G2_NSUB; ++sb) {
        uint32_t v[64];
        tmem_ld32(t_row + sb * G2_SUB, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(t_row + sb * G2_SUB + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
        uint8_t* buf_row = my_row + sb * G2_SUB_BYTES;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4* slot = reinterpret_cast<uint4*>(buf_row + ((c ^ sw) << 4));
          uint4 res = make_uint4(0, 0, 0, 0);
          if (HAS_RES) res = *slot;
          const int col = sb * G2_SUB + c * 8;
          *slot = gemm_epilogue_chunk<EPI>(&v[c * 8], p.bias ? p.bias + n0 + col : nullptr,
                                           EPI == EPI_GATE_RES ? grow + col : nullptr, res);
        }
        fence_proxy_async();
        named_barrier_sync(2, 128);
        if (warp == 2 && elect_one()) {
          tma_store_2d(omap, stage_out + sb * G2_SUB_BYTES, seg_col0 + sb * G2_SUB, row0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_empty0 + acc * 8);
    }
    if (warp == 2 && elect_one()) tma_store_wait_read<0>();
  }

  tc_fence_before();
  cluster_sync_all();   // the peer may still be reading this CTA's smem / TMEM through the pair MMA
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

struct PairMaps {
  CUtensorMap out[3];
  CUtensorMap res;
};

template <int EPI>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const PairMaps& pm, const GemmParams& p,
                        int num_sms, cudaStream_t stream) {
  auto kern = gemm2_bf16_kernel<EPI>;
  static int max_clusters = 0;
  if (max_clusters == 0) {
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, G2_SMEM_BYTES),
                           "cudaFuncSetAttribute(gemm2)"))
      return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(num_sms & ~1);
    cfg.blockDim = dim3(G2_THREADS);
    cfg.dynamicSmemBytes = G2_SMEM_BYTES;
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms / 2;
    }
    max_clusters = n < num_sms / 2 ? n : num_sms / 2;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  GemmParams q = p;
  // stream-K for badly filled last waves (e.g. 114 tiles on 74 pairs = 2 waves at 77 %).  OFF by default: measured on
  // B200 (profiles/, tools/gpu_microbench.py *_streamk) it does not win -- QKV 56.4 -> 58.3 us, O-proj 39.8 -> 46.0 us,
  // FFN2 112.7 -> 112.6 us even with the tail partial preloaded into TMEM ahead of time: a thinner last wave runs
  // faster per tile (shared L2 / power budget), and each split adds a partial hand-off plus one more epilogue.
  // block_n = 513 or SFB_GEMM_STREAMK=1 enables it.
  const int waves = (tiles + clusters - 1) / clusters;
  static const bool sk_env = getenv("SFB_GEMM_STREAMK") != nullptr;
  q.streamk = ((sk_env || p.streamk) && q.sk_ws != nullptr && tiles >= clusters && tiles % clusters != 0 && clusters <= G2_MAX_PAIRS &&
               (double)tiles / ((double)clusters * waves) < 0.95) ? 1 : 0;
  kern<<<2 * clusters, G2_THREADS, G2_SMEM_BYTES, stream>>>(ta, tb, pm.out[0], pm.out[1], pm.out[2], pm.res, q);
  return check_cuda(cudaGetLastError(), "gemm2 launch");
}

// Called by sfb_gemm_bf16 (gemm_tcgen05.cu) for N % 256 == 0 problems.  `p` carries pair-tile counts.
int launch_gemm_pair(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                     cudaStream_t stream) {
  PairMaps pm;
  const uint32_t box[2] = {G2_SUB, G2_ROWS};
  const int nseg = (p.N + p.seg_cols - 1) / p.seg_cols;
  for (int sgm = 0; sgm < 3; ++sgm) {
    const int src = sgm < nseg ? sgm : 0;   // unused slots alias segment 0 (never dereferenced)
    uint64_t dims[2] = {(uint64_t)p.seg_cols, (uint64_t)p.M};
    uint64_t strides[1] = {(uint64_t)p.ldo[src] * 2};
    if (int e = make_tmap_bf16(&pm.out[sgm], p.out[src], 2, dims, strides, box, true)) return e;
  }
  if (epi == EPI_RESIDUAL || epi == EPI_GATE_RES) {
    uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.M};
    uint64_t strides[1] = {(uint64_t)p.ldr * 2};
    if (int e = make_tmap_bf16(&pm.res, p.residual, 2, dims, strides, box, true)) return e;
  } else {
    pm.res = pm.out[0];
  }
  switch (epi) {
    case EPI_BIAS: return launch_gemm2<EPI_BIAS>(ta, tb, pm, p, num_sms, stream);
    case EPI_GELU: return launch_gemm2<EPI_GELU>(ta, tb, pm, p, num_sms, stream);
    case EPI_RESIDUAL: return launch_gemm2<EPI_RESIDUAL>(ta, tb, pm, p, num_sms, stream);
    case EPI_GATE_RES: return launch_gemm2<EPI_GATE_RES>(ta, tb, pm, p, num_sms, stream);
  }
  set_error("sfb_gemm_bf16: unknown epilogue %d", epi);
  return SFB_ERR_INVALID;
}

long long gemm_pair_workspace_bytes() {
  return 1024 + (long long)G2_MAX_PAIRS * 2 * (G2_BN * G2_ROWS * (long long)sizeof(float));   // flags | partial slots
}
// carve the caller's workspace: flags first (zero-initialised once by the caller), then the partial slots
void gemm_pair_workspace(void* ws, long long bytes, GemmParams& p) {
  static int epoch = 0;
  p.sk_ws = nullptr; p.sk_flags = nullptr; p.sk_epoch = 0;
  if (ws == nullptr || bytes < gemm_pair_workspace_bytes()) return;
  p.sk_flags = static_cast<int*>(ws);
  p.sk_ws = reinterpret_cast<float*>(static_cast<char*>(ws) + 1024);
  p.sk_epoch = ++epoch;
}

}  // namespace sfb
