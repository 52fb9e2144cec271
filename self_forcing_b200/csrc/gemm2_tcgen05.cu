// CTA-pair bf16 GEMM for sm_100a: one 256 x 256 output tile per pair of CTAs (tcgen05.mma.cta_group::2, M = 256,
// N = 256, K = 16 issued by the pair's leader).  Each CTA stages its own 128 rows of A and only HALF of B per 64-wide
// k-block (5-stage TMA ring, completion bytes of both CTAs on the leader's mbarrier): with one CTA per 128 x 256 tile
// the kernel was bound by the L2 -> SM fill rate (48 KB per 512 tensor cycles, 44 % tensor-pipe active, profiles/r01a);
// the pair moves 32 KB per SM for the same cycles.
//
// Cluster shape NP = pairs per cluster.  NP = 1 is the default.  NP = 2 (cluster of 4 CTAs) computes a 256 x 512 super
// tile whose two pairs SHARE the A operand: each CTA fetches 64 of its 128 A rows and TMA-multicasts them to the CTA
// with the same in-pair rank of the other pair (24 KB per SM per k-block).  Measured on B200 (profiles/r02c_*): only 33
// such clusters are co-resident (132 of 148 SMs), so it loses on the short-K projections (QKV 62.5 vs 56.3 us, FFN1
// 103 vs 93 us) and wins 4-5 % only on the K = 8960 FFN2 (103.5 vs 107.6 us), which is where the dispatcher uses it.
//
// What bounds it (clock64 timelines, SFB_GEMM_TIMING=1, profiles/r02b_gemm_timeline.log / r02c_*): with hot operands the
// main loop runs at 514 cycles per k-block (floor 512).  Round 1's 4-warp epilogue took 7.2 k (bias) .. 10.1 k (gate +
// residual) .. 14.8 k cycles (GELU) per 128 x 256 tile against a 12.3 k-cycle K = 1536 main loop: FFN1 was epilogue
// bound and the N = 1536 projections exposed a 10 k-cycle tail.  This version: 8 epilogue warps (two per scheduler),
// bias / gate staged in shared memory before the accumulator is ready, hardware tanh in the GELU, one TMA store per
// warp and sub-tile without cross-warp barriers -> 4.0 k / 6.0 k / 8.4 k cycles; FFN1 105 -> 93 us (cuBLAS 91),
// O-proj 34.8 -> 31.8 us.  The remaining gap on N = 1536 is wave quantisation (114 tiles on 74 pairs); a stream-K
// schedule balances the cycles (157 k -> 122 k for FFN2) but not the time: the chip is power-capped, a fully busy
// second wave runs at a lower clock (1.51 -> 1.18 GHz) and the fix-up traffic adds energy.  It was removed.
//
// Roles per CTA (320 threads): warp 0 TMA producer, warp 1 MMA issuer (pair leaders only), warps 2-9 epilogue (TMEM
// double buffered; bias / gate staged in shared memory, residual tiles TMA-prefetched, one TMA store per warp and
// 64-column sub-tile -- see the epilogue section for the measurements behind it).
// Barriers: `full` lives in each pair leader (64 KB per stage: both CTAs' A + B); `empty` lives in every CTA and needs
// one commit from EVERY pair leader of the cluster, because a stage slot is also written by the other pair's multicast.
#include <stdio.h>
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
constexpr int G2_AUX_BYTES = 128 + G2_BN * 8 + 128;   // barriers + TMEM slot | bias [256] + gate [2][256] (bf16)  OR  ln_sc [256] (float2)
// no slack for aligning the dynamic shared memory: the kernel declares it 1024-byte aligned and traps if it is not
constexpr int G2_SMEM_BYTES = G2_STAGES * G2_STAGE_BYTES + G2_NSUB * G2_SUB_BYTES + G2_AUX_BYTES;
constexpr int G2_THREADS = 64 + 256;   // producer warp, MMA warp, 8 epilogue warps
static_assert(G2_SMEM_BYTES <= 227 * 1024, "shared memory budget");

// TMA load issued by a CTA of a pair, multicast to the CTAs of `mask`: the data lands at the same smem offset in every
// destination and the completion bytes are signalled on the mbarrier at the same offset in each destination's PAIR
// LEADER (cta_group::2 semantics: the barrier address carries the leader's parity bit, bit 24 cleared).
__device__ __forceinline__ void tma_load_2d_pair_mc(void* dst, const CUtensorMap* m, uint32_t bar_addr, int c0, int c1,
                                                    uint16_t mask) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster"
      " [%0], [%1, {%3, %4}], [%2], %5;"
      ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(m)), "r"(bar_addr), "r"(c0), "r"(c1), "h"(mask)
      : "memory");
}

// super tiles of this cluster: a contiguous range, m fastest (consecutive clusters share the weight columns in L2)
struct ClusterRange {
  int cur, end;
  __device__ ClusterRange(int tiles, int cid, int num_clusters)
      : cur((int)((long long)tiles * cid / num_clusters)), end((int)((long long)tiles * (cid + 1) / num_clusters)) {}
};

template <int EPI, int NP>
__global__ void __launch_bounds__(G2_THREADS, 1)
gemm2_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                  const __grid_constant__ CUtensorMap tma_out0, const __grid_constant__ CUtensorMap tma_out1,
                  const __grid_constant__ CUtensorMap tma_out2, const __grid_constant__ CUtensorMap tma_res,
                  const GemmParams p) {
  constexpr int CL = 2 * NP;   // CTAs per cluster
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = smem_raw;
  if (smem_u32(smem) & 1023u) __trap();   // the 128-byte swizzle of the TMA boxes / UMMA descriptors needs 1024-byte aligned tiles
  uint8_t* stage_out = smem + G2_STAGES * G2_STAGE_BYTES;   // [G2_NSUB][128 rows][128 B], swizzled
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_out + G2_NSUB * G2_SUB_BYTES);
  uint64_t* empty_bar = full_bar + G2_STAGES;
  uint64_t* tmem_full = empty_bar + G2_STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]  (only the pair leader's copy is waited on)
  uint64_t* res_full = tmem_empty + 2;           // [1]  residual tile landed in the staging buffers
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 1);
  __nv_bfloat16* bias_s = reinterpret_cast<__nv_bfloat16*>(reinterpret_cast<uint8_t*>(full_bar) + 128);   // [256]
  __nv_bfloat16* gate_s = bias_s + G2_BN;                                                                 // [2][256]
  float2* ln_sc_s = reinterpret_cast<float2*>(bias_s);   // [256] (column sum, constant) of the LayerNorm fold; aliases bias / gate (unused then)

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();       // 0 .. CL-1; pairs are (0,1) and (2,3)
  const uint32_t prank = rank & 1;               // rank inside the pair
  const uint32_t pr = rank >> 1;                 // pair index inside the cluster
  const uint32_t leader_rank = rank & ~1u;
  const bool leader = prank == 0;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < G2_STAGES; ++s) {
      mbar_init(&full_bar[s], 1);     // pair leader: one arrive.expect_tx covering both CTAs' bytes
      mbar_init(&empty_bar[s], NP);   // one multicast commit per pair of the cluster per use
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 16);  // 8 epilogue warps x 2 CTAs arrive on the pair leader's barrier
    }
    mbar_init(res_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(tmem_slot, 512);
  tc_fence_before();
  cluster_sync_all();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int cid = blockIdx.x / CL;
  const int num_clusters = gridDim.x / CL;
  const int n_super = (p.num_n_blocks + NP - 1) / NP;
  const int tiles = p.num_m_blocks * n_super;
  // diagnostic timeline (SFB_GEMM_TIMING=1), clock64 relative to this point: [0] first operands landed, [1 + 2i] last
  // MMA of tile i issued, [2 + 2i] accumulator of tile i complete, [8 + i] epilogue of tile i done, [14] stores drained
  long long* dbg = p.dbg ? p.dbg + (long long)blockIdx.x * 16 : nullptr;
  const long long t_base = dbg ? clock64() : 0;
  if (dbg && threadIdx.x == 0) dbg[15] = t_base;

  if (warp == 0) {
    // ------------------------------ TMA producer (every CTA; warp-uniform, one lane issues) ----
    int stage = 0;
    uint32_t phase = 0;
    const uint32_t full0 = mapa_cluster(smem_u32(&full_bar[0]), leader_rank);   // == local address with bit 24 cleared
    const uint16_t a_mask = (uint16_t)((1u << rank) | (1u << (rank ^ 2u)));
    for (ClusterRange r(tiles, cid, num_clusters); r.cur < r.end; ++r.cur) {
      const int m_blk = r.cur % p.num_m_blocks, n_blk = (r.cur / p.num_m_blocks) * NP + (int)pr;
      const int a_row = m_blk * (2 * G2_ROWS) + (int)prank * G2_ROWS + (NP == 2 ? (int)pr * (G2_ROWS / 2) : 0);
      const int b_row = n_blk * G2_BN + (int)prank * (G2_BN / 2);   // beyond N for a phantom pair: TMA zero-fills
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + stage * G2_STAGE_BYTES;
          if (leader) mbar_expect_tx(&full_bar[stage], 2 * G2_STAGE_BYTES);
          const uint32_t bar = full0 + stage * 8;
          if (NP == 2)
            tma_load_2d_pair_mc(a_dst + pr * (G2_A_BYTES / 2), &tma_a, bar, kb * G2_BK, a_row, a_mask);
          else
            tma_load_2d_pair(a_dst, &tma_a, bar, kb * G2_BK, a_row);
          tma_load_2d_pair(a_dst + G2_A_BYTES, &tma_b, bar, kb * G2_BK, b_row);
        }
        __syncwarp();
        if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (pair leaders only) ------------------
    if (leader) {   // warp-uniform; one lane issues
      constexpr uint32_t idesc = umma_idesc_bf16(2 * G2_ROWS, G2_BN, 0, 0);
      constexpr uint16_t all_mask = (uint16_t)((1u << CL) - 1);
      const uint16_t pair_mask = (uint16_t)(3u << leader_rank);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (ClusterRange r(tiles, cid, num_clusters); r.cur < r.end; ++r.cur, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * G2_BN;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (dbg && it == 0 && kb == 0 && lane == 0) dbg[0] = clock64() - t_base;
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + stage * G2_STAGE_BYTES);
            const uint64_t a_desc = umma_desc_sw128(a_addr, 16, 1024);
            const uint64_t b_desc = umma_desc_sw128(a_addr + G2_A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < G2_BK / 16; ++k)
              umma_ss_pair(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb != 0 || k != 0) ? 1u : 0u);
            umma_commit_pair(&empty_bar[stage], all_mask);   // every CTA of the cluster writes into these slots
            if (kb == p.num_k_blocks - 1) umma_commit_pair(&tmem_full[acc], pair_mask);
          }
          __syncwarp();
          if (dbg && kb == p.num_k_blocks - 1 && it < 3 && lane == 0) dbg[1 + 2 * it] = clock64() - t_base;
          if (++stage == G2_STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------ epilogue: 8 warps (every CTA) ----------------
    // Warp (quarter, half) owns accumulator rows [32 quarter, +32) (its TMEM lane quarter) x columns [128 half, +128):
    // two 64-column sub-tiles, each staged in shared memory and written by the warp's OWN TMA store (32-row boxes), so
    // the sub-tile loop has no cross-warp synchronisation.  Two warps per scheduler hide the TMEM / smem latencies.
    constexpr bool HAS_RES = (EPI == EPI_RESIDUAL || EPI == EPI_GATE_RES);
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int r_local = quarter * 32 + lane;                 // accumulator row (TMEM lane) of this thread
    const int tid_e = threadIdx.x - 64;                      // 0 .. 255
    const uint32_t leader_empty0 = mapa_cluster(smem_u32(&tmem_empty[0]), leader_rank);
    uint8_t* my_row = stage_out + r_local * 128;
    const int sw = r_local & 7;                              // 128-byte swizzle: 16-byte chunk index ^ (row % 8)
    if (warp == 2 && elect_one()) {
      tma_prefetch_desc(&tma_out0);
      if (HAS_RES) tma_prefetch_desc(&tma_res);
    }
    uint32_t res_phase = 0;
    int it = 0;
    for (ClusterRange r(tiles, cid, num_clusters); r.cur < r.end; ++r.cur, ++it) {
      const int m_blk = r.cur % p.num_m_blocks, n_blk = (r.cur / p.num_m_blocks) * NP + (int)pr;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row0 = m_blk * (2 * G2_ROWS) + (int)prank * G2_ROWS;
      const int n0 = n_blk * G2_BN;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * G2_BN;
      const bool phantom = n_blk >= p.num_n_blocks;          // odd number of column blocks: this pair only feeds A

      const int seg = phantom ? 0 : n0 / p.seg_cols;
      const int seg_col0 = n0 - seg * p.seg_cols;
      const CUtensorMap* omap = seg == 0 ? &tma_out0 : (seg == 1 ? &tma_out1 : &tma_out2);
      // this warp's previous TMA stores have read their staging rows
      if (elect_one()) tma_store_wait_read<0>();
      __syncwarp();
      if (it > 0) named_barrier_sync(1, 256);   // every warp has finished reading the previous tile's bias / gate / ln_sc vectors
      // bias and gate vectors of this tile -> shared memory (the loads overlap the end of the main loop).  A 128-row tile
      // meets at most two gate vectors unless rows_per_gate is tiny; that rare case reads the gate from global memory.
      int g_lo = 0;
      bool gate_smem = false;
      if (EPI == EPI_GATE_RES) {
        const int rl = row0 < p.M ? row0 : p.M - 1, rh = row0 + G2_ROWS - 1 < p.M ? row0 + G2_ROWS - 1 : p.M - 1;
        g_lo = (rl + p.gate_row_offset) / p.rows_per_gate;
        const int g_hi = (rh + p.gate_row_offset) / p.rows_per_gate;
        gate_smem = g_hi - g_lo <= 1;
        if (gate_smem && !phantom) {
          gate_s[tid_e] = p.gate[(long long)g_lo * p.gate_stride + n0 + tid_e];
          gate_s[G2_BN + tid_e] = p.gate[(long long)g_hi * p.gate_stride + n0 + tid_e];
        }
      }
      if (p.bias != nullptr && !phantom) bias_s[tid_e] = p.bias[n0 + tid_e];
      // LayerNorm folded into this GEMM: mean / rstd of this thread's input row (overlaps the end of the main loop)
      const bool do_ln = EPI == EPI_BIAS && p.ln_stats != nullptr;
      float ln_mean = 0.f, ln_rstd = 1.f;
      if (do_ln && !phantom) {
        ln_sc_s[tid_e] = __ldg(p.ln_sc + n0 + tid_e);
        const int rr = row0 + r_local < p.M ? row0 + r_local : p.M - 1;
        stats_mean_rstd(p.ln_stats + (long long)rr * (p.K / STATS_CHUNK), p.K / STATS_CHUNK, p.ln_eps, ln_mean, ln_rstd);
      }
      named_barrier_sync(1, 256);    // staging buffers free in every warp; bias / gate visible
      if (HAS_RES && !phantom && warp == 2 && elect_one()) {
        mbar_expect_tx(res_full, G2_NSUB * G2_SUB_BYTES);
#pragma unroll
        for (int sb = 0; sb < G2_NSUB; ++sb)
          tma_load_2d(stage_out + sb * G2_SUB_BYTES, &tma_res, res_full, n0 + sb * G2_SUB, row0);
      }
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      if (dbg && it < 3 && threadIdx.x == 64) dbg[2 + 2 * it] = clock64() - t_base;
      if (!phantom) {
        if (HAS_RES) {
          mbar_wait(res_full, res_phase);
          res_phase ^= 1;
        }
        const int row = row0 + r_local;
        const int grow_idx = EPI == EPI_GATE_RES ? ((row < p.M ? row : p.M - 1) + p.gate_row_offset) / p.rows_per_gate : 0;
        const __nv_bfloat16* gsm = gate_s + (grow_idx - g_lo) * G2_BN;                    // gate_smem
        const __nv_bfloat16* ggl = EPI == EPI_GATE_RES ? p.gate + (long long)grow_idx * p.gate_stride + n0 : nullptr;
        const bool has_bias = p.bias != nullptr;
        // statistics of this thread's 128 output columns, shifted by the first value (x0) so that a large row mean
        // does not cancel: sd = sum (x - x0), sd2 = sum (x - x0)^2
        const bool do_stats = p.stats_out != nullptr;
        float x0 = 0.f, sd = 0.f, sd2 = 0.f;
#pragma unroll 1
        for (int j = 0; j < 2; ++j) {
          const int sb = half * 2 + j;
          uint32_t v[64];
          tmem_ld32(t_row + sb * G2_SUB, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
          tmem_ld32(t_row + sb * G2_SUB + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
          tmem_ld_wait();
          uint8_t* buf_row = my_row + sb * G2_SUB_BYTES;
#pragma unroll
          for (int c = 0; c < 8; ++c) {
            uint4* slot = reinterpret_cast<uint4*>(buf_row + ((c ^ sw) << 4));
            const int col = sb * G2_SUB + c * 8;
            uint4 res = make_uint4(0, 0, 0, 0), bq = make_uint4(0, 0, 0, 0), gq = make_uint4(0, 0, 0, 0);
            if (HAS_RES) res = *slot;
            if (has_bias) bq = *reinterpret_cast<const uint4*>(bias_s + col);
            if (EPI == EPI_GATE_RES)
              gq = gate_smem ? *reinterpret_cast<const uint4*>(gsm + col) : __ldg(reinterpret_cast<const uint4*>(ggl + col));
            if (do_ln) {
              const float4* scp = reinterpret_cast<const float4*>(ln_sc_s + col);   // (colsum, const) pairs, warp-uniform address: smem broadcast
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const float4 sc2 = scp[q4];
                v[c * 8 + 2 * q4] = __float_as_uint(fmaf(ln_rstd, fmaf(-ln_mean, sc2.x, __uint_as_float(v[c * 8 + 2 * q4])), sc2.y));
                v[c * 8 + 2 * q4 + 1] = __float_as_uint(fmaf(ln_rstd, fmaf(-ln_mean, sc2.z, __uint_as_float(v[c * 8 + 2 * q4 + 1])), sc2.w));
              }
            }
            const uint4 o8 = gemm_epilogue_vals<EPI, true>(&v[c * 8], has_bias, bq, gq, res);
            *slot = o8;
            if (do_stats) {
              const uint32_t ow[4] = {o8.x, o8.y, o8.z, o8.w};
              if (j == 0 && c == 0) x0 = bf_lo(ow[0]);
#pragma unroll
              for (int q4 = 0; q4 < 4; ++q4) {
                const float d0 = bf_lo(ow[q4]) - x0, d1 = bf_hi(ow[q4]) - x0;
                sd += d0 + d1;
                sd2 = fmaf(d0, d0, fmaf(d1, d1, sd2));
              }
            }
          }
          fence_proxy_async();
          __syncwarp();
          if (row0 + quarter * 32 < p.M && elect_one()) {
            tma_store_2d(omap, stage_out + sb * G2_SUB_BYTES + quarter * (32 * 128), seg_col0 + sb * G2_SUB, row0 + quarter * 32);
            tma_store_commit();
          }
        }
        if (do_stats && row < p.M) {
          const float inv = 1.0f / (float)STATS_CHUNK;
          p.stats_out[(long long)row * (p.N / STATS_CHUNK) + n_blk * (G2_BN / STATS_CHUNK) + half] =
              make_float2(x0 + sd * inv, fmaxf(sd2 - sd * sd * inv, 0.f));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive_cluster(leader_empty0 + acc * 8);
      if (dbg && it < 3 && threadIdx.x == 64) dbg[8 + it] = clock64() - t_base;
    }
    if (elect_one()) tma_store_wait_read<0>();
    if (dbg && threadIdx.x == 64) dbg[14] = clock64() - t_base;
  }

  tc_fence_before();
  cluster_sync_all();   // peers may still be reading this CTA's smem / TMEM through the pair MMA or writing its barriers
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc_pair(tmem_base, 512);
  }
}

struct PairMaps {
  CUtensorMap out[3];
  CUtensorMap res;
};

template <int EPI, int NP>
static int launch_gemm2(const CUtensorMap& ta, const CUtensorMap& tb, const PairMaps& pm, const GemmParams& p, int num_sms,
                        cudaStream_t stream) {
  constexpr int CL = 2 * NP;
  auto kern = gemm2_bf16_kernel<EPI, NP>;
  static SmemOptIn optin;
  if (int e = optin.ensure(kern, G2_SMEM_BYTES, "cudaFuncSetAttribute(gemm2)")) return e;
  static std::atomic<int> max_clusters_dev[64];   // co-resident clusters, per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) dev = 0;
  cudaLaunchConfig_t cfg = {};
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CL;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.blockDim = dim3(G2_THREADS);
  cfg.dynamicSmemBytes = G2_SMEM_BYTES;
  cfg.stream = stream;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int max_clusters = max_clusters_dev[dev].load(std::memory_order_acquire);
  if (max_clusters == 0) {
    cfg.gridDim = dim3((num_sms / CL) * CL);
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n <= 0) {
      cudaGetLastError();
      n = num_sms / CL;
    }
    max_clusters = n < num_sms / CL ? n : num_sms / CL;
    max_clusters_dev[dev].store(max_clusters, std::memory_order_release);
  }
  const int tiles = p.num_m_blocks * ((p.num_n_blocks + NP - 1) / NP);
  const int clusters = tiles < max_clusters ? tiles : max_clusters;
  GemmParams q = p;
  static long long* dbg_buf = nullptr;   // diagnostic mode only (SFB_GEMM_TIMING=1): synchronous, prints a timeline
  static const bool timing = getenv("SFB_GEMM_TIMING") != nullptr;
  if (timing && dbg_buf == nullptr) cudaMalloc(&dbg_buf, 512 * 16 * sizeof(long long));
  q.dbg = timing ? dbg_buf : nullptr;
  if (timing) cudaMemsetAsync(dbg_buf, 0, 512 * 16 * sizeof(long long), stream);
  if (int e = check_cuda(launch_cluster(kern, dim3(CL * clusters), dim3(G2_THREADS), G2_SMEM_BYTES, stream, CL, ta, tb, pm.out[0],
                                    pm.out[1], pm.out[2], pm.res, q), "gemm2 launch"))
    return e;
  if (timing) {
    static long long h[512 * 16];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_buf, sizeof(h), cudaMemcpyDeviceToHost);
    long long base_min = h[15];
    for (int c = 0; c < CL * clusters; ++c) base_min = h[c * 16 + 15] < base_min ? h[c * 16 + 15] : base_min;
    fprintf(stderr, "[gemm2 timing] M=%d N=%d K=%d epi=%d NP=%d super_tiles=%d clusters=%d of max %d (clock64 cycles since CTA start)\n",
            p.M, p.N, p.K, EPI, NP, tiles, clusters, max_clusters);
    for (int c : {0, CL * (clusters / 2), CL * (clusters - 1)}) {
      const long long* t = h + c * 16;
      fprintf(stderr, "  cta %3d start+%lld: first_data=%lld | mma_end %lld %lld %lld | acc_ready %lld %lld %lld | epi_end %lld %lld %lld | all_stores_read=%lld\n",
              c, t[15] - base_min, t[0], t[1], t[3], t[5], t[2], t[4], t[6], t[8], t[9], t[10], t[14]);
    }
  }
  return SFB_OK;
}

template <int NP>
static int dispatch_gemm2(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const PairMaps& pm, const GemmParams& p,
                          int num_sms, cudaStream_t stream) {
  switch (epi) {
    case EPI_BIAS: return launch_gemm2<EPI_BIAS, NP>(ta, tb, pm, p, num_sms, stream);
    case EPI_GELU: return launch_gemm2<EPI_GELU, NP>(ta, tb, pm, p, num_sms, stream);
    case EPI_RESIDUAL: return launch_gemm2<EPI_RESIDUAL, NP>(ta, tb, pm, p, num_sms, stream);
    case EPI_GATE_RES: return launch_gemm2<EPI_GATE_RES, NP>(ta, tb, pm, p, num_sms, stream);
  }
  set_error("sfb_gemm_bf16: unknown epilogue %d", epi);
  return SFB_ERR_INVALID;
}

// Called by sfb_gemm_bf16 (gemm_tcgen05.cu).  `p` carries pair-tile counts (256-row m blocks, 256-column n blocks);
// x / ldx are needed again because the A box height depends on the cluster shape.
int launch_gemm_cluster(int epi, int pairs_per_cluster, const void* x, long long ldx, const CUtensorMap& tb,
                        const GemmParams& p, int num_sms, cudaStream_t stream) {
  PairMaps pm;
  const uint32_t box[2] = {G2_SUB, G2_ROWS};
  const uint32_t obox[2] = {G2_SUB, 32};   // one store per epilogue warp
  const int nseg = (p.N + p.seg_cols - 1) / p.seg_cols;
  for (int sgm = 0; sgm < 3; ++sgm) {
    const int src = sgm < nseg ? sgm : 0;   // unused slots alias segment 0 (never dereferenced)
    uint64_t dims[2] = {(uint64_t)p.seg_cols, (uint64_t)p.M};
    uint64_t strides[1] = {(uint64_t)p.ldo[src] * 2};
    if (int e = make_tmap_bf16(&pm.out[sgm], p.out[src], 2, dims, strides, obox, true)) return e;
  }
  if (epi == EPI_RESIDUAL || epi == EPI_GATE_RES) {
    uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.M};
    uint64_t strides[1] = {(uint64_t)p.ldr * 2};
    if (int e = make_tmap_bf16(&pm.res, p.residual, 2, dims, strides, box, true)) return e;
  } else {
    pm.res = pm.out[0];
  }
  CUtensorMap ta;
  {
    uint64_t dims[2] = {(uint64_t)p.K, (uint64_t)p.M};
    uint64_t strides[1] = {(uint64_t)ldx * 2};
    uint32_t abox[2] = {G2_BK, (uint32_t)(G2_ROWS / pairs_per_cluster)};
    if (int e = make_tmap_bf16(&ta, x, 2, dims, strides, abox, true)) return e;
  }
  if (pairs_per_cluster == 2) return dispatch_gemm2<2>(epi, ta, tb, pm, p, num_sms, stream);
  return dispatch_gemm2<1>(epi, ta, tb, pm, p, num_sms, stream);
}

}  // namespace sfb
