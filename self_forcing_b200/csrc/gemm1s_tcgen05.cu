// One-CTA bf16 GEMM with the CTA-pair kernel's staged TMA-store epilogue: 128 x BN output tiles, BN in {192, 256}.
//
// Why it exists: the N = 1536 projections (O, cross-Q, cross-O, FFN2 -- half of the linear
// FLOPs, 57 % of the GEMM time) are 19 x 6 = 114 pair tiles on 74 CTA pairs = two waves at 77 %; with 128 x 192
// one-CTA tiles the same output is 37 x 8 = 296 tiles = EXACTLY two waves of 148 SMs.  The plain one-CTA kernel
// (gemm_tcgen05.cu) cannot show that because its per-thread row-strided epilogue costs 24-54 k cycles per tile; this
// kernel stages the accumulator rows in shared memory (BN / 64 sub-tiles of 128 x 64 bf16, 128-byte swizzle), prefetches
// the residual tile into the same buffers by TMA and writes with TMA stores, exactly like gemm2_tcgen05.cu.  The
// implicit-GEMM convolution (same one-CTA structure, long K) sustains 92 % tensor-pipe active, so the main loop itself is
// not the obstacle.
//
// Selected only on explicit request (sfb_gemm_bf16 block_n = 1192 / 1256) -- NOT YET VALIDATED ON HARDWARE: written
// after round 1's GPU budget was spent; parity checks sit in tests/gpu_checks.py: PENDING, timings in
// tools/gpu_microbench.py (*_1s192).
#include "gemm_common.cuh"

namespace sfb {

constexpr int G1_ROWS = 128;
constexpr int G1_BK = 64;
constexpr int G1_SUB = 64;                       // staging sub-tile width (one 128-byte swizzle row of bf16)
constexpr int G1_SUB_BYTES = G1_ROWS * G1_SUB * 2;
constexpr int G1_THREADS = 192;

template <int BN>
struct G1Cfg {
  static constexpr int NSUB = BN / G1_SUB;
  static constexpr int A_BYTES = G1_ROWS * G1_BK * 2;
  static constexpr int B_BYTES = BN * G1_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = BN == 256 ? 3 : 4;
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + NSUB * G1_SUB_BYTES + 1024 + 256;
  static_assert(BN % G1_SUB == 0 && 2 * BN <= 512, "tile does not fit");
  static_assert(SMEM_BYTES <= 227 * 1024, "shared memory budget");
};

template <int BN, int EPI>
__global__ void __launch_bounds__(G1_THREADS, 1)
gemm1s_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                   const __grid_constant__ CUtensorMap tma_out0, const __grid_constant__ CUtensorMap tma_out1,
                   const __grid_constant__ CUtensorMap tma_out2, const __grid_constant__ CUtensorMap tma_res,
                   const GemmParams p) {
  using Cfg = G1Cfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int NSUB = Cfg::NSUB;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stage_out = smem + STAGES * Cfg::STAGE_BYTES;     // [NSUB][128 rows][128 B], swizzled
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(stage_out + NSUB * G1_SUB_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint64_t* res_full = tmem_empty + 2;        // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(res_full + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_a);
    tma_prefetch_desc(&tma_b);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);   // one arrival per epilogue warp
    }
    mbar_init(res_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0) {
    // ------------------------------ TMA producer --------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m_blk = tile % p.num_m_blocks, n_blk = tile / p.num_m_blocks;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
          mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          tma_load_2d(a_dst, &tma_a, &full_bar[stage], kb * G1_BK, m_blk * G1_ROWS);
          tma_load_2d(a_dst + Cfg::A_BYTES, &tma_b, &full_bar[stage], kb * G1_BK, n_blk * BN);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ----------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(G1_ROWS, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BN;
      for (int kb = 0; kb < p.num_k_blocks; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t a_desc = umma_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_desc = umma_desc_sw128(a_addr + Cfg::A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < G1_BK / 16; ++k) umma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          umma_commit(&empty_bar[stage]);
          if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------ epilogue warps (same scheme as gemm2_tcgen05.cu) ----
    constexpr bool HAS_RES = (EPI == EPI_RESIDUAL || EPI == EPI_GATE_RES);
    const int quarter = warp & 3;
    const int r_local = quarter * 32 + lane;                 // accumulator row (TMEM lane) of this thread
    uint8_t* my_row = stage_out + r_local * 128;
    const int sw = r_local & 7;                              // 128-byte swizzle: 16-byte chunk index ^ (row % 8)
    if (warp == 2 && elect_one()) {
      tma_prefetch_desc(&tma_out0);
      if (HAS_RES) tma_prefetch_desc(&tma_res);
    }
    uint32_t res_phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile % p.num_m_blocks, n_blk = tile / p.num_m_blocks;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int row0 = m_blk * G1_ROWS;
      const int n0 = n_blk * BN;
      const uint32_t t_row = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BN;
      const int seg = n0 / p.seg_cols;
      const int seg_col0 = n0 - seg * p.seg_cols;
      const CUtensorMap* omap = seg == 0 ? &tma_out0 : (seg == 1 ? &tma_out1 : &tma_out2);
      // staging buffers are free once the previous tile's TMA stores have read them
      if (warp == 2 && elect_one()) tma_store_wait_read<0>();
      named_barrier_sync(1, 128);
      if (HAS_RES && warp == 2 && elect_one()) {
        mbar_expect_tx(res_full, NSUB * G1_SUB_BYTES);
#pragma unroll
        for (int sb = 0; sb < NSUB; ++sb)
          tma_load_2d(stage_out + sb * G1_SUB_BYTES, &tma_res, res_full, n0 + sb * G1_SUB, row0);
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
      for (int sb = 0; sb < NSUB; ++sb) {
        uint32_t v[64];
        tmem_ld32(t_row + sb * G1_SUB, *reinterpret_cast<uint32_t(*)[32]>(&v[0]));
        tmem_ld32(t_row + sb * G1_SUB + 32, *reinterpret_cast<uint32_t(*)[32]>(&v[32]));
        tmem_ld_wait();
        uint8_t* buf_row = my_row + sb * G1_SUB_BYTES;
#pragma unroll
        for (int c = 0; c < 8; ++c) {
          uint4* slot = reinterpret_cast<uint4*>(buf_row + ((c ^ sw) << 4));
          uint4 res = make_uint4(0, 0, 0, 0);
          if (HAS_RES) res = *slot;
          const int col = sb * G1_SUB + c * 8;
          *slot = gemm_epilogue_chunk<EPI>(&v[c * 8], p.bias ? p.bias + n0 + col : nullptr,
                                           EPI == EPI_GATE_RES ? grow + col : nullptr, res);
        }
        fence_proxy_async();
        named_barrier_sync(2, 128);
        if (warp == 2 && elect_one()) {
          tma_store_2d(omap, stage_out + sb * G1_SUB_BYTES, seg_col0 + sb * G1_SUB, row0);
          tma_store_commit();
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
    if (warp == 2 && elect_one()) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

template <int BN, int EPI>
static int launch_g1(const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* outs, const CUtensorMap& res,
                     const GemmParams& p, int num_sms, cudaStream_t stream) {
  using Cfg = G1Cfg<BN>;
  auto kern = gemm1s_bf16_kernel<BN, EPI>;
  static bool attr_set = false;
  if (!attr_set) {
    if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES),
                           "cudaFuncSetAttribute(gemm1s)"))
      return e;
    attr_set = true;
  }
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, G1_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, outs[0], outs[1], outs[2], res, p);
  return check_cuda(cudaGetLastError(), "gemm1s launch");
}

template <int BN>
static int dispatch_g1(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const CUtensorMap* outs, const CUtensorMap& res,
                       const GemmParams& p, int num_sms, cudaStream_t stream) {
  switch (epi) {
    case EPI_BIAS: return launch_g1<BN, EPI_BIAS>(ta, tb, outs, res, p, num_sms, stream);
    case EPI_GELU: return launch_g1<BN, EPI_GELU>(ta, tb, outs, res, p, num_sms, stream);
    case EPI_RESIDUAL: return launch_g1<BN, EPI_RESIDUAL>(ta, tb, outs, res, p, num_sms, stream);
    case EPI_GATE_RES: return launch_g1<BN, EPI_GATE_RES>(ta, tb, outs, res, p, num_sms, stream);
  }
  set_error("sfb_gemm_bf16: epilogue %d is not available with the staged one-CTA tiles", epi);
  return SFB_ERR_INVALID;
}

// Called by sfb_gemm_bf16 for block_n = 1192 / 1256.  `p` carries one-CTA tile counts for N tile `bn`; ta / tb are the
// operand maps with boxes {64, 128} / {64, bn}.
int launch_gemm_single_staged(int epi, int bn, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                              cudaStream_t stream) {
  CUtensorMap outs[3], res;
  const uint32_t box[2] = {G1_SUB, G1_ROWS};
  const int nseg = (p.N + p.seg_cols - 1) / p.seg_cols;
  for (int s = 0; s < 3; ++s) {
    const int src = s < nseg ? s : 0;
    uint64_t dims[2] = {(uint64_t)p.seg_cols, (uint64_t)p.M};
    uint64_t strides[1] = {(uint64_t)p.ldo[src] * 2};
    if (int e = make_tmap_bf16(&outs[s], p.out[src], 2, dims, strides, box, true)) return e;
  }
  if (epi == EPI_RESIDUAL || epi == EPI_GATE_RES) {
    uint64_t dims[2] = {(uint64_t)p.N, (uint64_t)p.M};
    uint64_t strides[1] = {(uint64_t)p.ldr * 2};
    if (int e = make_tmap_bf16(&res, p.residual, 2, dims, strides, box, true)) return e;
  } else {
    res = outs[0];
  }
  if (bn == 192) return dispatch_g1<192>(epi, ta, tb, outs, res, p, num_sms, stream);
  return dispatch_g1<256>(epi, ta, tb, outs, res, p, num_sms, stream);
}

}  // namespace sfb
