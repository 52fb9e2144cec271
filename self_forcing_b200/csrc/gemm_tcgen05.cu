// Persistent, warp-specialised bf16 GEMM for sm_100a:  Y[M,N] = epilogue(X[M,K] . W[N,K]^T + bias)
//
// Replaces the nn.Linear call sites of the reference hot path (cuBLAS there):
//   QKV / O / cross-Q / FFN / head / patch-embed / text-embed projections
//   (wan/modules/causal_model.py:80-83,112-114,240,277-279,351,458-462; wan/modules/model.py:172-193)
// with the elementwise ops that follow them fused into the epilogue, keeping the reference's bf16
// rounding points:
//   EPI_BIAS       y = bf16(acc + b)
//   EPI_GELU       y = bf16(gelu_tanh(bf16(acc + b)))                       (causal_model.py:278)
//   EPI_RESIDUAL   y = bf16(res + bf16(acc + b))                            (causal_model.py:324)
//   EPI_GATE_RES   y = bf16(res + bf16(bf16(acc + b) * gate[row / rows_per_gate]))  (:320, :331-332)
//
// Structure (one CTA per SM, 192 threads):
//   warp 0   : TMA producer  -- cp.async.bulk.tensor (128B swizzle) into a STAGES-deep smem ring
//   warp 1   : MMA issuer    -- tcgen05.mma cta_group::1, M=128, N=BLOCK_N, K=16, fp32 accum in TMEM
//   warps 2-5: epilogue      -- tcgen05.ld the accumulator (double-buffered in TMEM so the next
//                               tile's main loop overlaps), fused epilogue, 16-byte global stores
#include "gemm_common.cuh"

namespace sfb {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 64;   // 64 bf16 = one 128-byte swizzle row
constexpr int UMMA_K = 16;
constexpr int GEMM_THREADS = 192;

template <int BLOCK_N>
struct GemmCfg {
  static constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BLOCK_N == 256) ? 4 : (BLOCK_N == 128 ? 6 : 8);
  // two accumulator buffers; TMEM is allocated in powers of two
  static constexpr int TMEM_COLS = 2 * BLOCK_N <= 32 ? 32 : (2 * BLOCK_N <= 128 ? 128 : (2 * BLOCK_N <= 256 ? 256 : 512));
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
};

template <int BLOCK_N, int EPI>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tma_a, const __grid_constant__ CUtensorMap tma_b,
                 const GemmParams p) {
  using Cfg = GemmCfg<BLOCK_N>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tmem_full = empty_bar + STAGES;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;       // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

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
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.num_m_blocks * p.num_n_blocks;

  if (warp == 0) {
    // ------------------------------ TMA producer (warp-uniform, one lane issues) ----
    {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int m_blk = tile % p.num_m_blocks, n_blk = tile / p.num_m_blocks;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_2d(a_dst, &tma_a, &full_bar[stage], kb * BLOCK_K, m_blk * BLOCK_M);
            tma_load_2d(a_dst + Cfg::A_BYTES, &tma_b, &full_bar[stage], kb * BLOCK_K, n_blk * BLOCK_N);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer (warp-uniform, one lane issues) -----
    {
      constexpr uint32_t idesc = umma_idesc_bf16(BLOCK_M, BLOCK_N, 0, 0);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < p.num_k_blocks; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          if (elect_one()) {
            const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
            const uint64_t a_desc = umma_desc_sw128(a_addr, 16, 1024);
            const uint64_t b_desc = umma_desc_sw128(a_addr + Cfg::A_BYTES, 16, 1024);
#pragma unroll
            for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
              // advance 32 bytes (16 bf16) inside the 128-byte swizzle row: +2 in the >>4 address field
              umma_ss(d_tmem, a_desc + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
            }
            umma_commit(&empty_bar[stage]);   // smem slot reusable once these MMAs retire
            if (kb == p.num_k_blocks - 1) umma_commit(&tmem_full[acc]);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else {
    // ------------------------------ epilogue warps ----------------------------
    const int quarter = warp & 3;   // TMEM lane quarter this warp may access
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int m_blk = tile % p.num_m_blocks, n_blk = tile / p.num_m_blocks;
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const int row = m_blk * BLOCK_M + quarter * 32 + lane;
      gemm_epilogue_row<BLOCK_N, EPI>(p, row, n_blk * BLOCK_N, tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * BLOCK_N);
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

template <int BLOCK_N, int EPI>
static int launch_gemm(const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                       cudaStream_t stream) {
  using Cfg = GemmCfg<BLOCK_N>;
  auto kern = gemm_bf16_kernel<BLOCK_N, EPI>;
  static SmemOptIn optin;
  if (int e = optin.ensure(kern, Cfg::SMEM_BYTES, "cudaFuncSetAttribute(gemm)")) return e;
  const int tiles = p.num_m_blocks * p.num_n_blocks;
  const int grid = tiles < num_sms ? tiles : num_sms;
  kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(ta, tb, p);
  return check_cuda(cudaGetLastError(), "gemm launch");
}

template <int BLOCK_N>
static int dispatch_epi(int epi, const CUtensorMap& ta, const CUtensorMap& tb, const GemmParams& p, int num_sms,
                        cudaStream_t stream) {
  switch (epi) {
    case EPI_BIAS: return launch_gemm<BLOCK_N, EPI_BIAS>(ta, tb, p, num_sms, stream);
    case EPI_GELU: return launch_gemm<BLOCK_N, EPI_GELU>(ta, tb, p, num_sms, stream);
    case EPI_RESIDUAL: return launch_gemm<BLOCK_N, EPI_RESIDUAL>(ta, tb, p, num_sms, stream);
    case EPI_GATE_RES: return launch_gemm<BLOCK_N, EPI_GATE_RES>(ta, tb, p, num_sms, stream);
    case EPI_F32: return launch_gemm<BLOCK_N, EPI_F32>(ta, tb, p, num_sms, stream);
  }
  set_error("sfb_gemm_bf16: unknown epilogue %d", epi);
  return SFB_ERR_INVALID;
}

int device_sm_count();
int launch_gemm_cluster(int epi, int pairs_per_cluster, const void* x, long long ldx, const CUtensorMap& tb,
                        const GemmParams& p, int num_sms, cudaStream_t stream);   // gemm2_tcgen05.cu

}  // namespace sfb

// sfb_gemm_bf16 plus row statistics (include/sfb200.h): stats_out = statistics of the output rows, ln_stats / ln_sc =
// LayerNorm of the input folded into the epilogue.  Both need the CTA-pair kernel (N, seg_cols multiples of 256).
extern "C" int sfb_gemm_bf16_stats(const void* x, long long ldx, const void* w, long long ldw, const void* bias,
                                   int M, int N, int K, int epilogue, void* out0, long long ldo0, void* out1,
                                   long long ldo1, void* out2, long long ldo2, int seg_cols, const void* residual,
                                   long long ldr, const void* gate, long long gate_stride, int rows_per_gate,
                                   int gate_row_offset, int block_n, void* stats_out, const void* ln_stats,
                                   const void* ln_sc, float ln_eps, void* stream_) {
  using namespace sfb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (M <= 0 || N <= 0 || K <= 0) { set_error("sfb_gemm_bf16: empty problem M=%d N=%d K=%d", M, N, K); return SFB_ERR_INVALID; }
  if ((K % 8) || (N % 8) || (ldx % 8) || (ldw % 8)) {
    set_error("sfb_gemm_bf16: K, N, ldx, ldw must be multiples of 8 (16-byte rows); got K=%d N=%d ldx=%lld ldw=%lld",
              K, N, ldx, ldw);
    return SFB_ERR_INVALID;
  }
  if (seg_cols <= 0) seg_cols = N;
  // block_n: 0 = choose; 64 / 128 / 256 = one-CTA tiles of 128 x block_n; 512 = CTA-pair tiles of 256 x 256 (tcgen05
  // cta_group::2, gemm2_tcgen05.cu) -- the default whenever N and the segments are multiples of 256; 515 = pair tiles in
  // clusters of two pairs sharing A by TMA multicast (default for long-K, narrow-N problems: FFN2).
  if (block_n == 0 && N % 256 == 0 && seg_cols % 256 == 0 && M > 128 && epilogue != EPI_F32) {
    block_n = 512;
    if (K >= 4096 && N <= 2048 && M >= 2048) {
      // clusters of two pairs (A multicast) win 4-5 % on long-K problems at equal SM fill (measured: 33 of them are co-resident
      // on 148 SMs); choose them unless whole pairs fill the machine better (M = 9360, N = 1536: 222 tiles = 3 exact rounds
      // of 74 pairs against 111 super tiles on 33 clusters)
      const int sms_ = device_sm_count();
      const int pairs = sms_ / 2, quads = sms_ * 33 / 148 > 0 ? sms_ * 33 / 148 : 1;
      const int mb = (M + 255) / 256, nb = N / 256;
      const int t1 = mb * nb, t2 = mb * ((nb + 1) / 2);
      const double eff1 = (double)t1 / (((t1 + pairs - 1) / pairs) * pairs);
      const double eff2 = (double)t2 / (((t2 + quads - 1) / quads) * quads) * (4.0 * quads / sms_) * ((double)nb / (2 * ((nb + 1) / 2)));
      if (eff2 >= 0.98 * eff1) block_n = 515;
    }
  }
  if (block_n == 0) {
    // narrow problems take 128-wide tiles so the tile count fills the SMs
    block_n = (N % 256 == 0 && N >= 4096) ? 256 : (N % 128 == 0 ? 128 : 64);
  }
  const int cluster_pairs = block_n == 512 ? 1 : (block_n == 515 ? 2 : 0);
  const bool pair = cluster_pairs != 0;
  if (pair) block_n = 256;
  if (block_n != 64 && block_n != 128 && block_n != 256) { set_error("sfb_gemm_bf16: block_n must be 0/64/128/256/512/515"); return SFB_ERR_INVALID; }
  if (epilogue == EPI_F32 && (pair || seg_cols != N)) { set_error("sfb_gemm_bf16: the fp32-output epilogue takes one-CTA tiles and one output segment"); return SFB_ERR_INVALID; }
  if (pair && (N % 256 || seg_cols % 256)) { set_error("sfb_gemm_bf16: pair tiles need N and seg_cols to be multiples of 256"); return SFB_ERR_INVALID; }
  if (seg_cols % block_n && seg_cols != N) { set_error("sfb_gemm_bf16: seg_cols=%d not a multiple of the N tile %d", seg_cols, block_n); return SFB_ERR_INVALID; }
  if ((epilogue == EPI_RESIDUAL || epilogue == EPI_GATE_RES) && residual == nullptr) { set_error("sfb_gemm_bf16: residual epilogue without residual"); return SFB_ERR_INVALID; }
  if (epilogue == EPI_GATE_RES && (gate == nullptr || rows_per_gate <= 0)) { set_error("sfb_gemm_bf16: gate epilogue without gate"); return SFB_ERR_INVALID; }

  GemmParams p{};
  p.M = M; p.N = N; p.K = K;
  p.num_m_blocks = pair ? (M + 2 * BLOCK_M - 1) / (2 * BLOCK_M) : (M + BLOCK_M - 1) / BLOCK_M;
  p.num_n_blocks = (N + block_n - 1) / block_n;
  p.num_k_blocks = (K + BLOCK_K - 1) / BLOCK_K;
  p.bias = static_cast<const __nv_bfloat16*>(bias);
  p.out[0] = static_cast<__nv_bfloat16*>(out0); p.ldo[0] = ldo0;
  p.out[1] = static_cast<__nv_bfloat16*>(out1); p.ldo[1] = ldo1;
  p.out[2] = static_cast<__nv_bfloat16*>(out2); p.ldo[2] = ldo2;
  p.seg_cols = seg_cols;
  const int nseg = (N + seg_cols - 1) / seg_cols;
  if (nseg > 3) { set_error("sfb_gemm_bf16: at most 3 output segments"); return SFB_ERR_INVALID; }
  for (int s = 0; s < nseg; ++s)
    if (p.out[s] == nullptr || (p.ldo[s] % 8)) { set_error("sfb_gemm_bf16: output segment %d missing or ld not multiple of 8", s); return SFB_ERR_INVALID; }
  p.residual = static_cast<const __nv_bfloat16*>(residual); p.ldr = ldr;
  p.gate = static_cast<const __nv_bfloat16*>(gate); p.gate_stride = gate_stride; p.rows_per_gate = rows_per_gate;
  p.gate_row_offset = gate_row_offset;
  p.stats_out = static_cast<float2*>(stats_out);
  p.ln_stats = static_cast<const float2*>(ln_stats);
  p.ln_sc = static_cast<const float2*>(ln_sc);
  p.ln_eps = ln_eps;
  if ((stats_out != nullptr || ln_stats != nullptr) && !pair) { set_error("sfb_gemm_bf16_stats: row statistics need the CTA-pair kernel (N and seg_cols multiples of 256, M > 128)"); return SFB_ERR_INVALID; }
  if (ln_stats != nullptr && (epilogue != EPI_BIAS || bias != nullptr || ln_sc == nullptr || K % STATS_CHUNK)) {
    set_error("sfb_gemm_bf16_stats: the LayerNorm fold takes the plain epilogue, no bias (it lives in ln_sc), and K a multiple of %d", STATS_CHUNK);
    return SFB_ERR_INVALID;
  }

  CUtensorMap ta, tb;
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)M};
    uint64_t strides[1] = {(uint64_t)ldx * 2};
    uint32_t box[2] = {BLOCK_K, BLOCK_M};
    if (int e = make_tmap_bf16(&ta, x, 2, dims, strides, box, true)) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)K, (uint64_t)N};
    uint64_t strides[1] = {(uint64_t)ldw * 2};
    uint32_t box[2] = {BLOCK_K, (uint32_t)(pair ? block_n / 2 : block_n)};
    if (int e = make_tmap_bf16(&tb, w, 2, dims, strides, box, true)) return e;
  }
  const int sms = device_sm_count();
  if (sms <= 0) return SFB_ERR_CUDA;
  if (pair) return launch_gemm_cluster(epilogue, cluster_pairs, x, ldx, tb, p, sms, stream);
  switch (block_n) {
    case 64: return dispatch_epi<64>(epilogue, ta, tb, p, sms, stream);
    case 128: return dispatch_epi<128>(epilogue, ta, tb, p, sms, stream);
    default: return dispatch_epi<256>(epilogue, ta, tb, p, sms, stream);
  }
}

extern "C" int sfb_gemm_bf16(const void* x, long long ldx, const void* w, long long ldw, const void* bias,
                             int M, int N, int K, int epilogue, void* out0, long long ldo0, void* out1,
                             long long ldo1, void* out2, long long ldo2, int seg_cols, const void* residual,
                             long long ldr, const void* gate, long long gate_stride, int rows_per_gate,
                             int gate_row_offset, int block_n, void* workspace, long long workspace_bytes,
                             void* stream_) {
  (void)workspace; (void)workspace_bytes;
  return sfb_gemm_bf16_stats(x, ldx, w, ldw, bias, M, N, K, epilogue, out0, ldo0, out1, ldo1, out2, ldo2, seg_cols, residual,
                             ldr, gate, gate_stride, rows_per_gate, gate_row_offset, block_n, nullptr, nullptr, nullptr,
                             0.f, stream_);
}

// Kept for ABI stability: the GEMM needs no scratch any more (the stream-K schedule that used it was measured and
// removed, see gemm2_tcgen05.cu); `workspace` / `workspace_bytes` of sfb_gemm_bf16 are ignored.
extern "C" long long sfb_gemm_workspace_bytes(void) { return 0; }
