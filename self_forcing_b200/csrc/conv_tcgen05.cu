// Implicit-GEMM causal 3x3 convolution on channels-last frames for sm_100a: Y[voxel, Cout] = sum over taps of
// X[voxel + tap, Cin] . W[Cout, tap, Cin]^T, with the gathered operand NEVER materialised.
//
// Replaces (with sfb_causal_conv3d_cl's workspace == NULL) the gather + GEMM pair for CausalConv3d(.., 3, padding=1)
// and Conv2d(.., 3, padding=1) of the VAE decoder (wan/modules/vae.py:17-36, :75-83, :191-200).  In channels-last
// layout the A tile of one tap is a plain TMA box of the input: 128 output voxels = box_h rows x box_w pixels of one
// frame, shifted by (dt, dh - 1, dw - 1), 64 channels deep -- 128 rows of 128 bytes in shared memory, exactly the
// K-major swizzled operand the UMMA wants.  Spatial zero padding, the causal zero frames and the channel tail
// (Cin = 96 = 64 + 32) are all TMA out-of-bounds fills; taps that lie entirely in the zero frames are skipped.
// The gather version moves 2 x 27 x the input through HBM per convolution; this one reads the input ~once (the 27
// shifted boxes of a tile hit L2) and is bound by the L2 -> SM fill rate instead.
//
// Same roles as gemm_tcgen05.cu: warp 0 TMA producer, warp 1 MMA issuer (tcgen05.mma cta_group::1, M=128, N=BN),
// warps 2-5 epilogue (double-buffered TMEM accumulator; bias / residual fused, reference rounding points).
#include <stdlib.h>

#include "gemm_common.cuh"

namespace sfb {

constexpr int CV_BM = 128;
constexpr int CV_BK = 64;
constexpr int CV_THREADS = 192;

// MT = accumulator sub-tiles per CTA: with MT = 2 one CTA owns 256 voxels (two M = 128 MMAs per k-step against the
// SAME weight tile in shared memory), which halves the weight bytes crossing L2 -> SM per output voxel.
template <int BN, int MT>
struct ConvCfg {
  static constexpr int A_BYTES = MT * CV_BM * CV_BK * 2;
  static constexpr int B_BYTES = BN * CV_BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int STAGES = (BN == 256 || MT == 2) ? 4 : (BN == 192 ? 5 : 6);
  // TMEM is allocated in powers of two: 2 x 192 accumulator columns take a 512-column allocation
  static constexpr int TMEM_COLS = 2 * MT * BN <= 64 ? 64 : (2 * MT * BN <= 256 ? 256 : 512);
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
  static_assert(2 * MT * BN <= 512, "accumulators do not fit in TMEM");
};

struct ConvParams {
  int t_out, Ho, Wo, Cin;
  int kt, t_zero_pad;
  int box_w_log2, box_h;            // spatial tile: box_h rows of (1 << box_w_log2) pixels = MT * 128 voxels
  int tiles_w, tiles_h, num_n_blocks, c_chunks;
  GemmParams g;                     // bias, out[0], ldo[0], residual, ldr, N (= Cout), seg_cols (= N)
};

struct ConvTile {
  int n_blk, tw, th, to, dt0;
  __device__ ConvTile(const ConvParams& p, int tile) {
    n_blk = tile % p.num_n_blocks;                 // N tiles of one spatial tile are neighbours: they share A through L2
    int sp = tile / p.num_n_blocks;
    tw = sp % p.tiles_w; sp /= p.tiles_w;
    th = sp % p.tiles_h;
    to = sp / p.tiles_h;
    dt0 = p.t_zero_pad - to;                       // temporal taps below dt0 read only the causal zero frames
    if (dt0 < 0) dt0 = 0;
  }
};

template <int BN, int EPI, int MT>
__global__ void __launch_bounds__(CV_THREADS, 1)
conv3_implicit_kernel(const __grid_constant__ CUtensorMap tma_x, const __grid_constant__ CUtensorMap tma_w,
                      const ConvParams p) {
  using Cfg = ConvCfg<BN, MT>;
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
    tma_prefetch_desc(&tma_x);
    tma_prefetch_desc(&tma_w);
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int s = 0; s < 2; ++s) {
      mbar_init(&tmem_full[s], 1);
      mbar_init(&tmem_empty[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int num_tiles = p.t_out * p.tiles_h * p.tiles_w * p.num_n_blocks;
  const int box_w = 1 << p.box_w_log2;

  if (warp == 0) {
    // ------------------------------ TMA producer --------------------------------
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const ConvTile t(p, tile);
      const int w0 = t.tw * box_w - 1, h0 = t.th * p.box_h - 1;
      for (int tap = t.dt0 * 9; tap < p.kt * 9; ++tap) {
        const int dt = tap / 9, dh = (tap / 3) % 3, dw = tap % 3;
        for (int cc = 0; cc < p.c_chunks; ++cc) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          if (elect_one()) {
            uint8_t* a_dst = smem + stage * Cfg::STAGE_BYTES;
            mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
            tma_load_4d(a_dst, &tma_x, &full_bar[stage], cc * CV_BK, w0 + dw, h0 + dh, t.to + dt - p.t_zero_pad);
            tma_load_2d(a_dst + Cfg::A_BYTES, &tma_w, &full_bar[stage], tap * p.Cin + cc * CV_BK, t.n_blk * BN);
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer ----------------------------------
    constexpr uint32_t idesc = umma_idesc_bf16(CV_BM, BN, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const ConvTile t(p, tile);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * (MT * BN);
      const int iters = (p.kt - t.dt0) * 9 * p.c_chunks;
      for (int kb = 0; kb < iters; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (elect_one()) {
          const uint32_t a_addr = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint64_t a_desc = umma_desc_sw128(a_addr, 16, 1024);
          const uint64_t b_desc = umma_desc_sw128(a_addr + Cfg::A_BYTES, 16, 1024);
#pragma unroll
          for (int k = 0; k < CV_BK / 16; ++k) {
#pragma unroll
            for (int m = 0; m < MT; ++m)   // sub-tile m = rows [128 m, 128 m + 128) of the A box, 16 KB further on
              umma_ss(d_tmem + m * BN, a_desc + m * ((CV_BM * CV_BK * 2) >> 4) + 2 * k, b_desc + 2 * k, idesc, (kb | k) != 0);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == iters - 1) umma_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else {
    // ------------------------------ epilogue warps ------------------------------
    const int quarter = warp & 3;
    const int r = quarter * 32 + lane;                         // accumulator row = voxel (hh, ww) of the box
    int it = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const ConvTile t(p, tile);
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
#pragma unroll 1
      for (int m = 0; m < MT; ++m) {
        const int rr = m * CV_BM + r;                            // row of the (MT * 128)-voxel box
        const int h = t.th * p.box_h + (rr >> p.box_w_log2), w = t.tw * box_w + (rr & (box_w - 1));
        const bool ok = h < p.Ho && w < p.Wo;
        const long long voxel = ((long long)t.to * p.Ho + h) * p.Wo + w;
        gemm_epilogue_row_at<BN, EPI>(p.g, voxel, ok, t.n_blk * BN,
                                      tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * (MT * BN) + m * BN);
      }
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

int device_sm_count();

template <int BN, int EPI, int MT>
static int launch_conv3(const CUtensorMap& tx, const CUtensorMap& tw, const ConvParams& p, cudaStream_t stream) {
  using Cfg = ConvCfg<BN, MT>;
  auto kern = conv3_implicit_kernel<BN, EPI, MT>;
  static SmemOptIn optin;
  if (int e = optin.ensure(kern, Cfg::SMEM_BYTES, "cudaFuncSetAttribute(conv3)")) return e;
  const int sms = device_sm_count();
  if (sms <= 0) return SFB_ERR_CUDA;
  const long long tiles = (long long)p.t_out * p.tiles_h * p.tiles_w * p.num_n_blocks;
  if (tiles > 0x7fffffffLL) { set_error("conv3 implicit: too many tiles"); return SFB_ERR_INVALID; }
  const int grid = tiles < sms ? (int)tiles : sms;
  kern<<<grid, CV_THREADS, Cfg::SMEM_BYTES, stream>>>(tx, tw, p);
  return check_cuda(cudaGetLastError(), "conv3 implicit launch");
}

// x [t_in, H, W, Cin] channels-last; w packed [Cout, kt*9*Cin]; y [t_out*H*W, Cout] with row stride ldo.
int launch_conv3_implicit(const void* x, int t_in, int H, int W, int Cin, int t_zero_pad, const void* w, const void* bias,
                          int Cout, int kt, const void* residual, long long ldr, void* y, long long ldo,
                          cudaStream_t stream) {
  ConvParams p{};
  p.t_out = t_in + t_zero_pad - (kt - 1);
  p.Ho = H; p.Wo = W; p.Cin = Cin; p.kt = kt; p.t_zero_pad = t_zero_pad;
  // 32: the 3-channel head (padded to 8); 192-column tiles for the 192-channel layers (no padded MMA columns; measured
  // 0.467 -> 0.461 s per decoded video on B200, profiles/r02b_vae_exact_n.json)
  const int bn = Cout <= 32 ? 32 : (Cout <= 128 ? 128 : (Cout == 192 ? 192 : 256));
  // two accumulator sub-tiles per CTA (256 voxels against one weight tile) where TMEM has room and the frame is large
  // enough to keep every SM busy with 256-voxel tiles
  const int sms_hint = device_sm_count();
  const int mt = (bn == 128 && (long long)p.t_out * H * W >= 256LL * 4 * (sms_hint > 0 ? sms_hint : 148)) ? 2 : 1;
  // spatial box of mt * 128 voxels: the widest power-of-two row segment that wastes the least area
  long long best = -1;
  for (int lg = 3; lg <= 7; ++lg) {
    const int bw = 1 << lg, bh = mt * CV_BM / bw;
    const long long area = (long long)((W + bw - 1) / bw) * bw * ((H + bh - 1) / bh) * bh;
    if (best < 0 || area <= best) { best = area; p.box_w_log2 = lg; p.box_h = bh; }
  }
  const int bw = 1 << p.box_w_log2;
  p.tiles_w = (W + bw - 1) / bw;
  p.tiles_h = (H + p.box_h - 1) / p.box_h;
  p.c_chunks = (Cin + CV_BK - 1) / CV_BK;
  p.num_n_blocks = (Cout + bn - 1) / bn;
  GemmParams& g = p.g;
  g.M = 0; g.N = Cout; g.K = kt * 9 * Cin;
  g.bias = static_cast<const __nv_bfloat16*>(bias);
  g.out[0] = static_cast<__nv_bfloat16*>(y); g.ldo[0] = ldo;
  g.seg_cols = Cout;
  g.residual = static_cast<const __nv_bfloat16*>(residual); g.ldr = ldr;
  g.rows_per_gate = 1;

  CUtensorMap tx, tw;
  {
    uint64_t dims[4] = {(uint64_t)Cin, (uint64_t)W, (uint64_t)H, (uint64_t)t_in};
    uint64_t strides[3] = {(uint64_t)Cin * 2, (uint64_t)W * Cin * 2, (uint64_t)H * W * Cin * 2};
    uint32_t box[4] = {CV_BK, (uint32_t)bw, (uint32_t)p.box_h, 1};
    if (int e = make_tmap_bf16(&tx, x, 4, dims, strides, box, true)) return e;
  }
  {
    uint64_t dims[2] = {(uint64_t)g.K, (uint64_t)Cout};
    uint64_t strides[1] = {(uint64_t)g.K * 2};
    uint32_t box[2] = {CV_BK, (uint32_t)bn};
    if (int e = make_tmap_bf16(&tw, w, 2, dims, strides, box, true)) return e;
  }
  const bool res = residual != nullptr;
  if (bn == 32) return res ? launch_conv3<32, EPI_RESIDUAL, 1>(tx, tw, p, stream) : launch_conv3<32, EPI_BIAS, 1>(tx, tw, p, stream);
  if (bn == 128 && mt == 2)
    return res ? launch_conv3<128, EPI_RESIDUAL, 2>(tx, tw, p, stream) : launch_conv3<128, EPI_BIAS, 2>(tx, tw, p, stream);
  if (bn == 128) return res ? launch_conv3<128, EPI_RESIDUAL, 1>(tx, tw, p, stream) : launch_conv3<128, EPI_BIAS, 1>(tx, tw, p, stream);
  if (bn == 192) return res ? launch_conv3<192, EPI_RESIDUAL, 1>(tx, tw, p, stream) : launch_conv3<192, EPI_BIAS, 1>(tx, tw, p, stream);
  return res ? launch_conv3<256, EPI_RESIDUAL, 1>(tx, tw, p, stream) : launch_conv3<256, EPI_BIAS, 1>(tx, tw, p, stream);
}

}  // namespace sfb
