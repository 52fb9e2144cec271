// Shared pieces of the two tcgen05 GEMM kernels (1-CTA tiles in gemm_tcgen05.cu, CTA-pair tiles in
// gemm2_tcgen05.cu): parameters, the GELU used by the reference, and the fused epilogue of one
// accumulator row (TMEM -> registers -> bias / GELU / gate / residual -> 16-byte global stores).
#pragma once
#include "common.cuh"

namespace sfb {

enum : int { EPI_BIAS = 0, EPI_GELU = 1, EPI_RESIDUAL = 2, EPI_GATE_RES = 3, EPI_F32 = 4 };

struct GemmParams {
  int M, N, K;
  int num_m_blocks, num_n_blocks, num_k_blocks;
  const __nv_bfloat16* bias;   // [N] or nullptr
  __nv_bfloat16* out[3];       // output column segments (QKV writes three destinations)
  long long ldo[3];            // row stride (elements) of each segment
  int seg_cols;                // columns per segment (== N when there is a single destination)
  const __nv_bfloat16* residual;
  long long ldr;
  const __nv_bfloat16* gate;   // gate vector of row r lives at gate + ((r + gate_row_offset) / rows_per_gate) * gate_stride
  long long gate_stride;
  int rows_per_gate;
  int gate_row_offset;         // chunk-global index of row 0 (sequence-parallel callers hold a slice of the rows)
  long long* dbg;              // SFB_GEMM_TIMING=1: per-CTA clock64 timeline [grid][16], else nullptr
  // Row statistics (CTA-pair kernel only).  A statistics record is float2 (mean, M2 = sum (x - mean)^2) of one row over
  // one chunk of STATS_CHUNK consecutive columns, laid out [row][chunk]; chunks are merged with Chan's formula, so the
  // result does not depend on the size of the row mean.
  float2* stats_out;           // if set: statistics of the bf16 OUTPUT rows, [M][N / STATS_CHUNK]
  const float2* ln_stats;      // if set (EPI_BIAS, bias == nullptr): statistics of the INPUT rows, [M][K / STATS_CHUNK] -- the
  float ln_eps;                //   LayerNorm of the input is folded into the epilogue:
  const float2* ln_sc;         //   y[r][n] = rstd_r * (acc[r][n] - mean_r * ln_sc[n].x) + ln_sc[n].y   with the weights pre-scaled by the norm's affine weight
};

__device__ __forceinline__ float gelu_tanh_f(float x) {
  // 0.5 x (1 + tanh(sqrt(2/pi) (x + 0.044715 x^3))), tanh(u) = 1 - 2 / (1 + e^{2u})
  const float kBeta = 0.7978845608028654f, kKappa = 0.044715f;
  float u = kBeta * (x + kKappa * x * x * x);
  float t = 1.0f - __fdividef(2.0f, 1.0f + __expf(2.0f * u));
  return 0.5f * x * (1.0f + t);
}


// GELU(tanh) with the hardware tanh (one MUFU op instead of ex2 + rcp): used by the clustered GEMM whose FFN1 epilogue
// was XU-pipe / issue bound (14.8 k cycles per 128 x 256 tile against a 12.3 k-cycle main loop, profiles/r02b).
// tanh.approx.f32 has ~2^-11 relative error, an order below the bf16 rounding that follows.
__device__ __forceinline__ float gelu_tanh_fast(float x) {
  const float kBeta = 0.7978845608028654f, kKappa = 0.044715f;
  const float u = kBeta * (x + kKappa * x * x * x);
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(u));
  const float hx = 0.5f * x;
  return fmaf(hx, t, hx);
}

// Fused epilogue math on 8 consecutive accumulator columns (one 16-byte bf16 chunk), shared by all store paths.
// b / gq / res hold the chunk's bias, gate and residual values (ignored where the epilogue has none).
template <int EPI, bool FAST_GELU = false>
__device__ __forceinline__ uint4 gemm_epilogue_vals(const uint32_t* acc8, bool has_bias, uint4 b, uint4 gq, uint4 res) {
  float f[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) f[i] = __uint_as_float(acc8[i]);
  if (has_bias) {
    const uint32_t bw[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) { f[2 * i] += bf_lo(bw[i]); f[2 * i + 1] += bf_hi(bw[i]); }
  }
  // the Linear output is rounded to bf16 first (pairwise: one conversion per two values), then the fused op runs
  if (EPI == EPI_GELU) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t y2 = round_pair(f[2 * i], f[2 * i + 1]);
      f[2 * i] = FAST_GELU ? gelu_tanh_fast(bf_lo(y2)) : gelu_tanh_f(bf_lo(y2));
      f[2 * i + 1] = FAST_GELU ? gelu_tanh_fast(bf_hi(y2)) : gelu_tanh_f(bf_hi(y2));
    }
  }
  if (EPI == EPI_GATE_RES) {
    const uint32_t gw[4] = {gq.x, gq.y, gq.z, gq.w};
    const uint32_t rw[4] = {res.x, res.y, res.z, res.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t p2 = mul_bf16x2(round_pair(f[2 * i], f[2 * i + 1]), gw[i]);   // bf16(bf16(y) * gate), exact
      f[2 * i] = bf_lo(rw[i]) + bf_lo(p2);
      f[2 * i + 1] = bf_hi(rw[i]) + bf_hi(p2);
    }
  }
  if (EPI == EPI_RESIDUAL) {
    const uint32_t rw[4] = {res.x, res.y, res.z, res.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t y2 = round_pair(f[2 * i], f[2 * i + 1]);
      f[2 * i] = bf_lo(rw[i]) + bf_lo(y2);
      f[2 * i + 1] = bf_hi(rw[i]) + bf_hi(y2);
    }
  }
  uint4 o;
  o.x = pack_bf16(f[0], f[1]);
  o.y = pack_bf16(f[2], f[3]);
  o.z = pack_bf16(f[4], f[5]);
  o.w = pack_bf16(f[6], f[7]);
  return o;
}

// bias / gate point at the chunk's first column in global memory (nullptr = none)
template <int EPI>
__device__ __forceinline__ uint4 gemm_epilogue_chunk(const uint32_t* acc8, const __nv_bfloat16* bias,
                                                     const __nv_bfloat16* gate, uint4 res) {
  uint4 b = make_uint4(0, 0, 0, 0), gq = make_uint4(0, 0, 0, 0);
  if (bias != nullptr) b = __ldg(reinterpret_cast<const uint4*>(bias));
  if (EPI == EPI_GATE_RES) gq = __ldg(reinterpret_cast<const uint4*>(gate));
  return gemm_epilogue_vals<EPI>(acc8, bias != nullptr, b, gq, res);
}

// One thread owns accumulator row `row` (TMEM lane) of a tile that starts at column n0 and is
// TILE_N columns wide; t_row is the TMEM address of (lane, first column).
// `row` is the output row the thread writes (for the GEMM the accumulator row itself, for the implicit-GEMM
// convolution the voxel index of the accumulator row); row_ok = false only drains TMEM.
template <int TILE_N, int EPI>
__device__ __forceinline__ void gemm_epilogue_row_at(const GemmParams& p, long long row, bool row_ok, int n0, uint32_t t_row) {
  if (EPI == EPI_F32) {
    // fp32 result (attention scores): out[0] is a float matrix, ldo[0] counted in floats; one segment only
    float* frow = reinterpret_cast<float*>(p.out[0]) + row * p.ldo[0] + n0;
#pragma unroll 1
    for (int c = 0; c < TILE_N / 32; ++c) {
      uint32_t v[32];
      tmem_ld32(t_row + c * 32, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int g = 0; g < 8; ++g) {   // 4 columns (16 bytes of fp32) per step
          const int col = c * 32 + g * 4;
          if (n0 + col < p.N) {
            float4 o = make_float4(__uint_as_float(v[g * 4]), __uint_as_float(v[g * 4 + 1]), __uint_as_float(v[g * 4 + 2]),
                                   __uint_as_float(v[g * 4 + 3]));
            if (p.bias != nullptr) {
              const uint2 b = __ldg(reinterpret_cast<const uint2*>(p.bias + n0 + col));
              o.x += bf_lo(b.x); o.y += bf_hi(b.x); o.z += bf_lo(b.y); o.w += bf_hi(b.y);
            }
            *reinterpret_cast<float4*>(frow + col) = o;
          }
        }
      }
    }
    return;
  }
  const int seg = n0 / p.seg_cols;
  __nv_bfloat16* orow = p.out[seg] + row * p.ldo[seg] + (n0 - seg * p.seg_cols);
  const __nv_bfloat16* rrow = nullptr;
  const __nv_bfloat16* grow = nullptr;
  if (EPI == EPI_RESIDUAL || EPI == EPI_GATE_RES) rrow = p.residual + row * p.ldr + n0;
  if (EPI == EPI_GATE_RES) grow = p.gate + (long long)(row_ok ? ((int)row + p.gate_row_offset) / p.rows_per_gate : 0) * p.gate_stride + n0;
#pragma unroll 1
  for (int c = 0; c < TILE_N / 32; ++c) {
    uint32_t v[32];
    tmem_ld32(t_row + c * 32, v);
    tmem_ld_wait();
    if (row_ok) {
#pragma unroll
      for (int g = 0; g < 4; ++g) {   // 8 columns (16 bytes of bf16) per step
        const int col = c * 32 + g * 8;
        if (n0 + col < p.N) {
          uint4 rq = make_uint4(0u, 0u, 0u, 0u);
          if (EPI == EPI_RESIDUAL || EPI == EPI_GATE_RES) rq = *reinterpret_cast<const uint4*>(rrow + col);
          const uint4 o = gemm_epilogue_chunk<EPI>(v + g * 8, p.bias != nullptr ? p.bias + n0 + col : nullptr,
                                                   EPI == EPI_GATE_RES ? grow + col : nullptr, rq);
          *reinterpret_cast<uint4*>(orow + col) = o;
        }
      }
    }
  }
}

template <int TILE_N, int EPI>
__device__ __forceinline__ void gemm_epilogue_row(const GemmParams& p, int row, int n0, uint32_t t_row) {
  gemm_epilogue_row_at<TILE_N, EPI>(p, row, row < p.M, n0, t_row);
}

}  // namespace sfb
