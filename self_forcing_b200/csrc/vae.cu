// Wan VAE decoder: the kernels around the tensor-core GEMM, all on CHANNELS-LAST
// activations [T, H, W, C] bf16 (one voxel = one contiguous row of C channels), so that a causal 3-D convolution is
// a GEMM over gathered rows and the per-voxel channel norm is a row kernel.
//
//   sfb_vae_latent_in        z / (1/std) + mean, 1x1x1 conv2            wan_wrapper.py:101-102, vae.py:548-554
//   sfb_vae_norm_silu        RMS_norm over channels (+ SiLU)            vae.py:39-54, :195-198
//   sfb_causal_conv3d_cl     CausalConv3d / Conv2d(+nearest 2x) as      vae.py:17-36, :75-83
//                            gather (im2col) + sfb_gemm_bf16 (bias / residual fused in the GEMM epilogue)
//   sfb_softmax_rows         softmax of fp32 score rows -> bf16         vae.py:251-255 (single-head attention)
//   sfb_transpose_bf16       [R, C] -> [C, R]
//   sfb_vae_pixel_out        [T, H, W, 8] bf16 -> fp32 [T, 3, H, W], clamp(-1, 1)   wan_wrapper.py:110-116
//
// First correct version: the gathered operand is materialised in a caller-owned workspace (row chunks sized to it)
// and the existing tcgen05 GEMM consumes it.  The HBM traffic of that buffer (27 x the input for a 3x3x3 kernel) is
// the bound; an implicit-GEMM producer (shifted TMA boxes straight into the MMA pipeline) removes it -- see DESIGN.md.
#include <math.h>

#include "common.cuh"

// the tensor-core GEMM of gemm_tcgen05.cu (declared in include/sfb200.h)
extern "C" int sfb_gemm_bf16(const void* x, long long ldx, const void* w, long long ldw, const void* bias, int M, int N,
                             int K, int epilogue, void* out0, long long ldo0, void* out1, long long ldo1, void* out2,
                             long long ldo2, int seg_cols, const void* residual, long long ldr, const void* gate,
                             long long gate_stride, int rows_per_gate, int gate_row_offset, int block_n,
                             void* workspace, long long workspace_bytes, void* stream);

namespace sfb {

// conv_tcgen05.cu: implicit-GEMM 3x3 convolution (no gathered operand)
int launch_conv3_implicit(const void* x, int t_in, int H, int W, int Cin, int t_zero_pad, const void* w, const void* bias,
                          int Cout, int kt, const void* residual, long long ldr, void* y, long long ldo,
                          cudaStream_t stream);

// nearest-neighbour 2x upsampling of H and W, channels-last (vae.py:57-63)
__global__ void __launch_bounds__(256)
upsample2x_kernel(const uint4* __restrict__ x, uint4* __restrict__ y, int T, int H, int W, int cvec) {
  const long long total = (long long)T * 2 * H * 2 * W * cvec;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    long long v = i / cvec;
    const int wo = (int)(v % (2 * W)); v /= 2 * W;
    const int ho = (int)(v % (2 * H));
    const int t = (int)(v / (2 * H));
    y[i] = __ldg(x + (((long long)t * H + (ho >> 1)) * W + (wo >> 1)) * cvec + cv);
  }
}

// ------------------------------------------------------------------------------------
// latents in: out[v][o] = bf16(b[o] + sum_c w[o][c] * bf16(bf16(z[c][v] / inv_std[c]) + mean[c]))
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
vae_latent_in_kernel(const __nv_bfloat16* __restrict__ z, long long z_cstride, const __nv_bfloat16* __restrict__ mean,
                     const __nv_bfloat16* __restrict__ inv_std, const __nv_bfloat16* __restrict__ w,
                     const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ out, int voxels) {
  __shared__ float sw[16 * 16], sb[16], sm[16], si[16];
  for (int i = threadIdx.x; i < 256; i += blockDim.x) sw[i] = __bfloat162float(w[i]);
  if (threadIdx.x < 16) {
    sb[threadIdx.x] = __bfloat162float(bias[threadIdx.x]);
    sm[threadIdx.x] = __bfloat162float(mean[threadIdx.x]);
    si[threadIdx.x] = __bfloat162float(inv_std[threadIdx.x]);
  }
  __syncthreads();
  const int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= voxels) return;
  float x[16];
#pragma unroll
  for (int c = 0; c < 16; ++c)
    x[c] = bf16r(__fadd_rn(bf16r(__fdiv_rn(__bfloat162float(z[c * z_cstride + v]), si[c])), sm[c]));
  uint32_t packed[8];
#pragma unroll
  for (int o = 0; o < 16; o += 2) {
    float a0 = 0.f, a1 = 0.f;
#pragma unroll
    for (int c = 0; c < 16; ++c) { a0 = fmaf(sw[o * 16 + c], x[c], a0); a1 = fmaf(sw[(o + 1) * 16 + c], x[c], a1); }
    packed[o / 2] = pack_bf16(a0 + sb[o], a1 + sb[o + 1]);
  }
  uint4* dst = reinterpret_cast<uint4*>(out + (long long)v * 16);
  dst[0] = make_uint4(packed[0], packed[1], packed[2], packed[3]);
  dst[1] = make_uint4(packed[4], packed[5], packed[6], packed[7]);
}

// ------------------------------------------------------------------------------------
// RMS_norm (+SiLU): half a warp per voxel row, 16-byte vectors.  Reference chain on bf16 tensors:
//   d = bf16(sqrt(sum x^2)) clamped to 1e-12;  a = bf16(x / d);  b = bf16(a * sqrt(C));  c = bf16(b * gamma);
//   y = bf16(c / (1 + exp(-c)))
// The kernel is instruction-bound before it is HBM-bound (three IEEE divisions + an expf per element cost more issue
// slots than the 4 bytes of traffic), so x / d is one reciprocal per row plus a Newton correction per element
// (correctly rounded except in rare double-rounding cases) and the sigmoid uses ex2.approx + rcp; against the
// op-by-op chain the output differs in < 1e-4 of the elements, by one bf16 ulp.
// ------------------------------------------------------------------------------------
template <int NV>   // uint4 (8 x bf16) vectors per lane: C / 8 <= 16 * NV
__global__ void __launch_bounds__(256)
vae_norm_silu_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ gamma,
                     __nv_bfloat16* __restrict__ y, long long ldy, long long rows, int C, float sqrt_c, int silu) {
  const int lane = threadIdx.x & 15;
  const long long row = blockIdx.x * (long long)(blockDim.x >> 4) + (threadIdx.x >> 4);
  const bool live = row < rows;
  const int nvec = C >> 3;
  uint4 raw[NV];
  float ss = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 16 * i;
    raw[i] = make_uint4(0u, 0u, 0u, 0u);
    if (live && v < nvec) raw[i] = __ldg(reinterpret_cast<const uint4*>(x + row * ldx) + v);
    const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float a = bf_lo(w[j]), b = bf_hi(w[j]); ss = fmaf(a, a, ss); ss = fmaf(b, b, ss); }
  }
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);   // the 16 lanes of this row
  if (!live) return;
  const float denom = fmaxf(bf16r(sqrtf(ss)), 1e-12f);
  const float rinv = __frcp_rn(denom);
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int v = lane + 16 * i;
    if (v >= nvec) continue;
    const uint4 gq = __ldg(reinterpret_cast<const uint4*>(gamma) + v);
    const uint32_t w[4] = {raw[i].x, raw[i].y, raw[i].z, raw[i].w};
    const uint32_t g[4] = {gq.x, gq.y, gq.z, gq.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float e[2] = {bf_lo(w[j]), bf_hi(w[j])};
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        const float q0 = e[h] * rinv;
        e[h] = fmaf(fmaf(-q0, denom, e[h]), rinv, q0);          // x / d
      }
      float c0 = bf16r(__fmul_rn(bf16r(__fmul_rn(bf16r(e[0]), sqrt_c)), bf_lo(g[j])));
      float c1 = bf16r(__fmul_rn(bf16r(__fmul_rn(bf16r(e[1]), sqrt_c)), bf_hi(g[j])));
      if (silu) {
        c0 = c0 * __frcp_rn(1.0f + __expf(-c0));
        c1 = c1 * __frcp_rn(1.0f + __expf(-c1));
      }
      o[j] = pack_bf16(c0, c1);
    }
    reinterpret_cast<uint4*>(y + row * ldy)[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// ------------------------------------------------------------------------------------
// gather: A[r][tap * Cin + c] for output voxel rows [row0, row0 + rows) of a (kt, KS, KS) convolution with "same"
// spatial zero padding, causal temporal padding (t_zero_pad virtual zero frames in front of the t_in stored frames)
// and optional nearest 2x upsampling of H, W in front of the convolution.
// One warp per output row: the lanes walk the row's kt*KS*KS*CVEC 16-byte vectors, so stores are fully coalesced and
// loads are contiguous runs of KS*Cin channels (the dw taps of one (dt, dh) are neighbouring pixels in channels-last).
// CVEC = Cin / 8 and KS are compile-time so that the tap decomposition costs no integer division.
// ------------------------------------------------------------------------------------
template <int CVEC, int KS>
__global__ void __launch_bounds__(256)
vae_gather_kernel(const __nv_bfloat16* __restrict__ x, int t_in, int H, int W, int t_zero_pad, int up, int kt,
                  long long row0, int rows, __nv_bfloat16* __restrict__ A) {
  const int Ho = up ? 2 * H : H, Wo = up ? 2 * W : W;
  const int nvec = kt * KS * KS * CVEC;
  const int lane = threadIdx.x & 31;
  constexpr int HALF = KS / 2;
  const uint4* src = reinterpret_cast<const uint4*>(x);
  for (int row = blockIdx.x * 8 + (threadIdx.x >> 5); row < rows; row += gridDim.x * 8) {
    const long long r = row0 + row;
    const int wo = (int)(r % Wo);
    const int ho = (int)((r / Wo) % Ho);
    const int to = (int)(r / ((long long)Wo * Ho));
    uint4* dst = reinterpret_cast<uint4*>(A) + (long long)row * nvec;
    for (int j = lane; j < nvec; j += 32) {
      const int tap = j / CVEC, cv = j % CVEC;
      const int dw = tap % KS, dh = (tap / KS) % KS, dt = tap / (KS * KS);
      const int tf = to + dt - t_zero_pad;                     // stored frame index
      int hi = ho + dh - HALF, wi = wo + dw - HALF;
      uint4 val = make_uint4(0u, 0u, 0u, 0u);
      if (tf >= 0 && tf < t_in && hi >= 0 && hi < Ho && wi >= 0 && wi < Wo) {
        if (up) { hi >>= 1; wi >>= 1; }
        val = __ldg(src + (((long long)tf * H + hi) * W + wi) * CVEC + cv);
      }
      dst[j] = val;
    }
  }
}

template <int KS>
static int launch_gather(int cvec, int blocks, cudaStream_t st, const __nv_bfloat16* x, int t_in, int H, int W,
                         int t_zero_pad, int up, int kt, long long row0, int rows, __nv_bfloat16* A) {
  switch (cvec) {
    case 2: vae_gather_kernel<2, KS><<<blocks, 256, 0, st>>>(x, t_in, H, W, t_zero_pad, up, kt, row0, rows, A); return SFB_OK;
    case 12: vae_gather_kernel<12, KS><<<blocks, 256, 0, st>>>(x, t_in, H, W, t_zero_pad, up, kt, row0, rows, A); return SFB_OK;
    case 24: vae_gather_kernel<24, KS><<<blocks, 256, 0, st>>>(x, t_in, H, W, t_zero_pad, up, kt, row0, rows, A); return SFB_OK;
    case 48: vae_gather_kernel<48, KS><<<blocks, 256, 0, st>>>(x, t_in, H, W, t_zero_pad, up, kt, row0, rows, A); return SFB_OK;
  }
  set_error("sfb_causal_conv3d_cl: Cin=%d unsupported (16 / 96 / 192 / 384)", cvec * 8);
  return SFB_ERR_INVALID;
}

// ------------------------------------------------------------------------------------
// softmax over fp32 score rows (the fp32-output GEMM epilogue keeps the logits unrounded, like a fused attention
// kernel), probabilities written as bf16 for the P.V GEMM; one block per row
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
softmax_rows_kernel(const float* __restrict__ s, long long lds, __nv_bfloat16* __restrict__ p, long long ldp, int cols,
                    float scale) {
  __shared__ float red[8];
  const float* row = s + blockIdx.x * lds;
  __nv_bfloat16* prow = p + blockIdx.x * ldp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, row[c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) sum += expf(scale * (row[c] - m));
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  const float inv = 1.0f / sum;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) prow[c] = __float2bfloat16_rn(expf(scale * (row[c] - m)) * inv);
}

__global__ void __launch_bounds__(256)
transpose_bf16_kernel(const __nv_bfloat16* __restrict__ in, long long ldi, __nv_bfloat16* __restrict__ out,
                      long long ldo, int R, int C) {
  __shared__ __nv_bfloat16 tile[32][33];
  const int c0 = blockIdx.x * 32, r0 = blockIdx.y * 32;
  const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;      // 32 x 8
  for (int j = ty; j < 32; j += 8)
    if (r0 + j < R && c0 + tx < C) tile[j][tx] = in[(long long)(r0 + j) * ldi + c0 + tx];
  __syncthreads();
  for (int j = ty; j < 32; j += 8)
    if (c0 + j < C && r0 + tx < R) out[(long long)(c0 + j) * ldo + r0 + tx] = tile[tx][j];
}

__global__ void __launch_bounds__(256)
vae_pixel_out_kernel(const __nv_bfloat16* __restrict__ y, int ldy, float* __restrict__ out, int T, long long HW) {
  const long long total = (long long)T * HW;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < total; v += (long long)gridDim.x * blockDim.x) {
    const long long t = v / HW, p = v % HW;
#pragma unroll
    for (int c = 0; c < 3; ++c)
      out[(t * 3 + c) * HW + p] = fminf(fmaxf(__bfloat162float(y[v * ldy + c]), -1.0f), 1.0f);
  }
}

}  // namespace sfb

using namespace sfb;
typedef __nv_bfloat16 bf16;

static int grid_for(long long work, int per_block) {
  long long b = (work + per_block - 1) / per_block;
  if (b > 148LL * 16) b = 148LL * 16;
  return (int)(b < 1 ? 1 : b);
}

extern "C" int sfb_vae_latent_in(const void* z, long long z_channel_stride, const void* mean, const void* inv_std,
                                 const void* w, const void* bias, void* out, int voxels, void* stream) {
  if (voxels <= 0 || !z || !mean || !inv_std || !w || !bias || !out) { set_error("sfb_vae_latent_in: bad arguments"); return SFB_ERR_INVALID; }
  vae_latent_in_kernel<<<(voxels + 127) / 128, 128, 0, (cudaStream_t)stream>>>((const bf16*)z, z_channel_stride, (const bf16*)mean,
                                                                               (const bf16*)inv_std, (const bf16*)w, (const bf16*)bias,
                                                                               (bf16*)out, voxels);
  return check_cuda(cudaGetLastError(), "vae_latent_in launch");
}

extern "C" int sfb_vae_norm_silu(const void* x, long long ldx, const void* gamma, void* y, long long ldy, long long rows,
                                 int C, int silu, void* stream) {
  if (rows <= 0 || C <= 0 || (C % 8) || C > 512 || (ldx % 8) || (ldy % 8)) {
    set_error("sfb_vae_norm_silu: C=%d must be a multiple of 8 up to 512, row strides multiples of 8", C);
    return SFB_ERR_INVALID;
  }
  const long long blocks = (rows + 15) / 16;
  if (blocks > 0x7fffffffLL) { set_error("sfb_vae_norm_silu: too many rows"); return SFB_ERR_INVALID; }
  const float sc = sqrtf((float)C);
  cudaStream_t st = (cudaStream_t)stream;
  if (C <= 128) vae_norm_silu_kernel<1><<<(int)blocks, 256, 0, st>>>((const bf16*)x, ldx, (const bf16*)gamma, (bf16*)y, ldy, rows, C, sc, silu);
  else if (C <= 256) vae_norm_silu_kernel<2><<<(int)blocks, 256, 0, st>>>((const bf16*)x, ldx, (const bf16*)gamma, (bf16*)y, ldy, rows, C, sc, silu);
  else vae_norm_silu_kernel<4><<<(int)blocks, 256, 0, st>>>((const bf16*)x, ldx, (const bf16*)gamma, (bf16*)y, ldy, rows, C, sc, silu);
  return check_cuda(cudaGetLastError(), "vae_norm_silu launch");
}

extern "C" long long sfb_causal_conv3d_workspace_bytes(long long rows, int Cin, int kt, int ks) {
  return rows * (long long)kt * ks * ks * Cin * 2;
}

extern "C" int sfb_causal_conv3d_cl(const void* x, int t_in, int H, int W, int Cin, int t_zero_pad, int upsample2x,
                                    const void* w, const void* bias, int Cout, int kt, int ks,
                                    const void* residual, long long ldr, void* y0, void* y1, long long ldo, int seg_cols,
                                    void* workspace, long long workspace_bytes, void* stream) {
  if (!x || !w || !y0 || t_in <= 0 || H <= 0 || W <= 0 || (Cin % 8) || (Cout % 8) || kt < 1 || kt > 3 || (ks != 1 && ks != 3)) {
    set_error("sfb_causal_conv3d_cl: bad shape (Cin=%d, Cout=%d must be multiples of 8; kt=%d in 1..3; ks=%d in {1,3})", Cin, Cout, kt, ks);
    return SFB_ERR_INVALID;
  }
  const int t_out = t_in + t_zero_pad - (kt - 1);
  if (t_out <= 0 || t_zero_pad < 0 || t_zero_pad > kt - 1) { set_error("sfb_causal_conv3d_cl: t_in=%d + pad %d gives no output frame for kt=%d", t_in, t_zero_pad, kt); return SFB_ERR_INVALID; }
  const int Ho = upsample2x ? 2 * H : H, Wo = upsample2x ? 2 * W : W;
  const long long rows = (long long)t_out * Ho * Wo;
  const int K = kt * ks * ks * Cin;
  if (seg_cols <= 0) seg_cols = Cout;
  const int nseg = (Cout + seg_cols - 1) / seg_cols;
  if (nseg > 2 || (nseg == 2 && !y1)) { set_error("sfb_causal_conv3d_cl: at most two output segments"); return SFB_ERR_INVALID; }
  void* outs[2] = {y0, y1};
  const bool direct = (kt == 1 && ks == 1 && !upsample2x);      // 1x1x1: the activations are the operand
  if (!direct && workspace == nullptr) {
    // no staging buffer: implicit GEMM (conv_tcgen05.cu) -- 3x3 spatial kernels on un-upsampled input, one segment
    if (ks != 3 || upsample2x || nseg != 1) { set_error("sfb_causal_conv3d_cl: without a workspace only ks=3, no upsampling, one output segment"); return SFB_ERR_INVALID; }
    if (Cin % 16) { set_error("sfb_causal_conv3d_cl: implicit path needs Cin %% 16 == 0"); return SFB_ERR_INVALID; }
    return launch_conv3_implicit(x, t_in, H, W, Cin, t_zero_pad, w, bias, Cout, kt, residual, ldr, y0, ldo, (cudaStream_t)stream);
  }
  long long chunk = rows;
  if (!direct) {
    if (!workspace || workspace_bytes < (long long)K * 2 * 128) { set_error("sfb_causal_conv3d_cl: workspace missing or smaller than 128 gathered rows"); return SFB_ERR_INVALID; }
    chunk = workspace_bytes / ((long long)K * 2);
    if (chunk > rows) chunk = rows;
    if (chunk < rows) chunk &= ~127LL;
  }
  cudaStream_t st = (cudaStream_t)stream;
  // N tile of the GEMM: the gathered operand is the expensive one (HBM-resident, rows x K), so one tile should span all
  // output channels where possible -- it is then read exactly once.  Multiples of 256 take the CTA-pair kernel (0 = auto).
  int block_n = 0;
  if (Cout % 256 != 0 || seg_cols % 256 != 0) block_n = Cout <= 64 ? 64 : (Cout <= 128 ? 128 : (seg_cols == Cout ? 256 : 128));
  for (long long r0 = 0; r0 < rows; r0 += chunk) {
    const long long n = rows - r0 < chunk ? rows - r0 : chunk;
    const void* a = x;
    long long lda = Cin;
    if (!direct) {
      const int blocks = grid_for(n, 8);
      if (int e = ks == 3 ? launch_gather<3>(Cin >> 3, blocks, st, (const bf16*)x, t_in, H, W, t_zero_pad, upsample2x, kt, r0, (int)n, (bf16*)workspace)
                          : launch_gather<1>(Cin >> 3, blocks, st, (const bf16*)x, t_in, H, W, t_zero_pad, upsample2x, kt, r0, (int)n, (bf16*)workspace))
        return e;
      if (int e = check_cuda(cudaGetLastError(), "vae_gather launch")) return e;
      a = workspace;
      lda = K;
    } else {
      a = (const bf16*)x + r0 * Cin;
    }
    void* o0 = (bf16*)outs[0] + r0 * ldo;
    void* o1 = outs[1] ? (void*)((bf16*)outs[1] + r0 * ldo) : nullptr;
    const void* res = residual ? (const void*)((const bf16*)residual + r0 * ldr) : nullptr;
    if (n > 0x7fffffffLL) { set_error("sfb_causal_conv3d_cl: chunk too large"); return SFB_ERR_INVALID; }
    if (int e = sfb_gemm_bf16(a, lda, w, K, bias, (int)n, Cout, K, residual ? 2 : 0, o0, ldo, o1, ldo, nullptr, 0, seg_cols,
                              res, ldr, nullptr, 0, 1, 0, block_n, nullptr, 0, stream))
      return e;
  }
  return SFB_OK;
}

extern "C" int sfb_upsample2x_cl(const void* x, void* y, int T, int H, int W, int C, void* stream) {
  if (!x || !y || T <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 8)) { set_error("sfb_upsample2x_cl: bad arguments (C must be a multiple of 8)"); return SFB_ERR_INVALID; }
  const long long total = (long long)T * 4 * H * W * (C / 8);
  upsample2x_kernel<<<grid_for(total, 256 * 4), 256, 0, (cudaStream_t)stream>>>((const uint4*)x, (uint4*)y, T, H, W, C / 8);
  return check_cuda(cudaGetLastError(), "upsample2x launch");
}

extern "C" int sfb_softmax_rows(const void* s, long long lds, void* p, long long ldp, int rows, int cols, float scale,
                                void* stream) {
  if (!s || !p || rows <= 0 || cols <= 0) { set_error("sfb_softmax_rows: bad arguments"); return SFB_ERR_INVALID; }
  softmax_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>((const float*)s, lds, (bf16*)p, ldp, cols, scale);
  return check_cuda(cudaGetLastError(), "softmax_rows launch");
}

extern "C" int sfb_transpose_bf16(const void* in, long long ldi, void* out, long long ldo, int R, int C, void* stream) {
  if (!in || !out || R <= 0 || C <= 0) { set_error("sfb_transpose_bf16: bad arguments"); return SFB_ERR_INVALID; }
  dim3 grid((C + 31) / 32, (R + 31) / 32);
  transpose_bf16_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)in, ldi, (bf16*)out, ldo, R, C);
  return check_cuda(cudaGetLastError(), "transpose launch");
}

extern "C" int sfb_vae_pixel_out(const void* y, int ldy, void* out, int T, long long HW, void* stream) {
  if (!y || !out || T <= 0 || HW <= 0 || ldy < 3) { set_error("sfb_vae_pixel_out: bad arguments"); return SFB_ERR_INVALID; }
  vae_pixel_out_kernel<<<grid_for((long long)T * HW, 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)y, ldy, (float*)out, T, HW);
  return check_cuda(cudaGetLastError(), "vae_pixel_out launch");
}
