// UMT5 text encoder: the three element-wise kernels around the tensor-core GEMMs.
// The encoder runs once per prompt on 512 tokens (4.7 TFLOP of projections), so these are small, generic-width
// kernels; every one keeps the reference's bf16 rounding points (wan/modules/t5.py).
//
//   sfb_t5_rmsnorm          T5LayerNorm: w * bf16(x * rsqrt(mean(x^2) + eps)), fp32 statistics       t5.py:61-66
//   sfb_softmax_bias_rows   softmax(bf16(scores + bias)) with masked keys at finfo.min, fp32 softmax   t5.py:103-115
//   sfb_t5_gated_gelu       fc1 * gelu(gate) with the op-by-op tanh GELU of t5.py:46-50                t5.py:136-137
//
// NOT YET VALIDATED ON HARDWARE (written after the round's GPU budget was spent): parity checks are registered under
// tests/gpu_checks.py: PENDING, the host logic is tested on the CPU through the test double.
#include <math.h>

#include "common.cuh"

namespace sfb {

// one warp per row, any C that is a multiple of 8; the row is read twice (the second pass hits L1 / L2)
__global__ void __launch_bounds__(256)
t5_rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
                  int rows, int C, float eps, const __nv_bfloat16* __restrict__ w) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const uint4* src = reinterpret_cast<const uint4*>(x + (long long)row * ldx);
  const int nvec = C >> 3;
  float ss = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 q = __ldg(src + v);
    const uint32_t p[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) { const float a = bf_lo(p[j]), b = bf_hi(p[j]); ss = fmaf(a, a, ss); ss = fmaf(b, b, ss); }
  }
  ss = warp_sum(ss);
  const float rstd = rsqrtf(__fadd_rn(__fdiv_rn(ss, (float)C), eps));
  for (int v = lane; v < nvec; v += 32) {
    const uint4 q = __ldg(src + v);
    const uint4 g = __ldg(reinterpret_cast<const uint4*>(w) + v);
    const uint32_t p[4] = {q.x, q.y, q.z, q.w}, gw[4] = {g.x, g.y, g.z, g.w};
    uint32_t o[4];
#pragma unroll
    for (int j = 0; j < 4; ++j)   // bf16(x * rstd) pairwise, then the exact bf16 x bf16 product with the weight
      o[j] = mul_bf16x2(gw[j], round_pair(__fmul_rn(bf_lo(p[j]), rstd), __fmul_rn(bf_hi(p[j]), rstd)));
    reinterpret_cast<uint4*>(y + (long long)row * ldy)[v] = make_uint4(o[0], o[1], o[2], o[3]);
  }
}

// one block per query row: t = bf16(s + (masked ? finfo.min : bias)); p = bf16(softmax_fp32(t))
__global__ void __launch_bounds__(256)
softmax_bias_rows_kernel(const __nv_bfloat16* __restrict__ s, long long lds, const __nv_bfloat16* __restrict__ bias,
                         long long ldb, const int* __restrict__ key_mask, __nv_bfloat16* __restrict__ p, long long ldp,
                         int cols) {
  __shared__ float red[8];
  const __nv_bfloat16* srow = s + blockIdx.x * lds;
  const __nv_bfloat16* brow = bias + blockIdx.x * ldb;
  __nv_bfloat16* prow = p + blockIdx.x * ldp;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float kMin = -3.3895313892515355e38f;                  // torch.finfo(torch.bfloat16).min
  auto logit = [&](int c) {
    const float b = (key_mask != nullptr && key_mask[c] == 0) ? kMin : __bfloat162float(brow[c]);
    return bf16r(__fadd_rn(__bfloat162float(srow[c]), b));
  };
  float m = -INFINITY;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) m = fmaxf(m, logit(c));
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[warp] = m;
  __syncthreads();
  m = red[0];
#pragma unroll
  for (int i = 1; i < 8; ++i) m = fmaxf(m, red[i]);
  __syncthreads();
  float sum = 0.f;
  for (int c = threadIdx.x; c < cols; c += blockDim.x) sum += expf(logit(c) - m);
  sum = warp_sum(sum);
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  sum = 0.f;
#pragma unroll
  for (int i = 0; i < 8; ++i) sum += red[i];
  for (int c = threadIdx.x; c < cols; c += blockDim.x) prow[c] = __float2bfloat16_rn(__fdiv_rn(expf(logit(c) - m), sum));
}

// out = bf16(fc1 * f),  f = bf16(bf16(0.5 g) * bf16(1 + bf16(tanh(bf16(k * bf16(g + bf16(0.044715 * bf16(g^3))))))))
__global__ void __launch_bounds__(256)
t5_gated_gelu_kernel(const __nv_bfloat16* __restrict__ fc1, long long ld1, const __nv_bfloat16* __restrict__ gate,
                     long long ldg, __nv_bfloat16* __restrict__ out, long long ldo, int rows, int cols) {
  const float k = 0.7978845608028654f;                         // sqrt(2 / pi)
  const long long total = (long long)rows * cols;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int r = (int)(i / cols), c = (int)(i % cols);
    const float g = __bfloat162float(gate[r * ldg + c]);
    const float cube = bf16r(__fmul_rn(__fmul_rn(g, g), g));
    const float inner = bf16r(__fmul_rn(k, bf16r(__fadd_rn(g, bf16r(__fmul_rn(0.044715f, cube))))));
    const float one_plus = bf16r(__fadd_rn(1.0f, bf16r(tanhf(inner))));
    const float f = bf16r(__fmul_rn(bf16r(__fmul_rn(0.5f, g)), one_plus));
    out[r * ldo + c] = __float2bfloat16_rn(__fmul_rn(__bfloat162float(fc1[r * ld1 + c]), f));
  }
}

}  // namespace sfb

using namespace sfb;
typedef __nv_bfloat16 bf16;

extern "C" int sfb_t5_rmsnorm(const void* x, long long ldx, void* y, long long ldy, int rows, int C, float eps,
                              const void* weight, void* stream) {
  if (!x || !y || !weight || rows <= 0 || C <= 0 || (C % 8) || (ldx % 8) || (ldy % 8)) {
    set_error("sfb_t5_rmsnorm: bad arguments (C and the row strides must be multiples of 8)");
    return SFB_ERR_INVALID;
  }
  t5_rmsnorm_kernel<<<(rows + 7) / 8, 256, 0, (cudaStream_t)stream>>>((const bf16*)x, ldx, (bf16*)y, ldy, rows, C, eps, (const bf16*)weight);
  return check_cuda(cudaGetLastError(), "t5_rmsnorm launch");
}

extern "C" int sfb_softmax_bias_rows(const void* s, long long lds, const void* bias, long long ldb, const int* key_mask,
                                     void* p, long long ldp, int rows, int cols, void* stream) {
  if (!s || !bias || !p || rows <= 0 || cols <= 0) { set_error("sfb_softmax_bias_rows: bad arguments"); return SFB_ERR_INVALID; }
  softmax_bias_rows_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>((const bf16*)s, lds, (const bf16*)bias, ldb, key_mask, (bf16*)p, ldp, cols);
  return check_cuda(cudaGetLastError(), "softmax_bias_rows launch");
}

extern "C" int sfb_t5_gated_gelu(const void* fc1, long long ld1, const void* gate, long long ldg, void* out, long long ldo,
                                 int rows, int cols, void* stream) {
  if (!fc1 || !gate || !out || rows <= 0 || cols <= 0) { set_error("sfb_t5_gated_gelu: bad arguments"); return SFB_ERR_INVALID; }
  long long blocks = ((long long)rows * cols + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  t5_gated_gelu_kernel<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)fc1, ld1, (const bf16*)gate, ldg, (bf16*)out, ldo, rows, cols);
  return check_cuda(cudaGetLastError(), "t5_gated_gelu launch");
}
