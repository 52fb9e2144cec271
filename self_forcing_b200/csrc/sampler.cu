// One multistep-sampler step of the 50-step pipeline in ONE launch: classifier-free guidance, flow -> x0,
// the UniC corrector on the current sample and the UniP predictor for the next one.
//
// Replaces, per denoising step, the ~25 elementwise launches of
//   pipeline/causal_diffusion_inference.py:420-428   flow = uncond + g * (cond - uncond); scheduler.step(...)
//   wan/utils/fm_solvers_unipc.py:320-323            x0 = sample - sigma * flow
//   wan/utils/fm_solvers_unipc.py:549-626            UniC-B(h), order 1 / 2
//   wan/utils/fm_solvers_unipc.py:404-484            UniP-B(h), order 1 / 2
// The reference evaluates these expressions op by op on bf16 tensors, i.e. every intermediate is rounded to bf16;
// the kernel keeps exactly those rounding points (so it is bit-identical to the op-by-op chain), but reads each
// operand once and writes three tensors: 6 reads + 3 writes of n bf16 values instead of ~60 passes.
// All scalar coefficients (functions of the sigma table only) are computed on the host (unipc.py) in fp32.
#include "common.cuh"

namespace sfb {

struct UniPCCoef {
  float guidance;                                   // classifier-free guidance scale
  float sigma;                                      // sigma of this step (flow -> x0)
  float c_x, c_m0, c_b, c_inv_rk, c_rho0, c_rho1;   // corrector: sig_t/sig_s, alpha_t*h_phi_1, alpha_t*B_h, 1/rk, rho[0], rho[-1]
  float p_x, p_m0, p_b, p_inv_rk;                   // predictor (its rho is the constant 0.5)
};

__device__ __forceinline__ float rmul(float a, float b) { return bf16r(__fmul_rn(a, b)); }
__device__ __forceinline__ float radd(float a, float b) { return bf16r(__fadd_rn(a, b)); }
__device__ __forceinline__ float rsub(float a, float b) { return bf16r(__fsub_rn(a, b)); }

// one element through the whole chain; every r* helper is one bf16 tensor op of the reference
template <bool CFG, int CORR, int PRED>
__device__ __forceinline__ void unipc_element(const UniPCCoef& k, float fc, float fu, float x, float xl, float m0,
                                              float m1, float& m_new, float& x_corr, float& x_next) {
  float flow = fc;
  if (CFG) flow = radd(fu, rmul(k.guidance, rsub(fc, fu)));
  m_new = rsub(x, rmul(k.sigma, flow));
  x_corr = x;
  if (CORR > 0) {
    const float base = rsub(rmul(k.c_x, xl), rmul(k.c_m0, m0));
    float inner = rmul(k.c_rho1, rsub(m_new, m0));
    if (CORR == 2) inner = radd(rmul(k.c_rho0, rmul(rsub(m1, m0), k.c_inv_rk)), inner);
    x_corr = rsub(base, rmul(k.c_b, inner));
  }
  // the history shifts: the newest x0 prediction is m_new, the one before it m0
  const float base = rsub(rmul(k.p_x, x_corr), rmul(k.p_m0, m_new));
  x_next = base;
  if (PRED == 2) x_next = rsub(base, rmul(k.p_b, rmul(0.5f, rmul(rsub(m0, m_new), k.p_inv_rk))));
}

template <bool CFG, int CORR, int PRED>
__global__ void __launch_bounds__(256)
cfg_unipc_step_kernel(const __nv_bfloat16* __restrict__ flow_cond, const __nv_bfloat16* __restrict__ flow_uncond,
                      const __nv_bfloat16* __restrict__ sample, const __nv_bfloat16* last_sample,
                      const __nv_bfloat16* m0, const __nv_bfloat16* m1, __nv_bfloat16* m_out,
                      __nv_bfloat16* sample_out, __nv_bfloat16* __restrict__ prev_out, long long n, const UniPCCoef k) {
  // m_out / sample_out may alias m1 / last_sample: every element is read before it is written, by the same thread
  const long long nvec = n / 8;
  for (long long v = blockIdx.x * (long long)blockDim.x + threadIdx.x; v < nvec; v += (long long)gridDim.x * blockDim.x) {
    const uint4 zero = make_uint4(0u, 0u, 0u, 0u);
    const uint4 qc = reinterpret_cast<const uint4*>(flow_cond)[v];
    const uint4 qu = CFG ? reinterpret_cast<const uint4*>(flow_uncond)[v] : zero;
    const uint4 qx = reinterpret_cast<const uint4*>(sample)[v];
    const uint4 ql = CORR > 0 ? reinterpret_cast<const uint4*>(last_sample)[v] : zero;
    const uint4 q0 = (CORR > 0 || PRED == 2) ? reinterpret_cast<const uint4*>(m0)[v] : zero;
    const uint4 q1 = CORR == 2 ? reinterpret_cast<const uint4*>(m1)[v] : zero;
    const uint32_t wc[4] = {qc.x, qc.y, qc.z, qc.w}, wu[4] = {qu.x, qu.y, qu.z, qu.w}, wx[4] = {qx.x, qx.y, qx.z, qx.w};
    const uint32_t wl[4] = {ql.x, ql.y, ql.z, ql.w}, w0[4] = {q0.x, q0.y, q0.z, q0.w}, w1[4] = {q1.x, q1.y, q1.z, q1.w};
    uint32_t om[4], oc[4], on[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float ma, ca, na, mb, cb, nb;
      unipc_element<CFG, CORR, PRED>(k, bf_lo(wc[i]), bf_lo(wu[i]), bf_lo(wx[i]), bf_lo(wl[i]), bf_lo(w0[i]), bf_lo(w1[i]), ma, ca, na);
      unipc_element<CFG, CORR, PRED>(k, bf_hi(wc[i]), bf_hi(wu[i]), bf_hi(wx[i]), bf_hi(wl[i]), bf_hi(w0[i]), bf_hi(w1[i]), mb, cb, nb);
      om[i] = pack_bf16(ma, mb); oc[i] = pack_bf16(ca, cb); on[i] = pack_bf16(na, nb);
    }
    reinterpret_cast<uint4*>(m_out)[v] = make_uint4(om[0], om[1], om[2], om[3]);
    reinterpret_cast<uint4*>(sample_out)[v] = make_uint4(oc[0], oc[1], oc[2], oc[3]);
    reinterpret_cast<uint4*>(prev_out)[v] = make_uint4(on[0], on[1], on[2], on[3]);
  }
  // ragged tail (n % 8 elements), one thread each
  const long long tail = nvec * 8 + blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (tail < n) {
    auto ld = [&](const __nv_bfloat16* p, bool on) { return on ? __bfloat162float(p[tail]) : 0.0f; };
    float mn, xc, xn;
    unipc_element<CFG, CORR, PRED>(k, ld(flow_cond, true), ld(flow_uncond, CFG), ld(sample, true), ld(last_sample, CORR > 0),
                                   ld(m0, CORR > 0 || PRED == 2), ld(m1, CORR == 2), mn, xc, xn);
    m_out[tail] = __float2bfloat16_rn(mn);
    sample_out[tail] = __float2bfloat16_rn(xc);
    prev_out[tail] = __float2bfloat16_rn(xn);
  }
}

template <bool CFG, int CORR>
static void launch_pred(int pred, int blocks, cudaStream_t st, const __nv_bfloat16* fc, const __nv_bfloat16* fu,
                        const __nv_bfloat16* x, const __nv_bfloat16* xl, const __nv_bfloat16* m0, const __nv_bfloat16* m1,
                        __nv_bfloat16* mo, __nv_bfloat16* so, __nv_bfloat16* po, long long n, const UniPCCoef& k) {
  if (pred == 2) cfg_unipc_step_kernel<CFG, CORR, 2><<<blocks, 256, 0, st>>>(fc, fu, x, xl, m0, m1, mo, so, po, n, k);
  else cfg_unipc_step_kernel<CFG, CORR, 1><<<blocks, 256, 0, st>>>(fc, fu, x, xl, m0, m1, mo, so, po, n, k);
}

}  // namespace sfb

using namespace sfb;
typedef __nv_bfloat16 bf16;

extern "C" int sfb_cfg_unipc_step(const void* flow_cond, const void* flow_uncond, const void* sample,
                                  const void* last_sample, const void* m0, const void* m1, void* m_out,
                                  void* sample_out, void* prev_out, long long n, const float* coef,
                                  int corrector_order, int predictor_order, void* stream) {
  if (n <= 0 || coef == nullptr || flow_cond == nullptr || sample == nullptr || m_out == nullptr ||
      sample_out == nullptr || prev_out == nullptr) {
    set_error("sfb_cfg_unipc_step: null pointer or empty tensor");
    return SFB_ERR_INVALID;
  }
  if (corrector_order < 0 || corrector_order > 2 || predictor_order < 1 || predictor_order > 2) {
    set_error("sfb_cfg_unipc_step: corrector order %d (0..2) / predictor order %d (1..2) unsupported", corrector_order,
              predictor_order);
    return SFB_ERR_INVALID;
  }
  if ((corrector_order > 0 && (last_sample == nullptr || m0 == nullptr)) || (corrector_order == 2 && m1 == nullptr) ||
      (predictor_order == 2 && m0 == nullptr)) {
    set_error("sfb_cfg_unipc_step: history tensors missing for corrector order %d / predictor order %d", corrector_order,
              predictor_order);
    return SFB_ERR_INVALID;
  }
  const uintptr_t align = (uintptr_t)flow_cond | (uintptr_t)flow_uncond | (uintptr_t)sample | (uintptr_t)last_sample |
                          (uintptr_t)m0 | (uintptr_t)m1 | (uintptr_t)m_out | (uintptr_t)sample_out | (uintptr_t)prev_out;
  if (align & 15) { set_error("sfb_cfg_unipc_step: tensors must be 16-byte aligned"); return SFB_ERR_INVALID; }
  UniPCCoef k;
  k.guidance = coef[0]; k.sigma = coef[1];
  k.c_x = coef[2]; k.c_m0 = coef[3]; k.c_b = coef[4]; k.c_inv_rk = coef[5]; k.c_rho0 = coef[6]; k.c_rho1 = coef[7];
  k.p_x = coef[8]; k.p_m0 = coef[9]; k.p_b = coef[10]; k.p_inv_rk = coef[11];
  const long long work = (n + 7) / 8;
  long long blocks = (work + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  const bf16 *fc = (const bf16*)flow_cond, *fu = (const bf16*)flow_uncond, *x = (const bf16*)sample,
             *xl = (const bf16*)last_sample, *h0 = (const bf16*)m0, *h1 = (const bf16*)m1;
  bf16 *mo = (bf16*)m_out, *so = (bf16*)sample_out, *po = (bf16*)prev_out;
  cudaStream_t st = (cudaStream_t)stream;
  const int b = (int)blocks;
  if (fu != nullptr) {
    if (corrector_order == 0) launch_pred<true, 0>(predictor_order, b, st, fc, fu, x, xl, h0, h1, mo, so, po, n, k);
    else if (corrector_order == 1) launch_pred<true, 1>(predictor_order, b, st, fc, fu, x, xl, h0, h1, mo, so, po, n, k);
    else launch_pred<true, 2>(predictor_order, b, st, fc, fu, x, xl, h0, h1, mo, so, po, n, k);
  } else {
    if (corrector_order == 0) launch_pred<false, 0>(predictor_order, b, st, fc, fu, x, xl, h0, h1, mo, so, po, n, k);
    else if (corrector_order == 1) launch_pred<false, 1>(predictor_order, b, st, fc, fu, x, xl, h0, h1, mo, so, po, n, k);
    else launch_pred<false, 2>(predictor_order, b, st, fc, fu, x, xl, h0, h1, mo, so, po, n, k);
  }
  return check_cuda(cudaGetLastError(), "cfg_unipc_step launch");
}
