// HBM-bound fused kernels of the rollout hot path (one warp per token row, 16-byte vector
// loads/stores, warp-shuffle reductions, fp32 math with the reference's bf16 rounding points).
//
//   sfb_modulation_table   e = modulation + e0                  causal_model.py:310, :365
//   sfb_ln_modulate        LN(x) * (1 + scale) + shift          causal_model.py:315, :327-328, :366
//   sfb_ln_affine          LN(x) * w + b  (norm3)               causal_model.py:324 / model.py:89-99
//   sfb_qk_norm_rope       RMSNorm(q|k) -> 3-D RoPE -> q buffer / KV-cache slot (+ V copy)
//                                                               model.py:70-86, causal_model.py:28-56,222-229
//   sfb_rmsnorm            RMSNorm (cross-attention q / context k)   model.py:172,177
//   sfb_kv_roll            rolling-window eviction: shift the kept K/V rows left past the sink   causal_model.py:212-221
//   sfb_patchify           Conv3d(k=s=(1,2,2)) im2col gather    causal_model.py:775-778
//   sfb_sinusoid           timestep sinusoid table (f64)        model.py:15-25
//   sfb_skinny_linear      few-row Linear (+SiLU) for the time MLPs   causal_model.py:464-467
//   sfb_head_finish        unpatchify + flow -> x0 (f64)        causal_model.py:1081-1104, wan_wrapper.py:204-228
//   sfb_add_noise          (1 - sigma) x0 + sigma noise         scheduler.py:159-176
#include <math.h>

#include <type_traits>

#include "common.cuh"

namespace sfb {

int device_sm_count();

__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
__device__ __forceinline__ void unpack8(const uint4& q, float (&f)[8]) {
  f[0] = bf_lo(q.x); f[1] = bf_hi(q.x); f[2] = bf_lo(q.y); f[3] = bf_hi(q.y);
  f[4] = bf_lo(q.z); f[5] = bf_hi(q.z); f[6] = bf_lo(q.w); f[7] = bf_hi(q.w);
}
// Stops the compiler from keeping the unpacked fp32 copy of a packed row alive across passes (it would otherwise
// CSE the unpacks and need 2x the registers): after this the vector has to be unpacked again.
__device__ __forceinline__ void forget_unpacked(uint4& q) {
  asm volatile("" : "+r"(q.x), "+r"(q.y), "+r"(q.z), "+r"(q.w));
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 q;
  q.x = pack_bf16(f[0], f[1]); q.y = pack_bf16(f[2], f[3]);
  q.z = pack_bf16(f[4], f[5]); q.w = pack_bf16(f[6], f[7]);
  return q;
}

constexpr int ROW_WARPS = 8;   // warps (= rows) per block for the row kernels

// ------------------------------------------------------------------------------------
// modulation table: out[l][r][g][c] = bf16(mod[l][g][c] + e[r][g * e_group_stride + c])
// ------------------------------------------------------------------------------------
__global__ void modulation_table_kernel(const __nv_bfloat16* __restrict__ mod, const __nv_bfloat16* __restrict__ e,
                                        __nv_bfloat16* __restrict__ out, int NL, int R, int G, int C,
                                        long long e_row_stride, long long e_group_stride) {
  const long long total = (long long)NL * R * G * (C / 8);
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int c8 = idx % (C / 8);
    long long rest = idx / (C / 8);
    const int g = rest % G; rest /= G;
    const int r = rest % R;
    const int l = rest / R;
    float a[8], b[8];
    unpack8(ldg16(mod + ((long long)l * G + g) * C + c8 * 8), a);
    unpack8(ldg16(e + r * e_row_stride + g * e_group_stride + c8 * 8), b);
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] += b[i];
    *reinterpret_cast<uint4*>(out + (((long long)l * R + r) * G + g) * C + c8 * 8) = pack8(a);
  }
}

// ------------------------------------------------------------------------------------
// LayerNorm (no affine) + adaLN modulation, or LayerNorm with affine.  NV = C / 256.
// ------------------------------------------------------------------------------------
// resident blocks per SM asked of the compiler: 4 (64 registers) while a packed row fits, 1 (255 registers) for wide rows (C = 5120)
constexpr int row_kernel_blocks(int nv) { return nv <= 8 ? 4 : 1; }

template <int NV, bool AFFINE>
__global__ void __launch_bounds__(ROW_WARPS * 32, row_kernel_blocks(NV))   // NV <= 8: 32 rows per SM in flight, a 4680-row chunk is one wave
ln_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy, int rows,
          float eps, const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale,
          long long mod_stride, int rows_per_mod, int row_offset, const __nv_bfloat16* __restrict__ w,
          const __nv_bfloat16* __restrict__ b) {
  constexpr int C = NV * 256;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // the row stays in registers as packed bf16 (NV x 16 bytes per lane) and is unpacked per pass: half the registers
  // of an fp32 copy, which is what lets 32 rows per SM be resident
  uint4 raw[NV];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    raw[k] = ldg16(x + row * ldx + (k * 32 + lane) * 8);
    float v[8];
    unpack8(raw[k], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) sum += v[i];
  }
  const float mean = warp_sum(sum) * (1.0f / C);
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float v[8];
    forget_unpacked(raw[k]);
    unpack8(raw[k], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) { const float d = v[i] - mean; sq += d * d; }
  }
  const float rstd = rsqrtf(warp_sum(sq) * (1.0f / C) + eps);
#pragma unroll
  for (int k = 0; k < NV; ++k) forget_unpacked(raw[k]);
  const long long mrow = AFFINE ? 0 : (long long)((row + row_offset) / rows_per_mod) * mod_stride;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c0 = (k * 32 + lane) * 8;
    float v[8], o[8];
    unpack8(raw[k], v);
    if (AFFINE) {
      float ww[8], bb[8];
      unpack8(ldg16(w + c0), ww);
      unpack8(ldg16(b + c0), bb);
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (v[i] - mean) * rstd * ww[i] + bb[i];
    } else {
      const uint4 scq = ldg16(scale + mrow + c0), shq = ldg16(shift + mrow + c0);
      const uint32_t scw[4] = {scq.x, scq.y, scq.z, scq.w}, shw[4] = {shq.x, shq.y, shq.z, shq.w};
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        // bf16(bf16(LN(x)) * bf16(1 + scale)) + shift, two elements at a time
        const uint32_t n2 = round_pair((v[2 * i] - mean) * rstd, (v[2 * i + 1] - mean) * rstd);
        const uint32_t s2 = round_pair(1.0f + bf_lo(scw[i]), 1.0f + bf_hi(scw[i]));
        const uint32_t p2 = mul_bf16x2(n2, s2);
        o[2 * i] = bf_lo(p2) + bf_lo(shw[i]);
        o[2 * i + 1] = bf_hi(p2) + bf_hi(shw[i]);
      }
    }
    *reinterpret_cast<uint4*>(y + row * ldy + c0) = pack8(o);
  }
}

// Streaming form of LayerNorm + adaLN modulation: the row's (mean, rstd) come from the statistics records the GEMM that
// produced x wrote in its epilogue (sfb_gemm_bf16_stats), so there is no reduction pass over the data and every 16-byte
// vector is loaded, normalised, modulated and stored on its own.  (The resident-row kernel above is instruction-bound:
// three unpack passes, 52 % issue-active at 2.9 TB/s -- profiles/r02_ncu_kernels.json.)
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32, row_kernel_blocks(NV))
ln_stream_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy, int rows,
                 float eps, const __nv_bfloat16* __restrict__ shift, const __nv_bfloat16* __restrict__ scale,
                 long long mod_stride, int rows_per_mod, int row_offset, const float2* __restrict__ stats, int stats_ld) {
  constexpr int C = NV * 256, CHUNKS = C / STATS_CHUNK;
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // Chan's merge of the chunk records, one record (or a few) per lane
  const float2* rec = stats + (long long)row * stats_ld;
  float msum = 0.f;
  for (int j = lane; j < CHUNKS; j += 32) msum += __ldg(rec + j).x;
  const float mean = warp_sum(msum) * (1.0f / CHUNKS);
  float m2 = 0.f;
  for (int j = lane; j < CHUNKS; j += 32) {
    const float2 r = __ldg(rec + j);
    const float d = r.x - mean;
    m2 += r.y + (float)STATS_CHUNK * d * d;
  }
  const float rstd = rsqrtf(warp_sum(m2) * (1.0f / C) + eps);
  const long long mrow = (long long)((row + row_offset) / rows_per_mod) * mod_stride;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    const int c0 = (k * 32 + lane) * 8;
    float v[8], o[8];
    unpack8(ldg16(x + row * ldx + c0), v);
    const uint4 scq = ldg16(scale + mrow + c0), shq = ldg16(shift + mrow + c0);
    const uint32_t scw[4] = {scq.x, scq.y, scq.z, scq.w}, shw[4] = {shq.x, shq.y, shq.z, shq.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      // bf16(bf16(LN(x)) * bf16(1 + scale)) + shift, two elements at a time (same rounding chain as ln_kernel)
      const uint32_t n2 = round_pair((v[2 * i] - mean) * rstd, (v[2 * i + 1] - mean) * rstd);
      const uint32_t s2 = round_pair(1.0f + bf_lo(scw[i]), 1.0f + bf_hi(scw[i]));
      const uint32_t p2 = mul_bf16x2(n2, s2);
      o[2 * i] = bf_lo(p2) + bf_lo(shw[i]);
      o[2 * i + 1] = bf_hi(p2) + bf_hi(shw[i]);
    }
    *reinterpret_cast<uint4*>(y + row * ldy + c0) = pack8(o);
  }
}

// ------------------------------------------------------------------------------------
// RMSNorm over the full channel width (+ optional 3-D RoPE), NV = C / 256
// ------------------------------------------------------------------------------------
// Destinations of the rotated q / k rows (and the V copy) when the heads of a token are dealt to several head
// groups (Ulysses head-parallel attention: group g lives on rank g and the pointers are peer-mapped memory, so
// this kernel's stores ARE the all-to-all).  groups == 1 is the single-GPU case.
struct HeadScatter {
  __nv_bfloat16* q[8];
  __nv_bfloat16* k[8];
  __nv_bfloat16* v[8];
  int groups;
  int group_cols;     // heads_per_group * head_dim
  int token_offset;   // chunk-global token index of local row 0 (RoPE position and destination row)
};

struct RopeGeom {
  int L;            // tokens per sample in this call
  int Hh, Ww;       // token grid (height, width) of one frame
  int start_frame;  // frame offset of the chunk (current_start // (Hh*Ww))
  const int* start_frame_dev;   // if set, the frame offset is read from device memory instead (CUDA-graph replay across chunks)
  int n_f, n_h;     // complex pairs on the frame / height axes (22, 21 for d=128); rest is width
};

// Loads one row as packed bf16 and returns rsqrt(mean(x^2) + eps) (full-width RMSNorm statistics).
template <int NV>
__device__ __forceinline__ float rms_load(const __nv_bfloat16* __restrict__ src, float eps, int lane, uint4 (&raw)[NV]) {
  constexpr int C = NV * 256;
  float sq = 0.f;
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    raw[k] = ldg16(src + (k * 32 + lane) * 8);
    float v[8];
    unpack8(raw[k], v);
#pragma unroll
    for (int i = 0; i < 8; ++i) sq += v[i] * v[i];
  }
  return rsqrtf(warp_sum(sq) * (1.0f / C) + eps);
}
// k-th 16-byte vector of the normalised row: bf16(bf16(x * rstd) * w)
template <int NV>
__device__ __forceinline__ void rms_apply(const uint4& raw, const __nv_bfloat16* __restrict__ w, float rstd, int k, int lane,
                                          float (&v)[8]) {
  uint4 r = raw;
  forget_unpacked(r);
  unpack8(r, v);
  const uint4 wq4 = ldg16(w + (k * 32 + lane) * 8);
  const uint32_t ww[4] = {wq4.x, wq4.y, wq4.z, wq4.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const uint32_t p2 = mul_bf16x2(round_pair(v[2 * i] * rstd, v[2 * i + 1] * rstd), ww[i]);   // bf16(bf16(x * rstd) * w)
    v[2 * i] = bf_lo(p2);
    v[2 * i + 1] = bf_hi(p2);
  }
}

template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32, row_kernel_blocks(NV))
rmsnorm_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, __nv_bfloat16* __restrict__ y, long long ldy,
               int rows, float eps, const __nv_bfloat16* __restrict__ w) {
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  uint4 raw[NV];
  const float rstd = rms_load<NV>(x + row * ldx, eps, lane, raw);
#pragma unroll
  for (int k = 0; k < NV; ++k) {
    float v[8];
    rms_apply<NV>(raw[k], w, rstd, k, lane, v);
    *reinterpret_cast<uint4*>(y + row * ldy + (k * 32 + lane) * 8) = pack8(v);
  }
}

// One warp per token: q and k rows normalised + rotated; q -> q_out, k -> cache slot, v -> cache slot.
// Rows are (sample b, token n): source row = b * L + n; destinations use their own batch strides.
// SCATTER = false: one destination (single GPU) -- no dynamic indexing of the destination table (which costs a local
// copy of the parameter arrays and, under the 64-register cap, spills).
template <int NV, bool SCATTER>
__global__ void __launch_bounds__(ROW_WARPS * 32, row_kernel_blocks(NV))
qk_norm_rope_kernel(const __nv_bfloat16* __restrict__ q_in, long long ldq, const __nv_bfloat16* __restrict__ k_in,
                    long long ldk, const __nv_bfloat16* __restrict__ v_in, long long ldv,
                    const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk, float eps,
                    const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, int head_dim,
                    RopeGeom g, int rows, const HeadScatter hs, long long q_out_row, long long q_out_batch,
                    long long kv_out_row, long long kv_out_batch) {
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row / g.L, n = row - b * g.L + hs.token_offset;
  const int fhw = g.Hh * g.Ww;
  const int f = n / fhw, rem = n - f * fhw;
  const int hh = rem / g.Ww, ww = rem - hh * g.Ww;
  const int half = head_dim / 2;
  const int pos_f = (g.start_frame_dev != nullptr ? __ldg(g.start_frame_dev) : g.start_frame) + f;

  // A lane's 16-byte vectors all start at column (k * 32 + lane) * 8, i.e. at complex pair (lane % 16) * 4 of their head
  // whenever head_dim is 128: the 4 (cos, sin) pairs of this token are the same for every vector of q and of k, so
  // they are fetched once (8 scalar loads instead of 96 per lane -- the table lookups were most of the LSU work).
  const bool fixed_pairs = head_dim == 128;
  float cs4[4], sn4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pi = (lane & 15) * 4 + i;
    const int pos = pi < g.n_f ? pos_f : (pi < g.n_f + g.n_h ? hh : ww);
    cs4[i] = fixed_pairs ? __ldg(cos_tab + pos * half + pi) : 0.f;
    sn4[i] = fixed_pairs ? __ldg(sin_tab + pos * half + pi) : 0.f;
  }

  // q then k: load the row (packed), full-width RMS statistics, then per 16-byte vector: normalise, rotate the 4
  // complex pairs by the (frame | height | width) angles of this token, store to the vector's head-group destination
  uint4 raw[NV];
#pragma unroll 1
  for (int which = 0; which < 2; ++which) {
    const __nv_bfloat16* src = which == 0 ? q_in + row * ldq : k_in + row * ldk;
    const __nv_bfloat16* wgt = which == 0 ? wq : wk;
    const long long off = which == 0 ? b * q_out_batch + n * q_out_row : b * kv_out_batch + n * kv_out_row;
    const float rstd = rms_load<NV>(src, eps, lane, raw);
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = (k * 32 + lane) * 8;
      const int pair0 = (c0 % head_dim) / 2;
      float v[8], o[8];
      rms_apply<NV>(raw[k], wgt, rstd, k, lane, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        float cs = cs4[i], sn = sn4[i];
        if (!fixed_pairs) {
          const int pi = pair0 + i;
          const int pos = pi < g.n_f ? pos_f : (pi < g.n_f + g.n_h ? hh : ww);
          cs = __ldg(cos_tab + pos * half + pi);
          sn = __ldg(sin_tab + pos * half + pi);
        }
        o[2 * i] = v[2 * i] * cs - v[2 * i + 1] * sn;
        o[2 * i + 1] = v[2 * i] * sn + v[2 * i + 1] * cs;
      }
      const int grp = SCATTER ? c0 / hs.group_cols : 0;
      __nv_bfloat16* dst = (SCATTER ? (which == 0 ? hs.q[grp] : hs.k[grp]) : (which == 0 ? hs.q[0] : hs.k[0])) + off +
                           (c0 - grp * hs.group_cols);
      *reinterpret_cast<uint4*>(dst) = pack8(o);
    }
  }
  if (v_in != nullptr) {
    const long long kv_off = b * kv_out_batch + n * kv_out_row;
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      const int c0 = (k * 32 + lane) * 8;
      const int grp = SCATTER ? c0 / hs.group_cols : 0;
      *reinterpret_cast<uint4*>((SCATTER ? hs.v[grp] : hs.v[0]) + kv_off + (c0 - grp * hs.group_cols)) = ldg16(v_in + row * ldv + c0);
    }
  }
}

// Streaming form of qk_norm_rope: the RMS statistics of the q and k rows come from the statistics records the QKV
// projection's epilogue wrote (sfb_gemm_bf16_stats), so a row no longer has to sit in registers between a reduction pass
// and an apply pass -- every 16-byte vector is loaded, normalised, rotated and stored on its own.  (The resident-row
// kernel above spills 35 registers under its 64-register cap: profiles/r02_sass_summary.json.)
template <int NV>
__global__ void __launch_bounds__(ROW_WARPS * 32, row_kernel_blocks(NV))
qk_rope_stream_kernel(const __nv_bfloat16* __restrict__ q_in, long long ldq, const __nv_bfloat16* __restrict__ k_in,
                      long long ldk, const __nv_bfloat16* __restrict__ v_in, long long ldv,
                      const __nv_bfloat16* __restrict__ wq, const __nv_bfloat16* __restrict__ wk, float eps,
                      const float2* __restrict__ stats, int stats_ld, int q_chunk0, int k_chunk0,
                      const float* __restrict__ cos_tab, const float* __restrict__ sin_tab, RopeGeom g, int rows,
                      __nv_bfloat16* __restrict__ q_out, __nv_bfloat16* __restrict__ k_out, __nv_bfloat16* __restrict__ v_out,
                      long long q_out_row, long long q_out_batch, long long kv_out_row, long long kv_out_batch) {
  constexpr int C = NV * 256, CHUNKS = C / STATS_CHUNK, HALF = 64;   // head_dim 128
  const int row = blockIdx.x * ROW_WARPS + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  // sum of squares of the q and k rows from their chunk records (mean, M2): sum x^2 = M2 + n mean^2
  float ssq = 0.f, ssk = 0.f;
  const float2* rec = stats + (long long)row * stats_ld;
  for (int j = lane; j < CHUNKS; j += 32) {
    const float2 a = __ldg(rec + q_chunk0 + j), b = __ldg(rec + k_chunk0 + j);
    ssq += a.y + (float)STATS_CHUNK * a.x * a.x;
    ssk += b.y + (float)STATS_CHUNK * b.x * b.x;
  }
  const float rstd_q = rsqrtf(warp_sum(ssq) * (1.0f / C) + eps), rstd_k = rsqrtf(warp_sum(ssk) * (1.0f / C) + eps);

  const int b = row / g.L, n = row - b * g.L;
  const int fhw = g.Hh * g.Ww;
  const int f = n / fhw, rem = n - f * fhw;
  const int hh = rem / g.Ww, ww = rem - hh * g.Ww;
  const int pos_f = (g.start_frame_dev != nullptr ? __ldg(g.start_frame_dev) : g.start_frame) + f;
  // a lane's vectors all start at complex pair (lane % 16) * 4 of their head: 4 (cos, sin) pairs per token and lane
  float cs4[4], sn4[4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pi = (lane & 15) * 4 + i;
    const int pos = pi < g.n_f ? pos_f : (pi < g.n_f + g.n_h ? hh : ww);
    cs4[i] = __ldg(cos_tab + pos * HALF + pi);
    sn4[i] = __ldg(sin_tab + pos * HALF + pi);
  }
  const __nv_bfloat16* qs = q_in + row * ldq;
  const __nv_bfloat16* ks = k_in + row * ldk;
  __nv_bfloat16* qd = q_out + b * q_out_batch + n * q_out_row;
  __nv_bfloat16* kd = k_out + b * kv_out_batch + n * kv_out_row;
  uint4 rq[NV], rk[NV];
#pragma unroll
  for (int k = 0; k < NV; ++k) { rq[k] = ldg16(qs + (k * 32 + lane) * 8); rk[k] = ldg16(ks + (k * 32 + lane) * 8); }
#pragma unroll
  for (int which = 0; which < 2; ++which) {
#pragma unroll
    for (int k = 0; k < NV; ++k) {
      float v[8], o[8];
      rms_apply<NV>(which == 0 ? rq[k] : rk[k], which == 0 ? wq : wk, which == 0 ? rstd_q : rstd_k, k, lane, v);
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        o[2 * i] = v[2 * i] * cs4[i] - v[2 * i + 1] * sn4[i];
        o[2 * i + 1] = v[2 * i] * sn4[i] + v[2 * i + 1] * cs4[i];
      }
      *reinterpret_cast<uint4*>((which == 0 ? qd : kd) + (k * 32 + lane) * 8) = pack8(o);
    }
  }
  if (v_in != nullptr) {
    __nv_bfloat16* vd = v_out + b * kv_out_batch + n * kv_out_row;
#pragma unroll
    for (int k = 0; k < NV; ++k) *reinterpret_cast<uint4*>(vd + (k * 32 + lane) * 8) = ldg16(v_in + row * ldv + (k * 32 + lane) * 8);
  }
}

// ------------------------------------------------------------------------------------
// KV-cache roll (reference causal_model.py:212-221: `cache[:, sink:sink+keep] = cache[:, sink+ev:sink+ev+keep].clone()`)
// for ALL layers' K and V tensors in one launch per phase: rows [dst, dst + n) <- rows [dst + shift, dst + shift + n),
// n <= shift, so source and destination never overlap inside a launch; the host walks the kept range front to back.
// A token row of every (tensor, batch) is contiguous, so each (tensor, batch) moves one contiguous span of bytes.
// ------------------------------------------------------------------------------------
__global__ void kv_roll_kernel(const uint4* const* __restrict__ tensors, long long batch_stride_v, long long dst_v,
                               long long shift_v, long long n_v) {
  uint4* base = const_cast<uint4*>(tensors[blockIdx.y]) + blockIdx.z * batch_stride_v + dst_v;
  const uint4* src = base + shift_v;
  const long long stride = (long long)gridDim.x * blockDim.x;
  long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  // four independent 16-byte loads in flight per thread
  for (; i + 3 * stride < n_v; i += 4 * stride) {
    const uint4 a = __ldcs(src + i), b = __ldcs(src + i + stride), c = __ldcs(src + i + 2 * stride), d = __ldcs(src + i + 3 * stride);
    base[i] = a; base[i + stride] = b; base[i + 2 * stride] = c; base[i + 3 * stride] = d;
  }
  for (; i < n_v; i += stride) base[i] = __ldcs(src + i);
}

// ------------------------------------------------------------------------------------
// patchify: x[b][c][f][y][x] (element strides) -> tokens[(b,f,hh,ww)][c*4 + ph*2 + pw]
// ------------------------------------------------------------------------------------
__global__ void patchify_kernel(const __nv_bfloat16* __restrict__ x, long long sb, long long sc, long long sf,
                                long long sy, long long sx, __nv_bfloat16* __restrict__ out, int B, int Cin, int F,
                                int Hh, int Ww) {
  const long long total = (long long)B * F * Hh * Ww * Cin * 2;   // one thread per (token, c, ph): 2 pixels
  const long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const int cp = idx % (Cin * 2);
  long long tok = idx / (Cin * 2);
  const int c = cp >> 1, ph = cp & 1;
  const int ww = tok % Ww; long long r = tok / Ww;
  const int hh = r % Hh; r /= Hh;
  const int f = r % F;
  const int b = r / F;
  const __nv_bfloat16* src = x + b * sb + c * sc + f * sf + (2 * hh + ph) * sy + (2 * ww) * sx;
  __nv_bfloat16* dst = out + tok * (Cin * 4) + c * 4 + ph * 2;
  dst[0] = src[0];
  dst[1] = src[sx];
}

// ------------------------------------------------------------------------------------
// sinusoid: out[i][j] = cos(t_i w_j) (j < half) | sin(t_i w_{j-half}),  w_j = 10000^(-j/half), f64
// ------------------------------------------------------------------------------------
__device__ __forceinline__ double load_timestep(const void* t, int dtype, int i) {
  switch (dtype) {
    case 0: return (double)static_cast<const float*>(t)[i];
    case 1: return (double)static_cast<const long long*>(t)[i];
    case 2: return static_cast<const double*>(t)[i];
    default: return (double)__bfloat162float(static_cast<const __nv_bfloat16*>(t)[i]);
  }
}

__global__ void sinusoid_kernel(const void* __restrict__ t, int t_dtype, __nv_bfloat16* __restrict__ out, int n,
                                int freq_dim) {
  const int half = freq_dim / 2;
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * freq_dim) return;
  const int i = idx / freq_dim, j = idx % freq_dim;
  const int jj = j < half ? j : j - half;
  const double w = pow(10000.0, -((double)jj / (double)half));
  const double a = load_timestep(t, t_dtype, i) * w;
  out[idx] = __float2bfloat16_rn((float)(j < half ? cos(a) : sin(a)));   // torch: double -> float -> bf16
}

// ------------------------------------------------------------------------------------
// skinny linear: y[m][n] = bf16(sum_k in(x[m][k]) w[n][k] + b[n]), M <= 8 per pass; in = id | silu
// ------------------------------------------------------------------------------------
constexpr int SKINNY_M = 8;
// Each block stages the (activated) input rows once and its 8 warps then walk `cols_per_warp` output columns each: with
// one column per warp every block repeated the staging (8 x K loads + SiLU exponentials) for 24 KB of weights and the
// three time-MLP launches took 36 us each against ~6 us of weight streaming (profiles/r02q_ncu_launch_shares.json).
__global__ void __launch_bounds__(256)
skinny_linear_kernel(const __nv_bfloat16* __restrict__ x, long long ldx, const __nv_bfloat16* __restrict__ w,
                     long long ldw, const __nv_bfloat16* __restrict__ bias, __nv_bfloat16* __restrict__ y,
                     long long ldy, int M, int N, int K, int silu_in, int cols_per_warp) {
  extern __shared__ float xs[];   // [SKINNY_M][K] fp32 (already activated)
  const int m0 = blockIdx.y * SKINNY_M;
  const int mcount = min(SKINNY_M, M - m0);
  for (int i = threadIdx.x; i < mcount * K; i += blockDim.x) {
    const int m = i / K, k = i - m * K;
    float v = __bfloat162float(x[(m0 + m) * ldx + k]);
    if (silu_in) v = bf16r(v / (1.0f + __expf(-v)));
    xs[i] = v;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int c = 0; c < cols_per_warp; ++c) {
    const int n = (blockIdx.x * 8 + warp) * cols_per_warp + c;
    if (n >= N) return;
    float acc[SKINNY_M];
#pragma unroll
    for (int m = 0; m < SKINNY_M; ++m) acc[m] = 0.f;
    for (int k0 = lane * 8; k0 < K; k0 += 256) {
      float wv[8];
      unpack8(ldg16(w + n * ldw + k0), wv);
#pragma unroll
      for (int m = 0; m < SKINNY_M; ++m) {
        if (m < mcount) {   // block-uniform
          const float4 a = *reinterpret_cast<const float4*>(xs + m * K + k0);
          const float4 c4 = *reinterpret_cast<const float4*>(xs + m * K + k0 + 4);
          acc[m] += a.x * wv[0] + a.y * wv[1] + a.z * wv[2] + a.w * wv[3] + c4.x * wv[4] + c4.y * wv[5] + c4.z * wv[6] +
                    c4.w * wv[7];
        }
      }
    }
#pragma unroll
    for (int m = 0; m < SKINNY_M; ++m)
      if (m < mcount) acc[m] = warp_sum(acc[m]);
    if (lane == 0) {
      const float bv = bias ? __bfloat162float(bias[n]) : 0.f;
      for (int m = 0; m < mcount; ++m) y[(m0 + m) * ldy + n] = __float2bfloat16_rn(acc[m] + bv);
    }
  }
}

// ------------------------------------------------------------------------------------
// nearest-timestep lookup shared by the sampler kernels (first index of the minimum, like
// torch.argmin); returns sigma of that index.  All threads of the block must call it.
// ------------------------------------------------------------------------------------
template <typename T>
__device__ T block_sigma_lookup(const float* __restrict__ timesteps, const float* __restrict__ sigmas, int n_tab,
                                T t) {
  __shared__ double s_best[32];
  __shared__ int s_idx[32];
  double best = 1e300;
  int bi = 0x7fffffff;
  for (int i = threadIdx.x; i < n_tab; i += blockDim.x) {
    const T d = (T)timesteps[i] - t;
    const double a = (double)(d < 0 ? -d : d);
    if (a < best || (a == best && i < bi)) { best = a; bi = i; }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const double ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
    if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
  }
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = (blockDim.x + 31) >> 5;
  if (lane == 0) { s_best[warp] = best; s_idx[warp] = bi; }
  __syncthreads();
  if (warp == 0) {
    best = lane < nw ? s_best[lane] : 1e300;
    bi = lane < nw ? s_idx[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const double ob = __shfl_xor_sync(0xffffffffu, best, o);
      const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
      if (ob < best || (ob == best && oi < bi)) { best = ob; bi = oi; }
    }
    if (lane == 0) s_idx[0] = bi;
  }
  __syncthreads();
  return (T)sigmas[s_idx[0]];
}

// ------------------------------------------------------------------------------------
// head_finish: head GEMM output [B*L][ (ph*2+pw)*Cout + c ] -> flow[b][f][c][y][x] and
// x0 = x_t - sigma_t * flow in f64.  One block per (b, f, slab).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
head_finish_kernel(const __nv_bfloat16* __restrict__ head_out, long long ldh, const __nv_bfloat16* __restrict__ xt,
                   long long xs_b, long long xs_f, long long xs_c, long long xs_y, long long xs_x,
                   const void* __restrict__ timestep, int t_dtype, const float* __restrict__ timesteps,
                   const float* __restrict__ sigmas, int n_tab, __nv_bfloat16* __restrict__ flow,
                   __nv_bfloat16* __restrict__ x0, int B, int F, int Cout, int Hh, int Ww) {
  const int bf = blockIdx.x;
  const int b = bf / F, f = bf % F;
  const bool want_x0 = x0 != nullptr;   // block-uniform
  double sigma = 0.0;
  if (want_x0) sigma = block_sigma_lookup<double>(timesteps, sigmas, n_tab, load_timestep(timestep, t_dtype, bf));
  const int H = Hh * 2, W = Ww * 2;
  const int per_frame = Cout * H * W;
  const long long L = (long long)F * Hh * Ww;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < per_frame; idx += gridDim.y * blockDim.x) {
    const int xw = idx % W;
    int r = idx / W;
    const int yh = r % H;
    const int c = r / H;
    const long long tok = (long long)b * L + (long long)f * Hh * Ww + (yh >> 1) * Ww + (xw >> 1);
    const int feat = ((yh & 1) * 2 + (xw & 1)) * Cout + c;
    const __nv_bfloat16 fl = head_out[tok * ldh + feat];
    const double xv = want_x0 ? (double)__bfloat162float(xt[b * xs_b + f * xs_f + c * xs_c + yh * xs_y + xw * xs_x]) : 0.0;
    const long long o = ((long long)bf * Cout + c) * H * W + (long long)yh * W + xw;
    flow[o] = fl;
    if (want_x0)   // torch casts double -> float -> bf16; no fma contraction
      x0[o] = __float2bfloat16_rn((float)__dsub_rn(xv, __dmul_rn(sigma, (double)__bfloat162float(fl))));
  }
}

// ------------------------------------------------------------------------------------
// add_noise: out = bf16((1 - sigma) * x0 + sigma * noise), fp32, sigma by nearest timestep (fp32)
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
add_noise_kernel(const __nv_bfloat16* __restrict__ x0, const __nv_bfloat16* __restrict__ noise,
                 const void* __restrict__ timestep, int t_dtype, const float* __restrict__ timesteps,
                 const float* __restrict__ sigmas, int n_tab, __nv_bfloat16* __restrict__ out, int per_frame) {
  const int fr = blockIdx.x;
  const float sigma = block_sigma_lookup<float>(timesteps, sigmas, n_tab, (float)load_timestep(timestep, t_dtype, fr));
  const float one_minus = __fsub_rn(1.0f, sigma);
  const long long base = (long long)fr * per_frame;
  for (int idx = blockIdx.y * blockDim.x + threadIdx.x; idx < per_frame; idx += gridDim.y * blockDim.x) {
    const float a = __fmul_rn(one_minus, __bfloat162float(x0[base + idx]));
    const float c = __fmul_rn(sigma, __bfloat162float(noise[base + idx]));
    out[base + idx] = __float2bfloat16_rn(__fadd_rn(a, c));
  }
}

template <typename F>
static int dispatch_nv(int C, const char* who, F&& f) {
  switch (C) {
    case 256: return f(std::integral_constant<int, 1>{});
    case 512: return f(std::integral_constant<int, 2>{});
    case 1024: return f(std::integral_constant<int, 4>{});
    case 1536: return f(std::integral_constant<int, 6>{});
    case 2048: return f(std::integral_constant<int, 8>{});
    case 5120: return f(std::integral_constant<int, 20>{});
  }
  set_error("%s: channel width %d unsupported (256/512/1024/1536/2048/5120)", who, C);
  return SFB_ERR_INVALID;
}

}  // namespace sfb

using namespace sfb;
typedef __nv_bfloat16 bf16;

extern "C" int sfb_modulation_table(const void* mod, const void* e, void* out, int NL, int R, int G, int C,
                                    long long e_row_stride, long long e_group_stride, void* stream) {
  if (C % 8 || NL <= 0 || R <= 0 || G <= 0) { set_error("sfb_modulation_table: bad shape"); return SFB_ERR_INVALID; }
  const long long total = (long long)NL * R * G * (C / 8);
  const int blocks = (int)((total + 255) / 256);
  modulation_table_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>((const bf16*)mod, (const bf16*)e, (bf16*)out, NL, R,
                                                                   G, C, e_row_stride, e_group_stride);
  return check_cuda(cudaGetLastError(), "modulation_table launch");
}

extern "C" int sfb_ln_modulate(const void* x, long long ldx, void* y, long long ldy, int rows, int C, float eps,
                               const void* shift, const void* scale, long long mod_stride, int rows_per_mod,
                               int row_offset, void* stream) {
  if (rows <= 0 || rows_per_mod <= 0 || row_offset < 0 || (ldx % 8) || (ldy % 8) || (mod_stride % 8)) { set_error("sfb_ln_modulate: bad arguments"); return SFB_ERR_INVALID; }
  return dispatch_nv(C, "sfb_ln_modulate", [&](auto nv) {
    ln_kernel<decltype(nv)::value, false><<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const bf16*)x, ldx, (bf16*)y, ldy, rows, eps, (const bf16*)shift, (const bf16*)scale, mod_stride, rows_per_mod,
        row_offset, nullptr, nullptr);
    return check_cuda(cudaGetLastError(), "ln_modulate launch");
  });
}

// sfb_ln_modulate with the row statistics supplied by the GEMM that produced x (include/sfb200.h).
extern "C" int sfb_ln_modulate_stats(const void* x, long long ldx, void* y, long long ldy, int rows, int C, float eps,
                                     const void* shift, const void* scale, long long mod_stride, int rows_per_mod,
                                     int row_offset, const void* stats, int stats_ld, void* stream) {
  if (rows <= 0 || rows_per_mod <= 0 || row_offset < 0 || (ldx % 8) || (ldy % 8) || (mod_stride % 8)) { set_error("sfb_ln_modulate_stats: bad arguments"); return SFB_ERR_INVALID; }
  if (stats == nullptr || stats_ld < C / STATS_CHUNK) { set_error("sfb_ln_modulate_stats: statistics records [%d] do not cover C=%d", stats_ld, C); return SFB_ERR_INVALID; }
  return dispatch_nv(C, "sfb_ln_modulate_stats", [&](auto nv) {
    ln_stream_kernel<decltype(nv)::value><<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const bf16*)x, ldx, (bf16*)y, ldy, rows, eps, (const bf16*)shift, (const bf16*)scale, mod_stride, rows_per_mod,
        row_offset, (const float2*)stats, stats_ld);
    return check_cuda(cudaGetLastError(), "ln_modulate_stats launch");
  });
}

extern "C" int sfb_ln_affine(const void* x, long long ldx, void* y, long long ldy, int rows, int C, float eps,
                             const void* weight, const void* bias, void* stream) {
  if (rows <= 0 || (ldx % 8) || (ldy % 8) || !weight || !bias) { set_error("sfb_ln_affine: bad arguments"); return SFB_ERR_INVALID; }
  return dispatch_nv(C, "sfb_ln_affine", [&](auto nv) {
    ln_kernel<decltype(nv)::value, true><<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const bf16*)x, ldx, (bf16*)y, ldy, rows, eps, nullptr, nullptr, 0, 1, 0, (const bf16*)weight, (const bf16*)bias);
    return check_cuda(cudaGetLastError(), "ln_affine launch");
  });
}

extern "C" int sfb_rmsnorm(const void* x, long long ldx, void* y, long long ldy, int rows, int C, float eps,
                           const void* weight, void* stream) {
  if (rows <= 0 || (ldx % 8) || (ldy % 8) || !weight) { set_error("sfb_rmsnorm: bad arguments"); return SFB_ERR_INVALID; }
  return dispatch_nv(C, "sfb_rmsnorm", [&](auto nv) {
    rmsnorm_kernel<decltype(nv)::value><<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const bf16*)x, ldx, (bf16*)y, ldy, rows, eps, (const bf16*)weight);
    return check_cuda(cudaGetLastError(), "rmsnorm launch");
  });
}

extern "C" int sfb_qk_norm_rope(const void* q_in, long long ldq, const void* k_in, long long ldk, const void* v_in,
                                long long ldv, const void* wq, const void* wk, float eps, const float* cos_tab,
                                const float* sin_tab, int tab_rows, int B, int L, int C, int head_dim, int F, int Hh,
                                int Ww, int start_frame, const int* start_frame_dev, void* q_out, long long q_out_row,
                                long long q_out_batch, void* k_out, void* v_out, long long kv_out_row,
                                long long kv_out_batch, void* stream) {
  if (B <= 0 || L <= 0 || L != F * Hh * Ww) { set_error("sfb_qk_norm_rope: L=%d != F*H*W=%d*%d*%d", L, F, Hh, Ww); return SFB_ERR_INVALID; }
  if (head_dim % 16 || C % head_dim) { set_error("sfb_qk_norm_rope: bad head_dim %d for C=%d", head_dim, C); return SFB_ERR_INVALID; }
  if (start_frame < 0 || start_frame + F > tab_rows || Hh > tab_rows || Ww > tab_rows) { set_error("sfb_qk_norm_rope: position beyond the %d-row RoPE table (start_frame=%d F=%d)", tab_rows, start_frame, F); return SFB_ERR_INVALID; }
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (q_out_row % 8) || (kv_out_row % 8) || (q_out_batch % 8) || (kv_out_batch % 8)) { set_error("sfb_qk_norm_rope: strides must be multiples of 8"); return SFB_ERR_INVALID; }
  RopeGeom g;
  g.L = L; g.Hh = Hh; g.Ww = Ww; g.start_frame = start_frame; g.start_frame_dev = start_frame_dev;
  const int c = head_dim / 2;
  g.n_f = c - 2 * (c / 3);
  g.n_h = c / 3;
  const int rows = B * L;
  HeadScatter hs{};
  hs.q[0] = (bf16*)q_out; hs.k[0] = (bf16*)k_out; hs.v[0] = (bf16*)v_out;
  hs.groups = 1; hs.group_cols = C; hs.token_offset = 0;
  return dispatch_nv(C, "sfb_qk_norm_rope", [&](auto nv) {
    qk_norm_rope_kernel<decltype(nv)::value, false><<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const bf16*)q_in, ldq, (const bf16*)k_in, ldk, (const bf16*)v_in, ldv, (const bf16*)wq, (const bf16*)wk, eps,
        cos_tab, sin_tab, head_dim, g, rows, hs, q_out_row, q_out_batch, kv_out_row, kv_out_batch);
    return check_cuda(cudaGetLastError(), "qk_norm_rope launch");
  });
}

// sfb_qk_norm_rope with the row statistics supplied by the QKV projection (include/sfb200.h).
extern "C" int sfb_qk_norm_rope_stats(const void* q_in, long long ldq, const void* k_in, long long ldk, const void* v_in,
                                      long long ldv, const void* wq, const void* wk, float eps, const void* stats,
                                      int stats_ld, int q_chunk0, int k_chunk0, const float* cos_tab,
                                      const float* sin_tab, int tab_rows, int B, int L, int C, int head_dim, int F, int Hh,
                                      int Ww, int start_frame, const int* start_frame_dev, void* q_out, long long q_out_row,
                                      long long q_out_batch, void* k_out, void* v_out, long long kv_out_row,
                                      long long kv_out_batch, void* stream) {
  if (B <= 0 || L <= 0 || L != F * Hh * Ww) { set_error("sfb_qk_norm_rope_stats: L=%d != F*H*W=%d*%d*%d", L, F, Hh, Ww); return SFB_ERR_INVALID; }
  if (head_dim != 128 || C % head_dim) { set_error("sfb_qk_norm_rope_stats: head_dim %d (128 only) / C=%d", head_dim, C); return SFB_ERR_INVALID; }
  if (stats == nullptr || q_chunk0 < 0 || k_chunk0 < 0 || stats_ld < q_chunk0 + C / STATS_CHUNK || stats_ld < k_chunk0 + C / STATS_CHUNK) { set_error("sfb_qk_norm_rope_stats: statistics records [%d] do not cover chunks %d / %d + %d", stats_ld, q_chunk0, k_chunk0, C / STATS_CHUNK); return SFB_ERR_INVALID; }
  if (start_frame < 0 || start_frame + F > tab_rows || Hh > tab_rows || Ww > tab_rows) { set_error("sfb_qk_norm_rope_stats: position beyond the %d-row RoPE table (start_frame=%d F=%d)", tab_rows, start_frame, F); return SFB_ERR_INVALID; }
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (q_out_row % 8) || (kv_out_row % 8) || (q_out_batch % 8) || (kv_out_batch % 8)) { set_error("sfb_qk_norm_rope_stats: strides must be multiples of 8"); return SFB_ERR_INVALID; }
  RopeGeom g;
  g.L = L; g.Hh = Hh; g.Ww = Ww; g.start_frame = start_frame; g.start_frame_dev = start_frame_dev;
  const int c = head_dim / 2;
  g.n_f = c - 2 * (c / 3);
  g.n_h = c / 3;
  const int rows = B * L;
  return dispatch_nv(C, "sfb_qk_norm_rope_stats", [&](auto nv) {
    qk_rope_stream_kernel<decltype(nv)::value><<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const bf16*)q_in, ldq, (const bf16*)k_in, ldk, (const bf16*)v_in, ldv, (const bf16*)wq, (const bf16*)wk, eps,
        (const float2*)stats, stats_ld, q_chunk0, k_chunk0, cos_tab, sin_tab, g, rows, (bf16*)q_out, (bf16*)k_out,
        (bf16*)v_out, q_out_row, q_out_batch, kv_out_row, kv_out_batch);
    return check_cuda(cudaGetLastError(), "qk_norm_rope_stats launch");
  });
}

// Sequence-parallel (Ulysses) form: this rank holds `rows` consecutive tokens of the chunk starting at chunk token
// `token_offset`, with ALL heads; head group g (C / groups columns) of every row is stored to q_dst[g] / k_dst[g] /
// v_dst[g] at row (token_offset + local row).  With peer-mapped destination pointers the stores are the all-to-all
// of wan/distributed/xdit_context_parallel.py:179-184 (xFuserLongContextAttention), fused into the producer.
extern "C" int sfb_qk_norm_rope_sp(const void* q_in, long long ldq, const void* k_in, long long ldk, const void* v_in,
                                   long long ldv, const void* wq, const void* wk, float eps, const float* cos_tab,
                                   const float* sin_tab, int tab_rows, int rows, int C, int head_dim, int F, int Hh,
                                   int Ww, int start_frame, int token_offset, int groups, void* const* q_dst,
                                   long long q_dst_row, void* const* k_dst, void* const* v_dst, long long kv_dst_row,
                                   void* stream) {
  if (rows <= 0 || token_offset < 0 || token_offset + rows > F * Hh * Ww) { set_error("sfb_qk_norm_rope_sp: rows [%d, %d) outside the chunk of %d tokens", token_offset, token_offset + rows, F * Hh * Ww); return SFB_ERR_INVALID; }
  if (head_dim % 16 || C % head_dim) { set_error("sfb_qk_norm_rope_sp: bad head_dim %d for C=%d", head_dim, C); return SFB_ERR_INVALID; }
  if (groups < 1 || groups > 8 || (C / head_dim) % groups) { set_error("sfb_qk_norm_rope_sp: %d heads do not split into %d groups", C / head_dim, groups); return SFB_ERR_INVALID; }
  if (start_frame < 0 || start_frame + F > tab_rows || Hh > tab_rows || Ww > tab_rows) { set_error("sfb_qk_norm_rope_sp: position beyond the %d-row RoPE table", tab_rows); return SFB_ERR_INVALID; }
  if ((ldq % 8) || (ldk % 8) || (ldv % 8) || (q_dst_row % 8) || (kv_dst_row % 8)) { set_error("sfb_qk_norm_rope_sp: strides must be multiples of 8"); return SFB_ERR_INVALID; }
  RopeGeom g;
  g.L = F * Hh * Ww; g.Hh = Hh; g.Ww = Ww; g.start_frame = start_frame; g.start_frame_dev = nullptr;
  const int c = head_dim / 2;
  g.n_f = c - 2 * (c / 3);
  g.n_h = c / 3;
  HeadScatter hs{};
  for (int i = 0; i < groups; ++i) { hs.q[i] = (bf16*)q_dst[i]; hs.k[i] = (bf16*)k_dst[i]; hs.v[i] = (bf16*)v_dst[i]; }
  hs.groups = groups; hs.group_cols = C / groups; hs.token_offset = token_offset;
  return dispatch_nv(C, "sfb_qk_norm_rope_sp", [&](auto nv) {
    qk_norm_rope_kernel<decltype(nv)::value, true><<<(rows + ROW_WARPS - 1) / ROW_WARPS, ROW_WARPS * 32, 0, (cudaStream_t)stream>>>(
        (const bf16*)q_in, ldq, (const bf16*)k_in, ldk, (const bf16*)v_in, ldv, (const bf16*)wq, (const bf16*)wk, eps,
        cos_tab, sin_tab, head_dim, g, rows, hs, q_dst_row, 0, kv_dst_row, 0);
    return check_cuda(cudaGetLastError(), "qk_norm_rope_sp launch");
  });
}

// tensors_dev: DEVICE array of n_tensors pointers to [batch, rows, row_bytes] caches with the same geometry.
// Moves rows [src_row, src_row + n_rows) to [dst_row, dst_row + n_rows), dst_row < src_row (memmove semantics).
extern "C" int sfb_kv_roll(const void* const* tensors_dev, int n_tensors, int batch, long long batch_stride_bytes,
                           long long row_bytes, long long dst_row, long long src_row, long long n_rows, void* stream) {
  if (n_tensors <= 0 || batch <= 0 || n_rows <= 0) return SFB_OK;
  if (dst_row < 0 || src_row <= dst_row || (row_bytes % 16) || (batch_stride_bytes % 16)) {
    set_error("sfb_kv_roll: need 0 <= dst_row < src_row and 16-byte rows (dst %lld src %lld row_bytes %lld)", dst_row, src_row, row_bytes);
    return SFB_ERR_INVALID;
  }
  const int sms = device_sm_count();
  if (sms <= 0) return SFB_ERR_CUDA;
  const long long shift = src_row - dst_row;
  for (long long done = 0; done < n_rows; done += shift) {
    const long long n = (n_rows - done) < shift ? (n_rows - done) : shift;
    const long long n_v = n * row_bytes / 16;
    long long bx = (n_v + 256 * 4 - 1) / (256 * 4);
    const long long cap = (8LL * sms + (long long)n_tensors * batch - 1) / ((long long)n_tensors * batch);   // ~8 blocks per SM in total
    if (bx > cap) bx = cap;
    if (bx < 1) bx = 1;
    kv_roll_kernel<<<dim3((unsigned)bx, n_tensors, batch), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const uint4* const*>(tensors_dev), batch_stride_bytes / 16, (dst_row + done) * row_bytes / 16,
        shift * row_bytes / 16, n_v);
    if (int e = check_cuda(cudaGetLastError(), "kv_roll launch")) return e;
  }
  return SFB_OK;
}

extern "C" int sfb_patchify(const void* x, long long sb, long long sc, long long sf, long long sy, long long sx,
                            void* out, int B, int Cin, int F, int H, int W, void* stream) {
  if ((H % 2) || (W % 2) || B <= 0 || F <= 0) { set_error("sfb_patchify: H, W must be even"); return SFB_ERR_INVALID; }
  const long long total = (long long)B * F * (H / 2) * (W / 2) * Cin * 2;
  patchify_kernel<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>((const bf16*)x, sb, sc, sf, sy, sx,
                                                                              (bf16*)out, B, Cin, F, H / 2, W / 2);
  return check_cuda(cudaGetLastError(), "patchify launch");
}

extern "C" int sfb_sinusoid(const void* t, int t_dtype, void* out, int n, int freq_dim, void* stream) {
  if (n <= 0 || freq_dim % 2 || t_dtype < 0 || t_dtype > 3) { set_error("sfb_sinusoid: bad arguments"); return SFB_ERR_INVALID; }
  sinusoid_kernel<<<(n * freq_dim + 255) / 256, 256, 0, (cudaStream_t)stream>>>(t, t_dtype, (bf16*)out, n, freq_dim);
  return check_cuda(cudaGetLastError(), "sinusoid launch");
}

extern "C" int sfb_skinny_linear(const void* x, long long ldx, const void* w, long long ldw, const void* bias, void* y,
                                 long long ldy, int M, int N, int K, int silu_in, void* stream) {
  if (M <= 0 || N <= 0 || K <= 0 || (K % 256) || (ldw % 8)) { set_error("sfb_skinny_linear: need K %% 256 == 0 (got %d) and ldw %% 8 == 0", K); return SFB_ERR_INVALID; }
  const size_t smem = (size_t)SKINNY_M * K * sizeof(float);
  if (smem > 200 * 1024) { set_error("sfb_skinny_linear: K=%d too large", K); return SFB_ERR_INVALID; }
  static SmemOptIn optin;
  if (smem > 48 * 1024)
    if (int e = optin.ensure(skinny_linear_kernel, (int)smem, "cudaFuncSetAttribute(skinny)")) return e;
  // about two blocks per SM for the wide projection (N = 9216: 288 blocks of 32 columns), one column per warp when narrow
  const int sms = device_sm_count();
  int cpw = N / (8 * 2 * (sms > 0 ? sms : 148));
  cpw = cpw < 1 ? 1 : (cpw > 8 ? 8 : cpw);
  dim3 grid((N + 8 * cpw - 1) / (8 * cpw), (M + SKINNY_M - 1) / SKINNY_M);
  skinny_linear_kernel<<<grid, 256, smem, (cudaStream_t)stream>>>((const bf16*)x, ldx, (const bf16*)w, ldw,
                                                                (const bf16*)bias, (bf16*)y, ldy, M, N, K, silu_in, cpw);
  return check_cuda(cudaGetLastError(), "skinny_linear launch");
}

extern "C" int sfb_head_finish(const void* head_out, long long ldh, const void* xt, long long xs_b, long long xs_f,
                               long long xs_c, long long xs_y, long long xs_x, const void* timestep, int t_dtype,
                               const float* timesteps, const float* sigmas, int n_tab, void* flow, void* x0, int B,
                               int F, int Cout, int H, int W, void* stream) {
  if (B <= 0 || F <= 0 || (H % 2) || (W % 2) || (x0 != nullptr && (n_tab <= 0 || !xt || !timestep || !timesteps || !sigmas))) { set_error("sfb_head_finish: bad arguments"); return SFB_ERR_INVALID; }
  const int per_frame = Cout * H * W;
  dim3 grid(B * F, (per_frame + 256 * 8 - 1) / (256 * 8));
  head_finish_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)head_out, ldh, (const bf16*)xt, xs_b, xs_f,
                                                           xs_c, xs_y, xs_x, timestep, t_dtype, timesteps, sigmas, n_tab,
                                                           (bf16*)flow, (bf16*)x0, B, F, Cout, H / 2, W / 2);
  return check_cuda(cudaGetLastError(), "head_finish launch");
}

extern "C" int sfb_add_noise(const void* x0, const void* noise, const void* timestep, int t_dtype,
                             const float* timesteps, const float* sigmas, int n_tab, void* out, int n_frames,
                             int per_frame, void* stream) {
  if (n_frames <= 0 || per_frame <= 0 || n_tab <= 0) { set_error("sfb_add_noise: bad arguments"); return SFB_ERR_INVALID; }
  dim3 grid(n_frames, (per_frame + 256 * 8 - 1) / (256 * 8));
  add_noise_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const bf16*)x0, (const bf16*)noise, timestep, t_dtype,
                                                         timesteps, sigmas, n_tab, (bf16*)out, per_frame);
  return check_cuda(cudaGetLastError(), "add_noise launch");
}
