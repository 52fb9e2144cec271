// Dense (unmasked) flash attention forward for sm_100a, head_dim 128, bf16 in / fp32 accumulate.
//
// Replaces `attention()` -> flash_attn_varlen_func at wan/modules/attention.py:136-150 for the two
// call sites of the rollout:
//   * chunk self-attention over the rolling KV cache window  (wan/modules/causal_model.py:230-234)
//   * cross-attention against the cached T5 context          (wan/modules/model.py:189)
// K and V are read in place from the cache layout [B, S, H, 128] (row stride H*128) through TMA, so
// there is no varlen packing / cu_seqlens traffic and no copy of the window.
//
// One CTA per (pair of 128-row query tiles, head, batch); 320 threads:
//   warp 0     TMA producer: Q pair once, then K / V tiles through two 2-deep rings
//   warp 1     MMA issuer  : S_t = Q_t K^T (SS), O_t += P_t V (P from TMEM, V MN-major from smem)
//   warps 2-5  softmax for query tile 0      warps 6-9  softmax for query tile 1
// TMEM (512 columns): S0 | S1 | O0 | O1, P_t (bf16) aliases the first 64 columns of S_t.
// The two query tiles ping-pong on the tensor pipe: while one tile's softmax runs on the SIMT
// pipes the other tile's PV / next QK^T MMAs run.  O is rescaled lazily (only when a row max grows
// by more than 2^8), done by the softmax warps themselves between PV(j-1) and PV(j).
#include <math.h>
#include <stdlib.h>

#include "common.cuh"

namespace sfb {

struct AttnParams {
  int Lq, Skv, H, B;
  int n_kv_tiles;
  int kv_tail;            // valid columns in the last KV tile (1..128)
  float scale_log2;       // softmax_scale * log2(e)
  __nv_bfloat16* out;
  long long out_row_stride, out_batch_stride;   // elements
};

constexpr int ATT_BM = 128, ATT_BN = 128, ATT_D = 128;
constexpr int ATT_THREADS = 320;
constexpr int ATT_TILE_BYTES = 128 * 128 * 2;    // 32 KB: two 16 KB halves (d 0-63 | d 64-127)
constexpr int ATT_HALF_BYTES = 128 * 64 * 2;
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_DEFAULT_EMU = 1;
constexpr int ATT_SMEM_BYTES = 2 * ATT_TILE_BYTES + 2 * ATT_KV_STAGES * ATT_TILE_BYTES + 1024 + 256;

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x on the FMA pipe (Cody-Waite range reduction + degree-3 minimax polynomial, rel. error 8.8e-5 --
// far below the bf16 rounding of P).  MUFU.EX2 runs at 16 lanes/clk/SM, which for head_dim 128 is
// exactly the rate the tensor pipe consumes P at; moving a fraction of the exponentials here takes
// the softmax off the critical path.  x <= ~8 (lazy rescaling bound); clamped below for the exponent.
__device__ __forceinline__ float poly_exp2(float x) {
  x = fmaxf(x, -125.0f);
  float r;
  asm("add.rm.ftz.f32 %0, %1, %2;" : "=f"(r) : "f"(x), "f"(12582912.0f));   // 1.5 * 2^23 + floor(x)
  const float f = x - (r - 12582912.0f);                                       // frac(x) in [0, 1)
  const float p = fmaf(fmaf(fmaf(0.077119089663028717f, f, 0.227564394474029541f), f, 0.695146143436431885f), f, 1.0f);
  return __int_as_float(__float_as_int(p) + (__float_as_int(r) << 23));
}

// EMU: how many of every 4 exponentials run on the FMA pipe instead of MUFU (0, 1 or 2).
template <int EMU>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                     const __grid_constant__ CUtensorMap tma_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;                                    // [2 tiles][32 KB]
  uint8_t* k_smem = smem + 2 * ATT_TILE_BYTES;               // [stages][32 KB]
  uint8_t* v_smem = k_smem + ATT_KV_STAGES * ATT_TILE_BYTES; // [stages][32 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + ATT_KV_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;            // [1]
  uint64_t* k_full = bars + 1;        // [2]
  uint64_t* k_empty = bars + 3;       // [2]
  uint64_t* v_full = bars + 5;        // [2]
  uint64_t* v_empty = bars + 7;       // [2]
  uint64_t* s_full = bars + 9;        // [2] per query tile
  uint64_t* p_full = bars + 11;       // [2]
  uint64_t* o_final = bars + 13;      // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int q_row0 = blockIdx.x * (2 * ATT_BM);
  const int head = blockIdx.y;
  const int batch = blockIdx.z;
  const int n_tiles = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_full[s], 128);
      mbar_init(&o_final[s], 1);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ------------------------------ TMA producer ------------------------------
    if (lane == 0) {
      mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
      for (int t = 0; t < 2; ++t)
        for (int hf = 0; hf < 2; ++hf)
          tma_load_4d(q_smem + t * ATT_TILE_BYTES + hf * ATT_HALF_BYTES, &tma_q, q_full, hf * 64, head,
                      q_row0 + t * ATT_BM, batch);
      for (int j = 0; j < n_tiles; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_expect_tx(&k_full[st], ATT_TILE_BYTES);
        for (int hf = 0; hf < 2; ++hf)
          tma_load_4d(k_smem + st * ATT_TILE_BYTES + hf * ATT_HALF_BYTES, &tma_k, &k_full[st], hf * 64, head,
                      j * ATT_BN, batch);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_expect_tx(&v_full[st], ATT_TILE_BYTES);
        for (int hf = 0; hf < 2; ++hf)
          tma_load_4d(v_smem + st * ATT_TILE_BYTES + hf * ATT_HALF_BYTES, &tma_v, &v_full[st], hf * 64, head,
                      j * ATT_BN, batch);
      }
    }
  } else if (warp == 1) {
    // ------------------------------ MMA issuer --------------------------------
    if (lane == 0) {
      constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, 0, 0);   // A, B K-major
      constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D, 0, 1);    // B (=V) MN-major
      const uint32_t q_addr = smem_u32(q_smem), k_addr = smem_u32(k_smem), v_addr = smem_u32(v_smem);

      auto issue_qk = [&](int t, int kst) {
        const uint32_t qa = q_addr + t * ATT_TILE_BYTES, ka = k_addr + kst * ATT_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < ATT_D / 16; ++k) {
          const uint32_t off = (k >> 2) * ATT_HALF_BYTES + (k & 3) * 32;
          umma_ss(tmem_base + t * 128, umma_desc_sw128(qa + off, 16, 1024), umma_desc_sw128(ka + off, 16, 1024),
                  idesc_qk, k != 0);
        }
      };
      auto issue_pv = [&](int t, int vst, bool acc) {
        const uint32_t va = v_addr + vst * ATT_TILE_BYTES;
#pragma unroll
        for (int k = 0; k < ATT_BN / 16; ++k) {
          // A: P_t rows on TMEM lanes, 16 bf16 of K per 8 32-bit columns.
          // B: V tile [kv][d], d contiguous -> MN-major; 16 kv rows = 2048 B; d halves 16 KB apart.
          umma_ts(tmem_base + 256 + t * 128, tmem_base + t * 128 + k * 8,
                  umma_desc_sw128(va + k * 2048, ATT_HALF_BYTES, 1024), idesc_pv, (acc || k != 0) ? 1u : 0u);
        }
      };

      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_qk(0, 0);
      umma_commit(&s_full[0]);
      issue_qk(1, 0);
      umma_commit(&s_full[1]);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < n_tiles; ++j) {
        const int vst = j & 1;
        const uint32_t vph = (j >> 1) & 1;
        const bool has_next = (j + 1) < n_tiles;
        const int kst = (j + 1) & 1;
        const uint32_t kph = ((j + 1) >> 1) & 1;
        for (int t = 0; t < 2; ++t) {
          mbar_wait(&p_full[t], j & 1);
          if (t == 0) mbar_wait(&v_full[vst], vph);
          tc_fence_after();
          issue_pv(t, vst, j > 0);
          if (t == 1) umma_commit(&v_empty[vst]);
          if (has_next) {
            if (t == 0) {
              mbar_wait(&k_full[kst], kph);
              tc_fence_after();
            }
            issue_qk(t, kst);
            umma_commit(&s_full[t]);
            if (t == 1) umma_commit(&k_empty[kst]);
          } else {
            umma_commit(&o_final[t]);
          }
        }
      }
    }
  } else {
    // ------------------------------ softmax / correction / epilogue ------------
    const int t = (warp - 2) >> 2;
    const int quarter = warp & 3;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base + t * 128;
    const uint32_t o_addr = tmem_base + lane_base + 256 + t * 128;
    const float sl2 = p.scale_log2;
    float m_ref = -INFINITY;   // reference max (raw score units) the stored exponentials are relative to
    float l = 0.f;             // running sum of exponentials relative to m_ref

    for (int j = 0; j < n_tiles; ++j) {
      mbar_wait(&s_full[t], j & 1);
      tc_fence_after();
      uint32_t s[128];
#pragma unroll
      for (int c = 0; c < 4; ++c) tmem_ld32(s_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
      tmem_ld_wait();
      if (j == n_tiles - 1 && p.kv_tail < ATT_BN) {
#pragma unroll
        for (int i = 0; i < 128; ++i)
          if (i >= p.kv_tail) s[i] = 0xff800000u;   // -inf
      }
      float mx0 = __uint_as_float(s[0]), mx1 = __uint_as_float(s[1]), mx2 = __uint_as_float(s[2]),
            mx3 = __uint_as_float(s[3]);
#pragma unroll
      for (int i = 4; i < 128; i += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(s[i]));
        mx1 = fmaxf(mx1, __uint_as_float(s[i + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(s[i + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(s[i + 3]));
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      if (j == 0) {
        m_ref = mx;
      } else {
        const float m_new = fmaxf(m_ref, mx);
        const bool need = (m_new - m_ref) * sl2 > 8.0f;
        if (__any_sync(0xffffffffu, need)) {
          // PV(j-1) has retired (s_full(j) was committed after it), so O_t is quiescent here.
          const float m_upd = need ? m_new : m_ref;
          const float alpha = fast_exp2((m_ref - m_upd) * sl2);
          l *= alpha;
          m_ref = m_upd;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld32(o_addr + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * alpha);
            tmem_st32(o_addr + c * 32, o);
          }
          tmem_st_wait();
        }
      }
      const float neg_m = -m_ref * sl2;
      float sum0 = 0.f, sum1 = 0.f;
#pragma unroll
      for (int c = 0; c < 4; ++c) {
        uint32_t pk[16];
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          const float x0 = fmaf(__uint_as_float(s[c * 32 + 2 * i]), sl2, neg_m);
          const float x1 = fmaf(__uint_as_float(s[c * 32 + 2 * i + 1]), sl2, neg_m);
          // element index within its group of four: 2*(i&1) and 2*(i&1)+1
          const float p0 = (EMU >= 2 && (i & 1) == 0) ? poly_exp2(x0) : fast_exp2(x0);
          const float p1 = (EMU >= 1 && (i & 1) == 1) ? poly_exp2(x1) : fast_exp2(x1);
          sum0 += p0;
          sum1 += p1;
          pk[i] = pack_bf16(p0, p1);
        }
        tmem_st16(s_addr + c * 16, pk);   // P_t aliases the first 64 columns of S_t
      }
      l += sum0 + sum1;
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[t]);
    }

    // epilogue: O_t / l -> bf16 -> global, one query row per thread (256 contiguous bytes)
    mbar_wait(&o_final[t], 0);
    tc_fence_after();
    const int row = q_row0 + t * ATT_BM + quarter * 32 + lane;
    const float inv_l = 1.0f / l;
    __nv_bfloat16* orow = p.out + (long long)batch * p.out_batch_stride + (long long)row * p.out_row_stride +
                          head * ATT_D;
#pragma unroll 1
    for (int c = 0; c < 4; ++c) {
      uint32_t o[32];
      tmem_ld32(o_addr + c * 32, o);
      tmem_ld_wait();
      if (row < p.Lq) {
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          uint4 w;
          w.x = pack_bf16(__uint_as_float(o[g * 8 + 0]) * inv_l, __uint_as_float(o[g * 8 + 1]) * inv_l);
          w.y = pack_bf16(__uint_as_float(o[g * 8 + 2]) * inv_l, __uint_as_float(o[g * 8 + 3]) * inv_l);
          w.z = pack_bf16(__uint_as_float(o[g * 8 + 4]) * inv_l, __uint_as_float(o[g * 8 + 5]) * inv_l);
          w.w = pack_bf16(__uint_as_float(o[g * 8 + 6]) * inv_l, __uint_as_float(o[g * 8 + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + c * 32 + g * 8) = w;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

}  // namespace sfb

// q   : [B, Lq, H, 128] view with element strides (q_row_stride between tokens, q_batch_stride)
// k, v: cache window start (already offset to attn_start), [B, Skv, H, 128] with their strides
// out : [B, Lq, H, 128] with out_row_stride / out_batch_stride
extern "C" int sfb_attention_fwd(const void* q, long long q_row_stride, long long q_batch_stride, const void* k,
                                 const void* v, long long kv_row_stride, long long kv_batch_stride, void* out,
                                 long long out_row_stride, long long out_batch_stride, int B, int Lq, int Skv,
                                 int H, int head_dim, float softmax_scale, void* stream_) {
  using namespace sfb;
  cudaStream_t stream = reinterpret_cast<cudaStream_t>(stream_);
  if (head_dim != ATT_D) { set_error("sfb_attention_fwd: head_dim %d unsupported (128 only)", head_dim); return SFB_ERR_INVALID; }
  if (B <= 0 || Lq <= 0 || Skv <= 0 || H <= 0) { set_error("sfb_attention_fwd: empty problem B=%d Lq=%d Skv=%d H=%d", B, Lq, Skv, H); return SFB_ERR_INVALID; }
  if ((q_row_stride % 8) || (kv_row_stride % 8) || (out_row_stride % 8) || (q_batch_stride % 8) ||
      (kv_batch_stride % 8) || (out_batch_stride % 8)) {
    set_error("sfb_attention_fwd: strides must be multiples of 8 elements");
    return SFB_ERR_INVALID;
  }
  CUtensorMap tq, tk, tv;
  const uint32_t box[4] = {64, 1, 128, 1};
  {
    uint64_t dims[4] = {128, (uint64_t)H, (uint64_t)Lq, (uint64_t)B};
    uint64_t str[3] = {128 * 2, (uint64_t)q_row_stride * 2, (uint64_t)(B > 1 ? q_batch_stride : q_row_stride * Lq) * 2};
    if (int e = make_tmap_bf16(&tq, q, 4, dims, str, box, true)) return e;
  }
  {
    uint64_t dims[4] = {128, (uint64_t)H, (uint64_t)Skv, (uint64_t)B};
    uint64_t str[3] = {128 * 2, (uint64_t)kv_row_stride * 2, (uint64_t)(B > 1 ? kv_batch_stride : kv_row_stride * Skv) * 2};
    if (int e = make_tmap_bf16(&tk, k, 4, dims, str, box, true)) return e;
    if (int e = make_tmap_bf16(&tv, v, 4, dims, str, box, true)) return e;
  }
  AttnParams p{};
  p.Lq = Lq; p.Skv = Skv; p.H = H; p.B = B;
  p.n_kv_tiles = (Skv + ATT_BN - 1) / ATT_BN;
  p.kv_tail = Skv - (p.n_kv_tiles - 1) * ATT_BN;
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  p.out = static_cast<__nv_bfloat16*>(out);
  p.out_row_stride = out_row_stride;
  p.out_batch_stride = out_batch_stride;

  static int emu = -1;
  if (emu < 0) {
    const char* env = getenv("SFB_ATTN_EMU");   // tuning knob; the default is the measured best
    emu = env ? atoi(env) : ATT_DEFAULT_EMU;
    if (emu < 0 || emu > 2) emu = ATT_DEFAULT_EMU;
    for (auto kern : {attention_fwd_kernel<0>, attention_fwd_kernel<1>, attention_fwd_kernel<2>})
      if (int e = check_cuda(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, ATT_SMEM_BYTES),
                             "cudaFuncSetAttribute(attention)")) {
        emu = -1;
        return e;
      }
  }
  dim3 grid((Lq + 2 * ATT_BM - 1) / (2 * ATT_BM), H, B);
  auto kern = emu == 0 ? attention_fwd_kernel<0> : (emu == 1 ? attention_fwd_kernel<1> : attention_fwd_kernel<2>);
  kern<<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tq, tk, tv, p);
  return check_cuda(cudaGetLastError(), "attention launch");
}
