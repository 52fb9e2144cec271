// Dense (unmasked) flash attention forward for sm_100a, head_dim 128, bf16 in / fp32 accumulate.
//
// Replaces `attention()` -> flash_attn_varlen_func at wan/modules/attention.py:136-150 for the two
// call sites of the rollout:
//   * chunk self-attention over the rolling KV cache window  (wan/modules/causal_model.py:230-234)
//   * cross-attention against the cached T5 context          (wan/modules/model.py:189)
// K and V are read in place from the cache layout [B, S, H, 128] (row stride H*128) through TMA, so
// there is no varlen packing / cu_seqlens traffic and no copy of the window.
//
// Work decomposition (persistent, "stream-K" over the KV axis).  A work ITEM is (batch, head, pair of
// 128-row query tiles); it needs n_kv = ceil(S / 128) KV steps.  The B*H*ceil(Lq/256) items rarely divide
// by the 148 SMs (Lq = 4680, H = 12: 228 items = 1.54 waves, a quarter of the GPU idle), so the item x
// step space is linearised and cut into one CONTIGUOUS range of steps per CTA.  A CTA therefore runs a
// few SEGMENTS (item, [j0, j1)); a segment that covers its whole item writes bf16 output directly, a
// partial one parks (unnormalised O^T, row max, row sum) in the caller's workspace and
// `attention_combine_kernel` merges the (at most a few) partials of each split item.  Short KV (cross
// attention, 4 steps) is not split: CTAs take whole items.
//
// One CTA per SM, 384 threads = 3 warpgroups:
//   WG0  warp 0: TMA producer (Q pair per segment; K / V tiles through two 2-deep rings)
//        warp 1: MMA issuer   S_t = Q_t K^T (SS), O_t += P_t V (P from TMEM, V MN-major from smem)
//   WG1 / WG2    softmax of query tile 0 / 1, one row per thread (setmaxnreg moves registers to them)
// TMEM (512 columns): S0 | S1 | O0 | O1, P_t (bf16) aliases the first 64 columns of S_t.
// The two query tiles ping-pong on the tensor pipe: while one tile's softmax runs on the SIMT / XU
// pipes the other tile's PV / next QK^T MMAs run.  O is rescaled lazily (only when a row max grows
// by more than 2^8), by the softmax warps themselves between PV(j-1) and PV(j).
//
// Softmax arithmetic: x = s * scale_log2 - m with packed fp32x2 FMAs, 2^x on MUFU.EX2 (16 lanes/clk/SM --
// for head_dim 128 that is exactly the rate at which the tensor pipe consumes P, so the exponentials,
// not the MMAs, are the critical resource), row sums with packed adds.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "common.cuh"
#include "attention_sched.cuh"

namespace sfb {

__device__ __forceinline__ float fast_exp2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for two values on the FMA / ALU pipes only (no MUFU, no conversion-pipe instruction): round x to the nearest
// integer with the magic-number add (t = x + 1.5 * 2^23 holds round(x) in its low mantissa bits), evaluate a degree-3
// minimax polynomial of 2^f on f = x - round(x) in [-0.5, 0.5] (max relative error 7.5e-5, 26x below the bf16 rounding
// of P) and add round(x) to the exponent field.  Packed f32x2 arithmetic: 6 FMA-pipe + 4 ALU instructions per pair.
// (Round 1's polynomial used floorf + float-to-int conversion, which issue on the same quarter-rate pipe as MUFU.)
__device__ __forceinline__ float2 exp2_poly2(float2 x) {
  x.x = fmaxf(x.x, -125.0f);
  x.y = fmaxf(x.y, -125.0f);
  const float2 t = __fadd2_rn(x, make_float2(12582912.0f, 12582912.0f));
  const float2 xi = __fadd2_rn(t, make_float2(-12582912.0f, -12582912.0f));
  const float2 f = __ffma2_rn(xi, make_float2(-1.0f, -1.0f), x);
  float2 p = __ffma2_rn(f, make_float2(0.05517132207751274f, 0.05517132207751274f), make_float2(0.24261054396629333f, 0.24261054396629333f));
  p = __ffma2_rn(p, f, make_float2(0.6932609677314758f, 0.6932609677314758f));
  p = __ffma2_rn(p, f, make_float2(0.9999281167984009f, 0.9999281167984009f));
  float2 e;
  e.x = __int_as_float(__float_as_int(p.x) + (__float_as_int(t.x) << 23));
  e.y = __int_as_float(__float_as_int(p.y) + (__float_as_int(t.y) << 23));
  return e;
}
// Which of the 16 column pairs of a 32-column chunk take the polynomial (bit i = pair i); the others use MUFU.EX2.
// The exponentials (16 / clk / SM on the XU pipe) need exactly as many cycles per KV step as the MMAs; moving a share of
// them to the FMA pipe takes the softmax off the critical path.
// 64 bits: 16 per 32-column chunk of the KV tile (bit 16 c + i = pair i of chunk c).  Measured on B200, same box
// (tools/build_attn_variants.sh, profiles/r02w_attn_poly_share.json), Lq 4680 x S 18720 / 32760: no polynomial 443 / 752 us;
// every 8th pair 445 / 750; **every 4th pair (25 %) 433 / 730 us (+2.4 .. 3 %)**; 31 % concentrated in the later chunks
// 447 / 761; 37.5 % 468 / 789; 44 % 464 / 781; 50 % 480 / 814 -- beyond a quarter the extra FMA-pipe work costs more than
// the MUFU cycles it frees (and shares that touch the first chunk spill: the 128 score registers are all live there).
#ifndef ATT_POLY_MASK
#define ATT_POLY_MASK 0x8888888888888888ull
#endif

template <int REGS>
__device__ __forceinline__ void setmaxnreg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(REGS)); }
template <int REGS>
__device__ __forceinline__ void setmaxnreg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(REGS)); }

// ---- MMA issue helpers (one elected lane calls them) --------------------------------------------------------------
struct AttMmaCtx {
  uint32_t tmem_base, q_addr, k_addr, v_addr;   // TMEM base column, shared-memory addresses of the Q / K / V tiles
};
__device__ __forceinline__ void att_issue_qk(const AttMmaCtx& c, int t, int kst) {
  constexpr uint32_t idesc_qk = umma_idesc_bf16(ATT_BM, ATT_BN, 0, 0);   // A, B K-major
  const uint32_t qa = c.q_addr + t * ATT_TILE_BYTES, ka = c.k_addr + kst * ATT_TILE_BYTES;
#pragma unroll
  for (int k = 0; k < ATT_D / 16; ++k) {
    const uint32_t off = (k >> 2) * ATT_HALF_BYTES + (k & 3) * 32;
    umma_ss(c.tmem_base + t * 128, umma_desc_sw128(qa + off, 16, 1024), umma_desc_sw128(ka + off, 16, 1024), idesc_qk, k != 0);
  }
}
__device__ __forceinline__ void att_issue_pv(const AttMmaCtx& c, int t, int vst, bool acc, int half) {
  constexpr uint32_t idesc_pv = umma_idesc_bf16(ATT_BM, ATT_D, 0, 1);    // B (=V) MN-major
  const uint32_t va = c.v_addr + vst * ATT_TILE_BYTES;
#pragma unroll
  for (int k = half * 4; k < half * 4 + 4; ++k) {
    // A: P_t rows on TMEM lanes, 16 bf16 of K per 8 32-bit columns.
    // B: V tile [kv][d], d contiguous -> MN-major; 16 kv rows = 2048 B; d halves 16 KB apart.
    umma_ts(c.tmem_base + 256 + t * 128, c.tmem_base + t * 128 + k * 8, umma_desc_sw128(va + k * 2048, ATT_HALF_BYTES, 1024),
            idesc_pv, (acc || k != 0) ? 1u : 0u);
  }
}

// MMA schedule of a HALF item (called by the whole MMA warp): ONE query tile (in both Q slots); slot t takes the KV tiles
// kv_lo + t, kv_lo + t + 2, ... each its own ring entry; O_0 and O_1 are partial sums over disjoint KV tiles, merged by
// the softmax warps.  Kept out of line: one item in ~37 takes this path, and inlined it cost the main loop registers
// (the MMA warp runs under a 104-register cap).  g = ring entries consumed so far, sc[t] = score tiles committed per slot.
__device__ __forceinline__ void att_mma_half_item(AttMmaCtx c, uint64_t* bars, int kv_lo, int cnt, uint32_t seg, uint32_t g,
                                               uint32_t* sc) {
  uint64_t* q_empty = bars + 1;
  uint64_t* k_full = bars + 2;
  uint64_t* k_empty = bars + 4;
  uint64_t* v_full = bars + 6;
  uint64_t* v_empty = bars + 8;
  uint64_t* s_full = bars + 10;
  uint64_t* p_half = bars + 18;
  uint64_t* o_final = bars + 14;
  uint64_t* o_free = bars + 16;
  const int n_t[2] = {(cnt + 1) >> 1, cnt >> 1};
  for (int t = 0; t < 2; ++t) {
    if (n_t[t] == 0) continue;
    const uint32_t e = g + t;
    mbar_wait(&k_full[e & 1], (e >> 1) & 1);
    tc_fence_after();
    if (elect_one()) {
      att_issue_qk(c, t, e & 1);
      umma_commit(&s_full[t]);
      umma_commit(&k_empty[e & 1]);
      if ((int)t == cnt - 1) umma_commit(q_empty);
    }
    __syncwarp();
  }
  for (int i = 0; i < n_t[0]; ++i) {
    for (int t = 0; t < 2; ++t) {
      if (i >= n_t[t]) continue;
      const uint32_t e = g + 2 * i + t, e2 = e + 2;
      const int vst = e & 1;
      const bool has_next = (i + 1) < n_t[t];
      mbar_wait(&p_half[2 * t], sc[t] & 1);
      mbar_wait(&v_full[vst], (e >> 1) & 1);
      if (i == 0) mbar_wait(&o_free[t], (seg & 1) ^ 1);
      tc_fence_after();
      if (elect_one()) att_issue_pv(c, t, vst, i > 0, 0);
      __syncwarp();
      mbar_wait(&p_half[2 * t + 1], sc[t] & 1);
      tc_fence_after();
      if (has_next) {
        mbar_wait(&k_full[e2 & 1], (e2 >> 1) & 1);
        tc_fence_after();
      }
      if (elect_one()) {
        att_issue_pv(c, t, vst, true, 1);
        umma_commit(&v_empty[vst]);
        if (has_next) {
          att_issue_qk(c, t, e2 & 1);
          umma_commit(&s_full[t]);
          umma_commit(&k_empty[e2 & 1]);
          if (2 * (i + 1) + t == cnt - 1) umma_commit(q_empty);
        } else {
          umma_commit(&o_final[t]);
        }
      }
      __syncwarp();
      ++sc[t];
    }
  }
  if (n_t[1] == 0) {           // slot 1 had no KV tile in this segment: keep its per-segment barrier in step
    if (elect_one()) umma_commit(&o_final[1]);
    __syncwarp();
  }
}

// TIMING = true: diagnostic build with per-phase clock64 timers in one softmax warp (SFB_ATTN_TIMING=1); the timers cost
// ~15 registers in the softmax warps, so the shipping instantiation compiles them out.
template <bool TIMING>
__global__ void __launch_bounds__(ATT_THREADS, 1)
attention_fwd_kernel(const __grid_constant__ CUtensorMap tma_q, const __grid_constant__ CUtensorMap tma_k,
                     const __grid_constant__ CUtensorMap tma_v, const AttnParams p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* q_smem = smem;                                    // [2 tiles][32 KB]
  uint8_t* k_smem = smem + 2 * ATT_TILE_BYTES;               // [stages][32 KB]
  uint8_t* v_smem = k_smem + ATT_KV_STAGES * ATT_TILE_BYTES; // [stages][32 KB]
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + ATT_KV_STAGES * ATT_TILE_BYTES);
  uint64_t* q_full = bars;            // [1]  per segment
  uint64_t* q_empty = bars + 1;       // [1]  per segment
  uint64_t* k_full = bars + 2;        // [2]  ring
  uint64_t* k_empty = bars + 4;       // [2]
  uint64_t* v_full = bars + 6;        // [2]
  uint64_t* v_empty = bars + 8;       // [2]
  uint64_t* s_full = bars + 10;       // [2] per query tile, per KV step
  uint64_t* p_half = bars + 18;       // [2 tiles][2 halves of the KV step], one arrival per softmax warp
  uint64_t* o_final = bars + 14;      // [2] per query tile, per segment
  uint64_t* o_free = bars + 16;       // [2] per query tile, per segment (128 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 22);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_kv = p.n_kv_tiles;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tma_q);
    tma_prefetch_desc(&tma_k);
    tma_prefetch_desc(&tma_v);
    mbar_init(q_full, 1);
    mbar_init(q_empty, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1);
      mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1);
      mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1);
      mbar_init(&p_half[2 * s], 4);
      mbar_init(&p_half[2 * s + 1], 4);
      mbar_init(&o_final[s], 1);
      mbar_init(&o_free[s], 128);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, 512);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // per head group, this CTA owns one contiguous range of (item, KV step) work: see att_range_start

  if (warp < 4) {
    setmaxnreg_dec<104>();   // frees (168 - 104) x 128 = 8192 registers of the CTA pool = exactly the (200 - 168) x 256 the softmax warps take (112 here deadlocks the inc)
    if (warp == 0) {
      // ------------------------------ TMA producer (warp-uniform, one lane issues) ----
      {
        uint32_t seg = 0, g = 0;   // segment counter, ring counter (KV tiles issued so far)
        for (int grp = 0; grp < p.n_groups; ++grp) {
        const int range_end = att_range_start(blockIdx.x + 1, gridDim.x, grp, p);
        for (int cur = att_range_start(blockIdx.x, gridDim.x, grp, p); cur < range_end; ++seg) {
          const AttSeg sg = att_decode(cur, range_end, p);
          const int bh = grp * p.heads_per_group + sg.bh_local;
          const int head = bh % p.H, batch = bh / p.H;
          const int q_row0 = sg.q_tile * ATT_BM;
          mbar_wait(q_empty, (seg & 1) ^ 1);           // previous segment's QK^T MMAs have retired
          if (elect_one()) {
            mbar_expect_tx(q_full, 2 * ATT_TILE_BYTES);
#pragma unroll
            for (int t = 0; t < 2; ++t)
#pragma unroll
              for (int hf = 0; hf < 2; ++hf)   // half item: the one query tile in both slots
                tma_load_4d(q_smem + t * ATT_TILE_BYTES + hf * ATT_HALF_BYTES, &tma_q, q_full, hf * 64, head,
                            q_row0 + (sg.half ? 0 : t * ATT_BM), batch);
          }
          __syncwarp();
          // KV tiles of the segment, one ring entry each (a half item's step is two tiles)
          const int kv_lo = sg.half ? 2 * sg.j0 : sg.j0;
          const int kv_hi = sg.half ? (2 * sg.j1 < n_kv ? 2 * sg.j1 : n_kv) : sg.j1;
          auto load_k = [&](int j, uint32_t e) {
            mbar_wait(&k_empty[e & 1], ((e >> 1) & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&k_full[e & 1], ATT_TILE_BYTES);
#pragma unroll
              for (int hf = 0; hf < 2; ++hf)
                tma_load_4d(k_smem + (e & 1) * ATT_TILE_BYTES + hf * ATT_HALF_BYTES, &tma_k, &k_full[e & 1], hf * 64, head,
                            j * ATT_BN, batch);
            }
            __syncwarp();
          };
          auto load_v = [&](int j, uint32_t e) {
            mbar_wait(&v_empty[e & 1], ((e >> 1) & 1) ^ 1);
            if (elect_one()) {
              mbar_expect_tx(&v_full[e & 1], ATT_TILE_BYTES);
#pragma unroll
              for (int hf = 0; hf < 2; ++hf)
                tma_load_4d(v_smem + (e & 1) * ATT_TILE_BYTES + hf * ATT_HALF_BYTES, &tma_v, &v_full[e & 1], hf * 64, head,
                            j * ATT_BN, batch);
            }
            __syncwarp();
          };
          if (!sg.half) {
            for (int j = kv_lo; j < kv_hi; ++j, ++g) { load_k(j, g); load_v(j, g); }
          } else {
            // two ring entries per step: both K tiles first (they only wait for the QK^T two entries back), then the V
            // tiles (which wait for the PV two entries back -- a later event)
            for (int j = kv_lo; j < kv_hi; j += 2) {
              const bool two = j + 1 < kv_hi;
              load_k(j, g);
              if (two) load_k(j + 1, g + 1);
              load_v(j, g);
              if (two) load_v(j + 1, g + 1);
              g += two ? 2 : 1;
            }
          }
          cur += sg.advance;
        }
        }
      }
    } else if (warp == 1) {
      // ------------------------------ MMA issuer (warp-uniform, one lane issues) -----
      {
        const AttMmaCtx mc{tmem_base, smem_u32(q_smem), smem_u32(k_smem), smem_u32(v_smem)};

        uint32_t seg = 0, g = 0;     // segment counter, ring counter (KV tiles consumed so far)
        uint32_t sc[2] = {0, 0};     // score tiles committed so far per query-tile slot (parity of s_full / p_half)
        for (int grp = 0; grp < p.n_groups; ++grp) {
        const int range_end = att_range_start(blockIdx.x + 1, gridDim.x, grp, p);
        for (int cur = att_range_start(blockIdx.x, gridDim.x, grp, p); cur < range_end; ++seg) {
          const AttSeg sg = att_decode(cur, range_end, p);
          const int n = sg.j1 - sg.j0;
          mbar_wait(q_full, seg & 1);
          if (!sg.half) {
            // ---- full item: both query tiles walk the same KV tiles; one ring entry per step ----
            mbar_wait(&k_full[g & 1], (g >> 1) & 1);
            tc_fence_after();
            if (elect_one()) {
              att_issue_qk(mc, 0, g & 1);
              umma_commit(&s_full[0]);
              att_issue_qk(mc, 1, g & 1);
              umma_commit(&s_full[1]);
              umma_commit(&k_empty[g & 1]);
              if (n == 1) umma_commit(q_empty);
            }
            __syncwarp();
            for (int i = 0; i < n; ++i, ++g) {
              const int vst = g & 1;
              const uint32_t vph = (g >> 1) & 1;
              const bool has_next = (i + 1) < n;
              const int kst = (g + 1) & 1;
              const uint32_t kph = ((g + 1) >> 1) & 1;
              for (int t = 0; t < 2; ++t) {
                // P arrives in two halves (64 KV columns each): the first half's PV overlaps the softmax of the second
                mbar_wait(&p_half[2 * t], sc[t] & 1);
                if (t == 0) mbar_wait(&v_full[vst], vph);
                if (i == 0) mbar_wait(&o_free[t], (seg & 1) ^ 1);   // previous segment's epilogue has read O_t
                tc_fence_after();
                if (elect_one()) att_issue_pv(mc, t, vst, i > 0, 0);
                __syncwarp();
                mbar_wait(&p_half[2 * t + 1], sc[t] & 1);
                tc_fence_after();
                if (has_next && t == 0) {
                  mbar_wait(&k_full[kst], kph);
                  tc_fence_after();
                }
                if (elect_one()) {
                  att_issue_pv(mc, t, vst, true, 1);
                  if (t == 1) umma_commit(&v_empty[vst]);
                  if (has_next) {
                    att_issue_qk(mc, t, kst);
                    umma_commit(&s_full[t]);
                    if (t == 1) {
                      umma_commit(&k_empty[kst]);
                      if (i + 2 == n) umma_commit(q_empty);   // that was the segment's last QK^T: Q smem reusable
                    }
                  } else {
                    umma_commit(&o_final[t]);
                  }
                }
                __syncwarp();
                ++sc[t];
              }
            }
          } else {
            // ---- half item (out of line: att_mma_half_item) ----
            const int kv_lo = 2 * sg.j0, kv_hi = 2 * sg.j1 < n_kv ? 2 * sg.j1 : n_kv;
            const int cnt = kv_hi - kv_lo;
            att_mma_half_item(mc, bars, kv_lo, cnt, seg, g, sc);
            g += cnt;
          }
          cur += sg.advance;
        }
        }
      }
    }
  } else {
    // ------------------------------ softmax / correction / epilogue ------------
    setmaxnreg_inc<200>();
    const int t = (warp - 4) >> 2;
    const int quarter = warp & 3;
    const int r_local = quarter * 32 + lane;
    const uint32_t lane_base = (uint32_t)(quarter * 32) << 16;
    const uint32_t s_addr = tmem_base + lane_base + t * 128;
    const uint32_t o_addr = tmem_base + lane_base + 256 + t * 128;

    const bool timing = TIMING && p.dbg != nullptr && warp == 4 && lane == 0;
    long long tm[6] = {0, 0, 0, 0, 0, 0};
    long long t_prev = timing ? clock64() : 0;
    auto stamp = [&](int k) {
      if (TIMING && timing) {
        const long long now = clock64();
        tm[k] += now - t_prev;
        t_prev = now;
      }
    };
    uint32_t seg = 0, sc = 0;   // segment counter, score tiles of THIS query-tile slot consumed so far
    float2* xchg = reinterpret_cast<float2*>(bars + 32);   // [2 segment parities][128 rows] (m, l) of slot 1 -> slot 0's warps (half items)
    for (int grp = 0; grp < p.n_groups; ++grp) {
    const int range_begin = att_range_start(blockIdx.x, gridDim.x, grp, p);
    const int range_end = att_range_start(blockIdx.x + 1, gridDim.x, grp, p);
    for (int cur = range_begin; cur < range_end; ++seg) {
      const AttSeg sg = att_decode(cur, range_end, p);
      // KV tiles of this slot: every tile of the segment (full item) or every other one (half item)
      const int jb = sg.half ? 2 * sg.j0 + t : sg.j0, js = sg.half ? 2 : 1;
      const int je = sg.half ? (2 * sg.j1 < n_kv ? 2 * sg.j1 : n_kv) : sg.j1;
      float m_ref = -INFINITY;   // reference max (raw score units) the stored exponentials are relative to
      float l = 0.f;             // running sum of exponentials relative to m_ref
      float sl2 = p.scale_log2;  // softmax scale of this thread's row (x log2 e)
      if (p.q_stats != nullptr) {
        const int bh_s = grp * p.heads_per_group + sg.bh_local;
        const int row_s = sg.q_tile * ATT_BM + (sg.half ? 0 : t * ATT_BM) + r_local;
        sl2 *= stats_rms_rstd(p.q_stats + ((long long)(bh_s / p.H) * p.Lq + (row_s < p.Lq ? row_s : p.Lq - 1)) * p.q_chunks,
                              p.q_chunks, p.q_eps);
      }

      for (int j = jb; j < je; j += js, ++sc) {
        stamp(5);
        mbar_wait(&s_full[t], sc & 1);
        tc_fence_after();
        stamp(0);
        if (j == n_kv - 1 && p.kv_tail < ATT_BN) {
          // ragged last KV tile: overwrite the out-of-range score columns with -inf in TMEM (rare path,
          // kept out of the register-resident fast path below)
#pragma unroll 1
          for (int c = p.kv_tail & ~15; c < ATT_BN; c += 16) {
            uint32_t w[16];
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]), "=r"(w[4]), "=r"(w[5]), "=r"(w[6]), "=r"(w[7]),
                  "=r"(w[8]), "=r"(w[9]), "=r"(w[10]), "=r"(w[11]), "=r"(w[12]), "=r"(w[13]), "=r"(w[14]), "=r"(w[15])
                : "r"(s_addr + c)
                : "memory");
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 16; ++i)
              if (c + i >= p.kv_tail) w[i] = 0xff800000u;   // -inf
            tmem_st16(s_addr + c, w);
          }
          tmem_st_wait();
        }
        uint32_t s[128];
#pragma unroll
        for (int c = 0; c < 4; ++c) tmem_ld32(s_addr + c * 32, *reinterpret_cast<uint32_t(*)[32]>(&s[c * 32]));
        tmem_ld_wait();
        stamp(1);
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
        if (j == jb) {
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
        stamp(2);
        const float neg_m = -m_ref * sl2;
        const float2 sl2v = make_float2(sl2, sl2), negv = make_float2(neg_m, neg_m);
        float2 sum_a = make_float2(0.f, 0.f), sum_b = make_float2(0.f, 0.f);
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 16; ++i) {
            const float2 x = __ffma2_rn(make_float2(__uint_as_float(s[c * 32 + 2 * i]), __uint_as_float(s[c * 32 + 2 * i + 1])),
                                        sl2v, negv);
            float2 e;
            if ((ATT_POLY_MASK >> (16 * c + i)) & 1ull) {
              e = exp2_poly2(x);
            } else {
              e.x = fast_exp2(x.x);
              e.y = fast_exp2(x.y);
            }
            if (i & 1) sum_b = __fadd2_rn(sum_b, e); else sum_a = __fadd2_rn(sum_a, e);
            pk[i] = pack_bf16(e.x, e.y);
          }
          tmem_st16(s_addr + c * 16, pk);   // P_t aliases the first 64 columns of S_t
          if (c & 1) {                      // publish this half of P (64 KV columns)
            tmem_st_wait();
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&p_half[2 * t + (c >> 1)]);
          }
        }
        l += (sum_a.x + sum_a.y) + (sum_b.x + sum_b.y);
        stamp(3);
        stamp(4);
      }

      // segment epilogue
      mbar_wait(&o_final[t], seg & 1);
      tc_fence_after();
      const int bh = grp * p.heads_per_group + sg.bh_local;
      const int head = bh % p.H, batch = bh / p.H;
      const bool whole = sg.j0 == 0 && sg.j1 == sg.item_steps;
      if (sg.half && t == 1) {
        // half item, slot 1: hand (m, l) to slot 0's warp of the same TMEM lane quarter, which reads O_1 itself
        xchg[(seg & 1) * ATT_BM + r_local] = make_float2(m_ref, l);
        tc_fence_before();
        named_barrier_arrive(1, 2 * ATT_BM);
      } else {
        // rows of this thread: O_t scaled by w_own (+ O_1 scaled by w_oth for a half item), row sum l
        float w_own = 1.f, w_oth = 0.f;
        bool two = false;
        if (sg.half) {
          named_barrier_sync(1, 2 * ATT_BM);
          tc_fence_after();
          const float2 ml = xchg[(seg & 1) * ATT_BM + r_local];
          two = ((2 * sg.j1 < n_kv ? 2 * sg.j1 : n_kv) - 2 * sg.j0) > 1;   // slot 1 had at least one KV tile (segment-uniform)
          if (two) {
            const float m_new = fmaxf(m_ref, ml.x);
            w_own = fast_exp2((m_ref - m_new) * sl2);
            w_oth = fast_exp2((ml.x - m_new) * sl2);
            l = l * w_own + ml.y * w_oth;
            m_ref = m_new;
          }
        }
        const uint32_t o_oth = o_addr + 128;   // O_1 (same TMEM lanes)
        const int row = sg.q_tile * ATT_BM + t * ATT_BM + r_local;
        if (whole) {
          // whole item: O / l -> bf16 -> global, one query row per thread (256 contiguous bytes)
          const float inv_l = 1.0f / l;
          const float a_own = w_own * inv_l, a_oth = w_oth * inv_l;
          const int dst = row / p.rows_per_dst;
          __nv_bfloat16* orow = p.out[dst < 8 ? dst : 0] + (long long)batch * p.out_batch_stride +
                                (long long)(row - dst * p.rows_per_dst) * p.out_row_stride + head * ATT_D;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld32(o_addr + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * a_own);
            if (two) {
              uint32_t o2[32];
              tmem_ld32(o_oth + c * 32, o2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(fmaf(__uint_as_float(o2[i]), a_oth, __uint_as_float(o[i])));
            }
            if (row < p.Lq) {
#pragma unroll
              for (int gq = 0; gq < 4; ++gq) {
                uint4 w;
                w.x = pack_bf16(__uint_as_float(o[gq * 8 + 0]), __uint_as_float(o[gq * 8 + 1]));
                w.y = pack_bf16(__uint_as_float(o[gq * 8 + 2]), __uint_as_float(o[gq * 8 + 3]));
                w.z = pack_bf16(__uint_as_float(o[gq * 8 + 4]), __uint_as_float(o[gq * 8 + 5]));
                w.w = pack_bf16(__uint_as_float(o[gq * 8 + 6]), __uint_as_float(o[gq * 8 + 7]));
                *reinterpret_cast<uint4*>(orow + c * 32 + gq * 8) = w;
              }
            }
          }
        } else {
          // partial item: park ((O / l)^T as fp16, m, l) in this CTA's workspace slot (first segment -> 0, last -> 1).  The
          // normalised partial is a convex combination of V rows (no overflow) and fp16's 2^-11 is 8x below the bf16
          // rounding of the final output; it halves the bytes every split item sends through memory twice.
          float* slot = p.ws + ((((long long)grp * gridDim.x + blockIdx.x) * 2 + (cur == range_begin ? 0 : 1)) * 2 + t) *
                                   ATT_SLOT_FLOATS;
          __half2* slot_h = reinterpret_cast<__half2*>(slot);
          const float inv_l = 1.0f / l;
          const float a_own = w_own * inv_l, a_oth = w_oth * inv_l;
#pragma unroll 1
          for (int c = 0; c < 4; ++c) {
            uint32_t o[32];
            tmem_ld32(o_addr + c * 32, o);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(__uint_as_float(o[i]) * a_own);
            if (two) {
              uint32_t o2[32];
              tmem_ld32(o_oth + c * 32, o2);
              tmem_ld_wait();
#pragma unroll
              for (int i = 0; i < 32; ++i) o[i] = __float_as_uint(fmaf(__uint_as_float(o2[i]), a_oth, __uint_as_float(o[i])));
            }
#pragma unroll
            for (int i = 0; i < 16; ++i)   // column pair (2i, 2i + 1) of this row: 4 bytes per lane, coalesced over rows
              slot_h[(c * 16 + i) * ATT_BM + r_local] = __floats2half2_rn(__uint_as_float(o[2 * i]), __uint_as_float(o[2 * i + 1]));
          }
          slot[ATT_BM * ATT_D / 2 + r_local] = m_ref;
          slot[ATT_BM * ATT_D / 2 + ATT_BM + r_local] = l;
        }
        tc_fence_before();
        mbar_arrive(&o_free[t]);
        if (sg.half) mbar_arrive(&o_free[1]);   // slot 0's warps have read O_1 as well
      }
      cur += sg.advance;
    }
    }
    if (TIMING && timing) {
      for (int k = 0; k < 6; ++k) p.dbg[blockIdx.x * 8 + k] = tm[k];
      p.dbg[blockIdx.x * 8 + 6] = sc;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, 512);
  }
}

// Merges the partial segments of every split item: out = sum_s O_s 2^{(m_s - m) c} / sum_s l_s 2^{(m_s - m) c}.
// grid (items, 2 query tiles, 4 column quarters), 128 threads = one query row each.
__global__ void __launch_bounds__(128)
attention_combine_kernel(const AttnParams p, int grid_fwd) {
  const int t = blockIdx.y, r = threadIdx.x;
  const int qp = blockIdx.x % p.n_qpairs, bh = blockIdx.x / p.n_qpairs;
  const int head = bh % p.H, batch = bh / p.H;
  const int grp = bh / p.heads_per_group;
  const int item = (bh - grp * p.heads_per_group) * p.n_qpairs + qp;   // item index inside its head group
  const bool half = p.half_last && qp == p.n_qpairs - 1;
  if (half && t == 1) return;   // a half item has one query tile; its two slots were merged inside the CTA (slot 0)
  const long long s0 = att_item_first_step(item, p), s1 = s0 + (half ? p.n_half_steps : p.n_kv_tiles) - 1;
  const int c0 = att_step_owner(s0, grid_fwd, grp, p), c1 = att_step_owner(s1, grid_fwd, grp, p);
  if (c0 == c1) return;   // the item was computed whole by one CTA
  const int row = qp * (2 * ATT_BM) + t * ATT_BM + r;
  if (row >= p.Lq) return;
  auto slot_of = [&](int c) {
    const bool first = att_range_start(c, grid_fwd, grp, p) >= s0;   // CTA c's first segment belongs to this item
    return p.ws + ((((long long)grp * grid_fwd + c) * 2 + (first ? 0 : 1)) * 2 + t) * ATT_SLOT_FLOATS;
  };
  // an item is cut by at most a few CTA boundaries: keep the segment pointers and weights in registers
  constexpr int MAX_SEG = 8;
  constexpr int ML = ATT_BM * ATT_D / 2;   // float index of the (m, l) vectors behind the fp16 partial
  const int nseg = (c1 - c0 + 1) < MAX_SEG ? (c1 - c0 + 1) : MAX_SEG;
  const float* seg_ptr[MAX_SEG];
  float w[MAX_SEG];
  float m = -INFINITY;
#pragma unroll
  for (int i = 0; i < MAX_SEG; ++i) {
    seg_ptr[i] = i < nseg ? slot_of(c0 + i) : nullptr;
    if (i < nseg) m = fmaxf(m, seg_ptr[i][ML + r]);
  }
  float L = 0.f;
  float sl2 = p.scale_log2;
  if (p.q_stats != nullptr) sl2 *= stats_rms_rstd(p.q_stats + ((long long)batch * p.Lq + row) * p.q_chunks, p.q_chunks, p.q_eps);
#pragma unroll
  for (int i = 0; i < MAX_SEG; ++i) {
    w[i] = 0.f;
    if (i < nseg) {   // weight of the segment's NORMALISED partial: l_s 2^{(m_s - m) c}
      w[i] = exp2f((seg_ptr[i][ML + r] - m) * sl2) * seg_ptr[i][ML + ATT_BM + r];
      L += w[i];
    }
  }
  const float inv_l = 1.0f / L;
  const int dst = row / p.rows_per_dst;
  __nv_bfloat16* orow = p.out[dst] + (long long)batch * p.out_batch_stride +
                        (long long)(row - dst * p.rows_per_dst) * p.out_row_stride + head * ATT_D;
  // blockIdx.z = quarter of the head's columns: 4x the blocks of a row-per-thread kernel whose occupancy (about two
  // 128-thread blocks per SM do any work) left it latency-bound (38.7 us per launch in profiles/r02q_ncu_launch_shares.json)
  for (int col = blockIdx.z * (ATT_D / 4); col < (blockIdx.z + 1) * (ATT_D / 4); col += 8) {
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < MAX_SEG; ++i) {
      if (i < nseg) {
        const __half2* sh = reinterpret_cast<const __half2*>(seg_ptr[i]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 v2 = __half22float2(sh[(col / 2 + j) * ATT_BM + r]);
          acc[2 * j] = fmaf(v2.x, w[i], acc[2 * j]);
          acc[2 * j + 1] = fmaf(v2.y, w[i], acc[2 * j + 1]);
        }
      }
    }
    uint4 o;
    o.x = pack_bf16(acc[0] * inv_l, acc[1] * inv_l);
    o.y = pack_bf16(acc[2] * inv_l, acc[3] * inv_l);
    o.z = pack_bf16(acc[4] * inv_l, acc[5] * inv_l);
    o.w = pack_bf16(acc[6] * inv_l, acc[7] * inv_l);
    *reinterpret_cast<uint4*>(orow + col) = o;
  }
}

int device_sm_count();

}  // namespace sfb

// Bytes of scratch sfb_attention_fwd needs to balance long-KV problems across all SMs (0 on error).
extern "C" long long sfb_attention_workspace_bytes(void) {
  const int sms = sfb::device_sm_count();
  if (sms <= 0) return 0;
  return (long long)sfb::ATT_MAX_GROUPS * sms * 2 * 2 * sfb::ATT_SLOT_FLOATS * (long long)sizeof(float);
}

namespace sfb {

static int attention_launch(const void* q, long long q_row_stride, long long q_batch_stride, const void* k,
                            const void* v, long long kv_row_stride, long long kv_batch_stride, void* const* out_dst,
                            int n_dst, int rows_per_dst, long long out_row_stride, long long out_batch_stride, int B,
                            int Lq, int Skv, int H, int head_dim, float softmax_scale, void* workspace,
                            long long workspace_bytes, cudaStream_t stream, const void* q_stats = nullptr, int q_chunks = 0,
                            float q_eps = 0.f) {
  if (head_dim != ATT_D) { set_error("sfb_attention_fwd: head_dim %d unsupported (128 only)", head_dim); return SFB_ERR_INVALID; }
  if (B <= 0 || Lq <= 0 || Skv <= 0 || H <= 0) { set_error("sfb_attention_fwd: empty problem B=%d Lq=%d Skv=%d H=%d", B, Lq, Skv, H); return SFB_ERR_INVALID; }
  if ((q_row_stride % 8) || (kv_row_stride % 8) || (out_row_stride % 8) || (q_batch_stride % 8) ||
      (kv_batch_stride % 8) || (out_batch_stride % 8)) {
    set_error("sfb_attention_fwd: strides must be multiples of 8 elements");
    return SFB_ERR_INVALID;
  }
  if (n_dst < 1 || n_dst > 8 || rows_per_dst <= 0 || (long long)n_dst * rows_per_dst < Lq) {
    set_error("sfb_attention_fwd: %d destinations of %d rows do not cover Lq=%d", n_dst, rows_per_dst, Lq);
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
  const int sms = device_sm_count();
  if (sms <= 0) return SFB_ERR_CUDA;
  AttnParams p{};
  const int grid = att_plan(p, B, Lq, Skv, H, sms, workspace != nullptr ? workspace_bytes : 0);
  p.scale_log2 = softmax_scale * 1.4426950408889634f;
  for (int d = 0; d < 8; ++d) p.out[d] = static_cast<__nv_bfloat16*>(out_dst[d < n_dst ? d : 0]);
  p.rows_per_dst = rows_per_dst;
  p.out_row_stride = out_row_stride;
  p.out_batch_stride = out_batch_stride;
  p.ws = static_cast<float*>(workspace);
  p.q_stats = static_cast<const float2*>(q_stats);
  p.q_chunks = q_chunks;
  p.q_eps = q_eps;
  if (q_stats != nullptr && q_chunks <= 0) { set_error("sfb_attention_fwd_qnorm: q_chunks must be positive"); return SFB_ERR_INVALID; }

  static long long* dbg_buf = nullptr;
  static int timing = -1;
  if (timing < 0) {
    const char* env = getenv("SFB_ATTN_TIMING");
    timing = (env && env[0] == '1') ? 1 : 0;
    if (timing) cudaMalloc(&dbg_buf, 512 * 8 * sizeof(long long));   // diagnostic mode only
  }
  p.dbg = timing ? dbg_buf : nullptr;

  // (measured in round 1 and removed: integer-ALU bf16 packing of P; a polynomial exp2 built on floorf + float-to-int,
  // which issue on the quarter-rate conversion pipe -- the magic-number form above replaced it)
  if (timing) {
    static SmemOptIn optin_t;
    if (int e = optin_t.ensure(attention_fwd_kernel<true>, ATT_SMEM_BYTES, "cudaFuncSetAttribute(attention, timing)")) return e;
    attention_fwd_kernel<true><<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tq, tk, tv, p);
  } else {
    static SmemOptIn optin;
    if (int e = optin.ensure(attention_fwd_kernel<false>, ATT_SMEM_BYTES, "cudaFuncSetAttribute(attention)")) return e;
    attention_fwd_kernel<false><<<grid, ATT_THREADS, ATT_SMEM_BYTES, stream>>>(tq, tk, tv, p);
  }
  if (int e = check_cuda(cudaGetLastError(), "attention launch")) return e;
  if (timing) {   // diagnostic: per-phase cycles of one softmax warp, averaged over CTAs, per KV step
    static long long h[512 * 8];
    cudaStreamSynchronize(stream);
    cudaMemcpy(h, dbg_buf, sizeof(long long) * grid * 8, cudaMemcpyDeviceToHost);
    double acc[6] = {0, 0, 0, 0, 0, 0}, steps = 0;
    for (int c = 0; c < grid; ++c) { for (int kk = 0; kk < 6; ++kk) acc[kk] += (double)h[c * 8 + kk]; steps += (double)h[c * 8 + 6]; }
    fprintf(stderr, "[attn timing] Lq=%d S=%d H=%d grid=%d split=%d steps/cta=%.1f | clk/step: wait_s=%.0f ld=%.0f max=%.0f exp=%.0f st+arrive=%.0f other=%.0f total=%.0f\n",
            Lq, Skv, H, grid, p.split, steps / grid, acc[0] / steps, acc[1] / steps, acc[2] / steps, acc[3] / steps, acc[4] / steps,
            acc[5] / steps, (acc[0] + acc[1] + acc[2] + acc[3] + acc[4] + acc[5]) / steps);
  }
  if (p.split) {
    attention_combine_kernel<<<dim3(p.items, 2, 4), 128, 0, stream>>>(p, grid);
    return check_cuda(cudaGetLastError(), "attention combine launch");
  }
  return SFB_OK;
}

}  // namespace sfb

// q   : [B, Lq, H, 128] view with element strides (q_row_stride between tokens, q_batch_stride)
// k, v: cache window start (already offset to attn_start), [B, Skv, H, 128] with their strides
// out : [B, Lq, H, 128] with out_row_stride / out_batch_stride
// workspace: caller-owned scratch of sfb_attention_workspace_bytes() bytes (may be NULL: then long-KV problems
//            are not split across CTAs and run with whole-item granularity)
extern "C" int sfb_attention_fwd(const void* q, long long q_row_stride, long long q_batch_stride, const void* k,
                                 const void* v, long long kv_row_stride, long long kv_batch_stride, void* out,
                                 long long out_row_stride, long long out_batch_stride, int B, int Lq, int Skv,
                                 int H, int head_dim, float softmax_scale, void* workspace,
                                 long long workspace_bytes, void* stream_) {
  void* dst[1] = {out};
  return sfb::attention_launch(q, q_row_stride, q_batch_stride, k, v, kv_row_stride, kv_batch_stride, dst, 1,
                               Lq > 0 ? Lq : 1, out_row_stride, out_batch_stride, B, Lq, Skv, H, head_dim, softmax_scale,
                               workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream_));
}

// sfb_attention_fwd with WanRMSNorm of the query rows folded into the softmax scale: q holds the UN-normalised projection,
// q_stats its statistics records [B * Lq][q_chunks] over the full channel width (written by sfb_gemm_bf16_stats), and the
// caller has multiplied the norm's weight into K.  Replaces norm_q + attention of wan/modules/model.py:172,189.
extern "C" int sfb_attention_fwd_qnorm(const void* q, long long q_row_stride, long long q_batch_stride, const void* k,
                                       const void* v, long long kv_row_stride, long long kv_batch_stride, void* out,
                                       long long out_row_stride, long long out_batch_stride, int B, int Lq, int Skv,
                                       int H, int head_dim, float softmax_scale, const void* q_stats, int q_chunks,
                                       float q_eps, void* workspace, long long workspace_bytes, void* stream_) {
  void* dst[1] = {out};
  return sfb::attention_launch(q, q_row_stride, q_batch_stride, k, v, kv_row_stride, kv_batch_stride, dst, 1,
                               Lq > 0 ? Lq : 1, out_row_stride, out_batch_stride, B, Lq, Skv, H, head_dim, softmax_scale,
                               workspace, workspace_bytes, reinterpret_cast<cudaStream_t>(stream_), q_stats, q_chunks, q_eps);
}

// Head-parallel (Ulysses) form, one sample: this rank attends with ITS head group (H = heads per group) for ALL Lq
// query tokens; output rows [d * rows_per_dst, (d + 1) * rows_per_dst) belong to rank d and are stored straight into
// out_dst[d] (peer-mapped, already offset to this group's columns) at row (token - d * rows_per_dst) -- the reverse
// all-to-all of wan/distributed/xdit_context_parallel.py:179-184 fused into the attention epilogue.
extern "C" int sfb_attention_fwd_sp(const void* q, long long q_row_stride, const void* k, const void* v,
                                    long long kv_row_stride, void* const* out_dst, int n_dst, int rows_per_dst,
                                    long long out_row_stride, int Lq, int Skv, int H, int head_dim,
                                    float softmax_scale, void* workspace, long long workspace_bytes, void* stream_) {
  return sfb::attention_launch(q, q_row_stride, 0, k, v, kv_row_stride, 0, out_dst, n_dst, rows_per_dst, out_row_stride,
                               0, 1, Lq, Skv, H, head_dim, softmax_scale, workspace, workspace_bytes,
                               reinterpret_cast<cudaStream_t>(stream_));
}
