// Work decomposition of the attention kernel (attention_tcgen05.cu): launch parameters and the linearised
// (item, KV step) space with its per-CTA ranges.  Host- and device-callable so that the schedule can be checked
// exhaustively on the CPU (tests/test_attention_schedule.py).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace sfb {

struct AttnParams {
  int Lq, Skv, H, B;
  int n_kv_tiles;
  int kv_tail;            // valid columns in the last KV tile (1..128)
  int n_qpairs;           // ceil(Lq / 256)
  int n_qtiles;           // ceil(Lq / 128)
  int half_last;          // 1: the last query pair of every head has an empty second tile and runs as a HALF item
  int n_half_steps;       // steps of a half item = ceil(n_kv_tiles / 2) (two KV tiles per step, one per softmax warpgroup)
  int steps_per_head;     // (n_qpairs - half_last) * n_kv_tiles + half_last * n_half_steps
  int items;              // B * H * n_qpairs
  int split;              // 1: contiguous step ranges per CTA (partials in ws), 0: whole items per CTA
  int heads_per_group;    // (batch, head) pairs processed together; groups run one after the other so that the
  int n_groups;           //   K/V of the heads in flight stays L2-resident (every K/V tile is read by all q tiles)
  float scale_log2;       // softmax_scale * log2(e)
  const float2* q_stats;  // if set: statistics records of the (un-normalised) query rows [B * Lq][q_chunks] (common.cuh); the row's
  int q_chunks;           //   RMSNorm factor rsqrt(mean(q^2) + q_eps) over ALL heads multiplies its scores (WanRMSNorm folded into
  float q_eps;            //   the softmax scale; the norm's weight is folded into K by the caller)
  __nv_bfloat16* out[8];  // query rows [d * rows_per_dst, (d + 1) * rows_per_dst) go to out[d] (Ulysses: peer-mapped
  int rows_per_dst;       //   buffers, the epilogue stores are the reverse all-to-all); one destination otherwise
  long long out_row_stride, out_batch_stride;   // elements
  float* ws;              // [group][grid][2 slots][2 tiles][64*128 half2 (O/l)^T + 128 m + 128 l] (split mode)
  long long* dbg;         // diagnostic phase timers [grid][8] (SFB_ATTN_TIMING=1), else nullptr
};

constexpr int ATT_BM = 128, ATT_BN = 128, ATT_D = 128;
constexpr int ATT_THREADS = 384;
constexpr int ATT_TILE_BYTES = 128 * 128 * 2;    // 32 KB: two 16 KB halves (d 0-63 | d 64-127)
constexpr int ATT_HALF_BYTES = 128 * 64 * 2;
constexpr int ATT_KV_STAGES = 2;
constexpr int ATT_XCHG_BYTES = 2 * ATT_BM * 8;      // half items: (m, l) of slot 1's rows, double buffered by segment parity
constexpr int ATT_SMEM_BYTES = 2 * ATT_TILE_BYTES + 2 * ATT_KV_STAGES * ATT_TILE_BYTES + 1024 + 256 + ATT_XCHG_BYTES;
constexpr int ATT_SLOT_FLOATS = ATT_BM * ATT_D / 2 + 2 * ATT_BM;   // one (tile, segment) partial: O / l as fp16 pairs [64 column pairs][128 rows], then m, l (fp32)
constexpr int ATT_MIN_SPLIT_KV_TILES_MANY_TILES = 160;          // same, when there are at least two query tiles per SM (see att_plan)
constexpr int ATT_MIN_SPLIT_KV_TILES = 48;                     // shorter KV (measured: S = 4680 is better whole, S >= 9360 split): whole items per CTA
constexpr int ATT_MAX_GROUPS = 4;
constexpr long long ATT_L2_BUDGET = 72ll << 20;                // K + V bytes of the heads in flight (L2 is 126 MB; measured: 3 groups of 4 heads beat 4 x 3 and 6 x 2 at S = 32760)

// Linearised work of one head group.  An ITEM is (batch, head, pair of 128-row query tiles) and takes n_kv STEPS (one
// KV tile each, both query tiles).  When the last pair of a head has an EMPTY second tile (Lq = 4680: 36 full tiles + 72
// rows -- one tile in 38 would be pure padding) that pair is a HALF item: both softmax warpgroups work on the one
// query tile, each on every other KV tile, and the two partial results are merged inside the CTA; a half item takes
// n_half = ceil(n_kv / 2) steps of two KV tiles each.  Steps are linearised head-major, item-major.
__host__ __device__ __forceinline__ int att_group_heads(int grp, const AttnParams& p) {
  const int bh0 = grp * p.heads_per_group;
  return (p.B * p.H - bh0) < p.heads_per_group ? (p.B * p.H - bh0) : p.heads_per_group;
}
__host__ __device__ __forceinline__ int att_group_items(int grp, const AttnParams& p) { return att_group_heads(grp, p) * p.n_qpairs; }
__host__ __device__ __forceinline__ long long att_group_steps(int grp, const AttnParams& p) {
  return (long long)att_group_heads(grp, p) * p.steps_per_head;
}
// first step of item `item` (group-local index, item == items_g gives the group's step count)
__host__ __device__ __forceinline__ long long att_item_first_step(int item, const AttnParams& p) {
  return (long long)(item / p.n_qpairs) * p.steps_per_head + (long long)(item % p.n_qpairs) * p.n_kv_tiles;
}
struct AttSeg {
  int bh_local;       // head (group-local) of the item
  int q_tile;         // first 128-row query tile of the item
  int j0, j1;         // steps [j0, j1) of the item
  int item_steps;     // n_kv (full item) or n_half
  int advance;        // how far the segment moves the CTA's cursor
  bool half;
};
// WHOLE-ITEM mode (short KV, p.split == 0) linearises single query TILES, head-major: a CTA owns a contiguous range of
// tiles and walks it pair by pair -- two consecutive tiles of one head form a full item (the pair need not start at an
// even tile), a tile left alone at the end of the range or of the head runs as a half item.  Lq = 4680, H = 12: 444
// tiles = exactly 3 per SM (a pair + a half item) where whole pairs gave 228 items = two rounds on 148 SMs.
// SPLIT mode linearises (item, step) with fixed, even-aligned pairs (see att_item_first_step).
__host__ __device__ __forceinline__ AttSeg att_decode(int cur, int range_end, const AttnParams& p) {
  AttSeg s;
  if (!p.split) {
    s.bh_local = cur / p.n_qtiles;
    s.q_tile = cur - s.bh_local * p.n_qtiles;
    s.half = !((range_end - cur) >= 2 && s.q_tile + 1 < p.n_qtiles);
    s.item_steps = s.half ? p.n_half_steps : p.n_kv_tiles;
    s.j0 = 0;
    s.j1 = s.item_steps;
    s.advance = s.half ? 1 : 2;
    return s;
  }
  s.bh_local = cur / p.steps_per_head;
  const int r = cur - s.bh_local * p.steps_per_head;
  const int full_steps = (p.n_qpairs - p.half_last) * p.n_kv_tiles;
  s.half = r >= full_steps;
  const int qp = s.half ? p.n_qpairs - 1 : r / p.n_kv_tiles;
  s.q_tile = 2 * qp;
  s.j0 = s.half ? r - full_steps : r - qp * p.n_kv_tiles;
  s.item_steps = s.half ? p.n_half_steps : p.n_kv_tiles;
  s.j1 = (range_end - cur) < (s.item_steps - s.j0) ? s.j0 + (range_end - cur) : s.item_steps;
  s.advance = s.j1 - s.j0;
  return s;
}
// Split mode cuts the group's work into one contiguous range per CTA by COST, not by step count: a half-item step moves
// twice the K/V bytes of a full step for the same MMA work and measured 1.17x its time (profiles/r02i), so steps weigh
// ATT_COST_FULL : ATT_COST_HALF = 5 : 6.  cost_start(step) is the cost of everything before the step; CTA c starts at
// the first step whose cost_start >= floor(c * group cost / grid).
constexpr int ATT_COST_FULL = 5, ATT_COST_HALF = 6;
__host__ __device__ __forceinline__ long long att_head_cost(const AttnParams& p) {
  return (long long)ATT_COST_FULL * (p.n_qpairs - p.half_last) * p.n_kv_tiles + (long long)ATT_COST_HALF * p.half_last * p.n_half_steps;
}
__host__ __device__ __forceinline__ long long att_cost_start(long long step, const AttnParams& p) {
  const long long head = step / p.steps_per_head;
  const int local = (int)(step - head * p.steps_per_head);
  const int full_steps = (p.n_qpairs - p.half_last) * p.n_kv_tiles;
  return head * att_head_cost(p) + (local < full_steps ? (long long)ATT_COST_FULL * local
                                                        : (long long)ATT_COST_FULL * full_steps + (long long)ATT_COST_HALF * (local - full_steps));
}
// first step (in the group's linearised space) of CTA c, and the owner of a step
__host__ __device__ __forceinline__ int att_range_start(int c, int grid, int grp, const AttnParams& p) {
  if (!p.split) return (int)(((long long)c * att_group_heads(grp, p) * p.n_qtiles) / grid);   // tiles
  const long long hc = att_head_cost(p);
  const long long b = ((long long)c * att_group_heads(grp, p) * hc) / grid;
  const long long head = b / hc, r = b - head * hc;
  const long long full_cost = (long long)ATT_COST_FULL * (p.n_qpairs - p.half_last) * p.n_kv_tiles;
  const long long local = r <= full_cost ? (r + ATT_COST_FULL - 1) / ATT_COST_FULL
                                         : full_cost / ATT_COST_FULL + (r - full_cost + ATT_COST_HALF - 1) / ATT_COST_HALF;
  return (int)(head * p.steps_per_head + local);
}
__host__ __device__ __forceinline__ int att_step_owner(long long step, int grid, int grp, const AttnParams& p) {
  const long long G = att_group_heads(grp, p) * att_head_cost(p);
  return (int)(((att_cost_start(step, p) + 1) * grid - 1) / G);   // largest c with floor(c * G / grid) <= cost_start(step)   (split mode)
}

// Fills the schedule fields of `p` for a problem and returns the grid size.  workspace_bytes = caller scratch for
// split-KV partials (0: none).
inline int att_plan(AttnParams& p, int B, int Lq, int Skv, int H, int sms, long long workspace_bytes) {
  p.Lq = Lq; p.Skv = Skv; p.H = H; p.B = B;
  p.n_kv_tiles = (Skv + ATT_BN - 1) / ATT_BN;
  p.kv_tail = Skv - (p.n_kv_tiles - 1) * ATT_BN;
  p.n_qpairs = (Lq + 2 * ATT_BM - 1) / (2 * ATT_BM);
  p.n_qtiles = (Lq + ATT_BM - 1) / ATT_BM;
  p.items = B * H * p.n_qpairs;
  p.half_last = (Lq - (p.n_qpairs - 1) * 2 * ATT_BM) <= ATT_BM ? 1 : 0;
  if (const char* env = getenv("SFB_ATTN_NOHALF")) { if (env[0] == '1') p.half_last = 0; }   // diagnostic: pad the lone last tile to a pair
  p.n_half_steps = (p.n_kv_tiles + 1) / 2;
  p.steps_per_head = (p.n_qpairs - p.half_last) * p.n_kv_tiles + p.half_last * p.n_half_steps;
  // Long KV windows: one contiguous range of (item, KV step) work per SM, whatever the item count (also when there
  // are fewer items than SMs -- head-parallel ranks and frame-wise rollouts).  Short ones: whole items per CTA.
  const int tiles = B * H * p.n_qtiles;
  // whole-item mode: one tile per CTA (a half item: n_kv / 2 steps) if that fits the machine, else at least a pair per CTA
  int grid = tiles <= sms ? tiles : ((tiles + 1) / 2 < sms ? (tiles + 1) / 2 : sms);
  p.heads_per_group = B * H;
  p.n_groups = 1;
  const long long slot_bytes = (long long)sms * 2 * 2 * ATT_SLOT_FLOATS * (long long)sizeof(float);   // per group
  const char* nosplit = getenv("SFB_ATTN_NOSPLIT");
  // (with >= 8 items per SM whole items already balance to within a few per cent, and consecutive CTAs walk the same
  // head's K/V together -- no partials, natural L2 locality: the 14B teacher has 5120 items)
  // split threshold: with at least two tiles per SM the tile-granular whole-item mode fills the machine by itself and
  // measured faster than the split up to ~160 KV tiles (S = 9360 / 14040 / 18720: 238.6 / 335.0 / 429.1 -> 226.2 / 322.4 /
  // 422.9 us, equal at 23400: no partials, no combine pass; beyond that the head groups' L2 residency wins); problems with
  // fewer tiles (head-parallel ranks, small batches of short chunks) need the split to use all SMs
  int min_split = tiles >= 2 * sms ? ATT_MIN_SPLIT_KV_TILES_MANY_TILES : ATT_MIN_SPLIT_KV_TILES;
  if (const char* env = getenv("SFB_ATTN_MIN_SPLIT")) { if (atoi(env) > 0) min_split = atoi(env); }   // diagnostic
  p.split = (p.items % sms != 0 && p.items < 8 * sms && p.n_kv_tiles >= min_split &&
             workspace_bytes >= slot_bytes && !(nosplit && nosplit[0] == '1')) ? 1 : 0;
  if (p.split) {
    grid = sms;
    // head groups: keep K + V of the (batch, head) pairs in flight within the L2 budget
    const long long kv_bytes_per_head = 2ll * Skv * ATT_D * 2;
    long long hg_max = ATT_L2_BUDGET / kv_bytes_per_head;
    if (hg_max < 1) hg_max = 1;
    int groups = (int)((B * H + hg_max - 1) / hg_max);
    const int groups_fit = (int)(workspace_bytes / slot_bytes);
    if (groups > ATT_MAX_GROUPS) groups = ATT_MAX_GROUPS;
    if (groups > groups_fit) groups = groups_fit;
    if (const char* env = getenv("SFB_ATTN_GROUPS")) { const int gq = atoi(env); if (gq >= 1 && gq <= groups_fit && gq <= ATT_MAX_GROUPS) groups = gq; }
    p.n_groups = groups;
    p.heads_per_group = (B * H + groups - 1) / groups;
    p.n_groups = (B * H + p.heads_per_group - 1) / p.heads_per_group;
    // the merge kernel keeps at most 8 segments of an item in registers: never cut an item into more pieces
    // (an item of n_kv steps meets at most ceil(n_kv / (steps per CTA)) + 1 ranges)
    const long long g_min = att_group_heads(p.n_groups - 1, p) * att_head_cost(p);   // cost of the smallest (= last) group
    const long long item_cost = (long long)ATT_COST_FULL * p.n_kv_tiles;             // the most expensive item
    if ((item_cost * grid + g_min - 1) / g_min + 1 > 8) grid = (int)(6 * g_min / item_cost);
    if (grid < 1) grid = 1;
  }
  if (const char* cap = getenv("SFB_ATTN_GRID")) {   // diagnostic: run on fewer SMs
    const int g = atoi(cap);
    if (g > 0 && g < grid) grid = g;
  }
  return grid;
}

}  // namespace sfb
