// Exhaustive CPU check of the attention kernel's work decomposition (self_forcing_b200/csrc/attention_sched.cuh):
// every (item, step) is owned by exactly one CTA range, the segments a CTA decodes tile its range, half items carry
// ceil(n_kv / 2) steps, the combine kernel's owner arithmetic finds exactly the CTAs that parked a partial, and no
// item is cut into more pieces than the combine kernel holds.  Host code only (no GPU).
#include <stdio.h>
#include <vector>
#include <map>
#include "attention_sched.cuh"

using namespace sfb;

static int check(int B, int Lq, int Skv, int H, int sms, long long ws_bytes) {
  AttnParams p{};
  const int grid = att_plan(p, B, Lq, Skv, H, sms, ws_bytes);
  int fails = 0;
  auto fail = [&](const char* what, long long a, long long b) {
    if (fails++ < 5) printf("FAIL B%d Lq%d S%d H%d sms%d: %s (%lld, %lld)\n", B, Lq, Skv, H, sms, what, a, b);
  };
  if (grid < 1 || grid > sms) fail("grid", grid, sms);
  const int q_tiles = (Lq + ATT_BM - 1) / ATT_BM;
  if (p.half_last != (q_tiles % 2)) fail("half_last", p.half_last, q_tiles);
  for (int grp = 0; grp < p.n_groups; ++grp) {
    const long long G = p.split ? att_group_steps(grp, p) : (long long)att_group_heads(grp, p) * p.n_qtiles;
    if (att_range_start(0, grid, grp, p) != 0 || att_range_start(grid, grid, grp, p) != G) fail("range ends", G, 0);
    // every (head, query tile, KV tile) must be computed exactly once
    std::map<long long, int> seen;                 // key (head * n_qtiles + tile) * n_kv + kv tile
    std::map<long long, int> pieces;               // key head * n_qtiles + first tile of the item -> number of segments
    const int heads = att_group_heads(grp, p);
    for (int c = 0; c < grid; ++c) {
      const int r0 = att_range_start(c, grid, grp, p), r1 = att_range_start(c + 1, grid, grp, p);
      if (r1 < r0) fail("range order", r0, r1);
      for (int cur = r0; cur < r1;) {
        const AttSeg sg = att_decode(cur, r1, p);
        if (sg.j1 <= sg.j0 || sg.j1 > sg.item_steps || sg.advance < 1) { fail("segment", sg.j0, sg.j1); break; }
        if (sg.bh_local >= heads || sg.q_tile >= p.n_qtiles) { fail("item", sg.bh_local, sg.q_tile); break; }
        if (!sg.half && sg.q_tile + 1 >= p.n_qtiles) fail("pair runs past the head's tiles", sg.bh_local, sg.q_tile);
        if (sg.item_steps != (sg.half ? p.n_half_steps : p.n_kv_tiles)) fail("item steps", sg.q_tile, sg.item_steps);
        if (!p.split && !(sg.j0 == 0 && sg.j1 == sg.item_steps)) fail("whole-item mode cut an item", sg.q_tile, c);
        if (p.split) {
          const int item = sg.bh_local * p.n_qpairs + sg.q_tile / 2;
          if (sg.q_tile & 1) fail("split-mode pair not even-aligned", sg.q_tile, 0);
          if (att_item_first_step(item, p) + sg.j0 != cur) fail("first step", cur, item);
          if (sg.half != (p.half_last && sg.q_tile / 2 == p.n_qpairs - 1)) fail("half flag", item, sg.half);
          for (int s = cur; s < cur + sg.advance; ++s)
            if (att_step_owner(s, grid, grp, p) != c) fail("owner", s, c);
        }
        pieces[(long long)sg.bh_local * p.n_qtiles + sg.q_tile]++;
        const long long base = (long long)sg.bh_local * p.n_qtiles + sg.q_tile;
        if (sg.half) {
          const int lo = 2 * sg.j0, hi = 2 * sg.j1 < p.n_kv_tiles ? 2 * sg.j1 : p.n_kv_tiles;
          for (int j = lo; j < hi; ++j) seen[base * p.n_kv_tiles + j]++;
        } else {
          for (int j = sg.j0; j < sg.j1; ++j)
            for (int t = 0; t < 2; ++t) seen[(base + t) * p.n_kv_tiles + j]++;
        }
        cur += sg.advance;
      }
    }
    for (int h = 0; h < heads; ++h)
      for (int q = 0; q < p.n_qtiles; ++q)
        for (int j = 0; j < p.n_kv_tiles; ++j)
          if (seen[((long long)h * p.n_qtiles + q) * p.n_kv_tiles + j] != 1) fail("coverage", h * 1000 + q, j);
    for (auto& kv : pieces) {
      if (kv.second < 1 || kv.second > 8) fail("pieces", kv.first, kv.second);
      if (p.split) {   // the combine kernel's view
        const int h = (int)(kv.first / p.n_qtiles), q = (int)(kv.first % p.n_qtiles);
        const int item = h * p.n_qpairs + q / 2;
        const bool half = p.half_last && q / 2 == p.n_qpairs - 1;
        const long long s0 = att_item_first_step(item, p), s1 = s0 + (half ? p.n_half_steps : p.n_kv_tiles) - 1;
        const int c0 = att_step_owner(s0, grid, grp, p), c1 = att_step_owner(s1, grid, grp, p);
        if (c1 - c0 + 1 != kv.second) fail("combine piece count", item, c1 - c0 + 1);
      }
    }
    if (!p.split) {   // whole-item mode: tiles are dealt out evenly (to within one tile)
      int lo = 1 << 30, hi = 0;
      for (int c = 0; c < grid; ++c) {
        const int n = att_range_start(c + 1, grid, grp, p) - att_range_start(c, grid, grp, p);
        lo = n < lo ? n : lo; hi = n > hi ? n : hi;
      }
      if (hi - lo > 1) fail("tile balance", lo, hi);
    }
  }
  return fails;
}

int main() {
  int fails = 0, cases = 0;
  const long long big_ws = 1ll << 32;
  const int Lqs[] = {1, 72, 128, 129, 200, 256, 257, 300, 384, 385, 1170, 1300, 1560, 2340, 4000, 4680, 9360};
  const int Ss[] = {72, 128, 512, 520, 1000, 1560, 4680, 6143, 6144, 6145, 9360, 18720, 20000, 32760, 40000};
  const int Hs[] = {1, 2, 3, 12, 32, 40};
  for (int B = 1; B <= 2; ++B)
    for (int Lq : Lqs)
      for (int S : Ss)
        for (int H : Hs)
          for (int sms : {148, 132, 7})
            for (long long ws : {0ll, big_ws}) { fails += check(B, Lq, S, H, sms, ws); ++cases; }
  fails += check(1, 75600, 75600, 40, 148, big_ws);   // 14B teacher: 5120 items
  printf("%s: %d cases, %d failures\n", fails ? "FAILED" : "OK", cases + 1, fails);
  return fails ? 1 : 0;
}
