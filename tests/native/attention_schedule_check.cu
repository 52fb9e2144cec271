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
    const long long G = att_group_steps(grp, p);
    const int items_g = att_group_items(grp, p);
    if (att_range_start(0, grid, grp, p) != 0 || att_range_start(grid, grid, grp, p) != G) fail("range ends", G, 0);
    // KV tiles seen per (item, query-tile slot): every tile exactly once per slot for full items; for half items each
    // tile in exactly one slot (slot parity = tile parity relative to the segment start)
    std::map<long long, int> seen;   // key (item * 2 + slot) * n_kv + tile
    std::vector<int> pieces(items_g, 0);
    for (int c = 0; c < grid; ++c) {
      const int r0 = att_range_start(c, grid, grp, p), r1 = att_range_start(c + 1, grid, grp, p);
      if (r1 < r0) fail("range order", r0, r1);
      for (int cur = r0; cur < r1;) {
        const AttSeg sg = att_decode(cur, r1, p);
        if (sg.j1 <= sg.j0 || sg.j1 > sg.item_steps) { fail("segment", sg.j0, sg.j1); break; }
        const int item = sg.bh_local * p.n_qpairs + sg.qp;
        if (item >= items_g) { fail("item", item, items_g); break; }
        if (att_item_first_step(item, p) + sg.j0 != cur) fail("first step", cur, item);
        if (sg.half != (p.half_last && sg.qp == p.n_qpairs - 1)) fail("half flag", item, sg.half);
        if (sg.item_steps != (sg.half ? p.n_half_steps : p.n_kv_tiles)) fail("item steps", item, sg.item_steps);
        if (!p.split && !(sg.j0 == 0 && sg.j1 == sg.item_steps)) fail("whole-item mode cut an item", item, c);
        pieces[item]++;
        for (int s = cur; s < cur + (sg.j1 - sg.j0); ++s)
          if (att_step_owner(s, grid, grp, p) != c && p.split) fail("owner", s, c);
        if (sg.half) {
          const int lo = 2 * sg.j0, hi = 2 * sg.j1 < p.n_kv_tiles ? 2 * sg.j1 : p.n_kv_tiles;
          for (int j = lo; j < hi; ++j) seen[((long long)item * 2 + ((j - lo) & 1)) * p.n_kv_tiles + j]++;
        } else {
          for (int j = sg.j0; j < sg.j1; ++j)
            for (int t = 0; t < 2; ++t) seen[((long long)item * 2 + t) * p.n_kv_tiles + j]++;
        }
        cur += sg.j1 - sg.j0;
      }
    }
    for (int item = 0; item < items_g; ++item) {
      const bool half = p.half_last && (item % p.n_qpairs) == p.n_qpairs - 1;
      if (pieces[item] < 1 || pieces[item] > 8) fail("pieces", item, pieces[item]);
      for (int j = 0; j < p.n_kv_tiles; ++j) {
        const int a = seen[((long long)item * 2) * p.n_kv_tiles + j], b = seen[((long long)item * 2 + 1) * p.n_kv_tiles + j];
        if (half ? (a + b != 1) : (a != 1 || b != 1)) fail("coverage", item, j);
      }
      if (p.split) {   // the combine kernel's view
        const long long s0 = att_item_first_step(item, p), s1 = s0 + (half ? p.n_half_steps : p.n_kv_tiles) - 1;
        const int c0 = att_step_owner(s0, grid, grp, p), c1 = att_step_owner(s1, grid, grp, p);
        if (c1 - c0 + 1 != pieces[item]) fail("combine piece count", item, c1 - c0 + 1);
      }
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
