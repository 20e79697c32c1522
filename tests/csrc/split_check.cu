// Host-side check of stream_common.cuh's work split (no GPU needed): for random geometries, the per-CTA row ranges
// [cost_to_row(b * share), cost_to_row((b + 1) * share)) partition [0, total_rows) in order, and no CTA's COST
// (rows + seg_overhead per strip start it contains) exceeds the share by more than one strip's overhead.
#include <cstdio>
#include <cstdlib>
#include <random>

#include "stream_common.cuh"

using namespace bfcnn::stream;

int main() {
  std::mt19937_64 rng(1);
  int cases = 0;
  for (int it = 0; it < 20000; ++it) {
    Split p{};
    p.tiles_x = 1 + (int)(rng() % 200);
    p.rows_needed = 1 + (int)(rng() % 5000);
    p.seg_overhead = (int)(rng() % 40);
    const int sm = 1 + (int)(rng() % 148), min_share = 1 + (int)(rng() % 32);
    const int grid = plan_split(p, sm, min_share);
    if (grid < 1 || grid > sm) { printf("grid %d out of range (sm %d)\n", grid, sm); return 1; }
    long long prev = 0;
    for (int b = 0; b < grid; ++b) {
      long long r0 = cost_to_row(p, (long long)b * p.share), r1 = cost_to_row(p, ((long long)b + 1) * p.share);
      if (r0 > p.total_rows) r0 = p.total_rows;
      if (r1 > p.total_rows) r1 = p.total_rows;
      if (r0 != prev || r1 < r0) { printf("case %d: CTA %d range [%lld, %lld) does not continue %lld\n", it, b, r0, r1, prev); return 1; }
      // cost of the range: its rows plus one overhead per segment (a segment = the part of a strip inside the range)
      long long cost = 0;
      for (long long a = r0; a < r1;) {
        const long long strip = a / p.rows_needed, end = (strip + 1) * p.rows_needed < r1 ? (strip + 1) * p.rows_needed : r1;
        cost += (end - a) + p.seg_overhead;
        a = end;
      }
      if (cost > p.share + 2ll * p.seg_overhead) { printf("case %d: CTA %d cost %lld > share %lld + 2 x overhead %d\n", it, b, cost, p.share, p.seg_overhead); return 1; }
      prev = r1;
    }
    if (prev != p.total_rows) { printf("case %d: ranges end at %lld of %lld\n", it, prev, p.total_rows); return 1; }
    ++cases;
  }
  printf("ok %d geometries\n", cases);
  return 0;
}
