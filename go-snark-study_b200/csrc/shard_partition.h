// Index-range partition of the four MSMs of a Groth16 proving key (A, B1, B2, C||PTD) over `world` ranks
// (groth16/groth16.go:241-270 are the four sums).  Plain C++ — no CUDA — so that the CPU suite can
// compile it with g++ and hold it against its Python mirror (go-snark-study_b200/shard.py, tests/test_shard_partition.py).
//
// The sets lie end to end on a line weighted by cost per term (wgt[k], in G1 terms of the C||PTD set).  Rank g takes a
// contiguous stretch of that line, i.e. whole sets where it can and an index range where it must.
//   phase cost = 0   the line is cut into `world` equal pieces (closed form; the round-2 partition).
//   phase cost > 0   every PIECE a rank holds is one more MSM on that rank — its own digit sort, slice merge and bucket
//                    tail, none of which shrinks with the piece — so opening a piece of set k costs fix[k] on top of
//                    wgt[k] per term.  The ranks are filled greedily, in order, up to a common capacity T (a rank never
//                    opens a piece it cannot pay the fixed cost of), and T is the smallest capacity that covers the line
//                    (bisection; the sweep is monotone in T, and greedy filling is optimal for the largest load: a
//                    rank that opens a sliver spends capacity it would otherwise leave unused).  The last rank takes
//                    whatever is left.
// Both forms tile every set exactly: rank g's `lo` is rank g-1's `hi`.
#pragma once
#include <cstddef>

namespace b200 {

struct ShardCut {
  size_t lo[4], hi[4];
};

// Greedy sweep with capacity T.  force_last: the last rank ignores its capacity (used for the final assignment).
// Returns true when the ranks cover all four sets.
inline bool shard_sweep(const size_t len[4], const double wgt[4], const double fix[4], int world, double T, bool force_last,
                        ShardCut* out) {
  int k = 0;
  size_t pos = 0;
  for (int g = 0; g < world; g++) {
    ShardCut c;
    for (int j = 0; j < 4; j++) c.lo[j] = c.hi[j] = j < k ? len[j] : (j == k ? pos : 0);
    double cap = T;
    const bool unbounded = force_last && g == world - 1;
    while (k < 4) {
      const size_t rem = len[k] - pos;
      if (rem == 0) {   // empty set (or exactly finished): move on
        k++;
        pos = 0;
        continue;
      }
      size_t take = rem;
      if (!unbounded) {
        if (cap <= fix[k]) break;
        const double avail = (cap - fix[k]) / wgt[k];
        if (avail < (double)rem) take = (size_t)avail;
        if (take == 0) break;
      }
      c.lo[k] = pos;
      c.hi[k] = pos + take;
      cap -= fix[k] + wgt[k] * (double)take;
      pos += take;
      if (pos < len[k]) break;   // partial piece: this rank is full
      k++;
      pos = 0;
    }
    if (out) out[g] = c;
  }
  return k == 4;
}

// cuts[g] for g in [0, world).  fix[k] all zero -> equal pieces of the weighted line.
inline void shard_partition(const size_t len[4], const double wgt[4], const double fix[4], int world, ShardCut* cuts) {
  double off[5] = {0, 0, 0, 0, 0};
  for (int k = 0; k < 4; k++) off[k + 1] = off[k] + wgt[k] * (double)len[k];
  if (fix[0] == 0 && fix[1] == 0 && fix[2] == 0 && fix[3] == 0) {
    auto cut = [&](int g, int k) -> size_t {   // first index of set k at or after the g-th cut of the line
      if (g >= world) return len[k];
      double pos = off[4] * (double)g / (double)world;
      double x = (pos - off[k]) / wgt[k];
      if (x <= 0) return 0;
      if (x >= (double)len[k]) return len[k];
      return (size_t)x;
    };
    for (int g = 0; g < world; g++)
      for (int k = 0; k < 4; k++) {
        cuts[g].lo[k] = cut(g, k);
        cuts[g].hi[k] = cut(g + 1, k);
      }
    return;
  }
  double t_lo = 0, t_hi = off[4] + fix[0] + fix[1] + fix[2] + fix[3] + 1.0;   // one rank holding everything: feasible
  for (int it = 0; it < 64; it++) {
    double mid = 0.5 * (t_lo + t_hi);
    if (shard_sweep(len, wgt, fix, world, mid, false, nullptr)) t_hi = mid;
    else t_lo = mid;
  }
  shard_sweep(len, wgt, fix, world, t_hi, true, cuts);
}

}  // namespace b200
