// CPU test vehicle: C entry point over csrc/shard_partition.h for tests/test_shard_partition.py (g++, no CUDA).
#include <cstdint>
#include <vector>
#include "shard_partition.h"
extern "C" int shard_partition_c(const uint64_t len[4], const double wgt[4], int world, uint64_t* lo_hi) {
  size_t l[4];
  for (int k = 0; k < 4; k++) l[k] = (size_t)len[k];
  std::vector<b200::ShardCut> cuts((size_t)world);
  b200::shard_partition(l, wgt, world, cuts.data());
  for (int g = 0; g < world; g++)
    for (int k = 0; k < 4; k++) {
      lo_hi[8 * g + 2 * k] = cuts[(size_t)g].lo[k];
      lo_hi[8 * g + 2 * k + 1] = cuts[(size_t)g].hi[k];
    }
  return 0;
}
