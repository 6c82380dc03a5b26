// 1-D block-cyclic index math of the POTRF path (one tile per block, no tile offset) — the product-side counterpart of
// the reference's include/dlaf/matrix/util_distribution.h:82-196. The engine works in "virtual" coordinates (rank
// minus source rank, modulo the grid size), for which these reduce to g % P, g / P and cnt_tiles(); the functions
// below keep the reference's signatures (rank and source rank explicit) and are exported through the C ABI
// (dlaf_b200_*_tile*) so that tests can hold them against the reference's own table
// (test/unit/matrix/test_util_distribution.cpp:49-55 -> tests/golden/util_distribution_cases.json).
#pragma once

namespace dlaf_b200 {

// util_distribution.h:82-92
inline int rank_global_tile(long global_tile, int grid_size, int src_rank) {
  return static_cast<int>((global_tile + src_rank) % grid_size);
}
// util_distribution.h:103-126  (-1 if `rank` does not own the tile)
inline long local_tile_from_global_tile(long global_tile, int grid_size, int rank, int src_rank) {
  return rank_global_tile(global_tile, grid_size, src_rank) == rank ? global_tile / grid_size : -1;
}
// util_distribution.h:138-166: local index of the first tile of `rank` with global index >= global_tile
inline long next_local_tile_from_global_tile(long global_tile, int grid_size, int rank, int src_rank) {
  const int v = (rank - src_rank + grid_size) % grid_size;  // virtual coordinate
  // number of g' in [0, global_tile) with g' % grid_size == v
  return global_tile > v ? (global_tile - v + grid_size - 1) / grid_size : 0;
}
// util_distribution.h:177-196
inline long global_tile_from_local_tile(long local_tile, int grid_size, int rank, int src_rank) {
  const int v = (rank - src_rank + grid_size) % grid_size;
  return local_tile * grid_size + v;
}

// Number of elements of one dimension (global size n, block nb) held by virtual rank v of `grid_size`
// (src/matrix/distribution.cpp:117-150: local tiles x block minus the shortfall of the ragged last tile if it is mine).
inline long local_size_1d(long n, int nb, int grid_size, int v) {
  if (n <= 0)
    return 0;
  const long nt = (n + nb - 1) / nb;
  long s = next_local_tile_from_global_tile(nt, grid_size, v, 0) * nb;
  if ((nt - 1) % grid_size == v)
    s -= nt * nb - n;
  return s;
}

}  // namespace dlaf_b200
