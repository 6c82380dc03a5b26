// Host-only check of the header-only C++ surface (no GPU, no library call): matrix::Distribution against the reference's
// own table (tests/golden/util_distribution_cases.json is turned into the lines fed on stdin by tests/test_cpp_headers.py):
//   grid src rank global_tile  exp_rank exp_local exp_next_local   |   grid src rank local_tile exp_global
// and a few derived quantities in closed form. Also instantiates every public template of include/dlaf (syntax + ODR).
#include <cstdio>
#include <cstring>

#include <dlaf/eigensolver/gen_to_std.h>
#include <dlaf/factorization/cholesky.h>
#include <dlaf/inverse/cholesky.h>
#include <dlaf/inverse/triangular.h>
#include <dlaf/matrix/distribution.h>
#include <dlaf/solver/triangular.h>

using namespace dlaf;

int main() {
  char kind[8];
  long a, b, c, d, e, f, g;
  int bad = 0, n = 0;
  while (std::scanf("%7s", kind) == 1) {
    if (std::strcmp(kind, "G2L") == 0) {
      if (std::scanf("%ld %ld %ld %ld %ld %ld %ld", &a, &b, &c, &d, &e, &f, &g) != 7)
        return 2;
      // row flavour and column flavour must agree with the table
      matrix::Distribution dr(GlobalElementSize(1000, 7), TileElementSize(1, 1), comm::Size2D(a, 1), comm::Index2D(c, 0), comm::Index2D(b, 0));
      matrix::Distribution dc(GlobalElementSize(7, 1000), TileElementSize(1, 1), comm::Size2D(1, a), comm::Index2D(0, c), comm::Index2D(0, b));
      bad += dr.rank_global_tile<Coord::Row>(d) != e || dc.rank_global_tile<Coord::Col>(d) != e;
      bad += dr.local_tile_from_global_tile<Coord::Row>(d) != f || dc.local_tile_from_global_tile<Coord::Col>(d) != f;
      bad += dr.next_local_tile_from_global_tile<Coord::Row>(d) != g || dc.next_local_tile_from_global_tile<Coord::Col>(d) != g;
    }
    else {
      if (std::scanf("%ld %ld %ld %ld %ld", &a, &b, &c, &d, &e) != 5)
        return 2;
      matrix::Distribution dr(GlobalElementSize(1000, 7), TileElementSize(1, 1), comm::Size2D(a, 1), comm::Index2D(c, 0), comm::Index2D(b, 0));
      bad += dr.global_tile_from_local_tile<Coord::Row>(d) != e;
    }
    ++n;
  }
  // derived quantities: 34 x 34, tiles 13 x 13 on a 2 x 3 grid with source rank (1, 1)
  long tot_r = 0, tot_c = 0;
  for (int r = 0; r < 2; ++r)
    for (int cc = 0; cc < 3; ++cc) {
      matrix::Distribution dd(GlobalElementSize(34, 34), TileElementSize(13, 13), comm::Size2D(2, 3), comm::Index2D(r, cc), comm::Index2D(1, 1));
      bad += !(dd.nr_tiles() == GlobalTileSize(3, 3));
      if (cc == 0)
        tot_r += dd.local_size().rows();
      if (r == 0)
        tot_c += dd.local_size().cols();
      bad += !(dd.rank_global_tile(GlobalTileIndex(0, 0)) == comm::Index2D(1, 1));
      bad += !(dd.tile_size_of(GlobalTileIndex(2, 1)) == TileElementSize(8, 13));
      const auto lt = dd.local_nr_tiles();
      for (long li = 0; li < lt.rows(); ++li)
        bad += dd.local_tile_from_global_tile<Coord::Row>(dd.global_tile_from_local_tile<Coord::Row>(li)) != li;
    }
  bad += tot_r != 34 || tot_c != 34;
  std::printf("%d table lines, %d mismatches\n", n, bad);
  return bad ? 1 : 0;
}
