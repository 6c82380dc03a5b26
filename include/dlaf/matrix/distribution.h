// dlaf/matrix/distribution.h — matrix::Distribution as the Cholesky path and its tests use it (reference:
// include/dlaf/matrix/distribution.h:115-1054 and the free functions of matrix/util_distribution.h:82-196): the 2D
// block-cyclic bookkeeping of ONE matrix — size, tile size (one tile per block, no tile offset: the only case on this path,
// factorization/cholesky.h:45, :75), process grid, this process' rank and the source rank that holds tile (0, 0).
// Header-only, host-only; the same arithmetic the library uses internally (csrc/distribution.h) and exports through
// dlaf_b200_*_tile* — tests/test_cpp_headers.py holds this class against the reference's own table.
#pragma once

#include <algorithm>

#include <dlaf/common/index2d.h>
#include <dlaf/types.h>

namespace dlaf {

enum class Coord { Row, Col };  // reference: include/dlaf/types.h:64

struct GlobalTileTag;
struct LocalTileTag;
using GlobalTileIndex = common::Pair2D<GlobalTileTag>;
using GlobalTileSize = common::Pair2D<GlobalTileTag>;
using LocalTileIndex = common::Pair2D<LocalTileTag>;
using LocalTileSize = common::Pair2D<LocalTileTag>;

namespace matrix {

class Distribution {
public:
  Distribution() = default;
  // non-distributed matrix (distribution.h:122-126)
  Distribution(const LocalElementSize& size, const TileElementSize& tile_size)
      : Distribution(GlobalElementSize(size.rows(), size.cols()), tile_size, comm::Size2D(1, 1), comm::Index2D(0, 0),
                     comm::Index2D(0, 0)) {}
  // distribution.h:139-143
  Distribution(const GlobalElementSize& size, const TileElementSize& tile_size, const comm::Size2D& grid_size,
               const comm::Index2D& rank_index, const comm::Index2D& source_rank_index)
      : size_(size), tile_(tile_size), grid_(grid_size), rank_(rank_index), src_(source_rank_index) {}

  const GlobalElementSize& size() const noexcept { return size_; }
  const TileElementSize& tile_size() const noexcept { return tile_; }
  const TileElementSize& block_size() const noexcept { return tile_; }
  const TileElementSize& blockSize() const noexcept { return tile_; }
  const comm::Size2D& grid_size() const noexcept { return grid_; }
  const comm::Size2D& commGridSize() const noexcept { return grid_; }
  const comm::Index2D& rank_index() const noexcept { return rank_; }
  const comm::Index2D& rankIndex() const noexcept { return rank_; }
  const comm::Index2D& source_rank_index() const noexcept { return src_; }
  const comm::Index2D& sourceRankIndex() const noexcept { return src_; }

  GlobalTileSize nr_tiles() const noexcept { return GlobalTileSize(ntiles(size_.rows(), tile_.rows()), ntiles(size_.cols(), tile_.cols())); }
  GlobalTileSize nrTiles() const noexcept { return nr_tiles(); }
  LocalTileSize local_nr_tiles() const noexcept {
    return LocalTileSize(next_local_tile_from_global_tile<Coord::Row>(nr_tiles().rows()),
                         next_local_tile_from_global_tile<Coord::Col>(nr_tiles().cols()));
  }
  LocalTileSize localNrTiles() const noexcept { return local_nr_tiles(); }
  LocalElementSize local_size() const noexcept { return LocalElementSize(local_len<Coord::Row>(), local_len<Coord::Col>()); }
  LocalElementSize localSize() const noexcept { return local_size(); }

  // util_distribution.h:82-92
  template <Coord rc>
  int rank_global_tile(SizeType global_tile) const noexcept {
    return static_cast<int>((global_tile + get<rc>(src_)) % get<rc>(grid_));
  }
  comm::Index2D rank_global_tile(const GlobalTileIndex& t) const noexcept {
    return comm::Index2D(rank_global_tile<Coord::Row>(t.row()), rank_global_tile<Coord::Col>(t.col()));
  }
  // util_distribution.h:103-126 (-1 if this rank does not own the tile)
  template <Coord rc>
  SizeType local_tile_from_global_tile(SizeType global_tile) const noexcept {
    return rank_global_tile<rc>(global_tile) == get<rc>(rank_) ? global_tile / get<rc>(grid_) : -1;
  }
  // util_distribution.h:138-166: local index of this rank's first tile with global index >= global_tile
  template <Coord rc>
  SizeType next_local_tile_from_global_tile(SizeType global_tile) const noexcept {
    const SizeType g = get<rc>(grid_), v = (get<rc>(rank_) - get<rc>(src_) + g) % g;
    return global_tile > v ? (global_tile - v + g - 1) / g : 0;
  }
  // util_distribution.h:177-196
  template <Coord rc>
  SizeType global_tile_from_local_tile(SizeType local_tile) const noexcept {
    const SizeType g = get<rc>(grid_), v = (get<rc>(rank_) - get<rc>(src_) + g) % g;
    return local_tile * g + v;
  }
  GlobalTileIndex global_tile_index(const LocalTileIndex& t) const noexcept {
    return GlobalTileIndex(global_tile_from_local_tile<Coord::Row>(t.row()), global_tile_from_local_tile<Coord::Col>(t.col()));
  }
  LocalTileIndex local_tile_index(const GlobalTileIndex& t) const noexcept {
    return LocalTileIndex(local_tile_from_global_tile<Coord::Row>(t.row()), local_tile_from_global_tile<Coord::Col>(t.col()));
  }
  // distribution.h:614-627
  template <Coord rc>
  SizeType global_tile_size_of(SizeType global_tile) const noexcept {
    const SizeType n = get<rc>(size_), b = get<rc>(tile_);
    return std::min(b, n - global_tile * b);
  }
  TileElementSize tile_size_of(const GlobalTileIndex& t) const noexcept {
    return TileElementSize(global_tile_size_of<Coord::Row>(t.row()), global_tile_size_of<Coord::Col>(t.col()));
  }
  template <Coord rc>
  bool is_source_rank() const noexcept {
    return get<rc>(rank_) == get<rc>(src_);
  }
  bool operator==(const Distribution& o) const noexcept {
    return size_ == o.size_ && tile_ == o.tile_ && grid_ == o.grid_ && rank_ == o.rank_ && src_ == o.src_;
  }

private:
  template <Coord rc, class P>
  static SizeType get(const P& p) noexcept {
    return rc == Coord::Row ? p.row() : p.col();
  }
  static SizeType ntiles(SizeType n, SizeType b) noexcept { return n > 0 ? (n + b - 1) / b : 0; }
  // src/matrix/distribution.cpp:117-150
  template <Coord rc>
  SizeType local_len() const noexcept {
    const SizeType n = get<rc>(size_), b = get<rc>(tile_);
    if (n <= 0)
      return 0;
    const SizeType nt = ntiles(n, b);
    SizeType s = next_local_tile_from_global_tile<rc>(nt) * b;
    if (rank_global_tile<rc>(nt - 1) == get<rc>(rank_))
      s -= nt * b - n;
    return s;
  }

  GlobalElementSize size_;
  TileElementSize tile_{1, 1};
  comm::Size2D grid_{1, 1};
  comm::Index2D rank_, src_;
};

}  // namespace matrix
}  // namespace dlaf
