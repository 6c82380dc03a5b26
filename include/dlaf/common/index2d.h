// dlaf/common/index2d.h — minimal 2D size/index pairs (reference: include/dlaf/common/index2d.h,
// include/dlaf/matrix/index.h) as far as the Cholesky surface needs them.
#pragma once

#include <ostream>

#include <dlaf/types.h>

namespace dlaf::common {

template <class Tag>
class Pair2D {
public:
  constexpr Pair2D() = default;
  constexpr Pair2D(SizeType r, SizeType c) : row_(r), col_(c) {}
  constexpr SizeType rows() const { return row_; }
  constexpr SizeType cols() const { return col_; }
  constexpr SizeType row() const { return row_; }
  constexpr SizeType col() const { return col_; }
  constexpr bool operator==(const Pair2D& o) const { return row_ == o.row_ && col_ == o.col_; }
  friend std::ostream& operator<<(std::ostream& os, const Pair2D& p) {
    return os << "(" << p.row_ << ", " << p.col_ << ")";
  }

private:
  SizeType row_ = 0, col_ = 0;
};

enum class Ordering { RowMajor, ColumnMajor };

}  // namespace dlaf::common

namespace dlaf {
struct GlobalElementTag;
struct LocalElementTag;
struct TileElementTag;
struct GridTag;
using GlobalElementSize = common::Pair2D<GlobalElementTag>;
using LocalElementSize = common::Pair2D<LocalElementTag>;
using TileElementSize = common::Pair2D<TileElementTag>;
namespace comm {
using Size2D = common::Pair2D<GridTag>;
using Index2D = common::Pair2D<GridTag>;
}
}  // namespace dlaf
