// dlaf/communication/communicator_grid.h — the process grid (reference: include/dlaf/communication/
// communicator_grid.h:37-158). One rank = one GPU; row/column communicators are NCCL sub-communicators owned
// by the C layer, this class is the RAII handle of the grid context.
#pragma once

#include <dlaf/common/index2d.h>
#include <dlaf_c/b200_ext.h>
#include <dlaf_c/grid.h>

namespace dlaf::comm {

using common::Ordering;

class CommunicatorGrid {
public:
  // `comm` may be nullptr for a 1 x 1 grid.
  CommunicatorGrid(DLAF_Comm comm, int rows, int cols, Ordering ordering)
      : ctx_(dlaf_create_grid(comm, rows, cols, ordering == Ordering::ColumnMajor ? 'C' : 'R')) {
    int v[4];
    dlaf_b200_grid_info(ctx_, v);
    size_ = Size2D(v[0], v[1]);
    rank_ = Index2D(v[2], v[3]);
  }
  ~CommunicatorGrid() {
    if (ctx_ >= 0)
      dlaf_free_grid(ctx_);
  }
  CommunicatorGrid(const CommunicatorGrid&) = delete;
  CommunicatorGrid& operator=(const CommunicatorGrid&) = delete;

  Size2D size() const { return size_; }
  Index2D rank() const { return rank_; }
  int context() const { return ctx_; }  // what the C / ScaLAPACK-like API takes
  // reference: wait_all_communicators() drains the MPI pipelines (communicator_grid.cpp:84-95)
  void wait_all_communicators() { dlaf_b200_grid_barrier(ctx_); }

private:
  int ctx_ = -1;
  Size2D size_;
  Index2D rank_;
};

}  // namespace dlaf::comm
