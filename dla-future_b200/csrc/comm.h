// Process grid over NCCL — the replacement of the reference's MPI CommunicatorGrid
// (include/dlaf/communication/communicator_grid.h:37-158, src/communication/communicator_grid.cpp:28-96:
// three MPI_Comm_split + per-pipeline MPI_Comm_dup). One rank = one GPU; the row/column
// sub-communicators are ncclCommSplit children, and every collective is enqueued on a CUDA stream so
// it is ordered with the kernels by stream order/events instead of MPI request polling.
#pragma once

#include <cuda_runtime.h>
#include <nccl.h>

#include <cstdio>
#include <cstdlib>

namespace dlaf_b200 {

#define DLAF_NCCL_CHECK(expr)                                                                   \
  do {                                                                                          \
    ncclResult_t r_ = (expr);                                                                   \
    if (r_ != ncclSuccess) {                                                                    \
      std::fprintf(stderr, "[dlaf_b200] NCCL error %s at %s:%d: %s\n", ncclGetErrorString(r_),  \
                   __FILE__, __LINE__, #expr);                                                  \
      std::fflush(stderr);                                                                      \
      std::abort();                                                                             \
    }                                                                                           \
  } while (0)

}  // namespace dlaf_b200

// World communicator handle handed through the C ABI in place of MPI_Comm (include/dlaf_c/grid.h).
struct dlaf_b200_comm {
  ncclComm_t nccl = nullptr;
  int rank = 0;
  int size = 1;
};

namespace dlaf_b200 {
using Comm = ::dlaf_b200_comm;

Comm* comm_create(const void* unique_id, int rank, int size);  // collective (ncclCommInitRank)
void comm_destroy(Comm* c);

class CommGrid {
public:
  // Collective over `world` (may be nullptr for a 1x1 grid: then no NCCL object exists at all).
  // order: 'R' row-major rank -> (r / Q, r % Q); 'C' column-major rank -> (r % P, r / P)
  // (include/dlaf/common/index2d.h:301-332).
  CommGrid(Comm* world, int P, int Q, char order);
  ~CommGrid();
  CommGrid(const CommGrid&) = delete;
  CommGrid& operator=(const CommGrid&) = delete;

  int P, Q;
  int row = 0, col = 0;  // my coordinates
  bool in_grid = true;   // ranks >= P*Q are left out (communicator_grid.cpp:32, :53-54)
  int world_rank = 0, world_size = 1;
  ncclComm_t row_comm = nullptr;  // ranks of my process row, size Q, my rank = col
  ncclComm_t col_comm = nullptr;  // ranks of my process column, size P, my rank = row
  ncclComm_t grid_comm = nullptr; // all ranks of the grid (info reduction / barriers)
  // Second, independent pair for the critical-path stream (diagonal tile down the column, first panel tile along the
  // row): the reference keeps 3 round-robin clones per communicator so that independent collectives do not queue behind
  // each other (src/communication/communicator_grid.cpp:64-75, include/dlaf/tune.h:162); here one clone per STREAM that
  // issues collectives is what decouples them (NCCL orders operations per communicator).
  ncclComm_t row_comm_h = nullptr;
  ncclComm_t col_comm_h = nullptr;
};

template <class T>
struct NcclType;
template <>
struct NcclType<float> {
  static constexpr ncclDataType_t value = ncclFloat;
  static constexpr int mult = 1;
};
template <>
struct NcclType<double> {
  static constexpr ncclDataType_t value = ncclDouble;
  static constexpr int mult = 1;
};
template <>
struct NcclType<float2> {
  static constexpr ncclDataType_t value = ncclFloat;
  static constexpr int mult = 2;
};
template <>
struct NcclType<double2> {
  static constexpr ncclDataType_t value = ncclDouble;
  static constexpr int mult = 2;
};

}  // namespace dlaf_b200
