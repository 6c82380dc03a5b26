// See comm.h.
#include "comm.h"

#include <cstring>

namespace dlaf_b200 {

Comm* comm_create(const void* unique_id, int rank, int size) {
  ncclUniqueId id;
  static_assert(sizeof(ncclUniqueId) == NCCL_UNIQUE_ID_BYTES, "id size");
  std::memcpy(&id, unique_id, sizeof(id));
  Comm* c = new Comm;
  c->rank = rank;
  c->size = size;
  DLAF_NCCL_CHECK(ncclCommInitRank(&c->nccl, size, id, rank));
  return c;
}

void comm_destroy(Comm* c) {
  if (!c)
    return;
  if (c->nccl)
    ncclCommDestroy(c->nccl);
  delete c;
}

CommGrid::CommGrid(Comm* world, int P_, int Q_, char order) : P(P_), Q(Q_) {
  if (world == nullptr) {
    if (P * Q != 1) {
      std::fprintf(stderr, "[dlaf_b200] a %dx%d grid needs a communicator\n", P, Q);
      std::abort();
    }
    return;
  }
  world_rank = world->rank;
  world_size = world->size;
  if (P * Q > world_size) {
    std::fprintf(stderr, "[dlaf_b200] grid %dx%d larger than communicator (%d ranks)\n", P, Q, world_size);
    std::abort();
  }
  in_grid = world_rank < P * Q;
  const bool col_major = (order == 'C' || order == 'c');  // src/c_api/utils.cpp:51-54
  if (in_grid) {
    row = col_major ? world_rank % P : world_rank / Q;
    col = col_major ? world_rank / P : world_rank % Q;
  }
  if (P * Q == 1)
    return;  // a 1x1 grid never communicates
  if (world->nccl == nullptr)
    return;  // geometry-only communicator (dlaf_b200_comm_create_local): no collectives possible
  // ncclCommSplit is collective over the parent: left-out ranks pass NCCL_SPLIT_NOCOLOR.
  DLAF_NCCL_CHECK(ncclCommSplit(world->nccl, in_grid ? row : NCCL_SPLIT_NOCOLOR, col, &row_comm, nullptr));
  DLAF_NCCL_CHECK(ncclCommSplit(world->nccl, in_grid ? col : NCCL_SPLIT_NOCOLOR, row, &col_comm, nullptr));
  DLAF_NCCL_CHECK(ncclCommSplit(world->nccl, in_grid ? 0 : NCCL_SPLIT_NOCOLOR, world_rank, &grid_comm, nullptr));
  DLAF_NCCL_CHECK(ncclCommSplit(world->nccl, in_grid ? row : NCCL_SPLIT_NOCOLOR, col, &row_comm_h, nullptr));
  DLAF_NCCL_CHECK(ncclCommSplit(world->nccl, in_grid ? col : NCCL_SPLIT_NOCOLOR, row, &col_comm_h, nullptr));
}

CommGrid::~CommGrid() {
  if (row_comm)
    ncclCommDestroy(row_comm);
  if (col_comm)
    ncclCommDestroy(col_comm);
  if (grid_comm)
    ncclCommDestroy(grid_comm);
  if (row_comm_h)
    ncclCommDestroy(row_comm_h);
  if (col_comm_h)
    ncclCommDestroy(col_comm_h);
}

}  // namespace dlaf_b200
