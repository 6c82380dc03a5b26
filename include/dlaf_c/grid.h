/* dlaf_c/grid.h — replaces include/dlaf_c/grid.h:31-77 of the reference.
 *
 * The reference takes an MPI_Comm. This build has no MPI (one process per GPU, NCCL over NVLink), so
 * the communicator argument is an opaque DLAF_Comm created from an NCCL unique id that the launcher
 * distributes (torch.distributed / any bootstrap). Compiling against a real mpi.h
 * (-DDLAF_B200_WITH_MPI) keeps the reference's exact prototype; see INTEGRATION.md. */
#pragma once

#include <dlaf_c/utils.h>

#ifdef DLAF_B200_WITH_MPI
#include <mpi.h>
typedef MPI_Comm DLAF_Comm;
#else
struct dlaf_b200_comm;
typedef struct dlaf_b200_comm* DLAF_Comm; /* NULL is valid for a 1x1 grid */
#endif

#define DLAF_B200_UNIQUE_ID_BYTES 128

/* Bootstrap (replaces MPI_Init/MPI_COMM_WORLD): rank 0 calls dlaf_b200_get_unique_id and ships the
 * 128 bytes to every rank; then every rank calls dlaf_b200_comm_create (collective). */
DLAF_EXTERN_C void dlaf_b200_get_unique_id(void* id128) DLAF_NOEXCEPT;
DLAF_EXTERN_C struct dlaf_b200_comm* dlaf_b200_comm_create(const void* id128, int rank,
                                                            int nranks) DLAF_NOEXCEPT;
/* Geometry-only communicator (no NCCL object, no GPU needed): grids built on it answer layout queries
 * (local sizes, coordinates, input generator) but abort on any factorization. */
DLAF_EXTERN_C struct dlaf_b200_comm* dlaf_b200_comm_create_local(int rank, int nranks) DLAF_NOEXCEPT;
DLAF_EXTERN_C void dlaf_b200_comm_destroy(struct dlaf_b200_comm* comm) DLAF_NOEXCEPT;

/* Returns a context (counting down from INT_MAX like src/c_api/grid.cpp:28-40). order 'R' or 'C'. */
DLAF_EXTERN_C int dlaf_create_grid(DLAF_Comm comm, int nprow, int npcol, char order) DLAF_NOEXCEPT;
DLAF_EXTERN_C void dlaf_free_grid(int context) DLAF_NOEXCEPT;
DLAF_EXTERN_C void dlaf_free_all_grids(void) DLAF_NOEXCEPT;
/* 'R' / 'C': how ranks of `comm` map to (myprow, mypcol) (reference: grid.h:60-61). */
DLAF_EXTERN_C char grid_ordering(DLAF_Comm comm, int nprow, int npcol, int myprow,
                                 int mypcol) DLAF_NOEXCEPT;
