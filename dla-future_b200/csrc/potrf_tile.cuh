// Diagonal-block kernel of the POTRF path: Cholesky of one 128x128 block PLUS the inverse of its
// factor, in a single shared-memory sweep by one CTA.
//
// Replaces potrfDiagTile -> tile::potrf -> cusolverDn?potrf (+ bufferSize query, workspace
// allocation and the assert_info<<<1,1>>> kernel) of the reference
// (include/dlaf/factorization/cholesky/impl.h:46-53, include/dlaf/lapack/tile.h:696-725,
// src/cusolver/assert_info.cu:21-44). A 512x512 (nb) diagonal tile is factorised as 4 of these
// blocks + GEMM-shaped sub-steps (engine.cu), so the inverse is what turns every TRSM on the path
// into tensor-core GEMMs.
//
// Contract (same as LAPACK potrf, lower): only the lower triangle of T is read or written — the
// strictly upper part of the block is never touched (the reference's tests fill it with a sentinel,
// test/include/dlaf_test/matrix/util_generic_lapack.h:39-68). Non-SPD input: *info is set to
// info_offset + (1-based failing column) if it was 0; no trap.
#pragma once

#include <cuda_runtime.h>

namespace dlaf_b200 {

constexpr int kPotrfBlock = 128;

// T: 128x128 block, column-major ldt (lower triangle in/out). W: 128x128 column-major ldw, receives
// inv(L) (full square written: lower = inverse, strictly upper = 0).
void launch_potrf128_inv_f64(double* T, long ldt, double* W, long ldw, int* info, int info_offset,
                             cudaStream_t stream);

// Measurement aid (tools/): device buffer of 16*8*2 clock64 stamps per launch, nullptr switches it off.
void potrf_set_clock_trace(long long* dev_buffer);

// Whole diagonal tile in one cluster launch (potrf_tile_cluster.cu): T = nbp x nbp tile (lower triangle in/out), W =
// nbp / 128 inverted diagonal blocks (each 128 x 128 contiguous, ld 128). nbp in {128, 256, 384, 512}.
bool potrf_tile_cluster_supported(int nbp);
void launch_potrf_tile_cluster_f64(double* T, long ldt, double* W, int nbp, int* info, int info_offset,
                                   cudaStream_t stream);
// Measurement aid: device buffer of 64 * 12 clock64 stamps (CTA 0, thread 0, per panel step); nullptr = off.
void potrf_tile_set_clock_trace(long long* dev_buffer);

// Per element type entry point: Cholesky + inverse of one Gran<T> x Gran<T> diagonal block.
template <class T>
void launch_potrf_inv(T* t, long ldt, T* w, long ldw, int* info, int info_offset, cudaStream_t stream);

}  // namespace dlaf_b200
