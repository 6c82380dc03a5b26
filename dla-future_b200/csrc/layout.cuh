// Boundary layout conversion between the reference's local matrix layout and the engine's slab.
//
// Reference side (include/dlaf/matrix/matrix.h:268-283, matrix/col_major_layout.h:59-78): one
// column-major slab per rank with leading dimension ld; local tile (li, lj) starts at
// li*nb + lj*nb*ld; the last global tile may be ragged; only the `uplo` triangle is referenced.
//
// Engine side: always LOWER, every tile padded to nbp = round_up(nb, granularity) so no kernel ever
// sees a ragged tile. Padding rows/cols carry an identity diagonal, which leaves the factor of the
// real entries unchanged (chol of a symmetric permutation of [A 0; 0 I]). uplo == 'U' is mapped onto
// the lower algorithm by conjugate-transposing tiles at this boundary and swapping the roles of the
// process-grid rows and columns (tile (i,j), i<j, on rank (i%P, j%Q) becomes lower tile (j,i) of
// U^H on rank (j%Q, i%P) of the transposed Q x P grid).
#pragma once

#include <cuda_runtime.h>

namespace dlaf_b200 {

struct LayoutParams {
  long n;       // global matrix size
  int nb;       // user tile size
  int nbp;      // padded tile size
  int nt;       // global number of tiles
  int P, Q;     // ENGINE grid (swapped w.r.t. the user's when transposed)
  int prow, pcol;
  int ltr, ltc;  // engine local tile rows / cols
  long ld;       // slab leading dimension
  long ldu;      // user leading dimension
  int transposed;  // user holds the upper triangle
};

// user (device-visible pointer) -> slab, only the triangle that is referenced.
template <class T>
void launch_to_slab(T* slab, const T* user, const LayoutParams& p, cudaStream_t s);
// slab -> user, writes ONLY the referenced triangle (never the other one, not even inside diagonal tiles).
template <class T>
void launch_from_slab(const T* slab, T* user, const LayoutParams& p, cudaStream_t s);
// zero the slab padding and put ones on the padded part of the global diagonal.
template <class T>
void launch_pad_identity(T* slab, const LayoutParams& p, cudaStream_t s);
// column block (ntiles*nbp x nbp, leading dimension ld) -> ntiles contiguous nbp x nbp tiles.
template <class T>
void launch_pack_panel(const T* src, long ld, T* dst, int nbp, int ntiles, cudaStream_t s);

}  // namespace dlaf_b200
