// 3xTF32 tcgen05 GEMM for the fp32 trailing update (gemm_tf32_tcgen05.cu): host-side handles.
#pragma once

#include <cuda_runtime.h>

#include "gemm_args.h"

namespace dlaf_b200 {

// A panel split into exactly-TF32-representable hi / lo parts, stored K-major (rows x kdim, k contiguous),
// with the TMA tensor maps (CUtensorMap, opaque here) that describe them.
struct Tf32Split {
  float* hi = nullptr;
  float* lo = nullptr;
  long rows = 0;
  int kdim = 0;
  alignas(64) unsigned char map_hi[128];
  alignas(64) unsigned char map_lo[128];

  void allocate(long rows_max, int kdim);
  void release();
  // x: nrows x kdim, column-major (leading dimension ld) -> hi / lo rows [0, nrows).
  // tile_rows / tile_stride describe tile-contiguous panel workspaces: row r of x lives at
  // x + (r / tile_rows) * tile_stride + r % tile_rows (tile_stride == 0: plain column-major).
  void split(const float* x, long ld, long nrows, cudaStream_t s, int tile_rows = 0, long tile_stride = 0);
};

// C = beta C + alpha A B^T with A = rows [a_row, a_row + M) of `sa`, B = rows [b_row, b_row + N) of `sb`;
// mask / geometry / C / alpha / beta taken from `a` (its A, B pointers are ignored).
// b_tile_rows: distance in rows of `sb` between consecutive nbp-tiles of B (nbp when contiguous; Q * nbp when the
// transposed panel aliases every Q-th tile of the column panel, the P == 1 case).
void launch_gemm_tf32x3(const GemmArgsT<float>& a, const Tf32Split& sa, long a_row, const Tf32Split& sb, long b_row,
                        cudaStream_t stream, long b_tile_rows = 0);

}  // namespace dlaf_b200
