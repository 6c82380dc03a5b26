// dlaf/matrix/matrix.h — Matrix<T, Device> as seen by the Cholesky path (reference: include/dlaf/matrix/
// matrix.h:61-212, :349-501): the rank's local part of a 2D block-cyclic matrix in ONE column-major slab
// (AllocationLayout::ColMajor, matrix.h:268-283), either allocated here or wrapping caller memory
// (matrix.h:137-139, :519-539). The per-tile sender pipelines of the reference do not exist: operations on
// a matrix are ordered on its CUDA stream and waitLocalTiles() synchronises that stream.
#pragma once

#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>

#include <dlaf/common/index2d.h>
#include <dlaf/communication/communicator_grid.h>
#include <dlaf/matrix/distribution.h>
#include <dlaf_c/b200_ext.h>
#include <dlaf_c/desc.h>

namespace dlaf::matrix {

namespace internal {
inline void check(cudaError_t e, const char* what) {
  if (e != cudaSuccess) {
    std::fprintf(stderr, "[dlaf] CUDA error in %s: %s\n", what, cudaGetErrorString(e));
    std::abort();
  }
}
}  // namespace internal

template <class T, Device D>
class Matrix {
public:
  using ElementType = T;

  // Distributed over `grid`, source rank (0, 0) (matrix.h:112-114); allocates the local part
  // (Device::CPU: pinned host memory, Device::GPU: device memory; ld = local rows rounded up to even).
  Matrix(GlobalElementSize size, TileElementSize block, comm::CommunicatorGrid& grid)
      : Matrix(size, block, grid, comm::Index2D(0, 0)) {}
  // Same with an explicit source rank (the reference passes it through Distribution, matrix/distribution.h:115-160,
  // as in test/unit/factorization/test_cholesky.cpp:85-88).
  Matrix(GlobalElementSize size, TileElementSize block, comm::CommunicatorGrid& grid, comm::Index2D src_rank)
      : size_(size), block_(block), ctx_(grid.context()), grid_size_(grid.size()), src_rank_(src_rank) {
    init_geometry(static_cast<int>(src_rank.row()), static_cast<int>(src_rank.col()), 0);
    const std::size_t bytes = sizeof(T) * static_cast<std::size_t>(ld_) * (lcols_ > 0 ? lcols_ : 1);
    if (D == Device::GPU)
      internal::check(cudaMalloc(reinterpret_cast<void**>(&ptr_), bytes), "Matrix allocation");
    else
      internal::check(cudaMallocHost(reinterpret_cast<void**>(&ptr_), bytes), "Matrix allocation");
    owns_ = true;
    internal::check(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "stream");
  }
  // From a Distribution, as the reference's tests build their matrices (test/unit/factorization/test_cholesky.cpp:87-88:
  // `Distribution distribution(size, block_size, grid.size(), grid.rank(), src_rank_index); Matrix<T, D> mat(std::move(distribution));`).
  // The reference's constructor needs no grid because its communicators travel separately; here the matrix is bound to the
  // grid context, so the grid is the second argument (it must be the grid the distribution was built from).
  Matrix(const matrix::Distribution& distribution, comm::CommunicatorGrid& grid)
      : Matrix(distribution.size(), distribution.tile_size(), grid, distribution.source_rank_index()) {
    if (!(distribution.grid_size() == grid.size()) || !(distribution.rank_index() == grid.rank())) {
      std::fprintf(stderr, "[dlaf] Matrix: the distribution does not belong to the given communicator grid\n");
      std::abort();
    }
  }
  // the bookkeeping of this matrix (reference: Matrix::distribution(), matrix_base.h)
  matrix::Distribution distribution() const {
    int v[4];
    dlaf_b200_grid_info(ctx_, v);
    return matrix::Distribution(size_, block_, grid_size_, comm::Index2D(v[2], v[3]), src_rank_);
  }
  // Wraps caller memory: local part at `ptr`, column-major with leading dimension `ld`.
  Matrix(GlobalElementSize size, TileElementSize block, comm::CommunicatorGrid& grid, comm::Index2D src_rank,
         T* ptr, SizeType ld)
      : size_(size), block_(block), ctx_(grid.context()), grid_size_(grid.size()), src_rank_(src_rank), ptr_(ptr) {
    init_geometry(static_cast<int>(src_rank.row()), static_cast<int>(src_rank.col()), ld);
    internal::check(cudaStreamCreateWithFlags(&stream_, cudaStreamNonBlocking), "stream");
  }
  ~Matrix() {
    if (stream_) {
      cudaStreamSynchronize(stream_);
      cudaStreamDestroy(stream_);
    }
    if (owns_) {
      if (D == Device::GPU)
        cudaFree(ptr_);
      else
        cudaFreeHost(ptr_);
    }
  }
  Matrix(const Matrix&) = delete;
  Matrix& operator=(const Matrix&) = delete;

  GlobalElementSize size() const { return size_; }
  TileElementSize blockSize() const { return block_; }
  TileElementSize block_size() const { return block_; }
  LocalElementSize localSize() const { return LocalElementSize(lrows_, lcols_); }
  comm::Size2D commGridSize() const { return grid_size_; }
  comm::Index2D sourceRankIndex() const { return src_rank_; }
  T* ptr() { return ptr_; }
  const T* ptr() const { return ptr_; }
  SizeType ld() const { return ld_; }
  int context() const { return ctx_; }
  DLAF_descriptor descriptor() const { return desc_; }
  cudaStream_t stream() const { return stream_; }

  // reference: blocks until every local tile pipeline is drained (matrix.h waitLocalTiles)
  void waitLocalTiles() { internal::check(cudaStreamSynchronize(stream_), "waitLocalTiles"); }

private:
  void init_geometry(int isrc, int jsrc, SizeType ld) {
    desc_ = DLAF_descriptor{static_cast<int>(size_.rows()), static_cast<int>(size_.cols()),
                            static_cast<int>(block_.rows()), static_cast<int>(block_.cols()), isrc, jsrc, 0, 0, 1};
    lrows_ = dlaf_b200_local_rows(ctx_, desc_);
    lcols_ = dlaf_b200_local_cols(ctx_, desc_);
    ld_ = ld > 0 ? ld : (lrows_ + 1) / 2 * 2;
    if (ld_ < 1)
      ld_ = 2;
    desc_.ld = static_cast<int>(ld_);
  }

  GlobalElementSize size_;
  TileElementSize block_;
  int ctx_;
  comm::Size2D grid_size_;
  comm::Index2D src_rank_{0, 0};
  T* ptr_ = nullptr;
  SizeType ld_ = 1, lrows_ = 0, lcols_ = 0;
  DLAF_descriptor desc_{};
  bool owns_ = false;
  cudaStream_t stream_ = nullptr;
};

// copy(source, dest) between matrices with the same distribution (reference: matrix/copy.h); ordered on the
// destination's stream.
template <class T, Device S, Device D>
void copy(Matrix<T, S>& src, Matrix<T, D>& dst) {
  const auto ls = src.localSize();
  if (ls.rows() == 0 || ls.cols() == 0)
    return;
  src.waitLocalTiles();
  internal::check(cudaMemcpy2DAsync(dst.ptr(), sizeof(T) * dst.ld(), src.ptr(), sizeof(T) * src.ld(),
                                    sizeof(T) * ls.rows(), ls.cols(), cudaMemcpyDefault, dst.stream()),
                  "matrix copy");
}

// MatrixMirror<T, Target, Source> (reference: matrix/matrix_mirror.h:102-174): a twin of `source` on the Target
// device, filled on construction and copied back on destruction; a no-op alias when both devices agree.
template <class T, Device Target, Device Source>
class MatrixMirror {
public:
  explicit MatrixMirror(Matrix<T, Source>& source, comm::CommunicatorGrid& grid)
      : source_(source), twin_(source.size(), source.blockSize(), grid, source.sourceRankIndex()) {
    // same distribution (grid, block size AND source rank) as the source: the local parts have identical shapes
    copy(source_, twin_);
  }
  ~MatrixMirror() {
    copy(twin_, source_);
    source_.waitLocalTiles();
  }
  Matrix<T, Target>& get() { return twin_; }

private:
  Matrix<T, Source>& source_;
  Matrix<T, Target> twin_;
};

template <class T, Device Same>
class MatrixMirror<T, Same, Same> {
public:
  explicit MatrixMirror(Matrix<T, Same>& source, comm::CommunicatorGrid&) : source_(source) {}
  Matrix<T, Same>& get() { return source_; }

private:
  Matrix<T, Same>& source_;
};

}  // namespace dlaf::matrix

namespace dlaf {
using matrix::Matrix;
}
