// Device-memory pool for the per-call workspaces of the algorithms that follow POTRF (triangular solver, inverse,
// generalized -> standard): cudaMalloc / cudaFree of the slabs and digit-plane buffers cost tens of milliseconds per call —
// more than the whole computation at n = 8192 — so freed blocks are kept and handed out again (the role of the
// reference's Umpire pools, src/memory/memory_chunk.cpp, src/init.cpp:93-146). Blocks are returned only after the stream
// that used them has been synchronised by the caller. Everything is released by pool_trim() (dlaf_finalize,
// dlaf_free_grid) and when an allocation fails.
#pragma once

#include <cstddef>

namespace dlaf_b200 {

void* pool_alloc_bytes(size_t bytes);
void pool_free(void* p);  // nullptr is fine
void pool_trim();
size_t pool_cached_bytes();

template <class T>
T* pool_alloc(size_t n) {
  return static_cast<T*>(pool_alloc_bytes(sizeof(T) * (n > 0 ? n : 1)));
}

}  // namespace dlaf_b200
