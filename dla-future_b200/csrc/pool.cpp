// See pool.h.
#include "pool.h"

#include <cuda_runtime.h>

#include <cstdlib>
#include <map>
#include <mutex>
#include <unordered_map>

#include "common.h"

namespace dlaf_b200 {

namespace {
std::mutex g_mu;
std::multimap<size_t, void*> g_free;          // cached blocks by size
std::unordered_map<void*, size_t> g_size;     // every block handed out or cached
size_t g_cached = 0;

size_t cap_bytes() {
  static const size_t cap = [] {
    const char* e = std::getenv("DLAF_B200_POOL_MAX_GB");
    return static_cast<size_t>((e ? std::atof(e) : 64.0) * (1ull << 30));
  }();
  return cap;
}
size_t granule(size_t b) {
  const size_t g = b >= (1u << 20) ? (2u << 20) : 512;
  return (b + g - 1) / g * g;
}
void trim_locked() {
  for (auto& kv : g_free) {
    cudaFree(kv.second);
    g_size.erase(kv.second);
  }
  g_free.clear();
  g_cached = 0;
}
}  // namespace

void* pool_alloc_bytes(size_t bytes) {
  const size_t want = granule(bytes);
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_free.lower_bound(want);
  if (it != g_free.end() && it->first <= want + want / 4) {
    void* p = it->second;
    g_cached -= it->first;
    g_free.erase(it);
    return p;
  }
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, want);
  if (e != cudaSuccess) {
    cudaGetLastError();
    trim_locked();
    e = cudaMalloc(&p, want);
  }
  DLAF_CUDA_CHECK(e);
  g_size[p] = want;
  return p;
}

void pool_free(void* p) {
  if (p == nullptr)
    return;
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_size.find(p);
  if (it == g_size.end()) {  // not ours
    cudaFree(p);
    return;
  }
  if (g_cached + it->second > cap_bytes()) {
    cudaFree(p);
    g_size.erase(it);
    return;
  }
  g_free.emplace(it->second, p);
  g_cached += it->second;
}

void pool_trim() {
  std::lock_guard<std::mutex> lk(g_mu);
  trim_locked();
}

size_t pool_cached_bytes() {
  std::lock_guard<std::mutex> lk(g_mu);
  return g_cached;
}

}  // namespace dlaf_b200
