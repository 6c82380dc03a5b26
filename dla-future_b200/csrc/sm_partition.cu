// Spatial SM partition for the bulk stream (CUDA green contexts, driver API >= 12.4).
//
// The trailing update saturates every SM with long-lived CTAs (one 200 KB CTA per SM), so the short kernels of
// the panel chain (diagonal blocks, TRSM, splits) — the critical path once the bulk is fast — queue for an SM
// to free up. With DLAF_B200_RESERVE_SMS=n (multiple of 8) the bulk stream lives in a green context that owns all
// SMs but n; the chain streams stay in the primary context and always find the reserved SMs idle.
// Any failure (old driver, unsupported device) falls back to an ordinary stream.
#include <cuda.h>
#include <cuda_runtime.h>

#include <cstdio>

#include "common.h"

namespace dlaf_b200 {

namespace {
template <class F>
F driver_fn(const char* name) {
  void* f = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint(name, &f, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess)
    return nullptr;
  return reinterpret_cast<F>(f);
}

struct Partition {
  CUgreenCtx ctx = nullptr;
  int reserved = 0, total = 0;
  bool tried = false;
};
Partition g_part;
}  // namespace

// Stream (non-blocking, given priority) restricted to all SMs except `reserve_sms`; nullptr if not available.
cudaStream_t create_bulk_stream_with_reserved_sms(int reserve_sms, int priority) {
  using GetRes = CUresult (*)(CUdevice, CUdevResource*, CUdevResourceType);
  using Split = CUresult (*)(CUdevResource*, unsigned int*, const CUdevResource*, CUdevResource*, unsigned int, unsigned int);
  using GenDesc = CUresult (*)(CUdevResourceDesc*, CUdevResource*, unsigned int);
  using CtxCreate = CUresult (*)(CUgreenCtx*, CUdevResourceDesc, CUdevice, unsigned int);
  using StreamCreate = CUresult (*)(CUstream*, CUgreenCtx, unsigned int, int);
  if (reserve_sms <= 0)
    return nullptr;
  if (!g_part.tried) {
    g_part.tried = true;
    auto get_res = driver_fn<GetRes>("cuDeviceGetDevResource");
    auto split = driver_fn<Split>("cuDevSmResourceSplitByCount");
    auto gen = driver_fn<GenDesc>("cuDevResourceGenerateDesc");
    auto create = driver_fn<CtxCreate>("cuGreenCtxCreate");
    if (!get_res || !split || !gen || !create) {
      std::fprintf(stderr, "[dlaf_b200] SM reservation: green-context API not available, using ordinary streams\n");
      return nullptr;
    }
    int dev = 0;
    DLAF_CUDA_CHECK(cudaGetDevice(&dev));
    DLAF_CUDA_CHECK(cudaFree(nullptr));  // make sure the primary context exists
    CUdevResource all, group, rest;
    unsigned int ngroups = 1;
    CUresult r = get_res(static_cast<CUdevice>(dev), &all, CU_DEV_RESOURCE_TYPE_SM);
    if (r == CUDA_SUCCESS)
      r = split(&group, &ngroups, &all, &rest, 0, static_cast<unsigned int>(reserve_sms));
    CUdevResourceDesc desc = nullptr;
    if (r == CUDA_SUCCESS)
      r = gen(&desc, &rest, 1);
    if (r == CUDA_SUCCESS)
      r = create(&g_part.ctx, desc, static_cast<CUdevice>(dev), CU_GREEN_CTX_DEFAULT_STREAM);
    if (r != CUDA_SUCCESS) {
      std::fprintf(stderr, "[dlaf_b200] SM reservation failed (CUresult %d), using ordinary streams\n", static_cast<int>(r));
      g_part.ctx = nullptr;
      return nullptr;
    }
    g_part.reserved = static_cast<int>(group.sm.smCount);
    g_part.total = static_cast<int>(all.sm.smCount);
    std::fprintf(stderr, "[dlaf_b200] bulk stream restricted to %u of %u SMs (%u reserved for the panel chain)\n",
                 rest.sm.smCount, all.sm.smCount, group.sm.smCount);
  }
  if (g_part.ctx == nullptr)
    return nullptr;
  auto screate = driver_fn<StreamCreate>("cuGreenCtxStreamCreate");
  if (!screate)
    return nullptr;
  CUstream s = nullptr;
  if (screate(&s, g_part.ctx, CU_STREAM_NON_BLOCKING, priority) != CUDA_SUCCESS)
    return nullptr;
  return reinterpret_cast<cudaStream_t>(s);
}

}  // namespace dlaf_b200
