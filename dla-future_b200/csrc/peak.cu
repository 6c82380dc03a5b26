// Measured fp64 tensor-pipe peak of this device: a register-only loop of independent DMMA.8x8x4
// (mma.sync.m8n8k4.f64) on every SM. This is the roofline denominator bench.py reports for the fp64
// kernels: MEASURED_PEAKS.json (driver-written) only holds the HBM copy and bf16 cuBLAS figures, and
// tcgen05 has no f64 kind, so the fp64 tensor roofline IS the DMMA issue rate.
#include <cuda_runtime.h>

#include "common.h"

namespace {
__global__ void dmma_peak_kernel(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i)
    c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i)
    s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
}  // namespace

extern "C" double dlaf_b200_measure_fp64_tensor_peak_tflops(void) {
  using namespace dlaf_b200;
  int dev = 0, nsm = 0;
  DLAF_CUDA_CHECK(cudaGetDevice(&dev));
  DLAF_CUDA_CHECK(cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, dev));
  const int threads = 256, iters = 20000;
  double* out = nullptr;
  DLAF_CUDA_CHECK(cudaMalloc(&out, sizeof(double) * nsm * threads));
  cudaEvent_t e0, e1;
  DLAF_CUDA_CHECK(cudaEventCreate(&e0));
  DLAF_CUDA_CHECK(cudaEventCreate(&e1));
  dmma_peak_kernel<<<nsm, threads>>>(out, 200);
  double best = 0;
  for (int rep = 0; rep < 3; ++rep) {
    DLAF_CUDA_CHECK(cudaEventRecord(e0));
    dmma_peak_kernel<<<nsm, threads>>>(out, iters);
    DLAF_CUDA_CHECK(cudaEventRecord(e1));
    DLAF_CUDA_CHECK(cudaEventSynchronize(e1));
    float ms = 0;
    DLAF_CUDA_CHECK(cudaEventElapsedTime(&ms, e0, e1));
    const double fl = 2.0 * 256 * 16 * double(iters) * (threads / 32) * nsm;
    best = fl / ms / 1e9 > best ? fl / ms / 1e9 : best;
  }
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(out);
  return best;
}
