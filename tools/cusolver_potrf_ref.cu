// MEASUREMENT AID, not part of the product: the vendor-library GPU reference SURVEY.md §8(d) asks to be timed on
// the same box — monolithic cusolverDnDpotrf (the routine the reference's GPU backend calls per TILE,
// include/dlaf/lapack/tile.h:696-725, here on the whole matrix) for the bench sizes, device-resident, lower.
// Run on the GPU box:  tools/cusolver_potrf_ref [N ...]
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x)                                                                     \
  do {                                                                            \
    auto e_ = (x);                                                                \
    if (e_ != 0) {                                                                \
      std::printf("error %d at %s:%d (%s)\n", (int)e_, __FILE__, __LINE__, #x);   \
      return 1;                                                                   \
    }                                                                             \
  } while (0)

// symmetric, diagonally dominant: a(i,j) = hash(min,max) in (-1,1), a(i,i) = 2n (like the miniapp's generator)
__global__ void fill_spd(double* a, long n) {
  const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long j = blockIdx.y;
  if (i >= n)
    return;
  const unsigned long long lo = i < j ? i : j, hi = i < j ? j : i;
  unsigned long long h = lo * 0x9E3779B97F4A7C15ull + hi * 0xC2B2AE3D27D4EB4Full + 12345;
  h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
  const double u = (double)(h >> 11) / 9007199254740992.0 * 2.0 - 1.0;
  a[i + j * n] = (i == j) ? 2.0 * n : u;
}

int main(int argc, char** argv) {
  std::vector<long> sizes;
  for (int i = 1; i < argc; ++i) sizes.push_back(std::atol(argv[i]));
  if (sizes.empty()) sizes = {16384, 32768};
  cusolverDnHandle_t h;
  CK(cusolverDnCreate(&h));
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  for (long n : sizes) {
    double* a;
    CK(cudaMalloc(&a, sizeof(double) * n * n));
    int lwork = 0;
    CK(cusolverDnDpotrf_bufferSize(h, CUBLAS_FILL_MODE_LOWER, (int)n, a, (int)n, &lwork));
    double* work;
    int* info;
    CK(cudaMalloc(&work, sizeof(double) * (size_t)lwork));
    CK(cudaMalloc(&info, sizeof(int)));
    double best = 1e30;
    for (int rep = 0; rep < 4; ++rep) {
      fill_spd<<<dim3((unsigned)((n + 255) / 256), (unsigned)n), 256>>>(a, n);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      CK(cusolverDnDpotrf(h, CUBLAS_FILL_MODE_LOWER, (int)n, a, (int)n, work, lwork, info));
      cudaEventRecord(e1);
      CK(cudaEventSynchronize(e1));
      float ms;
      cudaEventElapsedTime(&ms, e0, e1);
      int hinfo = -1;
      cudaMemcpy(&hinfo, info, sizeof(int), cudaMemcpyDeviceToHost);
      if (rep > 0 && ms < best) best = ms;
      std::printf("cusolverDnDpotrf N=%ld rep %d: %.2f ms  %.1f GFLOP/s  info %d\n", n, rep, ms,
                  (double)n * n * n / 3.0 / ms / 1e6, hinfo);
    }
    std::printf("cusolverDnDpotrf N=%ld best: %.2f ms  %.1f GFLOP/s\n", n, best, (double)n * n * n / 3.0 / best / 1e6);
    cudaFree(a); cudaFree(work); cudaFree(info);
  }
  return 0;
}
