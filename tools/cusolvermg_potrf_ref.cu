// Vendor-library GPU reference #3 (measurement aid only): cusolverMgPotrf on N GPUs of one node, fp64, lower, the
// library's own 1-D column block-cyclic layout. Single process driving all GPUs. Usage: cusolvermg_potrf_ref N ngpus [T_A]
#include <cuda_runtime.h>
#include <cusolverMg.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { auto e_ = (x); if (e_ != 0) { std::printf("error %d at %s:%d\n", (int)e_, __FILE__, __LINE__); return 1; } } while (0)

__global__ void fill_cols(double* a, long n, long lda, long gcol0, int ncols) {
  const long jl = blockIdx.x;
  if (jl >= ncols) return;
  const long j = gcol0 + jl;
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned long long h = (unsigned long long)(i < j ? i * n + j : j * n + i) * 6364136223846793005ULL + 1442695040888963407ULL;
    h ^= h >> 33;
    double v = (double)(h >> 11) / 9007199254740992.0 * 2.0 - 1.0;
    a[i + jl * lda] = (i == j) ? v + 2.0 * n : v;
  }
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 16384;
  const int ng = argc > 2 ? atoi(argv[2]) : 2;
  const int TA = argc > 3 ? atoi(argv[3]) : 256;  // column block of the 1-D block-cyclic layout
  int ndev = 0;
  CK(cudaGetDeviceCount(&ndev));
  if (ndev < ng) { std::printf("needs %d GPUs, %d visible\n", ng, ndev); return 1; }
  std::vector<int> devs(ng);
  for (int i = 0; i < ng; ++i) devs[i] = i;
  for (int i = 0; i < ng; ++i) {  // peer access (the library's own examples enable it)
    cudaSetDevice(i);
    for (int j = 0; j < ng; ++j)
      if (i != j) cudaDeviceEnablePeerAccess(j, 0);
  }
  cudaGetLastError();
  cusolverMgHandle_t h;
  CK(cusolverMgCreate(&h));
  CK(cusolverMgDeviceSelect(h, ng, devs.data()));
  cudaLibMgGrid_t grid;
  CK(cusolverMgCreateDeviceGrid(&grid, 1, ng, devs.data(), CUDALIBMG_GRID_MAPPING_COL_MAJOR));
  cudaLibMgMatrixDesc_t desc;
  CK(cusolverMgCreateMatrixDesc(&desc, n, n, n, TA, CUDA_R_64F, grid));
  // local storage: column blocks b = 0.. of width TA dealt round-robin to the devices, each device's blocks contiguous
  const long nblk = (n + TA - 1) / TA;
  std::vector<double*> dA(ng);
  std::vector<long> lcols(ng, 0);
  for (long b = 0; b < nblk; ++b) lcols[b % ng] += TA;
  for (int d = 0; d < ng; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaMalloc(&dA[d], sizeof(double) * n * lcols[d]));
  }
  auto fill = [&]() {
    for (long b = 0; b < nblk; ++b) {
      const int d = (int)(b % ng);
      const long lb = b / ng;
      cudaSetDevice(d);
      fill_cols<<<TA, 256>>>(dA[d] + lb * TA * n, n, n, b * TA, (int)((b + 1) * TA <= n ? TA : n - b * TA));
    }
    for (int d = 0; d < ng; ++d) { cudaSetDevice(d); cudaDeviceSynchronize(); }
    return 0;
  };
  int64_t lwork = 0;
  CK(cusolverMgPotrf_bufferSize(h, CUBLAS_FILL_MODE_LOWER, (int)n, reinterpret_cast<void**>(dA.data()), 1, 1, desc, CUDA_R_64F, &lwork));
  std::vector<double*> dW(ng);
  for (int d = 0; d < ng; ++d) {
    CK(cudaSetDevice(d));
    CK(cudaMalloc(&dW[d], sizeof(double) * (lwork > 0 ? lwork : 1)));
  }
  double best = 1e30;
  for (int rep = 0; rep < 3; ++rep) {
    fill();
    cudaSetDevice(0);
    cudaEvent_t t0, t1;
    cudaEventCreate(&t0); cudaEventCreate(&t1);
    cudaEventRecord(t0);
    int info = -1;
    CK(cusolverMgPotrf(h, CUBLAS_FILL_MODE_LOWER, (int)n, reinterpret_cast<void**>(dA.data()), 1, 1, desc, CUDA_R_64F,
                       reinterpret_cast<void**>(dW.data()), lwork, &info));
    for (int d = 0; d < ng; ++d) { cudaSetDevice(d); cudaDeviceSynchronize(); }
    cudaSetDevice(0);
    cudaEventRecord(t1); cudaEventSynchronize(t1);
    float ms = 0; cudaEventElapsedTime(&ms, t0, t1);
    std::printf("rep %d: %.2f ms %.1f GFLOP/s info %d\n", rep, ms, (double)n * n * n / 3 / ms / 1e6, info);
    if (ms < best) best = ms;
  }
  std::printf("cusolverMgPotrf N %ld on %d GPUs (T_A %d) best: %.2f ms %.1f GFLOP/s\n", n, ng, TA, best, (double)n * n * n / 3 / best / 1e6);
  return 0;
}
