// Vendor-library GPU reference #2 (measurement aid only; nothing in the product links cuBLAS / cuSOLVER):
// the REFERENCE'S OWN SCHEDULE on one GPU — right-looking tiled Cholesky with one library call per tile task,
//   potrf: cusolverDnDpotrf      (include/dlaf/lapack/tile.h:696-725)
//   trsm : cublasDtrsm           (include/dlaf/blas/tile.h:337-349)
//   herk : cublasDsyrk           (blas/tile.h:293-304)
//   gemm : cublasDgemm           (blas/tile.h:249-261)
// issued in the loop order of Cholesky<Backend::GPU>::call_L (factorization/cholesky/impl.h:150-189) onto a pool of
// streams with one event per tile as the dependency carrier (what the pika sender DAG + per-tile pipelines amount to
// on the GPU backend, minus the host-side task overheads). Usage: cublas_tile_potrf_ref N nb [streams]
#include <cublas_v2.h>
#include <cuda_runtime.h>
#include <cusolverDn.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { auto e_ = (x); if (e_ != 0) { std::printf("error %d at %s:%d\n", (int)e_, __FILE__, __LINE__); return 1; } } while (0)

__global__ void fill_spd(double* a, long n) {
  const long j = blockIdx.x;
  for (long i = threadIdx.x; i < n; i += blockDim.x) {
    unsigned long long h = (unsigned long long)(i < j ? i * n + j : j * n + i) * 6364136223846793005ULL + 1442695040888963407ULL;
    h ^= h >> 33;
    double v = (double)(h >> 11) / 9007199254740992.0 * 2.0 - 1.0;
    a[i + j * n] = (i == j) ? v + 2.0 * n : v;
  }
}

int main(int argc, char** argv) {
  const long n = argc > 1 ? atol(argv[1]) : 16384;
  const int nb = argc > 2 ? atoi(argv[2]) : 512;
  const int ns = argc > 3 ? atoi(argv[3]) : 16;
  const int nt = (int)(n / nb);
  if (n % nb) { std::printf("N must be a multiple of nb\n"); return 1; }
  double* a;
  CK(cudaMalloc(&a, sizeof(double) * n * n));
  double* a0;
  CK(cudaMalloc(&a0, sizeof(double) * n * n));
  fill_spd<<<(unsigned)n, 256>>>(a0, n);
  std::vector<cudaStream_t> st(ns);
  std::vector<cublasHandle_t> bh(ns);
  std::vector<cusolverDnHandle_t> sh(ns);
  std::vector<double*> work(ns);
  std::vector<int*> dinfo(ns);
  int lwork = 0;
  for (int s = 0; s < ns; ++s) {
    CK(cudaStreamCreateWithFlags(&st[s], cudaStreamNonBlocking));
    CK(cublasCreate(&bh[s]));
    CK(cublasSetStream(bh[s], st[s]));
    CK(cusolverDnCreate(&sh[s]));
    CK(cusolverDnSetStream(sh[s], st[s]));
    if (s == 0) CK(cusolverDnDpotrf_bufferSize(sh[0], CUBLAS_FILL_MODE_LOWER, nb, a, (int)n, &lwork));
    CK(cudaMalloc(&work[s], sizeof(double) * (lwork > 0 ? lwork : 1)));
    CK(cudaMalloc(&dinfo[s], sizeof(int)));
  }
  std::vector<cudaEvent_t> ev((size_t)nt * nt);
  for (auto& e : ev) CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  auto tile = [&](int i, int j) { return a + (long)i * nb + (long)j * nb * n; };
  auto E = [&](int i, int j) -> cudaEvent_t& { return ev[(size_t)i + (size_t)j * nt]; };
  cudaEvent_t t0, t1;
  CK(cudaEventCreate(&t0));
  CK(cudaEventCreate(&t1));
  const double one = 1.0, mone = -1.0;
  double best = 1e30;
  for (int rep = 0; rep < 3; ++rep) {
    CK(cudaMemcpy(a, a0, sizeof(double) * n * n, cudaMemcpyDeviceToDevice));
    CK(cudaDeviceSynchronize());
    int rr = 0;
    std::vector<char> written((size_t)nt * nt, 0);
    CK(cudaEventRecord(t0, st[0]));
    for (int s = 1; s < ns; ++s) CK(cudaStreamWaitEvent(st[s], t0, 0));
    auto wait_tile = [&](int s, int i, int j) { if (written[(size_t)i + (size_t)j * nt]) cudaStreamWaitEvent(st[s], E(i, j), 0); };
    auto done_tile = [&](int s, int i, int j) { cudaEventRecord(E(i, j), st[s]); written[(size_t)i + (size_t)j * nt] = 1; };
    for (int k = 0; k < nt; ++k) {
      int s = rr++ % ns;
      wait_tile(s, k, k);
      CK(cusolverDnDpotrf(sh[s], CUBLAS_FILL_MODE_LOWER, nb, tile(k, k), (int)n, work[s], lwork, dinfo[s]));
      done_tile(s, k, k);
      for (int i = k + 1; i < nt; ++i) {
        s = rr++ % ns;
        wait_tile(s, k, k);
        wait_tile(s, i, k);
        CK(cublasDtrsm(bh[s], CUBLAS_SIDE_RIGHT, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_T, CUBLAS_DIAG_NON_UNIT, nb, nb, &one, tile(k, k),
                       (int)n, tile(i, k), (int)n));
        done_tile(s, i, k);
      }
      for (int j = k + 1; j < nt; ++j) {
        s = rr++ % ns;
        wait_tile(s, j, k);
        wait_tile(s, j, j);
        CK(cublasDsyrk(bh[s], CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, nb, nb, &mone, tile(j, k), (int)n, &one, tile(j, j), (int)n));
        done_tile(s, j, j);
        for (int i = j + 1; i < nt; ++i) {
          s = rr++ % ns;
          wait_tile(s, i, k);
          wait_tile(s, j, k);
          wait_tile(s, i, j);
          CK(cublasDgemm(bh[s], CUBLAS_OP_N, CUBLAS_OP_T, nb, nb, nb, &mone, tile(i, k), (int)n, tile(j, k), (int)n, &one, tile(i, j),
                         (int)n));
          done_tile(s, i, j);
        }
      }
    }
    for (int s = 1; s < ns; ++s) {
      cudaEvent_t e;
      CK(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
      CK(cudaEventRecord(e, st[s]));
      CK(cudaStreamWaitEvent(st[0], e, 0));
      CK(cudaEventDestroy(e));
    }
    CK(cudaEventRecord(t1, st[0]));
    CK(cudaEventSynchronize(t1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, t0, t1));
    int info = -1;
    CK(cudaMemcpy(&info, dinfo[0], sizeof(int), cudaMemcpyDeviceToHost));
    std::printf("rep %d: %.2f ms  %.1f GFLOP/s (info of last potrf on stream 0: %d)\n", rep, ms, (double)n * n * n / 3 / ms / 1e6, info);
    if (ms < best) best = ms;
  }
  std::printf("tile-schedule (cuBLAS/cuSOLVER per tile, %d streams) N %ld nb %d best: %.2f ms %.1f GFLOP/s\n", ns, n, nb, best,
              (double)n * n * n / 3 / best / 1e6);
  return 0;
}
