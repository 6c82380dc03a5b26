// GPU harness of the one-launch cluster kernel for the diagonal tile (potrf_tile_cluster.cu): factor and inverted
// 128-blocks against a host reference, untouched upper triangle, info on a non-SPD tile, timing and per-panel phase clocks.
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <random>
#include <vector>

#include "../dla-future_b200/csrc/common.h"
#include "../dla-future_b200/csrc/potrf_tile.cuh"

using namespace dlaf_b200;

int main() {
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> dist(-1, 1);
  int* dinfo;
  cudaMalloc(&dinfo, 4);
  bool all_ok = true;
  for (int nbp : {128, 256, 384, 512}) {
    const long ld = nbp + 6;
    std::vector<double> A(ld * nbp, -9.9), L(ld * nbp, 0.0);
    // SPD like the miniapp's tiles: U(-1,1) off-diagonal, diagonal + 2 n
    for (int j = 0; j < nbp; ++j)
      for (int i = j; i < nbp; ++i)
        A[i + j * ld] = (i == j) ? dist(rng) + 2.0 * nbp : dist(rng);
    // host reference (long double accumulation)
    for (int j = 0; j < nbp; ++j) {
      long double d = A[j + j * ld];
      for (int k = 0; k < j; ++k) d -= (long double)L[j + k * ld] * L[j + k * ld];
      L[j + j * ld] = (double)sqrtl(d);
      for (int i = j + 1; i < nbp; ++i) {
        long double v = A[i + j * ld];
        for (int k = 0; k < j; ++k) v -= (long double)L[i + k * ld] * L[j + k * ld];
        L[i + j * ld] = (double)(v / L[j + j * ld]);
      }
    }
    double *dT, *dW;
    const int ns = nbp / 128;
    cudaMalloc(&dT, A.size() * 8);
    cudaMalloc(&dW, (size_t)ns * 128 * 128 * 8);
    cudaMemcpy(dT, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dW, 0xff, (size_t)ns * 128 * 128 * 8);
    cudaMemset(dinfo, 0, 4);
    launch_potrf_tile_cluster_f64(dT, ld, dW, nbp, dinfo, 1000, 0);
    cudaError_t err = cudaDeviceSynchronize();
    if (err != cudaSuccess) { std::printf("nbp %d kernel FAILED: %s\n", nbp, cudaGetErrorString(err)); return 1; }
    std::vector<double> R(A.size()), Wh((size_t)ns * 128 * 128);
    int hinfo = -1;
    cudaMemcpy(R.data(), dT, R.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(Wh.data(), dW, Wh.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(&hinfo, dinfo, 4, cudaMemcpyDeviceToHost);
    double maxrel = 0; long upper_touched = 0;
    for (int j = 0; j < nbp; ++j)
      for (int i = 0; i < nbp; ++i) {
        if (i >= j) maxrel = std::fmax(maxrel, std::fabs(R[i + j * ld] - L[i + j * ld]) / std::fmax(std::fabs(L[i + j * ld]), 1e-3));
        else if (R[i + j * ld] != -9.9) ++upper_touched;
      }
    for (int i = nbp; i < ld; ++i) for (int j = 0; j < nbp; ++j) if (R[i + j * ld] != -9.9) ++upper_touched;
    double winv = 0, wupper = 0;
    for (int b = 0; b < ns; ++b) {
      const double* Wb = &Wh[(size_t)b * 128 * 128];
      for (int j = 0; j < 128; ++j)
        for (int i = 0; i < 128; ++i) {
          if (i < j) { wupper = std::fmax(wupper, std::fabs(Wb[i + j * 128])); continue; }
          long double s = 0;  // (W L)(i, j) = sum_k W(i,k) L(k,j), k in [j, i]
          for (int k = j; k <= i; ++k) s += (long double)Wb[i + k * 128] * R[(128 * b + k) + (long)(128 * b + j) * ld];
          winv = std::fmax(winv, std::fabs((double)s - (i == j ? 1.0 : 0.0)));
        }
    }
    const bool ok = hinfo == 0 && maxrel < 1e-12 && winv < 1e-12 && wupper == 0 && upper_touched == 0;
    all_ok &= ok;
    std::printf("nbp %d: info %d, max rel err of L %.3e, max |W L - I| %.3e, max |W upper| %.1e, untouched violations %ld  %s\n", nbp, hinfo,
                maxrel, winv, wupper, upper_touched, ok ? "OK" : "WRONG");
    // timing
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int reps = 50;
    float best = 1e9, total = 0;
    for (int rep = 0; rep < reps; ++rep) {
      cudaMemcpyAsync(dT, A.data(), A.size() * 8, cudaMemcpyHostToDevice, 0);
      cudaEventRecord(e0);
      launch_potrf_tile_cluster_f64(dT, ld, dW, nbp, dinfo, 0, 0);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      if (rep >= 5) { best = std::fmin(best, ms); total += ms; }
    }
    std::printf("nbp %d: cluster tile kernel %.1f us avg, %.1f us best (round 1: 4 block kernels + 6 GEMM launches = 342 us at 512)\n", nbp,
                total / (reps - 5) * 1000, best * 1000);
    if (nbp == 512) {
      long long* dtr; cudaMalloc(&dtr, 64 * 12 * 8); cudaMemset(dtr, 0, 64 * 12 * 8);
      potrf_tile_set_clock_trace(dtr);
      cudaMemcpy(dT, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
      launch_potrf_tile_cluster_f64(dT, ld, dW, nbp, dinfo, 0, 0);
      cudaDeviceSynchronize();
      potrf_tile_set_clock_trace(nullptr);
      std::vector<long long> tr(64 * 12);
      cudaMemcpy(tr.data(), dtr, tr.size() * 8, cudaMemcpyDeviceToHost);
      // stamps (CTA 0, thread 0): 0 step start | 1 own column updated + extracted | 2 factor done (thread 0 is the factor
      // thread when J % 16 == 0) | 3 barrier after the factor | 4 row solve + stores issued | 5 arrived | 6 rest updated | 7 waited + read back
      const char* nm[7] = {"upd+extract", "factor", "sync", "solve+stores", "arrive", "rest update", "wait+readback"};
      double sum_own[7] = {0}, sum_oth[7] = {0}; int nown = 0, noth = 0;
      for (int J = 0; J < 64; ++J) {
        const bool own = (J % 8 == 0);
        for (int q = 0; q < 7; ++q) {
          long long a = tr[J * 8 + q], b = tr[J * 8 + q + 1];
          if (!own && q < 4) continue;
          if (!own && q == 4) a = tr[J * 8 + 0];
          (own ? sum_own : sum_oth)[q] += double(b - a);
        }
        (own ? nown : noth)++;
      }
      std::printf("phase clocks, CTA 0 thread 0, whole phase 1 %.0f clk;  owned panels (avg of %d):", double(tr[63 * 8 + 7] - tr[0]), nown);
      for (int q = 0; q < 7; ++q) std::printf(" %s %.0f", nm[q], sum_own[q] / nown);
      std::printf(";  other panels (avg of %d):", noth);
      for (int q = 4; q < 7; ++q) std::printf(" %s %.0f", nm[q], sum_oth[q] / noth);
      std::printf("\n  owned panel J=0:");
      for (int q = 0; q < 7; ++q) std::printf(" %lld", tr[q + 1] - tr[q]);
      std::printf("   J=16:");
      for (int q = 0; q < 7; ++q) std::printf(" %lld", tr[16 * 8 + q + 1] - tr[16 * 8 + q]);
      std::printf("   J=8 (factor on another warp):");
      for (int q = 0; q < 7; ++q) std::printf(" %lld", tr[8 * 8 + q + 1] - tr[8 * 8 + q]);
      std::printf("\n  inside solve+stores (owned panels J=0,16,32,48: my solve | proxy fence | barrier): ");
      for (int J = 0; J < 64; J += 16)
        std::printf("[%lld | %lld | %lld] ", tr[512 + J * 4 + 0] - tr[J * 8 + 3], tr[512 + J * 4 + 1] - tr[512 + J * 4 + 0], tr[512 + J * 4 + 2] - tr[512 + J * 4 + 1]);
      std::printf("\n");
      cudaFree(dtr);
    }
    // non-SPD: diagonal matrix with one negative pivot
    if (nbp == 512 || nbp == 256) {
      std::vector<double> B(ld * nbp, 0.0);
      for (int j = 0; j < nbp; ++j) B[j + j * ld] = 4.0;
      const int bad = nbp - 57;
      B[bad + bad * ld] = -1.0;
      cudaMemcpy(dT, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
      cudaMemset(dinfo, 0, 4);
      launch_potrf_tile_cluster_f64(dT, ld, dW, nbp, dinfo, 7000, 0);
      cudaDeviceSynchronize();
      cudaMemcpy(&hinfo, dinfo, 4, cudaMemcpyDeviceToHost);
      std::printf("nbp %d non-SPD: info %d (expect %d) %s\n", nbp, hinfo, 7000 + bad + 1, hinfo == 7000 + bad + 1 ? "OK" : "WRONG");
      all_ok &= hinfo == 7000 + bad + 1;
    }
    cudaFree(dT); cudaFree(dW);
  }
  std::printf(all_ok ? "ALL OK\n" : "SOME WRONG\n");
  return all_ok ? 0 : 1;
}
