// GPU harness for the critical-path ("chain") kernels of the POTRF path: the fused panel TRSM
// (trsm_fused_f64_kernel, gemm_dmma.cuh) against a host substitution and against the 2*ns-1 separate GEMM
// launches it replaces. Run on the GPU box:  tools/gpu_chain_test
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../dla-future_b200/csrc/common.h"
#include "../dla-future_b200/csrc/gemm_dmma.cuh"

using namespace dlaf_b200;

static float time_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

// the launch sequence the engine used before (engine.cu: trsm_panel, DLAF_B200_TRSM=steps)
static void trsm_steps(double* b, long ldb, int m, const double* t, long ldt, const double* w, int ns, cudaStream_t st) {
  const int G = 128;
  for (int j = 0; j < ns; ++j) {
    double* bj = b + (long)j * G * ldb;
    if (j > 0) {
      GemmArgs a{};
      a.A = b; a.lda = ldb; a.B = t + (long)j * G; a.ldb = ldt; a.C = bj; a.ldc = ldb;
      a.M = m; a.N = G; a.K = j * G; a.alpha = -1.0; a.beta = 1.0; a.mask = kMaskNone; a.nbp = ns * G; a.P = a.Q = 1;
      launch_gemm_nt_f64(a, st);
    }
    GemmArgs s{};
    s.A = bj; s.lda = ldb; s.B = w + (long)j * G * G; s.ldb = G; s.C = bj; s.ldc = ldb;
    s.M = m; s.N = G; s.K = G; s.alpha = 1.0; s.beta = 0.0; s.mask = kMaskNone; s.nbp = ns * G; s.P = s.Q = 1;
    launch_gemm_nt_f64(s, st);
  }
}

int main() {
  cudaDeviceProp prop;
  DLAF_CUDA_CHECK(cudaGetDeviceProperties(&prop, 0));
  std::printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  std::mt19937_64 rng(7);
  std::uniform_real_distribution<double> dist(-1, 1);
  const int G = 128;

  // ---- correctness: ns = 1..4, ragged-looking leading dimensions, sentinel rows around B
  for (int ns = 1; ns <= 4; ++ns) {
    const int n = ns * G, m = 160;
    const long ldt = n + 2, ldb = m + 6;
    std::vector<double> L(ldt * n, 0.0), W((size_t)ns * G * G, 0.0), B(ldb * n), X(ldb * n), R(ldb * n);
    for (int c = 0; c < n; ++c)
      for (int r = 0; r < n; ++r)
        L[r + c * ldt] = (r > c) ? 0.25 * dist(rng) : (r == c ? 2.0 + dist(rng) * 0.5 : -7.7 /* never read */);
    // W_j = inv(L_jj), full square with zero upper part (what potrf_inv writes)
    for (int j = 0; j < ns; ++j) {
      double* w = W.data() + (size_t)j * G * G;
      for (int c = 0; c < G; ++c) {
        // column c of the inverse by forward substitution
        for (int r = 0; r < G; ++r) {
          double s = (r == c) ? 1.0 : 0.0;
          for (int q = c; q < r; ++q) s -= L[(j * G + r) + (long)(j * G + q) * ldt] * w[q + c * G];
          w[r + c * G] = (r >= c) ? s / L[(j * G + r) + (long)(j * G + r) * ldt] : 0.0;
        }
      }
    }
    for (auto& x : B) x = dist(rng);
    // host reference: x L^T = b  per row
    for (int r = 0; r < m; ++r)
      for (int c = 0; c < n; ++c) {
        double s = B[r + c * ldb];
        for (int q = 0; q < c; ++q) s -= X[r + q * ldb] * L[c + (long)q * ldt];
        X[r + c * ldb] = s / L[c + (long)c * ldt];
      }
    double *dL, *dW, *dB;
    cudaMalloc(&dL, L.size() * 8); cudaMalloc(&dW, W.size() * 8); cudaMalloc(&dB, B.size() * 8);
    cudaMemcpy(dL, L.data(), L.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dW, W.data(), W.size() * 8, cudaMemcpyHostToDevice);
    for (int variant = 0; variant < 2; ++variant) {
      cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
      if (variant == 0) {
        TrsmFusedArgs a{dB, ldb, dL, ldt, dW, ns};
        launch_trsm_fused_f64(a, m, 0);
      }
      else {
        // steps variant needs M % 128 == 0: run on the first 128 rows only
        trsm_steps(dB, ldb, 128, dL, ldt, dW, ns, 0);
      }
      DLAF_CUDA_CHECK(cudaDeviceSynchronize());
      cudaMemcpy(R.data(), dB, B.size() * 8, cudaMemcpyDeviceToHost);
      const int mchk = variant == 0 ? m : 128;
      double maxerr = 0, maxref = 0;
      long pad_touched = 0;
      for (int c = 0; c < n; ++c) {
        for (int r = 0; r < mchk; ++r) {
          maxerr = std::fmax(maxerr, std::fabs(R[r + c * ldb] - X[r + c * ldb]));
          maxref = std::fmax(maxref, std::fabs(X[r + c * ldb]));
        }
        for (int r = m; r < ldb; ++r)
          if (R[r + c * ldb] != B[r + c * ldb]) ++pad_touched;
      }
      std::printf("trsm %s ns=%d m=%d: max|X-ref| %.3e (max|ref| %.2e), padding rows modified %ld\n",
                  variant == 0 ? "fused" : "steps", ns, mchk, maxerr, maxref, pad_touched);
    }
    cudaFree(dL); cudaFree(dW); cudaFree(dB);
  }

  // ---- timing: fused vs steps, nb = 512
  {
    const int ns = 4, n = ns * G;
    double *dL, *dW, *dB;
    const int mmax = 32768;
    cudaMalloc(&dL, (size_t)n * n * 8); cudaMalloc(&dW, (size_t)ns * G * G * 8); cudaMalloc(&dB, (size_t)mmax * n * 8);
    cudaMemset(dL, 0, (size_t)n * n * 8); cudaMemset(dW, 0, (size_t)ns * G * G * 8); cudaMemset(dB, 0, (size_t)mmax * n * 8);
    for (int m : {512, 2048, 8192, 32768}) {
      for (int variant = 0; variant < 2; ++variant) {
        auto run = [&] {
          if (variant == 0) {
            TrsmFusedArgs a{dB, (long)m, dL, (long)n, dW, ns};
            launch_trsm_fused_f64(a, m, 0);
          }
          else
            trsm_steps(dB, m, m, dL, n, dW, ns, 0);
        };
        run();
        cudaEventRecord(e0);
        for (int i = 0; i < 10; ++i) run();
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        const double us = time_ms(e0, e1) * 100.0;
        const double fl = 2.0 * m * 128.0 * 128.0 * 10.0;  // (1+2+3+4) K-blocks of 128 per 128 columns
        std::printf("trsm %s m=%5d nb=512: %8.1f us  %.2f TFLOP/s\n", variant == 0 ? "fused" : "steps", m, us, fl / us / 1e6);
      }
    }
    cudaFree(dL); cudaFree(dW); cudaFree(dB);
  }
  std::printf("done\n");
  return 0;
}
