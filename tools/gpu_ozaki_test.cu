// GPU harness for the fp64-on-tcgen05 (int8 Ozaki scheme) trailing update: slicing exactness, GEMM correctness
// against a long-double host product (next to the error of the native DMMA GEMM), masks / row offsets, and
// timings against the DMMA kernel. Run on the GPU box:  tools/gpu_ozaki_test
#include <cuda_runtime.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <vector>

#include "../dla-future_b200/csrc/common.h"
#include "../dla-future_b200/csrc/gemm_dmma.cuh"
#include "../dla-future_b200/csrc/gemm_ozaki.h"

using namespace dlaf_b200;

static float time_ms(cudaEvent_t a, cudaEvent_t b) {
  float ms;
  cudaEventElapsedTime(&ms, a, b);
  return ms;
}

int main() {
  cudaDeviceProp prop;
  DLAF_CUDA_CHECK(cudaGetDeviceProperties(&prop, 0));
  std::printf("device %s sm_%d%d SMs %d\n", prop.name, prop.major, prop.minor, prop.multiProcessorCount);
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  std::mt19937_64 rng(11);
  std::uniform_real_distribution<double> dist(-1, 1);

  // ---- correctness
  for (int kdim : {512, 256}) {
    const int M = 256, N = 192, K = kdim;
    const int rowsA = M + 128, rowsB = N + 64;  // operands start at row offsets 128 / 64 of their split buffers
    const long lda = rowsA + 2, ldb = rowsB + 4, ldc = M + 6;
    for (int data = 0; data < 3; ++data) {
      std::vector<double> A(lda * K), B(ldb * K), C(ldc * N), R(ldc * N);
      for (auto& x : A) x = dist(rng);
      for (auto& x : B) x = dist(rng);
      if (data == 0) {  // single-digit operands: multiples of 1/64 in [-0.5, 0.5], row max exactly 0.5
        for (auto& x : A) x = std::nearbyint(x * 32) / 64.0;
        for (auto& x : B) x = std::nearbyint(x * 32) / 64.0;
        for (int r = 0; r < rowsA; ++r) A[r] = 0.5;
        for (int r = 0; r < rowsB; ++r) B[r] = -0.5;
      }
      if (data == 2) {  // wide dynamic range inside the rows
        std::uniform_int_distribution<int> ex(-20, 20);
        for (auto& x : A) x = std::ldexp(x, ex(rng));
        for (auto& x : B) x = std::ldexp(x, ex(rng));
      }
      for (auto& x : C) x = dist(rng);
      double *dA, *dB, *dC;
      cudaMalloc(&dA, A.size() * 8); cudaMalloc(&dB, B.size() * 8); cudaMalloc(&dC, C.size() * 8);
      cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
      cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
      OzakiSplit sa, sb;
      sa.allocate(rowsA, K);
      sb.allocate(rowsB, K);
      sa.split(dA, lda, rowsA, 0);
      sb.split(dB, ldb, rowsB, 0);
      DLAF_CUDA_CHECK(cudaDeviceSynchronize());
      if (data != 0 || kdim == 512) {
        // slicing: x == 2^e sum_t d_t 2^(-7-8t) up to 2^(e-56)
        std::vector<signed char> q((size_t)kOzakiSlices * rowsA * K);
        std::vector<double> sc(rowsA);
        cudaMemcpy(q.data(), sa.q, q.size(), cudaMemcpyDeviceToHost);
        cudaMemcpy(sc.data(), sa.scale, rowsA * 8, cudaMemcpyDeviceToHost);
        double worst = 0;
        int digit_max = 0;
        for (int r = 0; r < rowsA; ++r)
          for (int k = 0; k < K; ++k) {
            long double v = 0, w = 2;
            for (int t = 0; t < kOzakiSlices; ++t) {
              w /= 256;  // 2^(-7-8t)
              const int d = q[(size_t)t * rowsA * K + (size_t)r * K + k];
              digit_max = std::max(digit_max, std::abs(d));
              v += w * d;
            }
            worst = std::fmax(worst, (double)(fabsl(v * sc[r] - A[r + k * lda]) / sc[r]));
          }
        std::printf("split k=%d data %d: max |x - digits| / 2^e = %.3e (bound 2^-56 = %.3e), max |digit| %d\n", kdim, data,
                    worst, std::ldexp(1.0, -56), digit_max);
      }
      // reference in long double: C + alpha * A[128:,:] * B[64:,:]^T
      std::vector<long double> ref((size_t)M * N);
      std::vector<double> mag((size_t)M * N);
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < M; ++i) {
          long double s = 0, m = 0;
          for (int k = 0; k < K; ++k) {
            const long double p = (long double)A[128 + i + k * lda] * (long double)B[64 + j + k * ldb];
            s += p;
            m += fabsl(p);
          }
          ref[i + (size_t)j * M] = s;
          mag[i + (size_t)j * M] = (double)m;
        }
      for (int mode = 0; mode < 3; ++mode) {
        cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
        GemmArgs g{};
        g.A = dA + 128; g.lda = lda; g.B = dB + 64; g.ldb = ldb; g.C = dC; g.ldc = ldc;
        g.M = M; g.N = N; g.K = K; g.alpha = -1.0; g.beta = 1.0;
        g.mask = mode == 0 ? kMaskNone : kMaskLower; g.nbp = 128; g.P = g.Q = 1;
        const bool dmma = (mode == 2);
        if (dmma) { g.M = 256; g.N = 128; launch_gemm_nt_f64(g, 0); }  // native fp64 tensor-core GEMM on the same data
        else launch_gemm_ozaki_i8(g, sa, 128, sb, 64, 0);
        const cudaError_t err = cudaDeviceSynchronize();
        if (err != cudaSuccess) { std::printf("kernel FAILED: %s\n", cudaGetErrorString(err)); return 1; }
        cudaMemcpy(R.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
        double maxrel = 0, maxabs = 0;
        long wrong = 0;
        for (int j = 0; j < g.N; ++j)
          for (int i = 0; i < M; ++i) {
            const bool active = (g.mask == kMaskNone) || i >= j;
            if (active) {
              const long double want = (long double)C[i + j * ldc] - ref[i + (size_t)j * M];
              const double d = (double)fabsl(want - (long double)R[i + j * ldc]);
              maxabs = std::fmax(maxabs, d);
              maxrel = std::fmax(maxrel, d / (mag[i + (size_t)j * M] + 1e-300));
            }
            else if (R[i + j * ldc] != C[i + j * ldc]) ++wrong;
          }
        std::printf("k=%d data %d %s: max err %.3e, max err / sum|a||b| %.3e, masked elements modified %ld\n", kdim, data,
                    mode == 0 ? "ozaki full      " : (mode == 1 ? "ozaki lower mask" : "dmma  lower mask"), maxabs, maxrel, wrong);
      }
      sa.release(); sb.release();
      cudaFree(dA); cudaFree(dB); cudaFree(dC);
    }
  }

  // ---- guard: flag raised only for rows spanning too many binades; guarded kernels are mutually exclusive
  {
    const int n = 256, K = 512;
    std::vector<double> X((size_t)n * K);
    double* dX; int* dflag; double* dC;
    cudaMalloc(&dX, X.size() * 8); cudaMalloc(&dflag, 8); cudaMalloc(&dC, (size_t)n * n * 8);
    OzakiSplit sp; sp.allocate(n, K);
    for (int variant = 0; variant < 4; ++variant) {
      for (auto& x : X) x = dist(rng);
      if (variant == 1) X[17 + 33 * n] = std::ldexp(X[17 + 33 * n], -45);   // one entry 2^-45 below its row: rounded, < 16 bits kept
      if (variant == 2) X[17 + 33 * n] = std::ldexp(1.0, -50);               // tiny but EXACT (a power of two): no loss, no flag
      if (variant == 3) X[5 + 7 * n] = 0.0;                                  // zeros never raise the flag
      cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice);
      cudaMemset(dflag, 0, 8);
      sp.split(dX, n, n, 0, 0, 0, dflag);
      std::vector<double> C0((size_t)n * n, 1.0), C1(C0.size());
      cudaMemcpy(dC, C0.data(), C0.size() * 8, cudaMemcpyHostToDevice);
      GemmArgs g{};
      g.A = dX; g.lda = n; g.B = dX; g.ldb = n; g.C = dC; g.ldc = n; g.M = n; g.N = n; g.K = K; g.alpha = -1.0; g.beta = 1.0;
      g.mask = kMaskLower; g.nbp = 128; g.P = g.Q = 1;
      launch_gemm_ozaki_i8(g, sp, 0, sp, 0, 0, 0, dflag);
      DLAF_CUDA_CHECK(cudaDeviceSynchronize());
      cudaMemcpy(C1.data(), dC, C1.size() * 8, cudaMemcpyDeviceToHost);
      const bool oz_ran = C1[n - 1] != 1.0;
      launch_gemm_nt_f64_if(g, dflag, 0);
      DLAF_CUDA_CHECK(cudaDeviceSynchronize());
      std::vector<double> C2(C0.size());
      cudaMemcpy(C2.data(), dC, C2.size() * 8, cudaMemcpyDeviceToHost);
      const bool native_ran = C2[n - 1] != C1[n - 1];
      int hf = -1; cudaMemcpy(&hf, dflag, 4, cudaMemcpyDeviceToHost);
      std::printf("guard variant %d: flag %d, int8 kernel ran %d, native kernel ran %d  (expect %s)\n", variant, hf, (int)oz_ran,
                  (int)native_ran, variant == 1 ? "1 0 1" : "0 1 0");
    }
    sp.release(); cudaFree(dX); cudaFree(dflag); cudaFree(dC);
  }

  // ---- phase clocks of one launch (4096 x 4096 full update: 2048 tiles)
  {
    const int K = 512, n = 4096;
    double *dP, *dC;
    cudaMalloc(&dC, (size_t)n * n * 8); cudaMalloc(&dP, (size_t)n * K * 8);
    cudaMemset(dC, 0, (size_t)n * n * 8);
    std::vector<double> hp((size_t)n * K);
    for (auto& x : hp) x = dist(rng);
    cudaMemcpy(dP, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice);
    OzakiSplit sp; sp.allocate(n, K); sp.split(dP, n, n, 0);
    GemmArgs g{};
    g.C = dC; g.ldc = n; g.M = n; g.N = n; g.K = K; g.alpha = -1.0; g.beta = 1.0; g.mask = kMaskNone; g.nbp = 512; g.P = g.Q = 1;
    launch_gemm_ozaki_i8(g, sp, 0, sp, 0, 0);
    long long* dtr; cudaMalloc(&dtr, 4096 * 8 * 8); cudaMemset(dtr, 0, 4096 * 8 * 8);
    ozaki_set_clock_trace(dtr);
    launch_gemm_ozaki_i8(g, sp, 0, sp, 0, 0);
    DLAF_CUDA_CHECK(cudaDeviceSynchronize());
    ozaki_set_clock_trace(nullptr);
    std::vector<long long> tr(4096 * 8);
    cudaMemcpy(tr.data(), dtr, tr.size() * 8, cudaMemcpyDeviceToHost);
    const char* names[6] = {"setup", "first_stage_wait", "mma_issue", "mma_drain", "epilogue", "teardown"};
    for (int range = 0; range < 2; ++range) {  // first wave (cold) / later CTAs
      double sum[6] = {0, 0, 0, 0, 0, 0}, tot = 0; int cnt = 0;
      for (int c = range == 0 ? 0 : 148; c < (range == 0 ? 148 : 2048); ++c) {
        const long long* t = &tr[c * 8];
        if (t[6] == 0) continue;
        for (int q = 0; q < 6; ++q) sum[q] += (double)(t[q + 1] - t[q]);
        tot += (double)(t[6] - t[0]); ++cnt;
      }
      std::printf("ozaki phase clocks (%s, %d CTAs), avg clk:", range == 0 ? "first wave" : "later waves", cnt);
      for (int q = 0; q < 6; ++q) std::printf(" %s %.0f", names[q], sum[q] / cnt);
      std::printf(" | total %.0f clk = %.2f us @1.965 GHz\n", tot / cnt, tot / cnt / 1965.0);
    }
    sp.release(); cudaFree(dP); cudaFree(dC); cudaFree(dtr);
  }

  // ---- timing
  {
    const int K = 512;
    for (int n : {4096, 16384, 32256}) {
      double *dP, *dC;
      if (cudaMalloc(&dC, (size_t)n * n * 8) != cudaSuccess) { std::printf("skip %d (alloc)\n", n); continue; }
      cudaMalloc(&dP, (size_t)n * K * 8);
      cudaMemset(dC, 0, (size_t)n * n * 8);
      std::vector<double> hp((size_t)n * K);
      for (auto& x : hp) x = dist(rng);
      cudaMemcpy(dP, hp.data(), hp.size() * 8, cudaMemcpyHostToDevice);
      OzakiSplit sp;
      sp.allocate(n, K);
      sp.split(dP, n, n, 0);
      cudaEventRecord(e0);
      for (int i = 0; i < 5; ++i) sp.split(dP, n, n, 0);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      std::printf("split_i8 %d x %d: %.1f us\n", n, K, time_ms(e0, e1) * 200.0);
      for (int mask = 0; mask < 2; ++mask) {
        GemmArgs g{};
        g.A = dP; g.lda = n; g.B = dP; g.ldb = n; g.C = dC; g.ldc = n;
        g.M = n; g.N = n; g.K = K; g.alpha = -1.0; g.beta = 1.0;
        g.mask = mask ? kMaskLower : kMaskNone; g.nbp = 512; g.P = g.Q = 1;
        const double fl = (mask ? 1.0 : 2.0) * (double)n * n * K;
        for (int variant = 0; variant < 2; ++variant) {
          auto run = [&] {
            if (variant == 0) launch_gemm_ozaki_i8(g, sp, 0, sp, 0, 0);
            else launch_gemm_nt_f64(g, 0);
          };
          run();
          cudaEventRecord(e0);
          for (int i = 0; i < 3; ++i) run();
          cudaEventRecord(e1);
          const cudaError_t err = cudaEventSynchronize(e1);
          if (err != cudaSuccess) { std::printf("timing FAILED: %s\n", cudaGetErrorString(err)); return 1; }
          const double ms = time_ms(e0, e1) / 3;
          std::printf("%s %dx%dx%d mask %d: %.3f ms  %.2f TFLOP/s (fp64-equivalent%s)\n", variant == 0 ? "ozaki_i8" : "dmma    ",
                      n, n, K, mask, ms, fl / ms / 1e9, variant == 0 ? "; x28 int8 TOP/s" : "");
        }
      }
      sp.release();
      cudaFree(dP); cudaFree(dC);
    }
  }
  std::printf("done\n");
  return 0;
}
