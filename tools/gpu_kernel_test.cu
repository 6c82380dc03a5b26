// Standalone GPU harness for the hot kernels: peak microbenchmarks (DMMA / DFMA), correctness of the
// DMMA GEMM (plain, masked, in-place) and of the 128-block potrf+inverse against host loops, and
// timings next to cuBLAS DGEMM on the same shapes. Run on the GPU box:  tools/gpu_kernel_test
#include <cublas_v2.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <cstdio>
#include <random>
#include <vector>

#include "../dla-future_b200/csrc/common.h"
#include "../dla-future_b200/csrc/gemm_dmma.cuh"
#include "../dla-future_b200/csrc/potrf_tile.cuh"
#include "../dla-future_b200/csrc/gemm_tf32.h"

using namespace dlaf_b200;

__global__ void dmma_peak_kernel(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = threadIdx.x * 2e-3;
  double c[16][2];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i][0] = c[i][1] = 0.0;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i)
      asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                   : "+d"(c[i][0]), "+d"(c[i][1])
                   : "d"(a), "d"(b));
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i][0] + c[i][1];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

__global__ void dfma_peak_kernel(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0000001;
  double c[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) c[i] = i;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) c[i] = fma(c[i], b, a);
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += c[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

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
  const int nsm = prop.multiProcessorCount;

  // ---- peaks
  {
    double* out;
    cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024);
    for (int threads : {128, 256, 512, 1024}) {
      const int iters = 20000;
      dmma_peak_kernel<<<nsm, threads>>>(out, 100);
      cudaEventRecord(e0);
      dmma_peak_kernel<<<nsm, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      double fl = 2.0 * 256 * 16 * double(iters) * (threads / 32) * nsm;
      std::printf("DMMA peak  %4d thr/SM: %.2f TFLOP/s\n", threads, fl / time_ms(e0, e1) / 1e9);
      dfma_peak_kernel<<<nsm, threads>>>(out, 100);
      cudaEventRecord(e0);
      dfma_peak_kernel<<<nsm, threads>>>(out, iters);
      cudaEventRecord(e1);
      cudaEventSynchronize(e1);
      fl = 2.0 * 16 * double(iters) * threads * nsm;
      std::printf("DFMA peak  %4d thr/SM: %.2f TFLOP/s\n", threads, fl / time_ms(e0, e1) / 1e9);
    }
    cudaFree(out);
  }

  std::mt19937_64 rng(42);
  std::uniform_real_distribution<double> dist(-1, 1);

  // ---- GEMM correctness
  {
    const int M = 384, N = 256, K = 144;  // K multiple of 16
    const long lda = M + 6, ldb = N + 2, ldc = M + 10;
    std::vector<double> A(lda * K), B(ldb * K), C(ldc * N), R(ldc * N);
    for (auto& x : A) x = dist(rng);
    for (auto& x : B) x = dist(rng);
    for (auto& x : C) x = dist(rng);
    double *dA, *dB, *dC;
    cudaMalloc(&dA, A.size() * 8);
    cudaMalloc(&dB, B.size() * 8);
    cudaMalloc(&dC, C.size() * 8);
    cudaMemcpy(dA, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemcpy(dB, B.data(), B.size() * 8, cudaMemcpyHostToDevice);
    for (int mode = 0; mode < 3; ++mode) {
      cudaMemcpy(dC, C.data(), C.size() * 8, cudaMemcpyHostToDevice);
      GemmArgs g{};
      g.A = dA; g.lda = lda; g.B = dB; g.ldb = ldb; g.C = dC; g.ldc = ldc;
      g.M = M; g.N = N; g.K = K;
      g.alpha = -1.0; g.beta = 1.0;
      g.mask = kMaskNone; g.nbp = 128; g.P = g.Q = 1;
      if (mode == 1) { g.mask = kMaskLower; }
      if (mode == 2) { g.mask = kMaskLower; g.P = 2; g.Q = 3; g.prow = 1; g.pcol = 0; g.ti0 = 0; g.tj0 = 0; g.alpha = 0.5; g.beta = 0.0; }
      launch_gemm_nt_f64(g, 0);
      DLAF_CUDA_CHECK(cudaDeviceSynchronize());
      cudaMemcpy(R.data(), dC, C.size() * 8, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      long touched_wrong = 0;
      for (int j = 0; j < N; ++j)
        for (int i = 0; i < M; ++i) {
          long gi = i, gj = j;
          if (mode == 2) { gi = (long)(i / 128) * g.P * 128 + g.prow * 128 + i % 128; gj = (long)(j / 128) * g.Q * 128 + g.pcol * 128 + j % 128; }
          bool active = (g.mask == kMaskNone) || gi >= gj;
          double ref = C[i + j * ldc];
          if (active) {
            double s = 0;
            for (int k = 0; k < K; ++k) s += A[i + k * lda] * B[j + k * ldb];
            ref = g.alpha * s + (g.beta != 0 ? g.beta * C[i + j * ldc] : 0.0);
            maxerr = std::fmax(maxerr, std::fabs(ref - R[i + j * ldc]));
          }
          else if (R[i + j * ldc] != C[i + j * ldc]) touched_wrong++;
        }
      std::printf("GEMM correctness mode %d: max err %.3e, masked elements modified %ld\n", mode, maxerr, touched_wrong);
    }
    // in-place: C aliases A, N = 128, K = 128
    {
      const int M2 = 256, K2 = 128;
      std::vector<double> X(M2 * K2), Wm(128 * 128), Out(M2 * K2);
      for (auto& x : X) x = dist(rng);
      for (auto& x : Wm) x = dist(rng);
      double *dX, *dW;
      cudaMalloc(&dX, X.size() * 8);
      cudaMalloc(&dW, Wm.size() * 8);
      cudaMemcpy(dX, X.data(), X.size() * 8, cudaMemcpyHostToDevice);
      cudaMemcpy(dW, Wm.data(), Wm.size() * 8, cudaMemcpyHostToDevice);
      GemmArgs g{};
      g.A = dX; g.lda = M2; g.B = dW; g.ldb = 128; g.C = dX; g.ldc = M2;
      g.M = M2; g.N = 128; g.K = K2; g.alpha = 1.0; g.beta = 0.0; g.mask = kMaskNone; g.nbp = 128; g.P = g.Q = 1;
      launch_gemm_nt_f64(g, 0);
      DLAF_CUDA_CHECK(cudaDeviceSynchronize());
      cudaMemcpy(Out.data(), dX, X.size() * 8, cudaMemcpyDeviceToHost);
      double maxerr = 0;
      for (int j = 0; j < 128; ++j)
        for (int i = 0; i < M2; ++i) {
          double s = 0;
          for (int k = 0; k < K2; ++k) s += X[i + k * M2] * Wm[j + k * 128];
          maxerr = std::fmax(maxerr, std::fabs(s - Out[i + j * M2]));
        }
      std::printf("GEMM in-place (C aliases A): max err %.3e\n", maxerr);
      cudaFree(dX); cudaFree(dW);
    }
    cudaFree(dA); cudaFree(dB); cudaFree(dC);
  }

  // ---- potrf128 + inverse
  {
    const int n = 128; const long ld = 200;
    std::vector<double> X(n * n), A(ld * n, -9.9), L(n * n, 0.0), R(ld * n), Wh(n * n);
    for (auto& x : X) x = dist(rng);
    for (int j = 0; j < n; ++j)
      for (int i = j; i < n; ++i) {
        double s = (i == j) ? n : 0.0;
        for (int k = 0; k < n; ++k) s += X[i + k * n] * X[j + k * n];
        A[i + j * ld] = s;
      }
    // host Cholesky
    std::vector<double> H(n * n, 0.0);
    for (int j = 0; j < n; ++j) for (int i = j; i < n; ++i) H[i + j * n] = A[i + j * ld];
    for (int j = 0; j < n; ++j) {
      double d = std::sqrt(H[j + j * n]); H[j + j * n] = d;
      for (int i = j + 1; i < n; ++i) H[i + j * n] /= d;
      for (int s = j + 1; s < n; ++s) for (int i = s; i < n; ++i) H[i + s * n] -= H[i + j * n] * H[s + j * n];
    }
    double *dT, *dW; int* dinfo;
    cudaMalloc(&dT, A.size() * 8); cudaMalloc(&dW, n * n * 8); cudaMalloc(&dinfo, 4);
    cudaMemset(dinfo, 0, 4);
    cudaMemcpy(dT, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    launch_potrf128_inv_f64(dT, ld, dW, n, dinfo, 0, 0);
    DLAF_CUDA_CHECK(cudaDeviceSynchronize());
    cudaMemcpy(R.data(), dT, A.size() * 8, cudaMemcpyDeviceToHost);
    cudaMemcpy(Wh.data(), dW, n * n * 8, cudaMemcpyDeviceToHost);
    int info; cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
    double errL = 0, errI = 0; long sentinel_bad = 0;
    for (int j = 0; j < n; ++j)
      for (int i = 0; i < n; ++i) {
        if (i >= j) errL = std::fmax(errL, std::fabs(R[i + j * ld] - H[i + j * n]));
        else if (R[i + j * ld] != -9.9) sentinel_bad++;
        double s = 0;  // (W * L)(i,j)
        for (int k = 0; k < n; ++k) s += Wh[i + k * n] * ((k >= j) ? H[k + j * n] : 0.0);
        errI = std::fmax(errI, std::fabs(s - (i == j ? 1.0 : 0.0)));
      }
    std::printf("potrf128: info %d, max|L-ref| %.3e, max|W*L-I| %.3e, upper-triangle writes %ld\n", info, errL, errI, sentinel_bad);
    // non-SPD -> info
    std::vector<double> Z(ld * n, 0.0);
    cudaMemcpy(dT, Z.data(), Z.size() * 8, cudaMemcpyHostToDevice);
    launch_potrf128_inv_f64(dT, ld, dW, n, dinfo, 0, 0);
    cudaDeviceSynchronize();
    cudaMemcpy(&info, dinfo, 4, cudaMemcpyDeviceToHost);
    std::printf("potrf128 zero matrix: info %d (expect 1)\n", info);
    // per-phase clock trace (thread 0 and thread 255)
    {
      cudaMemcpy(dT, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
      long long* dtr; cudaMalloc(&dtr, 16 * 8 * 2 * 8); cudaMemset(dtr, 0, 16 * 8 * 2 * 8);
      potrf_set_clock_trace(dtr);
      launch_potrf128_inv_f64(dT, ld, dW, n, dinfo, 0, 0);
      cudaDeviceSynchronize();
      potrf_set_clock_trace(nullptr);
      std::vector<long long> tr(16 * 8 * 2); cudaMemcpy(tr.data(), dtr, tr.size() * 8, cudaMemcpyDeviceToHost);
      const char* names[7] = {"write_panel", "bar1", "solve", "bar2", "update", "factor", "bar3"};
      for (int th = 0; th < 2; ++th) {
        long long sum[7] = {0};
        for (int J = 0; J < 16; ++J) for (int q = 0; q < 7; ++q) sum[q] += tr[(J * 8 + q + 1) * 2 + th] - tr[(J * 8 + q) * 2 + th];
        std::printf("potrf128 phase clocks thread %3d (sum over 16 steps):", th ? 255 : 0);
        for (int q = 0; q < 7; ++q) std::printf(" %s %lld", names[q], sum[q]);
        std::printf(" | total %lld\n", tr[(15 * 8 + 7) * 2 + th] - tr[th]);
      }
      for (int J : {0, 4, 8, 12, 15}) {
        std::printf("  step %2d thread0:", J);
        for (int q = 0; q < 7; ++q) std::printf(" %lld", tr[(J * 8 + q + 1) * 2] - tr[(J * 8 + q) * 2]);
        std::printf("   thread255:");
        for (int q = 0; q < 7; ++q) std::printf(" %lld", tr[(J * 8 + q + 1) * 2 + 1] - tr[(J * 8 + q) * 2 + 1]);
        std::printf("\n");
      }
      cudaFree(dtr);
    }
    // timing
    cudaMemcpy(dT, A.data(), A.size() * 8, cudaMemcpyHostToDevice);
    cudaMemset(dinfo, 0, 4);
    launch_potrf128_inv_f64(dT, ld, dW, n, dinfo, 0, 0);
    cudaEventRecord(e0);
    for (int i = 0; i < 20; ++i) launch_potrf128_inv_f64(dT, ld, dW, n, dinfo, 0, 0);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    std::printf("potrf128+inv: %.1f us per call\n", time_ms(e0, e1) * 1000 / 20);
    cudaFree(dT); cudaFree(dW); cudaFree(dinfo);
  }

  // ---- GEMM timing vs cuBLAS
  {
    cublasHandle_t h; cublasCreate(&h);
    struct Shape { int M, N, K; int mask; };
    Shape shapes[] = {{16384, 16384, 512, 0}, {16384, 16384, 512, 1}, {32256, 32256, 512, 1}, {8192, 8192, 512, 0}, {4096, 128, 512, 0}, {16384, 512, 512, 0}};
    for (auto s : shapes) {
      double *dP, *dC;
      size_t cbytes = (size_t)s.M * s.N * 8;
      if (cudaMalloc(&dC, cbytes) != cudaSuccess) { std::printf("skip %d (alloc)\n", s.M); continue; }
      cudaMalloc(&dP, (size_t)s.M * s.K * 8);
      cudaMemset(dC, 0, cbytes); cudaMemset(dP, 0, (size_t)s.M * s.K * 8);
      GemmArgs g{};
      g.A = dP; g.lda = s.M; g.B = dP; g.ldb = s.M; g.C = dC; g.ldc = s.M; g.M = s.M; g.N = s.N; g.K = s.K;
      g.alpha = -1; g.beta = 1; g.mask = s.mask; g.nbp = 512; g.P = g.Q = 1;
      launch_gemm_nt_f64(g, 0);
      cudaDeviceSynchronize();
      const int reps = 3;
      cudaEventRecord(e0);
      for (int i = 0; i < reps; ++i) launch_gemm_nt_f64(g, 0);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      DLAF_CUDA_CHECK(cudaGetLastError());
      double ms = time_ms(e0, e1) / reps;
      double fl = 2.0 * s.M * (double)s.N * s.K * (s.mask ? 0.5 * (1.0 + 128.0 / s.N) : 1.0);
      std::printf("gemm_nt_f64 %dx%dx%d mask %d: %.3f ms  %.2f TFLOP/s\n", s.M, s.N, s.K, s.mask, ms, fl / ms / 1e9);
      for (int cfg : {1, 5, 6}) {
        if (s.M * (long)s.N < 8192L * 8192L) break;
        g.dbg_stagger_ns = 0;
        if (cfg == 101) { g.dbg_stagger_ns = 30000; cfg = 1; }
        launch_gemm_nt_f64_cfg(g, cfg, 0);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i) launch_gemm_nt_f64_cfg(g, cfg, 0);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        DLAF_CUDA_CHECK(cudaGetLastError());
        double msc = time_ms(e0, e1) / reps;
        std::printf("   cfg %d stagger %d ns: %.3f ms  %.2f TFLOP/s\n", cfg, g.dbg_stagger_ns, msc, fl / msc / 1e9);
        g.dbg_stagger_ns = 0;
      }
      if (!s.mask) {
        const double al = -1, be = 1;
        cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, s.M, s.N, s.K, &al, dP, s.M, dP, s.M, &be, dC, s.M);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i)
          cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, s.M, s.N, s.K, &al, dP, s.M, dP, s.M, &be, dC, s.M);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        ms = time_ms(e0, e1) / reps;
        std::printf("cublasDgemm  %dx%dx%d       : %.3f ms  %.2f TFLOP/s\n", s.M, s.N, s.K, ms, 2.0 * s.M * (double)s.N * s.K / ms / 1e9);
      } else {
        const double al = -1, be = 1;
        cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, s.M, s.K, &al, dP, s.M, &be, dC, s.M);
        cudaEventRecord(e0);
        for (int i = 0; i < reps; ++i)
          cublasDsyrk(h, CUBLAS_FILL_MODE_LOWER, CUBLAS_OP_N, s.M, s.K, &al, dP, s.M, &be, dC, s.M);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        ms = time_ms(e0, e1) / reps;
        std::printf("cublasDsyrk  %dx%d           : %.3f ms  %.2f TFLOP/s\n", s.M, s.K, ms, (double)s.M * s.M * s.K / ms / 1e9);
      }
      cudaFree(dC); cudaFree(dP);
    }
    // big square DGEMM as the measured fp64 tensor peak
    {
      const int n = 8192;
      double *a, *b, *c;
      cudaMalloc(&a, (size_t)n * n * 8); cudaMalloc(&b, (size_t)n * n * 8); cudaMalloc(&c, (size_t)n * n * 8);
      cudaMemset(a, 0, (size_t)n * n * 8); cudaMemset(b, 0, (size_t)n * n * 8); cudaMemset(c, 0, (size_t)n * n * 8);
      const double al = 1, be = 0;
      cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, a, n, b, n, &be, c, n);
      cudaEventRecord(e0);
      for (int i = 0; i < 3; ++i) cublasDgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, n, n, n, &al, a, n, b, n, &be, c, n);
      cudaEventRecord(e1); cudaEventSynchronize(e1);
      std::printf("cublasDgemm 8192^3: %.2f TFLOP/s\n", 2.0 * n * (double)n * n / (time_ms(e0, e1) / 3) / 1e9);
      cudaFree(a); cudaFree(b); cudaFree(c);
    }
    cublasDestroy(h);
  }
  // ---- tcgen05 3xTF32 GEMM (fp32)
  if (std::getenv("SKIP_TF32") == nullptr) {
    std::mt19937_64 rng2(7);
    std::uniform_real_distribution<float> df(-1.f, 1.f);
    for (int mode = 0; mode < 2; ++mode) {
      const int M = 384, N = 256, K = 96;
      const long lda = M + 4, ldc = M + 8;
      std::vector<float> P(lda * K), C(ldc * N), R(ldc * N);
      for (auto& x : P) x = df(rng2);
      for (auto& x : C) x = df(rng2);
      float *dP, *dC;
      cudaMalloc(&dP, P.size() * 4); cudaMalloc(&dC, C.size() * 4);
      cudaMemcpy(dP, P.data(), P.size() * 4, cudaMemcpyHostToDevice);
      cudaMemcpy(dC, C.data(), C.size() * 4, cudaMemcpyHostToDevice);
      Tf32Split sp; sp.allocate(M, K);
      sp.split(dP, lda, M, 0);
      GemmArgsT<float> g{};
      g.C = dC; g.ldc = ldc; g.M = M; g.N = N; g.K = K; g.alpha = -1.0; g.beta = 1.0;
      g.mask = mode ? kMaskLower : kMaskNone; g.nbp = 128; g.P = g.Q = 1;
      launch_gemm_tf32x3(g, sp, 0, sp, 0, 0);   // C -= P[0:M] * P[0:N]^T
      cudaError_t e = cudaDeviceSynchronize();
      if (e != cudaSuccess) { std::printf("tf32x3 gemm FAILED: %s\n", cudaGetErrorString(e)); return 1; }
      cudaMemcpy(R.data(), dC, C.size() * 4, cudaMemcpyDeviceToHost);
      double maxerr = 0, e1 = 0, e2a = 0, e2b = 0, e3 = 0; long wrong = 0;
      auto trunc = [](float v) { uint32_t u; memcpy(&u, &v, 4); u &= 0xFFFFE000u; float r; memcpy(&r, &u, 4); return r; };
      for (int j = 0; j < N; ++j) for (int i = 0; i < M; ++i) {
        const bool active = !mode || i >= j;
        if (active) {
          double s = 0, s1 = 0, s2a = 0, s2b = 0, s3 = 0;
          for (int k = 0; k < K; ++k) {
            const float a = P[i + k * lda], b = P[j + k * lda];
            const float ah = trunc(a), al = trunc(a - ah), bh = trunc(b), bl = trunc(b - bh);
            s += (double)a * b; s1 += (double)ah * bh; s2a += (double)ah * bh + (double)ah * bl; s2b += (double)ah * bh + (double)al * bh;
            s3 += (double)ah * bh + (double)ah * bl + (double)al * bh;
          }
          const double got = (double)C[i + j * ldc] - (double)R[i + j * ldc];  // = computed product
          maxerr = std::fmax(maxerr, std::fabs(got - s)); e1 = std::fmax(e1, std::fabs(got - s1));
          e2a = std::fmax(e2a, std::fabs(got - s2a)); e2b = std::fmax(e2b, std::fabs(got - s2b)); e3 = std::fmax(e3, std::fabs(got - s3));
        } else if (R[i + j * ldc] != C[i + j * ldc]) wrong++;
      }
      std::printf("tf32x3 tcgen05 GEMM mode %d: max err vs exact %.3e | vs hi*hi %.3e | vs +hi*lo %.3e | vs +lo*hi %.3e | vs all three %.3e ; masked modified %ld\n", mode, maxerr, e1, e2a, e2b, e3, wrong);
      sp.release(); cudaFree(dP); cudaFree(dC);
    }
    {
      cublasHandle_t h; cublasCreate(&h);
      const int M = 16384, K = 1024;
      float *dP, *dC; cudaMalloc(&dP, (size_t)M * K * 4); cudaMalloc(&dC, (size_t)M * M * 4);
      cudaMemset(dP, 0, (size_t)M * K * 4); cudaMemset(dC, 0, (size_t)M * M * 4);
      Tf32Split sp; sp.allocate(M, K); sp.split(dP, M, M, 0);
      GemmArgsT<float> g{}; g.C = dC; g.ldc = M; g.M = M; g.N = M; g.K = K; g.alpha = -1; g.beta = 1; g.mask = kMaskNone; g.nbp = 1024; g.P = g.Q = 1;
      launch_gemm_tf32x3(g, sp, 0, sp, 0, 0); cudaDeviceSynchronize();
      cudaEventRecord(e0); for (int i = 0; i < 3; ++i) launch_gemm_tf32x3(g, sp, 0, sp, 0, 0); cudaEventRecord(e1); cudaEventSynchronize(e1);
      double ms = time_ms(e0, e1) / 3; double fl = 2.0 * M * (double)M * K;
      std::printf("tf32x3 tcgen05 %dx%dx%d: %.3f ms  %.1f TFLOP/s (fp32-equivalent; %.1f TF/s of TF32 MMAs)\n", M, M, K, ms, fl / ms / 1e9, 3 * fl / ms / 1e9);
      const float al = -1, be = 1;
      cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, M, M, K, &al, dP, M, dP, M, &be, dC, M);
      cudaEventRecord(e0); for (int i = 0; i < 3; ++i) cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, M, M, K, &al, dP, M, dP, M, &be, dC, M); cudaEventRecord(e1); cudaEventSynchronize(e1);
      ms = time_ms(e0, e1) / 3; std::printf("cublasSgemm (fp32)       : %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
      cublasSetMathMode(h, CUBLAS_TF32_TENSOR_OP_MATH);
      cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, M, M, K, &al, dP, M, dP, M, &be, dC, M);
      cudaEventRecord(e0); for (int i = 0; i < 3; ++i) cublasSgemm(h, CUBLAS_OP_N, CUBLAS_OP_T, M, M, K, &al, dP, M, dP, M, &be, dC, M); cudaEventRecord(e1); cudaEventSynchronize(e1);
      ms = time_ms(e0, e1) / 3; std::printf("cublasSgemm (1xTF32 mode): %.3f ms  %.1f TFLOP/s\n", ms, fl / ms / 1e9);
      sp.release(); cudaFree(dP); cudaFree(dC); cublasDestroy(h);
    }
  }
  std::printf("done\n");
  return 0;
}
