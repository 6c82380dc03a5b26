// Do DMMA (tensor pipe) and DFMA (fp64 pipe) share execution units on B200? Runs (a) DMMA only,
// (b) DFMA only, (c) both interleaved in every warp, (d) half of the warps each.
#include <cuda_runtime.h>
#include <cstdio>

template <int MODE>
__global__ void k(double* out, int iters) {
  double a = threadIdx.x * 1e-3, b = 1.0000001;
  double c[8][2], f[16];
#pragma unroll
  for (int i = 0; i < 8; ++i) c[i][0] = c[i][1] = 0.0;
#pragma unroll
  for (int i = 0; i < 16; ++i) f[i] = i;
  const int warp = threadIdx.x >> 5;
  const bool do_mma = (MODE == 0) || (MODE == 2) || (MODE == 3 && (warp & 1) == 0);
  const bool do_fma = (MODE == 1) || (MODE == 2) || (MODE == 3 && (warp & 1) == 1);
  for (int it = 0; it < iters; ++it) {
    if (do_mma) {
#pragma unroll
      for (int i = 0; i < 8; ++i)
        asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                     : "+d"(c[i][0]), "+d"(c[i][1]) : "d"(a), "d"(b));
    }
    if (do_fma) {
#pragma unroll
      for (int i = 0; i < 16; ++i) f[i] = fma(f[i], b, a);
    }
  }
  double s = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) s += c[i][0] + c[i][1];
#pragma unroll
  for (int i = 0; i < 16; ++i) s += f[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE>
void run(const char* name, int nsm, double* out) {
  const int threads = 512, iters = 20000;
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<nsm, threads>>>(out, 100);
  cudaEventRecord(e0);
  k<MODE><<<nsm, threads>>>(out, iters);
  cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  const double warps = threads / 32.0 * nsm;
  double mma_w = (MODE == 0 || MODE == 2) ? warps : (MODE == 3 ? warps / 2 : 0);
  double fma_w = (MODE == 1 || MODE == 2) ? warps : (MODE == 3 ? warps / 2 : 0);
  double fl_mma = 2.0 * 256 * 8 * iters * mma_w, fl_fma = 2.0 * 16 * 32 * iters * fma_w;
  std::printf("%-28s %.3f ms  DMMA %.2f + DFMA %.2f = %.2f TFLOP/s\n", name, ms, fl_mma / ms / 1e9, fl_fma / ms / 1e9, (fl_mma + fl_fma) / ms / 1e9);
}

int main() {
  int nsm; cudaDeviceGetAttribute(&nsm, cudaDevAttrMultiProcessorCount, 0);
  double* out; cudaMalloc(&out, sizeof(double) * nsm * 1024);
  run<0>("DMMA only", nsm, out);
  run<1>("DFMA only", nsm, out);
  run<2>("both, every warp", nsm, out);
  run<3>("both, alternating warps", nsm, out);
  return 0;
}
