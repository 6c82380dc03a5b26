// Dependent-chain latencies of the FP64 operations the diagonal kernel relies on (one warp, clock64).
#include <cuda_runtime.h>
#include <cstdio>
__global__ void lat(double* out, long long* cyc, double seed) {
  double x = seed + threadIdx.x * 1e-9, y = 1.0000001, z = 0.5;
  long long t0, t1;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { x = fma(x, y, z); x = fma(x, y, z); x = fma(x, y, z); x = fma(x, y, z); }
  t1 = clock64(); if (threadIdx.x == 0) cyc[0] = (t1 - t0) / 1024;
  double a = x * 1e-300 + 2.0;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) { a = rsqrt(a) + 1.5; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[1] = (t1 - t0) / 64;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) { a = sqrt(a) + 1.5; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[2] = (t1 - t0) / 64;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) { a = 1.0 / a + 1.5; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[3] = (t1 - t0) / 64;
  float f = (float)a;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) { f = rsqrtf(f) + 1.5f; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[4] = (t1 - t0) / 64;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 64; ++i) { double y0 = (double)rsqrtf((float)a); y0 = y0 * fma(-0.5 * a * y0, y0, 1.5); y0 = y0 * fma(-0.5 * a * y0, y0, 1.5); a = y0 + 1.5; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[5] = (t1 - t0) / 64;
  __shared__ double sh[64];
  sh[threadIdx.x] = a; __syncthreads();
  int idx = threadIdx.x;
  t0 = clock64();
#pragma unroll 1
  for (int i = 0; i < 256; ++i) { idx = (int)sh[idx & 31] & 31; }
  t1 = clock64(); if (threadIdx.x == 0) cyc[6] = (t1 - t0) / 256;
  out[threadIdx.x] = x + a + f + idx;
}
int main() {
  double* o; long long* c; cudaMalloc(&o, 256 * 8); cudaMalloc(&c, 64);
  lat<<<1, 32>>>(o, c, 1.0); lat<<<1, 32>>>(o, c, 1.0); cudaDeviceSynchronize();
  long long h[8]; cudaMemcpy(h, c, 56, cudaMemcpyDeviceToHost);
  printf("dependent DFMA %lld clk | rsqrt(double)+add %lld | sqrt(double)+add %lld | 1.0/x+add %lld | rsqrtf+add %lld | rsqrtf seed + 2 Newton (fp64) + add %lld | LDS.64 + cvt chain %lld\n", h[0], h[1], h[2], h[3], h[4], h[5], h[6]);
  return 0;
}
