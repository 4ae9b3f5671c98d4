// microbenchmark: is packed FADD2/FFMA2 (sm_100 f32x2) cheaper in issue slots than two scalar ops?
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE> __global__ void k(float2 *out, int iters) {
  float2 a[8], b = make_float2(1.0001f, 0.9999f), c = make_float2(1e-3f, -1e-3f);
  for (int i = 0; i < 8; i++) a[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) {
      if (MODE == 0) { a[i].x = fmaf(a[i].x, b.x, c.x); a[i].y = fmaf(a[i].y, b.y, c.y); }
      if (MODE == 1) a[i] = __ffma2_rn(a[i], b, c);
      if (MODE == 2) { a[i].x = a[i].x + c.x; a[i].y = a[i].y + c.y; }
      if (MODE == 3) a[i] = __fadd2_rn(a[i], c);
    }
  }
  float2 s = make_float2(0, 0);
  for (int i = 0; i < 8; i++) { s.x += a[i].x; s.y += a[i].y; }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> float run(float2 *d, int iters) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  k<MODE><<<148 * 8, 256>>>(d, iters); cudaDeviceSynchronize();
  cudaEventRecord(e0); k<MODE><<<148 * 8, 256>>>(d, iters); cudaEventRecord(e1); cudaEventSynchronize(e1);
  float ms; cudaEventElapsedTime(&ms, e0, e1); return ms;
}
int main() {
  float2 *d; cudaMalloc(&d, sizeof(float2) * 148 * 8 * 256);
  int iters = 20000;
  double flop = 148.0 * 8 * 256 * iters * 8 * 2;  // per mode: 16 scalar results per inner step
  printf("scalar FFMA x2 : %.3f ms  %.1f Gop/s\n", run<0>(d, iters), flop / run<0>(d, iters) / 1e6);
  printf("FFMA2          : %.3f ms  %.1f Gop/s\n", run<1>(d, iters), flop / run<1>(d, iters) / 1e6);
  printf("scalar FADD x2 : %.3f ms  %.1f Gop/s\n", run<2>(d, iters), flop / run<2>(d, iters) / 1e6);
  printf("FADD2          : %.3f ms  %.1f Gop/s\n", run<3>(d, iters), flop / run<3>(d, iters) / 1e6);
  return 0;
}
