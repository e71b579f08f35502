// Microbenchmark: scalar FFMA vs packed FFMA2 (fma.rn.f32x2, sm_100) issue throughput on B200.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void k(float* out, int iters, float a, float b) {
  float2 acc[8];
  for (int i = 0; i < 8; ++i) acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f);
  const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.9999f);
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, A.x, B.x); acc[i].y = fmaf(acc[i].y, A.y, B.y); }
      else acc[i] = __ffma2_rn(acc[i], A, B);
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = 148 * 4, threads = 512;
  for (int mode = 0; mode < 2; ++mode) for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    if (mode == 0) k<0><<<blocks, threads>>>(d, iters, 1.0001f, 0.0001f); else k<1><<<blocks, threads>>>(d, iters, 1.0001f, 0.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    double fma = double(blocks) * threads * iters * 16.0;
    printf("mode %s: %.3f ms  %.2f TFMA/s  (%.1f FMA/clk/SM @1.965GHz)\n", mode ? "FFMA2" : "FFMA ", ms, fma / ms * 1e-9, fma / (ms * 1e-3) / 148 / 1.965e9);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
