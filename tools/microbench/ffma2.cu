// Microbenchmark: scalar vs packed (f32x2, sm_100) FP32 issue throughput on B200, by operand form.
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
      if (MODE == 0) { acc[i].x = fmaf(acc[i].x, A.x, B.x); acc[i].y = fmaf(acc[i].y, A.y, B.y); }          // FFMA r,r,r
      else if (MODE == 1) acc[i] = __ffma2_rn(acc[i], A, B);                                                 // FFMA2 r,r,r
      else if (MODE == 2) { acc[i].x = fmaf(acc[i].x, 1.0001f, B.x); acc[i].y = fmaf(acc[i].y, 0.9999f, B.y); } // FFMA r,imm,r
      else if (MODE == 3) { acc[i].x = acc[i].x + B.x; acc[i].y = acc[i].y + B.y; }                          // FADD r,r
      else if (MODE == 4) acc[i] = __fadd2_rn(acc[i], B);                                                    // FADD2
      else if (MODE == 5) { acc[i].x = acc[i].x * 1.0001f; acc[i].y = acc[i].y * 0.9999f; }                  // FMUL r,imm
      else if (MODE == 6) acc[i] = __fmul2_rn(acc[i], A);                                                    // FMUL2
      else if (MODE == 7) acc[i] = __ffma2_rn(acc[i], make_float2(1.0001f, 1.0001f), B);                     // FFMA2 with constant pair
      else if (MODE == 8) { acc[i].x = fmaf(acc[i].x, 1.0001f, acc[(i + 1) & 7].y); acc[i].y = fmaf(acc[i].y, 0.9999f, acc[(i + 1) & 7].x); } // FFMA r,imm,r(other)
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(float* d, const char* name) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = 148 * 4, threads = 512;
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(d, iters, 1.0001f, 0.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  double ops = double(blocks) * threads * iters * 16.0;
  printf("%-28s %.3f ms  %6.1f flop-lanes/clk/SM @1.965GHz\n", name, best, ops / (best * 1e-3) / 148 / 1.965e9);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  run<0>(d, "FFMA  r,r,r"); run<1>(d, "FFMA2 r,r,r"); run<2>(d, "FFMA  r,imm,r"); run<8>(d, "FFMA  r,imm,r2");
  run<3>(d, "FADD  r,r"); run<4>(d, "FADD2 r,r"); run<5>(d, "FMUL  r,imm"); run<6>(d, "FMUL2 r,r"); run<7>(d, "FFMA2 r,constpair,r");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
