// Microbenchmark: do packed f32x2 instructions (FFMA2/FADD2/FMUL2, sm_100a) leave issue slots free for
// ALU / LSU instructions?  Each kernel runs a loop of NF packed (or 2*NF scalar) FP ops on independent
// accumulators plus NA integer ALU ops and NL shared-memory loads per iteration.
#include <cstdio>
#include <cuda_runtime.h>

template <int FP, int NA, int NL>   // FP: 0 none, 1 scalar FFMA r,imm,r (2 per slot pair), 2 FFMA2 r,r,r, 3 FADD2 r,r, 4 scalar FADD, 5 FFMA2 r,const,r, 6 scalar FFMA r,r,r
__global__ void k(float* out, int iters, float a, float b) {
  __shared__ float sm[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) sm[i] = i;
  __syncthreads();
  float2 acc[8];
  unsigned ia[8];
  float ls = 0.f;
  for (int i = 0; i < 8; ++i) { acc[i] = make_float2(threadIdx.x * 0.001f + i, i * 0.5f); ia[i] = threadIdx.x + i; }
  const float2 A = make_float2(a, a * 1.0001f), B = make_float2(b, b * 0.9999f);
  unsigned off = threadIdx.x & 31;
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (FP == 1) { acc[i].x = fmaf(acc[i].x, 1.0001f, B.x); acc[i].y = fmaf(acc[i].y, 0.9999f, B.y); }
      else if (FP == 2) acc[i] = __ffma2_rn(acc[i], A, B);
      else if (FP == 3) acc[i] = __fadd2_rn(acc[i], B);
      else if (FP == 4) { acc[i].x = acc[i].x + B.x; acc[i].y = acc[i].y + B.y; }
      else if (FP == 5) acc[i] = __ffma2_rn(acc[i], make_float2(1.0001f, 1.0001f), B);
      else if (FP == 6) { acc[i].x = fmaf(acc[i].x, A.x, B.x); acc[i].y = fmaf(acc[i].y, A.y, B.y); }
      if (i < NA) ia[i] = (ia[i] ^ (ia[i] >> 3)) + 0x9e3779b9u;   // LOP3/SHF + IADD: 2-3 ALU ops
      if (i < NL) { ls += sm[(off + i * 32) & 1023]; }
    }
    off += 7;
  }
  float s = ls;
  for (int i = 0; i < 8; ++i) s += acc[i].x + acc[i].y + float(ia[i]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int FP, int NA, int NL> void run(float* d, const char* name) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 20000, blocks = 148 * 4, threads = 512;
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<FP, NA, NL><<<blocks, threads>>>(d, iters, 1.0001f, 0.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  // cycles per loop iteration per SMSP: warps per SMSP = 4 blocks * 16 warps / 4 = 16
  double cyc = best * 1e-3 * 1.965e9 / iters / 16.0;
  printf("%-44s %.3f ms  %6.2f clk per warp-iteration (8 FP slots, %d ALU groups, %d LDS)\n", name, best, cyc, NA, NL);
}
int main() {
  float* d; cudaMalloc(&d, 148 * 8 * 1024 * 4);
  run<0, 8, 0>(d, "ALU only (8 groups)");
  run<0, 0, 8>(d, "LDS only (8)");
  run<1, 0, 0>(d, "16 scalar FFMA r,imm,r");
  run<6, 0, 0>(d, "16 scalar FFMA r,r,r");
  run<4, 0, 0>(d, "16 scalar FADD");
  run<2, 0, 0>(d, "8 FFMA2 r,r,r");
  run<5, 0, 0>(d, "8 FFMA2 r,const,r");
  run<3, 0, 0>(d, "8 FADD2");
  run<1, 8, 0>(d, "16 scalar FFMA imm + ALU 8");
  run<2, 8, 0>(d, "8 FFMA2 + ALU 8");
  run<3, 8, 0>(d, "8 FADD2 + ALU 8");
  run<2, 4, 0>(d, "8 FFMA2 + ALU 4");
  run<1, 0, 8>(d, "16 scalar FFMA imm + LDS 8");
  run<2, 0, 8>(d, "8 FFMA2 + LDS 8");
  run<3, 0, 8>(d, "8 FADD2 + LDS 8");
  run<2, 4, 4>(d, "8 FFMA2 + ALU 4 + LDS 4");
  run<1, 4, 4>(d, "16 scalar FFMA imm + ALU 4 + LDS 4");
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
