// Microbenchmark: instruction-fetch behaviour of straight-line code on B200.
// A kernel body of K independent-chain FFMAs (K * 16 bytes of SASS) is executed `iters` times by W warps per SM.
//   mode 0: every warp runs the whole body            (shared code: later warps hit what the first one fetched)
//   mode 1: warp w runs only slice w of 8 slices      (each instruction is used by one warp per pass, like a per-warp switch)
#include <cstdio>
#include <cuda_runtime.h>

template <int K>
__device__ __forceinline__ void body(float (&a)[8], float x, float y) {
#pragma unroll
  for (int i = 0; i < K; ++i) a[i & 7] = fmaf(a[i & 7], x, y);
}

template <int K, int MODE>
__global__ void __launch_bounds__(1024) k(float* out, int iters, float x, float y) {
  float a[8];
  for (int i = 0; i < 8; ++i) a[i] = threadIdx.x * 0.001f + i;
  const int w = (threadIdx.x >> 5) & 7;
  for (int it = 0; it < iters; ++it) {
    if (MODE == 0) {
      body<K>(a, x, y);
    } else {
      switch (w) {   // 8 distinct copies of K/8 instructions (different immediates keep them distinct)
        case 0: body<K / 8>(a, x, y + 0.f); break;
        case 1: body<K / 8>(a, x, y + 1.f); break;
        case 2: body<K / 8>(a, x, y + 2.f); break;
        case 3: body<K / 8>(a, x, y + 3.f); break;
        case 4: body<K / 8>(a, x, y + 4.f); break;
        case 5: body<K / 8>(a, x, y + 5.f); break;
        case 6: body<K / 8>(a, x, y + 6.f); break;
        default: body<K / 8>(a, x, y + 7.f); break;
      }
    }
  }
  float s = 0;
  for (int i = 0; i < 8; ++i) s += a[i];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int K, int MODE>
void run(float* d, int warps) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int per_warp = MODE == 0 ? K : K / 8;
  const int iters = (4 << 20) / per_warp;  // ~4M instructions per warp
  float best = 1e9;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k<K, MODE><<<148, warps * 32>>>(d, iters, 1.0001f, 0.0001f);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double instr_per_smsp = double(iters) * per_warp * warps / 4.0;
  const double cyc = best * 1e-3 * 1.965e9;
  printf("K=%6d (%4d KB) mode %d warps/SM %2d : %.3f ms  IPC/SMSP %.3f\n", K, K * 16 / 1024, MODE, warps, best, instr_per_smsp / cyc);
}

int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  for (int warps : {4, 8, 16, 32}) {
    run<512, 0>(d, warps); run<1024, 0>(d, warps); run<2048, 0>(d, warps); run<4096, 0>(d, warps); run<8192, 0>(d, warps); run<16384, 0>(d, warps);
  }
  for (int warps : {8, 16, 32}) {
    run<1024, 1>(d, warps); run<2048, 1>(d, warps); run<4096, 1>(d, warps); run<8192, 1>(d, warps); run<16384, 1>(d, warps);
  }
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
}
