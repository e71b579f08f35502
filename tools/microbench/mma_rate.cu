// Microbenchmark: issue rate of the legacy warp-level tensor-core path (mma.sync) on B200, next to the scalar work a
// split-precision tensor-core DFT needs per element.  Decides DESIGN.md section 6 ("why no tensor cores for the 400-point
// DFT"): a folded DFT-as-GEMM front end needs ~59 HMMA.16816 per frame (3-term fp16/bf16 split for fp32-level accuracy) PLUS
// window / fold / hi-lo split work on the CUDA cores for every element of every frame.
//   part 1: mma.sync.m16n8k16 (f16 -> f32 and bf16 -> f32) and m16n8k8 tf32, ACC independent accumulators per warp,
//           W warps per SM: clk per MMA per sub-partition and the implied dense TFLOP/s of the whole GPU
//   part 2: the per-element operand preparation (2 loads, window multiply, fold add/sub, fp16 hi + bf16 lo split, packed
//           shared-memory stores) in instructions per element
#include <cstdio>
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

template <int KIND>
__device__ __forceinline__ void mma(float (&c)[4], unsigned a0, unsigned a1, unsigned a2, unsigned a3, unsigned b0, unsigned b1) {
  if (KIND == 0)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else if (KIND == 1)
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
  else
    asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

template <int KIND, int ACC>
__global__ void __launch_bounds__(1024) k_mma(float* out, int iters, unsigned seed) {
  float c[ACC][4];
#pragma unroll
  for (int i = 0; i < ACC; ++i) c[i][0] = c[i][1] = c[i][2] = c[i][3] = 0.0f;
  unsigned a0 = seed + threadIdx.x, a1 = a0 * 3u, a2 = a0 * 5u, a3 = a0 * 7u, b0 = a0 * 11u, b1 = a0 * 13u;
  a0 &= 0x3c003c00u; a1 &= 0x3c003c00u; a2 &= 0x3c003c00u; a3 &= 0x3c003c00u; b0 &= 0x3c003c00u; b1 &= 0x3c003c00u;   // finite operands
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < ACC; ++i) mma<KIND>(c[i], a0, a1, a2, a3, b0, b1);
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < ACC; ++i) s += c[i][0] + c[i][1] + c[i][2] + c[i][3];
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int KIND, int ACC>
void run_mma(float* d, int warps, const char* name) {
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = (1 << 19) / ACC;
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_mma<KIND, ACC><<<148, warps * 32>>>(d, iters, 12345u);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  const double mma_per_smsp = double(iters) * ACC * warps / 4.0;
  const double cyc = best * 1e-3 * 1.965e9;
  const double flop_per_mma = KIND == 2 ? 2.0 * 16 * 8 * 8 : 2.0 * 16 * 8 * 16;
  const double tflops = double(iters) * ACC * warps * 148.0 * flop_per_mma / (best * 1e-3) / 1e12;
  printf("%-18s acc %d warps/SM %2d : %.3f ms  clk/MMA/SMSP %.2f  -> %.0f TFLOP/s dense (whole GPU, at 1.965 GHz clocks assumed for clk only)\n",
         name, ACC, warps, best, cyc / mma_per_smsp, tflops);
}

// part 2: operand preparation of a folded, windowed, split-precision A tile.  One element pair per loop trip:
//   e = (x[n] + x[N-1-n]) * w[n],  o = (x[n] - x[N-1-n]) * w[n]   (symmetric window)
//   hi = fp16(v), lo = bf16(v - float(hi))   for v in {e, o}; packed pairs stored to shared memory
__global__ void __launch_bounds__(256) k_prep(const float* __restrict__ x, unsigned* out, int iters) {
  __shared__ float s_x[400 * 8 + 32];
  __shared__ float s_w[200];
  __shared__ unsigned s_a[4][100 * 8];
  for (int i = threadIdx.x; i < 400 * 8 + 32; i += 256) s_x[i] = x[i];
  for (int i = threadIdx.x; i < 200; i += 256) s_w[i] = 0.5f - 0.5f * cospif(2.0f * i / 399.0f);
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  unsigned acc = 0;
  for (int it = 0; it < iters; ++it) {
    const float* fr = s_x + warp * 400 + (it & 31);   // (one frame per warp and trip: the point is the instruction count per element)
#pragma unroll
    for (int j = 0; j < 3; ++j) {   // lanes cover element pairs (n, n+1) of 200 folded positions: 100 pairs -> 3 trips + tail
      const int n = 2 * (lane + 32 * j);
      if (n < 200) {
        const float a0 = fr[n], a1 = fr[n + 1], b0 = fr[399 - n], b1 = fr[398 - n], w0 = s_w[n], w1 = s_w[n + 1];
        const float e0 = (a0 + b0) * w0, e1 = (a1 + b1) * w1, o0 = (a0 - b0) * w0, o1 = (a1 - b1) * w1;
        const __half2 eh = __floats2half2_rn(e0, e1), oh = __floats2half2_rn(o0, o1);
        const float2 ef = __half22float2(eh), of = __half22float2(oh);
        const __nv_bfloat162 el = __floats2bfloat162_rn(e0 - ef.x, e1 - ef.y), ol = __floats2bfloat162_rn(o0 - of.x, o1 - of.y);
        s_a[0][warp * 100 + n / 2] = *reinterpret_cast<const unsigned*>(&eh);
        s_a[1][warp * 100 + n / 2] = *reinterpret_cast<const unsigned*>(&el);
        s_a[2][warp * 100 + n / 2] = *reinterpret_cast<const unsigned*>(&oh);
        s_a[3][warp * 100 + n / 2] = *reinterpret_cast<const unsigned*>(&ol);
      }
    }
    __syncwarp();
    acc += s_a[it & 3][warp * 100 + lane];
  }
  out[blockIdx.x * 256 + threadIdx.x] = acc;
}

int main() {
  float* d; cudaMalloc(&d, 148 * 1024 * 4);
  for (int warps : {4, 8, 16}) {
    run_mma<0, 4>(d, warps, "m16n8k16 f16");
    run_mma<0, 8>(d, warps, "m16n8k16 f16");
    run_mma<1, 8>(d, warps, "m16n8k16 bf16");
    run_mma<2, 8>(d, warps, "m16n8k8 tf32");
  }
  // part 2
  float* x; cudaMalloc(&x, (400 * 8 + 32) * 4); cudaMemset(x, 0, (400 * 8 + 32) * 4);
  unsigned* o; cudaMalloc(&o, 148 * 4 * 256 * 4);
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  const int iters = 1 << 16;
  float best = 1e9f;
  for (int rep = 0; rep < 3; ++rep) {
    cudaEventRecord(e0);
    k_prep<<<148 * 4, 256>>>(x, o, iters);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
  }
  // 8 warps per CTA, 4 CTAs per SM = 32 warps per SM = 8 per sub-partition; each warp prepares one 400-sample frame per trip
  const double frames = double(iters) * 8 * 4 * 148;
  printf("operand preparation (fold + window + fp16/bf16 split + store): %.3f ms for %.3g frames -> %.1f clk per frame per SMSP, "
         "%.2f ms for the 3.07e6 frames of a 1024 x 30 s step\n", best, frames, best * 1e-3 * 1.965e9 / (frames / (148.0 * 4)),
         best * 3.07e6 / frames);
  if (cudaDeviceSynchronize() != cudaSuccess) { printf("CUDA error\n"); return 1; }
  return 0;
}
