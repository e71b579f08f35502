// Tensor-core Whisper front end for sm_100a: the 400-point DFT of every frame as split-precision GEMMs on the 5th-generation
// tensor cores (tcgen05.mma, accumulators in TMEM), with the window / folding / operand split in front of it and
// |X|^2 -> mel -> log10 -> (x + 4) / 4 -> (T', M) store behind it, in ONE kernel.
//
// Replaces the same chain as frontend_kernel<Plan400, ...> (STT/Whisper/WhisperAudio.swift:78-137; stft(),
// Codec/S3Tokenizer/S3TokenizerUtils.swift:224-263).  Nothing here is derived from MLX source.
//
// Math (tools/tcgen05_dft/emulate.py is the NumPy statement of exactly this):
//   xw[n] = x[n] * w[n];  s1 = xw[n], s2 = xw[400 - n], s3 = xw[200 - n], s4 = xw[200 + n], n = 0 .. 100 (taps that do not exist are 0)
//   a1 = s1 + s2, b1 = s1 - s2, a2 = s3 + s4, b2 = s3 - s4            (fold about n = 200, then about n = 100)
//   ee = a1 + a2 -> Re X[2j],  eo = a1 - a2 -> Re X[2j+1],  oe = b1 - b2 -> Im X[2j],  oo = b1 + b2 -> Im X[2j+1]
//   each of the four is a (frames x 112) x (112 x 112) GEMM against a constant cos / sin matrix (K and N zero-padded from ~100)
//   Operands: v * 2^12 and F * 2^4 are split as hi = fp16(v), lo = fp16(v - hi) (the global pre-scale keeps the lo terms in fp16's
//   normal range), D = Ah Bh + Al Bh + Ah Bl in one fp32 TMEM accumulator: ~2^-24 per operand, fp32 accumulation.
//
// Structure: persistent CTA (512 threads, 1 per SM), tile = 128 consecutive frames of one clip = the 128 TMEM lanes;
// THREAD t & 127 == FRAME == TMEM LANE in the operand preparation and in the epilogue.
//   * PCM of the tile staged once with cp.async into skewed rows (161-word pitch: the stride-160 frame starts of a warp's lanes
//     fall in 32 different banks);
//   * K is walked in 7 slices of 16: all threads build the slice's A operands (4 quadrants x hi / lo, canonical no-swizzle K-major
//     UMMA layout, 32 KB) while the tensor core works on the previous slice (2-slot ring); the matching 28 KB slice of the constant
//     B operands arrives by TMA (cp.async.bulk.tensor) on an mbarrier; one thread issues the slice's 12 tcgen05.mma and commits them
//     to the slot's "free" mbarrier;
//   * epilogue: tcgen05.ld of Re / Im, power into a [bin][frame] tile in shared memory, sparse mel projection from it
//     (lane == frame), log / scale, per-warp transposing staging, coalesced 128-byte row segments of the (T', M) output, per-clip
//     max and per-32-frame-tile min for the max - 8 clamp (the existing whisper_clamp_kernel runs afterwards).
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200audio.h"
#include "internal.h"

namespace b2a {

namespace {

constexpr int kN = 400, kHop = 160, kM = 128, kKP = 112, kNP = 112, kSlices = kKP / 16, kThreads = 512, kWarps = kThreads / 32;
constexpr int kQuads = 4;                               // ee, eo, oe, oo
constexpr int kABlock = kM * 16 * 2;                    // one (quadrant, hi|lo) A block of a K slice: 128 rows x 16 halves = 4096 B
constexpr int kBBlock = kNP * 16 * 2;                   // the same for B: 112 rows x 16 halves = 3584 B
constexpr int kASlice = kQuads * 2 * kABlock;           // 32768 B
constexpr int kBSlice = kQuads * 2 * kBBlock;           // 28672 B
constexpr int kSpan = (kM - 1) * kHop + kN + 1;         // samples a tile touches (incl. the zero-weight tap x[400])
constexpr int kPcmRows = (kSpan - 1 + 3) / kHop + 1;    // skewed rows: row R holds tile samples [160 R - 3, 160 R + 157)
constexpr int kPcmWords = (kSpan + kPcmRows + 3) & ~3;
constexpr float kScaleX = 4096.0f, kScaleF = 16.0f;     // 2^12, 2^4
constexpr float kPowerScale = 1.0f / (4096.0f * 4096.0f * 16.0f * 16.0f);   // 2^-32

// shared-memory map (bytes): the operand rings are dead once the last slice's MMAs are done; the epilogue reuses everything
constexpr int kOffPcm = 0;
constexpr int kOffA = ((kPcmWords * 4 + 1023) / 1024) * 1024;
constexpr int kOffB = kOffA + 2 * kASlice;
constexpr int kOffEnd = kOffB + 2 * kBSlice;
constexpr int kOffSpec = 0;                             // [201][128] floats
constexpr int kOffStage = ((201 * kM * 4 + 1023) / 1024) * 1024;   // 16 warps x 32 x 33 floats
constexpr int kSmemBytes = kOffEnd;
static_assert(kOffStage + kWarps * 32 * 33 * 4 <= kOffEnd, "epilogue regions fit the operand rings");
static_assert(kSmemBytes <= 227 * 1024, "shared memory");

struct TcTable {            // per (slice, sub): word offsets of the four sample streams and their 4 x 4 window taps (pre-scaled by 2^12)
  int off[kSlices][4][4];
  float tap[kSlices][4][16];   // [stream][i]
};

struct TcParams {
  const float* x;
  float* out;
  float* dbg_power;          // optional (frames x 201), debug entry only
  long long n_samples, n_frames, out_clip_stride;
  int n_clips, tiles_per_clip, tiles32_per_clip, n_mels;
  long long total_tiles;
  const int4* fb_desc;       // per filter: (first bin, number of bins, offset into fb_w, 0)
  const float* fb_w;
  int* clip_max;
  int* tile_min;
  float log_floor_scaled;    // log floor * 2^32 is NOT used: the mel value is scaled back before the log (kept for clarity)
  TcTable tab;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "WAIT_LOOP:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra WAIT_DONE;\n\t"
      "bra WAIT_LOOP;\n\t"
      "WAIT_DONE:\n\t"
      "}" ::"r"(bar), "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
               "l"(map), "r"(bar), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void cp_async4(uint32_t dst, const float* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(dst), "l"(src));
}

// K-major, no swizzle ("interleave") canonical layout: 8 rows x 16 bytes contiguous (row r of the core matrix at r * 16),
// 8-row groups SBO bytes apart, the two 16-byte K chunks of one MMA LBO bytes apart (cute/arch/mma_sm100_desc.hpp: SmemDescriptor)
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= uint64_t((saddr & 0x3FFFFu) >> 4);
  d |= uint64_t(lbo_bytes >> 4) << 16;
  d |= uint64_t(sbo_bytes >> 4) << 32;
  d |= uint64_t(1) << 46;      // descriptor version (Blackwell)
  return d;                    // base offset 0, layout type 0 = SWIZZLE_NONE
}
// kind::f16 instruction descriptor: D = f32, A = B = f16, both K-major, M = 128, N = 112 (cute/arch/mma_sm100_desc.hpp: InstrDescriptor)
constexpr uint32_t kIdesc = (1u << 4) | (uint32_t(kNP >> 3) << 17) | (uint32_t(kM >> 4) << 24);

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t"
      ".reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
      "}" ::"r"(d_tmem),
      "l"(adesc), "l"(bdesc), "r"(kIdesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tmem_ld8(uint32_t taddr, float (&v)[8]) {
  uint32_t r[8];
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 8; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ int enc_ordered(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
__device__ __forceinline__ float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// reflectPad index map (S3TokenizerUtils.swift:266-298; same statement as padded_index() in frontend.cu), pad 200, reflect mode
__device__ __forceinline__ long long reflect_index(long long j, long long n) {
  if (j < 0 || j >= n) {
    if (n == 1) return 0;
    long long t = j < 0 ? -j - 1 : j - n;
    if (t >= n - 1) t %= n - 1;
    j = j < 0 ? t + 1 : n - 2 - t;
  }
  return j;
}

__global__ void __launch_bounds__(kThreads, 1) tc_whisper_kernel(const __grid_constant__ TcParams prm, const __grid_constant__ CUtensorMap bmap) {
  extern __shared__ __align__(1024) unsigned char smem[];
  __shared__ __align__(8) unsigned long long s_bar[5];   // 0,1: A/B slot free; 2,3: B slice landed; 4: accumulators complete
  __shared__ uint32_t s_tmem;
  __shared__ __align__(16) TcTable s_tab;
  float* s_pcm = reinterpret_cast<float*>(smem + kOffPcm);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int t = tid & (kM - 1);        // frame of the tile == TMEM lane
  const int sub = tid >> 7;            // which quarter of every K slice / of the epilogue this thread takes
  const uint32_t bar_free0 = smem_u32(&s_bar[0]), bar_b0 = smem_u32(&s_bar[2]), bar_done = smem_u32(&s_bar[4]);

  // ---- one-time setup ----
  for (int i = tid; i < int(sizeof(TcTable) / 4); i += kThreads) reinterpret_cast<int*>(&s_tab)[i] = reinterpret_cast<const int*>(&prm.tab)[i];
  if (tid == 0) {
    mbar_init(bar_free0, 1);
    mbar_init(bar_free0 + 8, 1);
    mbar_init(bar_b0, 1);
    mbar_init(bar_b0 + 8, 1);
    mbar_init(bar_done, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    mbar_arrive(bar_free0);        // both ring slots start out free
    mbar_arrive(bar_free0 + 8);
  }
  if (warp == 0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = s_tmem;
  uint32_t ph_free[2] = {0u, 0u}, ph_b[2] = {0u, 0u}, ph_done = 0u;

  const long long n = prm.n_samples;
  for (long long g = blockIdx.x; g < prm.total_tiles; g += gridDim.x) {
    const int clip = int(g / prm.tiles_per_clip);
    const int tile = int(g - (long long)clip * prm.tiles_per_clip);
    const int f0 = tile * kM;
    const float* __restrict__ xc = prm.x + (long long)clip * n;

    // ---- 1. PCM of the tile -> skewed rows (tile sample j at word j + (j + 3) / 160) ----
    {
      const long long i0 = (long long)f0 * kHop - 200;      // signal index of tile sample 0
      for (int R = warp; R < kPcmRows; R += kWarps) {
        const int jr = R * kHop - 3;
        const uint32_t dst = smem_u32(s_pcm + jr + R + lane);
        const long long ir = i0 + jr;
        if (jr >= 0 && jr + kHop <= kSpan && ir >= 0 && ir + kHop <= n) {     // (warp-uniform) whole row inside the tile and the clip
          const float* src = xc + ir + lane;
#pragma unroll
          for (int q = 0; q < kHop / 32; ++q) cp_async4(dst + q * 128, src + q * 32);
        } else {
          for (int q = 0; q < kHop / 32; ++q) {
            const int j = jr + lane + q * 32;
            if (j < 0 || j >= kSpan) continue;
            const long long i = reflect_index(i0 + j, n);
            cp_async4(dst + q * 128, xc + i);
          }
        }
      }
      asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
    }
    __syncthreads();

    // ---- 2. K slices: operand preparation (all threads) overlapped with the previous slice's MMAs ----
    const float* fpcm = s_pcm + t * (kHop + 1);
#pragma unroll 1
    for (int s = 0; s < kSlices; ++s) {
      const int slot = s & 1;
      unsigned char* a_slot = smem + kOffA + slot * kASlice;
      unsigned char* b_slot = smem + kOffB + slot * kBSlice;
      mbar_wait(bar_free0 + 8 * slot, ph_free[slot]);      // the MMAs that read this slot two slices ago are done
      ph_free[slot] ^= 1u;
      if (tid == 0) {
        mbar_expect_tx(bar_b0 + 8 * slot, kBSlice);
        tma_load_2d(smem_u32(b_slot), &bmap, bar_b0 + 8 * slot, 0, s * (kBSlice / 256));
      }
      {
        const int4 o = *reinterpret_cast<const int4*>(&s_tab.off[s][sub][0]);
        float w[16];
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 v = *reinterpret_cast<const float4*>(&s_tab.tap[s][sub][4 * q]);
          w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
        }
        float ee[4], eo[4], oe[4], oo[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float x1 = fpcm[o.x + i], x2 = fpcm[o.y - i], x3 = fpcm[o.z - i], x4 = fpcm[o.w + i];
          const float t1 = x1 * w[i], t3 = x3 * w[8 + i];
          const float a1 = fmaf(x2, w[4 + i], t1), b1 = fmaf(-x2, w[4 + i], t1);
          const float a2 = fmaf(x4, w[12 + i], t3), b2 = fmaf(-x4, w[12 + i], t3);
          ee[i] = a1 + a2; eo[i] = a1 - a2; oe[i] = b1 - b2; oo[i] = b1 + b2;
        }
        // hi / lo split and store: 4 halves = 8 bytes per (quadrant, part) at row t, K offset 4 * sub of the slice
        unsigned char* row = a_slot + (t & 7) * 16 + (t >> 3) * 128 + (sub >> 1) * (kM * 16) + (sub & 1) * 8;
        auto put = [&](int q, const float (&v)[4]) {
          const __half2 h01 = __floats2half2_rn(v[0], v[1]), h23 = __floats2half2_rn(v[2], v[3]);
          const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
          const __half2 l01 = __floats2half2_rn(v[0] - f01.x, v[1] - f01.y), l23 = __floats2half2_rn(v[2] - f23.x, v[3] - f23.y);
          uint2 hi, lo;
          hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
          lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
          *reinterpret_cast<uint2*>(row + (2 * q) * kABlock) = hi;
          *reinterpret_cast<uint2*>(row + (2 * q + 1) * kABlock) = lo;
        };
        put(0, ee); put(1, eo); put(2, oe); put(3, oo);
      }
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy stores -> visible to the tensor core's reads
      __syncthreads();
      if (tid == 0) {
        mbar_wait(bar_b0 + 8 * slot, ph_b[slot]);
        tc_fence_after();
        const uint32_t a0 = smem_u32(a_slot), b0 = smem_u32(b_slot);
#pragma unroll
        for (int q = 0; q < kQuads; ++q) {
          const uint64_t ah = umma_desc(a0 + (2 * q) * kABlock, kM * 16, 128), al = umma_desc(a0 + (2 * q + 1) * kABlock, kM * 16, 128);
          const uint64_t bh = umma_desc(b0 + (2 * q) * kBBlock, kNP * 16, 128), bl = umma_desc(b0 + (2 * q + 1) * kBBlock, kNP * 16, 128);
          const uint32_t d = tmem_base + q * kNP;
          umma_f16(d, ah, bh, s > 0 ? 1u : 0u);
          umma_f16(d, al, bh, 1u);
          umma_f16(d, ah, bl, 1u);
        }
        umma_commit(bar_free0 + 8 * slot);                  // the slot is free again when these MMAs have read it
        if (s == kSlices - 1) umma_commit(bar_done);        // ... and the accumulators are complete
      }
      ph_b[slot] ^= 1u;   // (every thread tracks the parity; only thread 0 waits on it)
    }

    // ---- 3. epilogue: power spectrum tile [bin][frame] ----
    mbar_wait(bar_done, ph_done);
    ph_done ^= 1u;
    tc_fence_after();
    float* s_p = reinterpret_cast<float*>(smem + kOffSpec);
    {
      // sub 0: even bins 2j, j = 0..55;  sub 1: even bins, j = 56..100;  sub 2: odd bins 2j+1, j = 0..55;  sub 3: odd bins, j = 56..99
      const int odd = sub >> 1, j0 = (sub & 1) * 56, jn = (sub & 1) ? (odd ? 100 : 101) : 56;
      const uint32_t lane_base = tmem_base + (uint32_t(32 * (warp & 3)) << 16);
      const uint32_t re_col = (odd ? 1 : 0) * kNP, im_col = (odd ? 3 : 2) * kNP;
      for (int j = j0; j < jn; j += 8) {
        float re[8], im[8];
        tmem_ld8(lane_base + re_col + j, re);
        tmem_ld8(lane_base + im_col + j, im);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          if (j + i < jn) {
            const int bin = 2 * (j + i) + odd;
            const float pw = re[i] * re[i] + im[i] * im[i];
            s_p[bin * kM + t] = pw;
            if (prm.dbg_power != nullptr && f0 + t < prm.n_frames)
              prm.dbg_power[((long long)clip * prm.n_frames + f0 + t) * 201 + bin] = pw * kPowerScale;
          }
        }
      }
    }
    tc_fence_before();
    __syncthreads();

    // ---- 4. sparse mel projection (lane == frame), log10 / scale, transposing staging, (T', M) rows ----
    {
      const int g4 = warp & 3;                              // 32-frame group of this warp
      float* s_st = reinterpret_cast<float*>(smem + kOffStage) + warp * (32 * 33);
      const bool frame_ok = f0 + t < prm.n_frames;
      float lmax = -3.0e38f, vmin = 3.0e38f;
      const int m0 = sub * 32;
      for (int mm = 0; mm < 32; ++mm) {
        const int m = m0 + mm;
        float v = 0.0f;
        if (m < prm.n_mels) {
          const int4 d = __ldg(prm.fb_desc + m);
          const float* __restrict__ wv = prm.fb_w + d.z;
          const float* p = s_p + d.x * kM + t;
          for (int i = 0; i < d.y; ++i) v = fmaf(__ldg(wv + i), p[i * kM], v);
          v *= kPowerScale;
          v = fmaf(lg2_ftz(fmaxf(v, 1e-10f)), 0.25f * 0.30102999566398120f, 1.0f);   // (log10(max(v, 1e-10)) + 4) / 4
          if (frame_ok) {
            lmax = fmaxf(lmax, v);
            vmin = fminf(vmin, v);
          }
        }
        s_st[mm * 33 + lane] = v;
      }
      __syncwarp();
      const int rows = int(prm.n_frames - (f0 + 32 * g4) < 32 ? prm.n_frames - (f0 + 32 * g4) : 32);
      if (m0 + lane < prm.n_mels) {
        float* dst = prm.out + (long long)clip * prm.out_clip_stride + (long long)(f0 + 32 * g4) * prm.n_mels + m0 + lane;
        for (int r = 0; r < rows; ++r) dst[(long long)r * prm.n_mels] = s_st[lane * 33 + r];
      }
      const int wmax = __reduce_max_sync(0xffffffffu, enc_ordered(lmax));
      const int wnmin = __reduce_max_sync(0xffffffffu, enc_ordered(-vmin));
      if (lane == 0 && rows > 0) {
        atomicMax(prm.clip_max + clip, wmax);
        atomicMax(prm.tile_min + (long long)clip * prm.tiles32_per_clip + (f0 + 32 * g4) / 32, wnmin);   // negated minimum
      }
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // epilogue (generic) writes before the next tile's TMA writes
    __syncthreads();   // the next tile's staging overwrites the epilogue's regions
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
}

// ---- host side ------------------------------------------------------------------------------------------------------------
struct DeviceOperands {
  void* b = nullptr;          // the constant B operands: kSlices x (quadrant, hi|lo) blocks in canonical UMMA layout
  CUtensorMap map;
  bool ok = false;
};
std::mutex g_mu;
DeviceOperands g_ops[64];

double dft_entry(int q, int nrow, int j) {
  // D_q[frame][j] = sum_n A_q[frame][n] * M_q[n][j]; see the header comment (rows n > 100 and unused columns are zero)
  const double two_pi = 6.283185307179586476925286766559;
  if (nrow > 100) return 0.0;
  double v = 0.0;
  if (q == 0) v = j <= 100 ? cos(two_pi * double((2 * j * nrow) % kN) / kN) : 0.0;
  else if (q == 1) v = j <= 99 ? cos(two_pi * double(((2 * j + 1) * nrow) % kN) / kN) : 0.0;
  else if (q == 2) v = j <= 100 ? -sin(two_pi * double((2 * j * nrow) % kN) / kN) : 0.0;
  else v = j <= 99 ? -sin(two_pi * double(((2 * j + 1) * nrow) % kN) / kN) : 0.0;
  if (nrow == 100) v = (q == 0 || q == 3) ? 0.5 * v : 0.0;
  if (nrow == 0 && q >= 2) v = 0.0;
  if (fabs(v) < 1e-15) v = 0.0;
  return v * double(kScaleF);
}

int build_operands(int dev, std::string* err) {
  DeviceOperands& o = g_ops[dev];
  if (o.ok) return B2A_OK;
  std::vector<__half> host(size_t(kSlices) * kBSlice / 2);
  for (int s = 0; s < kSlices; ++s)
    for (int q = 0; q < kQuads; ++q)
      for (int k = 0; k < 16; ++k)
        for (int j = 0; j < kNP; ++j) {
          const float v = float(dft_entry(q, 16 * s + k, j));
          const __half hi = __float2half_rn(v);
          const __half lo = __float2half_rn(v - __half2float(hi));
          const size_t within = size_t(j % 8) * 16 + size_t(j / 8) * 128 + size_t(k / 8) * (kNP * 16) + size_t(k % 8) * 2;
          const size_t base = size_t(s) * kBSlice + size_t(2 * q) * kBBlock;
          host[(base + within) / 2] = hi;
          host[(base + kBBlock + within) / 2] = lo;
        }
  cudaError_t e = cudaMalloc(&o.b, host.size() * 2);
  if (e != cudaSuccess) {
    if (err) *err = std::string("cudaMalloc: ") + cudaGetErrorString(e);
    return B2A_E_NOMEM;
  }
  if ((e = cudaMemcpy(o.b, host.data(), host.size() * 2, cudaMemcpyHostToDevice)) != cudaSuccess) {
    if (err) *err = std::string("operand upload: ") + cudaGetErrorString(e);
    return B2A_E_CUDA;
  }
  // 2-D view of the operand array: rows of 256 bytes (64 x uint32), one K slice = 112 rows = one TMA box
  typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                               const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if ((e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres)) != cudaSuccess || fn == nullptr) {
    if (err) *err = "cuTensorMapEncodeTiled is not available";
    return B2A_E_CUDA;
  }
  const cuuint64_t dims[2] = {64, cuuint64_t(kSlices) * (kBSlice / 256)};
  const cuuint64_t strides[1] = {256};
  const cuuint32_t box[2] = {64, cuuint32_t(kBSlice / 256)};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = reinterpret_cast<EncodeFn>(fn)(&o.map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 2, o.b, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                                    CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) *err = "cuTensorMapEncodeTiled failed (" + std::to_string(int(r)) + ")";
    return B2A_E_CUDA;
  }
  o.ok = true;
  return B2A_OK;
}

void build_table(const float* window, TcTable& tb) {
  for (int s = 0; s < kSlices; ++s)
    for (int sub = 0; sub < 4; ++sub) {
      const int n0 = 16 * s + 4 * sub;
      tb.off[s][sub][0] = n0;
      tb.off[s][sub][1] = (400 - n0) + (n0 <= 83 ? 2 : 1);
      tb.off[s][sub][2] = (200 - n0) + (n0 <= 43 ? 1 : 0);
      tb.off[s][sub][3] = 200 + n0 + 1;
      for (int i = 0; i < 4; ++i) {
        const int nn = n0 + i;
        const bool in = nn <= 100;
        tb.tap[s][sub][i] = in ? window[nn] * kScaleX : 0.0f;
        tb.tap[s][sub][4 + i] = in && nn >= 1 ? window[400 - nn] * kScaleX : 0.0f;
        tb.tap[s][sub][8 + i] = in ? window[200 - nn] * kScaleX : 0.0f;
        tb.tap[s][sub][12 + i] = in && nn >= 1 ? window[200 + nn] * kScaleX : 0.0f;
      }
    }
}

}  // namespace

// debug hook (tests / bring-up only): when set, the kernel also writes |X|^2 (frames x 201) of every clip there
static float* g_dbg_power = nullptr;
void tc_debug_set_power_buffer(float* device_ptr) { g_dbg_power = device_ptr; }

// The tensor-core path is opt-in (B2A_WHISPER_TC=1 in the environment, or b2a_debug_whisper_tc(1)): see DESIGN.md section 6 for the
// measured go / no-go.
static int g_tc_enabled = -1;
bool tc_whisper_enabled() {
  if (g_tc_enabled < 0) {
    const char* v = getenv("B2A_WHISPER_TC");
    g_tc_enabled = (v != nullptr && v[0] == '1') ? 1 : 0;
  }
  return g_tc_enabled == 1;
}
void tc_whisper_enable(int on) { g_tc_enabled = on ? 1 : 0; }

bool tc_whisper_applicable(const FrontendArgs& a) {
  return a.n_fft == 400 && a.hop == 160 && a.win_len == 400 && a.pre_mode == PRE_NONE && a.spec_mode == SPEC_POWER && a.whisper_norm &&
         a.log_mode == LOG_LOG10 && a.out_mode == OUT_TM && !a.out_f16 && a.clip_tab == nullptr && a.zero_tail == 0 && a.pad_mode == PAD_REFLECT &&
         a.pad_left == 200 && !a.post_affine && a.bank.desc != nullptr && a.bank.n_mels <= 128 && a.log_floor == 1e-10f &&
         a.n_samples >= 2 && a.n_frames > 0;
}

int launch_tc_whisper(const FrontendArgs& a, void* stream, int* launches, std::string* err, float* dbg_power) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) {
    if (err) *err = "device index out of range";
    return B2A_E_CUDA;
  }
  static int n_sm[64] = {};
  {
    std::lock_guard<std::mutex> lk(g_mu);
    int rc = build_operands(dev, err);
    if (rc != B2A_OK) return rc;
    if (n_sm[dev] == 0) {
      cudaError_t e = cudaFuncSetAttribute(tc_whisper_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes);
      if (e != cudaSuccess) {
        if (err) *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
        return B2A_E_CUDA;
      }
      cudaDeviceGetAttribute(&n_sm[dev], cudaDevAttrMultiProcessorCount, dev);
    }
  }
  TcParams prm;
  std::memset(&prm, 0, sizeof(prm));
  prm.x = a.x;
  prm.out = a.out;
  prm.dbg_power = dbg_power != nullptr ? dbg_power : g_dbg_power;
  prm.n_samples = a.n_samples;
  prm.n_frames = a.n_frames;
  prm.n_mels = a.bank.n_mels;
  prm.out_clip_stride = a.n_frames * (long long)a.bank.n_mels;
  prm.n_clips = int(a.batch);
  prm.tiles_per_clip = int((a.n_frames + kM - 1) / kM);
  prm.tiles32_per_clip = frontend_tiles_per_clip(400, a.n_frames);
  prm.total_tiles = (long long)prm.tiles_per_clip * a.batch;
  prm.fb_desc = reinterpret_cast<const int4*>(a.bank.desc);
  prm.fb_w = a.bank.weights;
  prm.clip_max = a.clip_max;
  prm.tile_min = reinterpret_cast<int*>(a.tile_min);
  build_table(a.window, prm.tab);
  cudaError_t e;
  const long long n_tiles32 = (long long)prm.tiles32_per_clip * a.batch;
  if (reinterpret_cast<int*>(a.tile_min) == a.clip_max + a.batch) {
    if ((e = cudaMemsetAsync(a.clip_max, 0x80, sizeof(int) * size_t(a.batch + n_tiles32), st)) != cudaSuccess) goto fail;
  } else {
    if ((e = cudaMemsetAsync(a.clip_max, 0x80, sizeof(int) * size_t(a.batch), st)) != cudaSuccess) goto fail;
    if ((e = cudaMemsetAsync(a.tile_min, 0x80, sizeof(int) * size_t(n_tiles32), st)) != cudaSuccess) goto fail;
  }
  {
    const long long blocks = std::min<long long>(prm.total_tiles, n_sm[dev]);
    tc_whisper_kernel<<<unsigned(blocks), kThreads, kSmemBytes, st>>>(prm, g_ops[dev].map);
    if ((e = cudaGetLastError()) != cudaSuccess) goto fail;
    *launches += 1;
  }
  return launch_whisper_clamp(a.out, a.clip_max, reinterpret_cast<int*>(a.tile_min), a.batch, a.n_frames, a.bank.n_mels, st, launches, err);
fail:
  if (err) *err = std::string("tc_whisper_kernel: ") + cudaGetErrorString(e);
  return B2A_E_CUDA;
}

}  // namespace b2a
