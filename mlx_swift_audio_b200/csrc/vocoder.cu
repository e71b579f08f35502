// Vocoder-side STFT / iSTFT kernels for sm_100a: n_fft 16 / hop 4 (HiFT, CosyVoice3) and
// n_fft 20 / hop 5 (Kokoro iSTFTNet).
//
// Replaces istftHiFiGAN (Codec/S3Gen/HiFiGAN.swift:298-367), cosyVoice3Istft
// (TTS/CosyVoice3/HiFiGAN/CausalHiFTGenerator.swift:463-514), MLXSTFT.inverse
// (TTS/Kokoro/Decoder/MLXSTFT.swift:115-163,211-235) and the matching forward transforms
// (HiFiGAN.swift:257-295, CausalHiFTGenerator.swift:435-460, MLXSTFT.swift:69-113,181-209).
//
// iSTFT design: THREAD == FRAME, lanes are consecutive frames, so the (batch, F, frames) inputs are
// read with perfectly coalesced 128-byte warp loads.  Each thread clips, converts polar->rectangular
// with an in-register sincos (Cody-Waite + minimax polynomials, ~1 ulp), runs a generated
// straight-line half-complex-to-real codelet, applies window/N, and the overlap-add is a GATHER:
// output segment s (hop samples) = y_s[0:h] + y_{s-1}[h:2h] + y_{s-2}[2h:3h] + y_{s-3}[3h:4h], fetched
// from the neighbouring lanes with warp shuffles (cross-warp halo through shared memory).  No
// atomics, no index tensors, deterministic summation order (the frame order of the reference's
// scatter-add), window-envelope normalisation folded into the same pass.
#include <cuda_runtime.h>

#include <cmath>
#include <string>
#include <type_traits>

#include "../../include/b200audio.h"
#include "codelets.h"
#include "internal.h"

namespace b2a {

#define B2A_DEV __device__ __forceinline__

B2A_DEV void c2r(const float (&xr)[9], const float (&xi)[9], float (&y)[16]) { b2a_c2r16(xr, xi, y); }
B2A_DEV void c2r(const float (&xr)[11], const float (&xi)[11], float (&y)[20]) { b2a_c2r20(xr, xi, y); }
B2A_DEV void rdft_small(const float (&x)[16], float (&yr)[9], float (&yi)[9]) { b2a_rdft16(x, yr, yi); }
B2A_DEV void rdft_small(const float (&x)[20], float (&yr)[11], float (&yi)[11]) { b2a_rdft20(x, yr, yi); }

// sin and cos of x, ~1 ulp for |x| < 4.8e4: 3-term Cody-Waite reduction by pi/2 with FMA (quadrant taken
// from the mantissa of the magic-number rounding, no float->int conversion), cephes-style minimax
// polynomials on [-pi/4, pi/4], quadrant fix-up with sign-bit arithmetic.  Branch-free; the caller checks
// the argument range once per frame and uses the libdevice slow path otherwise.
template <bool WANT_SIN>
B2A_DEV void sincos_fast(float x, float* s, float* c) {
  float q = fmaf(x, 0.636619772f, 12582912.0f);  // 1.5 * 2^23: round-to-nearest integer lands in the mantissa
  const int i = __float_as_int(q);
  q -= 12582912.0f;
  float r = fmaf(q, -1.57079601e+00f, x);
  r = fmaf(q, -3.13916473e-07f, r);
  r = fmaf(q, -5.39030253e-15f, r);
  const float r2 = r * r;
  float ps = fmaf(r2, -1.9515295891e-4f, 8.3321608736e-3f);
  ps = fmaf(ps, r2, -1.6666654611e-1f);
  ps = fmaf(ps * r2, r, r);
  float pc = fmaf(r2, 2.443315711809948e-5f, -1.388731625493765e-3f);
  pc = fmaf(pc, r2, 4.166664568298827e-2f);
  pc = fmaf(pc * r2, r2, fmaf(r2, -0.5f, 1.0f));
  const bool odd = (i & 1) != 0;
  const float cc = odd ? ps : pc;
  *c = __int_as_float(__float_as_int(cc) ^ (((i + 1) & 2) << 30));
  if (WANT_SIN) {
    const float ss = odd ? pc : ps;
    *s = __int_as_float(__float_as_int(ss) ^ ((i & 2) << 30));
  }
}

// sin and cos for |x| < pi/2 (what the vocoders feed: phase = sin(.) in [-1, 1]): no range reduction, no quadrant logic.
// Minimax polynomials in x^2 fitted on [0, (1.0005 pi/2)^2]: |err| < 7e-9 (sin, degree 9), 5e-8 (cos, degree 8) before fp32
// rounding; 1.7e-7 / 1.9e-7 evaluated in fp32.
B2A_DEV float cos_small(float u) {  // u = x*x
  float p = fmaf(u, 2.315233542e-05f, -1.385363517e-03f);
  p = fmaf(p, u, 4.166357592e-02f);
  p = fmaf(p, u, -4.999990463e-01f);
  return fmaf(p, u, 9.999999404e-01f);
}
B2A_DEV float sin_small(float x, float u) {
  float p = fmaf(u, 2.605078407e-06f, -1.980901143e-04f);
  p = fmaf(p, u, 8.333050646e-03f);
  p = fmaf(p, u, -1.666665822e-01f);
  return fmaf(p * u, x, x);
}

// sin(x) for the fused vocoder head (phase = sin(conv output), any argument): the Cody-Waite path below 4.8e4, libdevice above
__device__ __noinline__ float sin_huge(float x) { return sinf(x); }   // out of line: keeps the Payne-Hanek code off the hot path
// sin(x), |x| < 4.8e4: x = k*pi + r with r in [-pi/2, pi/2] (magic-number rounding, 3-term Cody-Waite), sin(x) = (-1)^k sin(r)
// with the degree-9 polynomial of sin_small; the parity of k goes straight into the sign bit.
// CHECK = false: the caller has established |x| < 4.8e4 (one test per frame instead of one per bin)
template <bool CHECK = true>
B2A_DEV float sin_any(float x) {
  if (CHECK && !(fabsf(x) < 48000.0f)) return sin_huge(x);
  float q = fmaf(x, 0.318309886f, 12582912.0f);   // 1.5 * 2^23: round(x / pi) lands in the mantissa
  const int k = __float_as_int(q);
  q -= 12582912.0f;
  float r = fmaf(q, -3.14159203e+00f, x);
  r = fmaf(q, -6.27832947e-07f, r);
  r = fmaf(q, -1.07806051e-14f, r);
  const float s = sin_small(r, r * r);
  return __int_as_float(__float_as_int(s) ^ (k << 31));
}
// exp(x) = 2^(x log2 e) on the MUFU, with the rounding error of the product x*log2(e) fed back (first-order): ~1.5 ulp for
// |x| < 88, 0 below (flush to zero), +inf above -- the magnitudes are clipped at 100 right after.
B2A_DEV float exp_fast(float x) {
  const float t = x * 1.44269504f;
  float e = fmaf(x, 1.44269504f, -t);          // exact residual of the product
  e = fmaf(x, 1.92596299e-8f, e);              // + x * (log2 e - fl(log2 e))
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(t));
  return fmaf(y, e * 0.693147182f, y);
}

// atan2(y, x) without libdevice's slow path: octant reduction a = min / max in [0, 1], atan(a) = a * P(a^2) with a degree-7
// near-minimax P (1.4e-7 rad in fp32 over [0, 1]), then the octant / quadrant / sign fix-ups.  atan2(+-0, x >= +0) = +-0,
// atan2(+-0, x <= -0) = +-pi as in IEEE; MLX's arctan2 (MLXSTFT.swift:204) is compared at 2e-5 on mag * exp(i phase).
B2A_DEV float atan2_poly(float y, float x) {
  const float ax = fabsf(x), ay = fabsf(y);
  const float mx = fmaxf(ax, ay), mn = fminf(ax, ay);
  float a = __fdividef(mn, mx);
  if (!(mx > 0.0f)) a = 0.0f;            // atan2(0, 0)
  const float u = a * a;
  float p = fmaf(u, -0.004054217599332333f, 0.021861661225557327f);
  p = fmaf(p, u, -0.05591040849685669f);
  p = fmaf(p, u, 0.0964205339550972f);
  p = fmaf(p, u, -0.13908571004867554f);
  p = fmaf(p, u, 0.1994655430316925f);
  p = fmaf(p, u, -0.33329859375953674f);
  p = fmaf(p, u, 0.9999993443489075f);
  float r = p * a;
  if (ay > ax) r = 1.57079637f - r;
  if (__float_as_int(x) < 0) r = 3.14159274f - r;
  return copysignf(r, y);
}
B2A_DEV float sqrt_fast(float v) {   // MUFU square root (~2 ulp): the magnitude is compared at 1e-5 absolute
  float r;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

template <int NFFT, int HOP>
struct IstftParams {
  const float* mag;
  const float* phase;
  float* out;
  int n_frames;           // < 2^31 / (NFFT/2+1) (checked on the host): per-clip offsets fit 32 bits, one IMAD.WIDE per load address
  long long out_len;
  float clip_lo, clip_hi;
  int use_clip_lo, norm;
  int* unwrap_flag;       // set to 1 if any |dphi| >= pi was seen (Kokoro optimistic path); may be null
  float out_limit;        // HEAD: clip(output, -out_limit, out_limit) (HiFiGAN.swift:587); <= 0 = none
  const float* fade;      // HEAD: fade-in window multiplying the first fade_len output samples of every clip (S3Gen.swift:284-289)
  int fade_len;           //       0 = none
  float wn[NFFT];         // window[n] / NFFT
  float wenv[NFFT];       // window^2 (HiFT/CosyVoice3) or window (Kokoro): envelope contributions
  float inv_env[HOP];     // interior 1 / envelope
};

constexpr int kIstftThreads = 256;

// HEAD: the input is the vocoder's last convolution output h (batch, 2F, frames); magnitude = exp(h[:, :F]),
// phase = sin(h[:, F:]) are formed in registers (HiFiGAN.swift:577-580, Generator.swift:182-183) and the waveform is clipped to
// +-out_limit (HiFiGAN.swift:587) -- the exp / sin pass, the iSTFT and the limiter are one pass over 88 bytes per frame.
// MODE 2 (RECT): the input is the complex spectrum itself, complex64 (batch, F, frames) -- stand-alone mlxIstft
// (MLXSTFT.swift:115-163): no clip, no polar conversion, no unwrap.
enum IstftMode { IM_POLAR = 0, IM_HEAD = 1, IM_RECT = 2 };
template <int NFFT, int HOP, int MODE>
__global__ void __launch_bounds__(kIstftThreads) istft_kernel(const __grid_constant__ IstftParams<NFFT, HOP> prm) {
  constexpr bool HEAD = MODE == IM_HEAD, RECT = MODE == IM_RECT;
  constexpr int F = NFFT / 2 + 1;
  constexpr int R = NFFT / HOP;          // overlapping frames per output segment
  constexpr int HALO = R - 1;
  constexpr int SEG_PER_BLOCK = kIstftThreads - HALO;
  // row pitch of the frame tile: hop 4 uses 128-bit accesses (pitch odd in 16-byte units), hop 5 scalar ones (odd pitch)
  constexpr bool VEC = HOP == 4;
  // hop 5 (20-point frames): rows of five 16-byte units (odd pitch: conflict-free) written as float4; the gather reads the two
  // aligned float4 that cover the five floats it needs from each row (8 LDS.128 instead of 20 LDS)
  constexpr bool VEC5 = HOP == 5 && NFFT == 20;
  constexpr int YP = (VEC || VEC5) ? (((NFFT / 4) % 2 == 1) ? NFFT : NFFT + 4) : NFFT + 1;
  static_assert(NFFT % HOP == 0 && R == 4 && NFFT % 4 == 0, "built for 4x overlap");
  __shared__ __align__(16) float s_y[kIstftThreads * YP];  // windowed frames of the block
  __shared__ float s_stage[HOP == 4 ? 1 : kIstftThreads * HOP];   // hop 5: per-warp output staging

  const int tid = threadIdx.x, lane = tid & 31;
  const long long clip = blockIdx.y;
  const int nF = prm.n_frames;
  // Row stride of the (F, frames) inputs as a load-address operand.  32 bits (one IMAD.WIDE per address) for the 20-point
  // transform; 64 bits for the 16-point one, whose loads then issue interleaved with the longer address arithmetic --
  // measured 10 % faster there (the kernel runs at ~90 % of the HBM roofline and is sensitive to how its 18 loads are paced).
  using stride_t = typename std::conditional<NFFT == 16, long long, unsigned>::type;
  const stride_t nFu = stride_t(nF);
  const int nSeg = nF + HALO;                            // segments of the untrimmed OLA buffer
  const int seg0 = int(blockIdx.x) * SEG_PER_BLOCK;      // first segment this block emits
  const int f = seg0 - HALO + tid;                       // frame (== segment) of this thread
  const bool has_frame = f >= 0 && f < nF;

  // ---- per-frame inverse real FFT, windowed ------------------------------------------------------
  {
    float y[NFFT];
    const long long clip_rows = HEAD ? 2 * F : F;   // rows of `frames` floats per clip in the source tensor(s)
    const float* __restrict__ mp = prm.mag + (clip * clip_rows * nF + f);
    const float* __restrict__ pp = prm.phase + (clip * clip_rows * nF + f);
    float xr[F], xi[F], ph[F];
    float amax = 0.0f;
    if (RECT) {
      const float2* __restrict__ cp = reinterpret_cast<const float2*>(prm.mag) + (clip * F * nF + f);
#pragma unroll
      for (int k = 0; k < F; ++k) {
        const float2 v = has_frame ? __ldg(cp + k * nFu) : make_float2(0.0f, 0.0f);
        xr[k] = v.x;
        xi[k] = v.y;   // (DC / Nyquist imaginary parts are ignored by the codelet, as by irfft)
      }
    } else if (HEAD) {
      // all 2F loads first (memory-level parallelism), then the exp / sin arithmetic
#pragma unroll
      for (int k = 0; k < F; ++k) {
        xr[k] = has_frame ? __ldg(mp + k * nFu) : -200.0f;   // exp(-200) flushes to 0: frames outside the clip contribute nothing
        ph[k] = has_frame ? __ldg(pp + k * nFu) : 0.0f;
      }
      float hmax = 0.0f;
#pragma unroll
      for (int k = 0; k < F; ++k) hmax = fmaxf(hmax, fabsf(ph[k]));
      // (fmaxf drops NaN operands: a NaN bin rides the fast path, which propagates it like sinf does; +-inf takes the slow one)
      const bool in_range = hmax < 48000.0f;
#pragma unroll
      for (int k = 0; k < F; ++k) {
        float m = fminf(exp_fast(xr[k]), prm.clip_hi);
        if (prm.use_clip_lo) m = fmaxf(m, prm.clip_lo);
        xr[k] = m;
      }
      if (in_range) {
#pragma unroll
        for (int k = 0; k < F; ++k) ph[k] = sin_any<false>(ph[k]);
      } else {
#pragma unroll
        for (int k = 0; k < F; ++k) ph[k] = sin_any<true>(ph[k]);
      }
    } else {
#pragma unroll
      for (int k = 0; k < F; ++k) {
        float m = has_frame ? __ldg(mp + k * nFu) : 0.0f;
        ph[k] = has_frame ? __ldg(pp + k * nFu) : 0.0f;
        m = fminf(m, prm.clip_hi);
        if (prm.use_clip_lo) m = fmaxf(m, prm.clip_lo);
        xr[k] = m;
        amax = fmaxf(amax, fabsf(ph[k]));
      }
    }
    // every phase of the warp's 32 frames inside (-pi/2, pi/2)?  (false for NaN)  Then (a) sin / cos need no range
    // reduction and (b) no phase step inside the warp can reach pi.
    // (HEAD: |phase| <= 1 by construction, or NaN, which the polynomials propagate like the reference's cos / sin)
    const bool small = RECT || HEAD || __all_sync(0xffffffffu, amax < 1.5707963f);
    if (RECT) {
      // rectangular input: nothing to convert
    } else
    if (NFFT == 20 && prm.unwrap_flag != nullptr && !small) {   // (only Kokoro's 20 / 5 transform unwraps)
      // Kokoro's unwrap (MLXSTFT.swift:23-46) is the identity unless some |phase[t] - phase[t-1]| >= pi.  A pair (t-1, t) is
      // examined by the warp of frame t (left neighbour: shuffle, lane 0 from global memory) and, when t-1 is a warp's last
      // lane, also by that warp (right neighbour from global memory) -- so warps whose phases are all small can skip the test
      // even when their neighbour warp's are not.  (warp-uniform branch; every lane takes part in the shuffles)
      bool bad = false;
#pragma unroll
      for (int k = 0; k < F; ++k) {
        float prev = __shfl_up_sync(0xffffffffu, ph[k], 1);
        if (has_frame && f > 0) {
          if (lane == 0) prev = __ldg(pp + k * nFu - 1);
          bad |= !(fabsf(ph[k] - prev) < 3.14159274f);
        }
        if (lane == 31 && has_frame && f + 1 < nF) bad |= !(fabsf(__ldg(pp + k * nFu + 1) - ph[k]) < 3.14159274f);
      }
      if (bad) *prm.unwrap_flag = 1;
    }
    if (RECT) {
    } else if (small) {
#pragma unroll
      for (int k = 0; k < F; ++k) {
        const float u = ph[k] * ph[k];
        const float m = xr[k];
        xr[k] = m * cos_small(u);
        // imaginary parts of DC / Nyquist are ignored by the codelet, as by irfft: no sine needed there
        xi[k] = (k == 0 || k == F - 1) ? 0.0f : m * sin_small(ph[k], u);
      }
    } else if (amax < 48000.0f) {  // false for NaN too
      float sn, cs;
      sincos_fast<false>(ph[0], &sn, &cs);
      xr[0] *= cs; xi[0] = 0.0f;
      sincos_fast<false>(ph[F - 1], &sn, &cs);
      xr[F - 1] *= cs; xi[F - 1] = 0.0f;
#pragma unroll
      for (int k = 1; k < F - 1; ++k) {
        sincos_fast<true>(ph[k], &sn, &cs);
        const float m = xr[k];
        xr[k] = m * cs;
        xi[k] = m * sn;
      }
    } else {
#pragma unroll
      for (int k = 0; k < F; ++k) {  // rare: huge / non-finite phases (fully unrolled: keeps the arrays in registers)
        float sn, cs;
        sincosf(ph[k], &sn, &cs);
        const float m = xr[k];
        xr[k] = m * cs;
        xi[k] = m * sn;
      }
    }
    c2r(xr, xi, y);
    // frames outside [0, nF) have magnitude 0 -> exact zeros
    if (VEC || VEC5) {
      float4* row = reinterpret_cast<float4*>(s_y + tid * YP);
#pragma unroll
      for (int q = 0; q < NFFT / 4; ++q)
        row[q] = make_float4(y[4 * q] * prm.wn[4 * q], y[4 * q + 1] * prm.wn[4 * q + 1], y[4 * q + 2] * prm.wn[4 * q + 2],
                             y[4 * q + 3] * prm.wn[4 * q + 3]);
    } else {
#pragma unroll
      for (int n = 0; n < NFFT; ++n) s_y[tid * YP + n] = y[n] * prm.wn[n];
    }
  }
  __syncthreads();

  // ---- overlap-add as a gather: segment s = y_s[0:h] + y_{s-1}[h:2h] + y_{s-2}[2h:3h] + y_{s-3}[3h:4h] -------------
  // (added in that order: the frame order of the reference's scatter-add)
  const int s = f;
  const bool emit = tid >= HALO && s >= R / 2 && s < nSeg - R / 2;
  float o[HOP];
  if (emit) {
    if (VEC) {
      const float4 a0 = *reinterpret_cast<const float4*>(s_y + tid * YP);
      o[0] = a0.x; o[1] = a0.y; o[2] = a0.z; o[3] = a0.w;
#pragma unroll
      for (int r = 1; r <= HALO; ++r) {
        const float4 a = *reinterpret_cast<const float4*>(s_y + (tid - r) * YP + r * HOP);
        o[0] += a.x; o[1] += a.y; o[2] += a.z; o[3] += a.w;
      }
    } else if (VEC5) {
#pragma unroll
      for (int r = 0; r <= HALO; ++r) {
        const float4* row = reinterpret_cast<const float4*>(s_y + (tid - r) * YP) + (r * HOP) / 4;
        const float4 a = row[0], b = row[1];
        const float t[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
        for (int i = 0; i < HOP; ++i) o[i] = r == 0 ? t[i] : o[i] + t[(r * HOP) % 4 + i];
      }
    } else {
#pragma unroll
      for (int i = 0; i < HOP; ++i) o[i] = s_y[tid * YP + i];
#pragma unroll
      for (int r = 1; r <= HALO; ++r) {
#pragma unroll
        for (int i = 0; i < HOP; ++i) o[i] += s_y[(tid - r) * YP + r * HOP + i];
      }
    }
    const bool interior = s >= HALO && s <= nF - 1;
    if (interior) {
#pragma unroll
      for (int i = 0; i < HOP; ++i) o[i] *= prm.inv_env[i];
    } else {
#pragma unroll
      for (int i = 0; i < HOP; ++i) {
        float e = 0.0f;
#pragma unroll
        for (int r = 0; r < R; ++r)
          if (s - r >= 0 && s - r < nF) e += prm.wenv[r * HOP + i];
        if (prm.norm == NORM_WSQ_FLOOR) o[i] = o[i] / fmaxf(e, 1e-8f);
        else o[i] = e != 0.0f ? o[i] / e : o[i];
      }
    }
  }
  if (HEAD && prm.out_limit > 0.0f) {
#pragma unroll
    for (int i = 0; i < HOP; ++i) o[i] = fminf(fmaxf(o[i], -prm.out_limit), prm.out_limit);
  }
  if (HEAD && emit && (s - R / 2) * HOP < prm.fade_len) {   // only the first blocks of a clip get here
    const int t0 = (s - R / 2) * HOP;
#pragma unroll
    for (int i = 0; i < HOP; ++i)
      if (t0 + i < prm.fade_len) o[i] *= __ldg(prm.fade + t0 + i);
  }
  float* __restrict__ dst = prm.out + clip * prm.out_len;
  if (HOP == 4) {
    if (emit) {
      float* d = dst + (long long)(s - R / 2) * HOP;
      if ((reinterpret_cast<uintptr_t>(d) & 15) == 0) {
        *reinterpret_cast<float4*>(d) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int i = 0; i < HOP; ++i) d[i] = o[i];
      }
    }
  } else {
    // hop 5: a segment is 20 bytes, so the warp stages its 32 segments (160 floats, lane stride 5: conflict-free) in its own
    // rows of a separate staging buffer and writes them out as contiguous 128-byte runs -- no block barrier (the frame tile
    // s_y is still being read by other warps), only __syncwarp.  The emitting lanes of a warp are a contiguous lane range.
    if (emit) {
#pragma unroll
      for (int i = 0; i < HOP; ++i) s_stage[tid * HOP + i] = o[i];
    }
    __syncwarp();
    const unsigned em = __ballot_sync(0xffffffffu, emit);
    if (em != 0u) {
      const int la = __ffs(em) - 1, n = __popc(em) * HOP;
      const float* src = s_stage + ((tid - lane) + la) * HOP;
      float* d = dst + (long long)(s - lane + la - R / 2) * HOP;   // first emitted segment of the warp
      for (int i = lane; i < n; i += 32) d[i] = src[i];
    }
  }
}

// ---- Kokoro slow path: numpy-style unwrap along time (MLXSTFT.swift:23-46), one warp per (clip, bin) row
__global__ void unwrap_kernel(const float* __restrict__ phase, float* __restrict__ out, long long n_frames, long long n_rows) {
  const long long row = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= n_rows) return;
  const int lane = threadIdx.x & 31;
  const float* p = phase + row * n_frames;
  float* o = out + row * n_frames;
  const float period = 6.28318548f, hi = 3.14159274f, lo = -3.14159274f;
  float carry = 0.0f;   // running sum of corrections
  float prev_last = 0.0f;
  for (long long t0 = 0; t0 < n_frames; t0 += 32) {
    const long long t = t0 + lane;
    const float v = t < n_frames ? p[t] : 0.0f;
    float prev = __shfl_up_sync(0xffffffffu, v, 1);
    if (lane == 0) prev = prev_last;
    float corr = 0.0f;
    if (t > 0 && t < n_frames) {
      const float d = v - prev;
      float dm = d - lo;
      dm = fmodf(fmodf(dm, period) + period, period) + lo;
      if (dm == lo && d > 0.0f) dm = hi;
      corr = fabsf(d) < hi ? 0.0f : dm - d;
    }
    // inclusive warp scan of the corrections
    float sc = corr;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const float n = __shfl_up_sync(0xffffffffu, sc, d);
      if (lane >= d) sc += n;
    }
    sc += carry;
    if (t < n_frames) o[t] = t == 0 ? v : v + sc;
    carry = __shfl_sync(0xffffffffu, sc, 31);
    prev_last = __shfl_sync(0xffffffffu, v, 31);
  }
}

// ------------------------------------------------------------------------------------------------
// small forward STFT: thread == frame
// ------------------------------------------------------------------------------------------------
template <int NFFT, int HOP>
struct SmallStftParams {
  const float* x;
  float* out0;
  float* out1;
  long long n_samples, n_frames;
  int pad_mode, out_kind;
  float w[NFFT];
};

// POLAR: magnitude / phase outputs (Kokoro's transform) instead of real / imaginary parts -- a template parameter so that the two
// forms do not share one register allocation
template <int NFFT, int HOP, bool POLAR>
__global__ void __launch_bounds__(256) small_stft_kernel(const __grid_constant__ SmallStftParams<NFFT, HOP> prm) {
  constexpr int F = NFFT / 2 + 1;
  constexpr int PAD = NFFT / 2;
  const long long clip = blockIdx.y;
  const long long f = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (f >= prm.n_frames) return;
  const long long T = prm.n_samples;
  const float* __restrict__ xc = prm.x + clip * T;
  float in[NFFT];
  const long long j0 = f * HOP - PAD;
  if (j0 >= 0 && j0 + NFFT <= T) {
    if (HOP % 4 == 0 && NFFT % 4 == 0 && (reinterpret_cast<uintptr_t>(xc + j0) & 15) == 0) {
      // hop 4: a frame starts on a 16-byte boundary whenever the clip does -- NFFT/4 coalesced 16-byte loads per thread
      const float4* __restrict__ x4 = reinterpret_cast<const float4*>(xc + j0);
#pragma unroll
      for (int q = 0; q < NFFT / 4; ++q) {
        const float4 v = __ldg(x4 + q);
        in[4 * q] = v.x * prm.w[4 * q];
        in[4 * q + 1] = v.y * prm.w[4 * q + 1];
        in[4 * q + 2] = v.z * prm.w[4 * q + 2];
        in[4 * q + 3] = v.w * prm.w[4 * q + 3];
      }
    } else {
#pragma unroll
      for (int n = 0; n < NFFT; ++n) in[n] = __ldg(xc + j0 + n) * prm.w[n];
    }
  } else {
#pragma unroll
    for (int n = 0; n < NFFT; ++n) {
      long long j = j0 + n;
      float v = 0.0f;
      if (j >= 0 && j < T) v = __ldg(xc + j);
      else if (prm.pad_mode == PAD_REFLECT) {
        j = j < 0 ? -j : 2 * (T - 1) - j;   // T > PAD is checked on the host
        v = __ldg(xc + j);
      }
      in[n] = v * prm.w[n];
    }
  }
  float yr[F], yi[F];
  rdft_small(in, yr, yi);
  float* o0 = prm.out0 + clip * F * prm.n_frames + f;
  float* o1 = prm.out1 + clip * F * prm.n_frames + f;
#pragma unroll
  for (int k = 0; k < F; ++k) {
    if (!POLAR) {
      o0[k * prm.n_frames] = yr[k];
      o1[k * prm.n_frames] = yi[k];
    } else {
      o0[k * prm.n_frames] = sqrt_fast(yr[k] * yr[k] + yi[k] * yi[k]);
      o1[k * prm.n_frames] = atan2_poly(yi[k], yr[k]);
    }
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int cuda_fail2(cudaError_t e, const char* what, std::string* err) {
  if (err) *err = std::string(what) + ": " + cudaGetErrorString(e);
  return B2A_E_CUDA;
}

template <int NFFT, int HOP>
static int launch_istft_t(const IstftArgs& a, cudaStream_t st, int* launches, std::string* err, const float* phase,
                          int* unwrap_flag) {
  constexpr int F = NFFT / 2 + 1;
  constexpr int R = NFFT / HOP;
  IstftParams<NFFT, HOP> prm;
  prm.mag = a.mag;
  prm.phase = phase;
  prm.out = a.out;
  if (a.n_frames >= (1LL << 31) / (4 * (NFFT / 2 + 1)) - 1024) {
    if (err) *err = "iSTFT: too many frames per clip";
    return B2A_E_BAD_ARG;
  }
  prm.n_frames = int(a.n_frames);
  prm.out_len = (a.n_frames - 1) * HOP;
  prm.clip_lo = a.clip_lo;
  prm.clip_hi = a.clip_hi;
  prm.use_clip_lo = a.use_clip_lo;
  prm.norm = a.norm;
  prm.unwrap_flag = unwrap_flag;
  prm.out_limit = a.out_limit;
  prm.fade = a.fade;
  prm.fade_len = a.head == 1 && a.fade != nullptr ? a.fade_len : 0;
  if (a.head == 1) {   // one tensor (batch, 2F, frames): the phase rows follow the magnitude rows of the same clip
    prm.phase = a.mag + (long long)F * a.n_frames;
  }
  for (int n = 0; n < NFFT; ++n) {
    prm.wn[n] = a.window[n] / float(NFFT);
    prm.wenv[n] = a.norm == NORM_WSQ_FLOOR ? a.window[n] * a.window[n] : a.window[n];
  }
  for (int i = 0; i < HOP; ++i) {
    float e = 0.0f;
    for (int r = 0; r < R; ++r) e += prm.wenv[r * HOP + i];  // same order as the scatter-add by frame
    if (a.norm == NORM_WSQ_FLOOR) e = e > 1e-8f ? e : 1e-8f;
    prm.inv_env[i] = e != 0.0f ? 1.0f / e : 1.0f;
  }
  const long long n_seg = a.n_frames + R - 1;
  const int per_block = kIstftThreads - (R - 1);
  dim3 grid(unsigned((n_seg + per_block - 1) / per_block), unsigned(a.batch));
  // (residency: 5 blocks per SM by registers; capping it lower with unused dynamic shared memory only costs -- measured 1.371 ms at 5
  // blocks, 1.373 at 4, 1.74 at 3 for the 16 / 4 transform)
  if (a.head == 2) istft_kernel<NFFT, HOP, IM_RECT><<<grid, kIstftThreads, 0, st>>>(prm);
  else if (a.head) istft_kernel<NFFT, HOP, IM_HEAD><<<grid, kIstftThreads, 0, st>>>(prm);
  else istft_kernel<NFFT, HOP, IM_POLAR><<<grid, kIstftThreads, 0, st>>>(prm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail2(e, "istft_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

int launch_istft(const IstftArgs& a, void* stream, int* launches, std::string* err) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a.n_frames < 2) {
    if (err) *err = "iSTFT needs at least 2 frames";
    return B2A_E_TOO_SHORT;
  }
  auto run = [&](const float* phase, int* flag) -> int {
    if (a.n_fft == 16 && a.hop == 4) return launch_istft_t<16, 4>(a, st, launches, err, phase, flag);
    if (a.n_fft == 20 && a.hop == 5) return launch_istft_t<20, 5>(a, st, launches, err, phase, flag);
    return -1;
  };
  auto run_unwrapped = [&]() -> int {
    const long long rows = a.batch * (a.n_fft / 2 + 1);
    unwrap_kernel<<<unsigned((rows + 7) / 8), 256, 0, st>>>(a.phase, a.scratch_phase, a.n_frames, rows);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return cuda_fail2(e, "unwrap_kernel launch", err);
    *launches += 1;
    return run(a.scratch_phase, nullptr);
  };
  int rc;
  if (a.unwrap == 0) {
    rc = run(a.phase, nullptr);
  } else if (a.unwrap == 2 && a.d_flag != nullptr && a.h_flag != nullptr) {
    // Optimistic: unwrap is the identity unless a phase step >= pi exists.  The kernel raises a flag when it
    // sees one; only then is the slow path (fp32 cumsum of corrections, then iSTFT again) taken.  The host
    // reads the flag, i.e. this entry point synchronises -- as the reference's own eval() calls do
    // (MLXSTFT.swift:219,229).
    cudaError_t e = cudaMemsetAsync(a.d_flag, 0, sizeof(int), st);
    if (e != cudaSuccess) return cuda_fail2(e, "memset", err);
    rc = run(a.phase, a.d_flag);
    if (rc != B2A_OK) goto done;
    if ((e = cudaMemcpyAsync(a.h_flag, a.d_flag, sizeof(int), cudaMemcpyDeviceToHost, st)) != cudaSuccess) return cuda_fail2(e, "flag copy", err);
    if ((e = cudaStreamSynchronize(st)) != cudaSuccess) return cuda_fail2(e, "sync", err);
    if (*a.h_flag != 0) rc = run_unwrapped();
  } else {
    rc = run_unwrapped();
  }
done:
  if (rc >= 0) return rc;
  if (err) *err = "iSTFT built for (n_fft, hop) = (16, 4) and (20, 5)";
  return B2A_E_UNSUPPORTED;
}

template <int NFFT, int HOP>
static int launch_small_stft_t(const SmallStftArgs& a, cudaStream_t st, int* launches, std::string* err) {
  SmallStftParams<NFFT, HOP> prm;
  prm.x = a.x;
  prm.out0 = a.out0;
  prm.out1 = a.out1;
  prm.n_samples = a.n_samples;
  prm.n_frames = a.n_frames;
  prm.pad_mode = a.pad_mode;
  prm.out_kind = a.out_kind;
  for (int n = 0; n < NFFT; ++n) prm.w[n] = a.window[n];
  dim3 grid(unsigned((a.n_frames + 255) / 256), unsigned(a.batch));
  // Residency cap of the real / imaginary hop-4 kernel (HiFT source STFT, 88 B per frame, no arithmetic to speak of): with all 8
  // blocks per SM resident its 19 address streams per block thrash the DRAM pages -- measured on 512 x 30 s: 8 blocks 1.857 ms,
  // 7: 1.62, 6: 1.61, 5: 1.52, 4: 1.51, 3: 1.68.  The unused dynamic shared memory only limits the residency to 4 blocks.
  // (The magnitude / phase kernel has the arithmetic to hide behind and is best left alone: 1.71 ms at 8 blocks, 1.76 at 4.)
  size_t pad = 0;
  if (NFFT == 16 && HOP == 4 && a.out_kind == SOUT_REAL_IMAG) {
    pad = 56000;
    static bool attr_set[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !attr_set[dev]) {
      cudaError_t ea = cudaFuncSetAttribute(small_stft_kernel<NFFT, HOP, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(pad));
      if (ea != cudaSuccess) return cuda_fail2(ea, "cudaFuncSetAttribute", err);
      attr_set[dev] = true;
    }
  }
  if (a.out_kind == SOUT_REAL_IMAG) small_stft_kernel<NFFT, HOP, false><<<grid, 256, pad, st>>>(prm);
  else small_stft_kernel<NFFT, HOP, true><<<grid, 256, 0, st>>>(prm);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail2(e, "small_stft_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

int launch_unwrap(const float* phase, float* out, int64_t n_rows, int64_t n_frames, void* stream, int* launches, std::string* err) {
  unwrap_kernel<<<unsigned((n_rows + 7) / 8), 256, 0, static_cast<cudaStream_t>(stream)>>>(phase, out, n_frames, n_rows);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail2(e, "unwrap_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

int launch_small_stft(const SmallStftArgs& a, void* stream, int* launches, std::string* err) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (a.n_fft == 16 && a.hop == 4) return launch_small_stft_t<16, 4>(a, st, launches, err);
  if (a.n_fft == 20 && a.hop == 5) return launch_small_stft_t<20, 5>(a, st, launches, err);
  if (err) *err = "vocoder STFT built for (n_fft, hop) = (16, 4) and (20, 5)";
  return B2A_E_UNSUPPORTED;
}

}  // namespace b2a
