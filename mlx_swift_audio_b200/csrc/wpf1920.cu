// Warp-per-frame front end for n_fft = 1920, hop = 480 (S3Gen 24 kHz mel, S3GenMel.swift:43-102) on sm_100a.
//
// Why a second layout.  frontend.cu keeps LANE == FRAME: a CTA tile is 16 frames of 1920 points, its exchange buffer alone is
// 123 KB, so one CTA of 16 warps per SM, four all-to-all CTA barriers per tile, 25 % occupancy and 49 % issue (0.465 ms for
// 256 x 10 s, profiles/r01_frontend_s3gen_ncu.txt).  Here ONE WARP owns ONE FRAME and nothing is shared between warps:
//   * 1920 = 60 x 32.  Stage A: lane n2 runs the real 60-point DFT of samples 32 n1 + n2 -- the 32 lanes read 32 consecutive
//     samples, so the PCM comes straight from global memory as aligned 128-byte lines (the 4x frame overlap of a strip of
//     consecutive frames is served by the L1 / L2), no shared-memory staging of PCM at all;
//   * inter-stage twiddles and window values differ per LANE (not per warp), so they come from conflict-free shared-memory tables
//     ([k1][lane] float2, [lane][60] read as float4) instead of immediates;
//   * one exchange through the warp's own 8.4 KB of shared memory (row pitch 34 float2: 8-byte stores stride-1 across lanes, 16-byte
//     loads at a lane pitch of 68 words, both conflict-free), __syncwarp only -- the kernel has NO CTA barrier after its prologue;
//   * stage B: lane k1 (0..30) runs the complex 32-point DFT of row k1 -> bins k1 + 60 k2 (k2 < 16) and, mirrored, 1920 - k1 - 60 k2:
//     every lane executes the same codelet (k1 = 0 with zero imaginary parts, k1 = 30 from the twiddled Nyquist value), lane 31 idles;
//   * |X| (MUFU sqrt) or |X|^2 goes bin-major into the same shared memory; the mel projection runs over host-built balanced segments
//     (build_wpf_mel: 17 steps of four products per lane and frame for the 80-mel bank, 16-byte loads, conflict-free), then log /
//     affine and 4-byte stores into the (M, T') rows;
//   * work units are strips of 8 consecutive frames of one clip, dealt round-robin to the warps of the 148 persistent CTAs (one
//     contiguous range of frames per warp balances the tail better but measured 7 % slower: 2368 streams 104 KB apart keep as many
//     DRAM rows open; requesting the next frame's samples ahead of the mel projection spills at 128 registers and measured 7 % slower;
//     strips handed to the warps by a launch-wide atomic counter -- what the tiled kernels of frontend.cu gain 5 % from -- measured 3 % slower
//     here, 0.350 vs 0.339 ms: the 16 warps of a CTA then no longer work on 16 neighbouring strips whose frames overlap in the L1; one
//     contiguous range of strips per CTA with a shared-memory counter for its warps -- neighbours kept, ranges equal to within one strip --
//     measured the same as the round-robin deal, 0.337 ms either way: the tail is not what bounds this kernel);
//   * the frame is rotated by rot = pad_left mod 32 samples (x'[m] = x[(m + rot) mod 1920], window rotated alike) so that the 128-byte
//     lines are aligned: a circular shift changes the phase of X[k] only, and only |X| is used.
// Reflect / zero padded edge frames (4 of 500) are staged sample by sample through the padding index map (pad_index.cuh).
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200audio.h"
#include "codelets.h"
#include "internal.h"
#include "pad_index.cuh"

#ifndef B2A_WPF_WARPS
#define B2A_WPF_WARPS 16
#endif
// inter-stage twiddles are loaded in batches of this many, all in flight before the first complex multiply of the batch
// (measured on 256 x 10 s: 1 -> 0.368 ms, 6 -> 0.362 ms, 10 -> 0.359 ms)
#ifndef B2A_WPF_TW_BATCH
#define B2A_WPF_TW_BATCH 10
#endif

namespace b2a {
namespace {

constexpr int kN = 1920, kN1 = 60, kHop = 480, kStrip = 8, kRaggedTile = 16;
constexpr int kExPitch = 34;                      // float2 per exchange row (k1)
constexpr int kExWords = 31 * kExPitch * 2;       // 2108 floats: rows k1 = 0..30; later the 961-bin spectrum
constexpr int kPartWords = kWpfRounds * 32 + 4;   // segment sums + the zero slot
constexpr int kWarpWords = kExWords + kPartWords;
static_assert(kWarpWords % 4 == 0 && kExWords % 4 == 0, "16-byte aligned warp regions");
static_assert(kPlanShapes[2].n_fft == kN && kPlanShapes[2].frame_tile == kRaggedTile, "ragged work units are the tile table's 16-frame tiles of the 1920-point plan");
constexpr int kTableWords = kN + 2 * 30 * 32 + kWpfMelMaxWords;   // window, twiddles (k1 = 1..30), mel schedule

__constant__ float2 c_wpf_tw[30 * 32];   // W_1920^{n2 k1} as (cos, sin of the negative angle), [k1 - 1][n2], k1 = 1..30

struct WpfParams {
  const float* x;
  float* out;
  const uint32_t* mel;
  const int4* tile_tab;   // ragged batches: work unit g = (clip, 16-frame tile of the clip, the clip's n_samples, n_frames); null = equal lengths
  long long clip_stride, n_samples, zero_tail, pad_left, n_frames, out_clip_stride, total_strips;
  int strips_per_clip, pad_mode, log_mode, post_affine, n_mels, rot, mel_words;
  float log_floor, post_sub, post_div;
  float window[kN];
};

// L2 prefetch of a 128-byte line (no register, no scoreboard): the next frame's samples are requested one frame (~2.5 us) ahead
#ifndef B2A_WPF_PREFETCH
#define B2A_WPF_PREFETCH 1
#endif
B2A_DEV void wpf_prefetch(const float* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
B2A_DEV float wpf_sqrt(float x) {
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
B2A_DEV float wpf_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Edge frames: every sample through the padding index map into the warp's exchange buffer, in the order the interior path reads
// global memory (buf[i] = frame sample (rot + i) mod 1920).  Out of line: the 64-bit index arithmetic stays out of the hot loop.
__device__ __noinline__ void wpf_stage_edge(const float* __restrict__ xc, float* __restrict__ buf, long long p0, long long pad_left,
                                            long long n_samples, long long n_eff, int pad_mode, int rot, int lane) {
  for (int i = lane; i < kN; i += 32) {
    int m = rot + i;
    if (m >= kN) m -= kN;
    buf[i] = fetch_padded(xc, p0 + m, pad_left, n_samples, n_eff, pad_mode);
  }
}

// PRUNE: the bank reads no bin above kPruneBins - 1 (S3Gen's 80 filters end at 8 kHz = bin 640 of 961).  Bin k1 + 60 k2 or its mirror
// image 1920 - k1 - 60 k2 then lies above that limit for k2 = 11..20 in EVERY lane: those ten outputs of the 32-point codelet are never
// formed (the compiler drops the butterflies that only feed them) and their |X| / store disappear: -6 % instructions per frame.  The mel
// schedule's 16-byte reads may still touch words above the limit (zero weights): they hold this frame's exchange values or the
// never-written pad words of the exchange rows, zeroed once per kernel -- finite, so 0 * v == 0.
constexpr int kPruneBins = 641, kPruneLo = 11, kPruneHi = 20, kPruneDummy = 2848;
static_assert(kPruneDummy % 32 == 0 && kPruneDummy - 60 * 31 >= kN / 2 + 4 && kPruneDummy + 2 - 60 * (kPruneHi + 1) < kExWords, "dummy mirror stores stay behind the spectrum, inside the warp's buffer");
static_assert((9 * kExPitch * 2 + 64) / 4 * 4 >= (kPruneBins & ~3) && 14 * kExPitch * 2 + 64 >= kN / 2 + 4, "pad words of rows 9..13 are the ones inside [640, 964)");
static_assert(60 * kPruneLo >= kPruneBins && kN - 29 - 60 * kPruneHi >= kPruneBins, "pruned outputs are dead in every lane");
static_assert(60 * (kPruneLo - 1) + 30 < kPruneBins, "kept direct outputs are alive in every lane (no partial prune below the band)");

template <int NW, bool MAG, bool RAGGED, bool PRUNE>
__global__ void __launch_bounds__(NW * 32, 1) wpf1920_kernel(const __grid_constant__ WpfParams prm) {
  extern __shared__ __align__(16) float smem[];
  float* s_win = smem;                                              // [lane][60]: window of the lane's samples (rot + 32 n1 + lane) mod 1920
  float2* s_tw = reinterpret_cast<float2*>(smem + kN);              // [k1 - 1][lane]
  uint32_t* s_mel = reinterpret_cast<uint32_t*>(smem + kN + 2 * 30 * 32);
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  float* s_ex = smem + kTableWords + warp * kWarpWords;
  float* s_part = s_ex + kExWords;
  const int rot = prm.rot;

  // ---- tables (once per persistent CTA) ----
  // window: warp-uniform reads of the kernel parameter (a lane-dependent index into the parameter space is served one lane at a time:
  // ncu attributed 4 % of the first version's stall samples to that loop); sample m of the frame belongs to lane (m - rot) mod 32, slot (m - rot) / 32
  for (int m = warp; m < kN; m += NW) {
    const float v = prm.window[m];
    int i = m - rot;
    if (i < 0) i += kN;
    if (lane == 0) s_win[(i & 31) * kN1 + (i >> 5)] = v;
  }
  for (int i = tid; i < 30 * 32; i += NW * 32) s_tw[i] = c_wpf_tw[i];
  for (int i = tid; i < prm.mel_words; i += NW * 32) s_mel[i] = __ldg(prm.mel + i);
  if (PRUNE)
    for (int i = lane; i < kExWords; i += 32) s_ex[i] = 0.0f;   // (the pad words of the exchange rows are never written again)
  __syncthreads();

  const int M = prm.n_mels, rounds = int(s_mel[1]);
  if (lane < 4) s_part[rounds * 32 + lane] = 0.0f;   // the zero slot of the segment table (never written by a round)
  __syncwarp();
  const int4* s_fin = reinterpret_cast<const int4*>(s_mel + s_mel[11]);
  const uint32_t* s_start = s_mel + s_mel[10];
  const int log_mode = prm.log_mode;
  const float log_floor = prm.log_floor;
  const bool affine = prm.post_affine != 0;
  const long long n_frames = prm.n_frames;   // row stride of the (M, T') output
  const int spc = prm.strips_per_clip;
  // sample n1 = 59 of the lanes whose rotated position runs past the frame wraps to the frame's first samples
  const int last_off = lane >= 32 - rot ? 59 * 32 - kN : 59 * 32;
  const int k1 = lane < 31 ? lane : 30;   // (lane 31 repeats row 30 and stores nothing)

  // Work units: strips of kStrip consecutive frames of one clip (equal lengths), or the 16-frame tiles of the ragged batch's tile table
  // (tile_table_kernel of frontend.cu: one 16-byte entry per unit carries the clip's own length and frame count); output strides stay
  // those of the longest clip.
  long long n_frames_c = n_frames, n_samples = prm.n_samples;
  for (long long s = (long long)blockIdx.x * NW + warp; s < prm.total_strips; s += (long long)gridDim.x * NW) {
    long long clip;
    int t0;
    if (RAGGED) {
      const int4 u = __ldg(prm.tile_tab + s);
      clip = u.x;
      t0 = u.y * kRaggedTile;
      n_samples = u.z;
      n_frames_c = u.w;
    } else {
      clip = s / spc;
      t0 = int(s - clip * spc) * kStrip;
    }
    const int unit = RAGGED ? kRaggedTile : kStrip;
    const int nf = int(n_frames_c - t0 < unit ? n_frames_c - t0 : unit);
    const float* __restrict__ xc = prm.x + clip * prm.clip_stride;
    float* __restrict__ oc = prm.out + clip * prm.out_clip_stride;
    for (int f = 0; f < nf; ++f) {
      // ---- stage A: lane n2 = real 60-point DFT over samples 32 n1 + n2 of the rotated frame ----
      const long long p0 = (long long)(t0 + f) * kHop, j0 = p0 - prm.pad_left;
      float in[kN1];
      if (j0 >= 0 && j0 + kN <= n_samples) {
        const float* __restrict__ p = xc + j0 + rot + lane;
#pragma unroll
        for (int n1 = 0; n1 < kN1 - 1; ++n1) in[n1] = __ldg(p + 32 * n1);
        in[kN1 - 1] = __ldg(p + last_off);
        // the 480 samples the next frame of this strip adds (15 aligned lines; every line start is a sample of that frame), if that frame is interior too
        if (B2A_WPF_PREFETCH && f + 1 < nf && lane < 15 && j0 + kHop + kN <= n_samples) wpf_prefetch(xc + j0 + rot + kN + 32 * lane);
      } else {
        wpf_stage_edge(xc, s_ex, p0, prm.pad_left, n_samples, n_samples + prm.zero_tail, prm.pad_mode, rot, lane);
        __syncwarp();
#pragma unroll
        for (int n1 = 0; n1 < kN1; ++n1) in[n1] = s_ex[32 * n1 + lane];
        __syncwarp();
      }
      {
        const float4* wl = reinterpret_cast<const float4*>(s_win + lane * kN1);
#pragma unroll
        for (int q = 0; q < kN1 / 4; ++q) {
          const float4 w4 = wl[q];
          in[4 * q] *= w4.x; in[4 * q + 1] *= w4.y; in[4 * q + 2] *= w4.z; in[4 * q + 3] *= w4.w;
        }
      }
      float2* ex = reinterpret_cast<float2*>(s_ex);
      {
        float yr[kN1 / 2 + 1], yi[kN1 / 2 + 1];
        b2a_rdft60(in, yr, yi);
        ex[lane] = make_float2(yr[0], 0.0f);
        constexpr int CH = B2A_WPF_TW_BATCH;
#pragma unroll
        for (int c = 0; c < 30; c += CH) {
          float2 tw[CH];
#pragma unroll
          for (int j = 0; j < CH; ++j)
            if (c + j < 30) tw[j] = s_tw[(c + j) * 32 + lane];
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const int k = c + j + 1;
            if (k < 30) ex[k * kExPitch + lane] = make_float2(yr[k] * tw[j].x - yi[k] * tw[j].y, yr[k] * tw[j].y + yi[k] * tw[j].x);
            else if (k == 30) ex[30 * kExPitch + lane] = make_float2(yr[30] * tw[j].x, yr[30] * tw[j].y);
          }
        }
      }
      __syncwarp();

      // ---- stage B: lane k1 = complex 32-point DFT of row k1; bins k1 + 60 k2 and their mirror images ----
      {
        float xr[32], xi[32], ur[32], ui[32];
        const float4* zr = reinterpret_cast<const float4*>(ex + k1 * kExPitch);
#pragma unroll
        for (int q = 0; q < 16; ++q) {
          const float4 v = zr[q];
          xr[2 * q] = v.x; xi[2 * q] = v.y; xr[2 * q + 1] = v.z; xi[2 * q + 1] = v.w;
        }
        __syncwarp();   // every lane holds its row: the buffer becomes the spectrum
        // the pad words of exchange rows 9..13 lie in the pruned band and may hold an edge frame's staged samples: one store keeps them
        // zero (a NaN sample must not outlive its frame through 0 * NaN in the mel schedule's slack reads)
        if (PRUNE && lane < 20) s_ex[(9 + (lane >> 2)) * kExPitch * 2 + 64 + (lane & 3)] = 0.0f;
        b2a_cdft32(xr, xi, ur, ui);
        float* pd = s_ex + lane;             // bin lane + 60 k2
        float* pm = s_ex + kN - lane;        // bin 1920 - lane - 60 k2
        const bool direct = lane < 31, mirror = lane >= 1 && lane < 30;
        if (PRUNE) {
          // No predicates on the stores (with only two loop-invariant conditions left the compiler builds one copy of stage B per
          // condition and a warp runs both): lane 31 repeats lane 30's stores word for word, and the three lanes without mirror images
          // write theirs into the exchange rows behind the spectrum, dead since the row loads (banks 0, 1, 2: the ones lanes 1..29 leave free).
          pd = s_ex + k1;
          if (!mirror) pm = s_ex + kPruneDummy + (lane == 0 ? 0 : 32 - lane);
        }
#pragma unroll
        for (int k2 = 0; k2 < 32; ++k2) {
          if (PRUNE && k2 >= kPruneLo && k2 <= kPruneHi) continue;
          float pw = ur[k2] * ur[k2] + ui[k2] * ui[k2];
          if (MAG) pw = wpf_sqrt(pw);
          if (PRUNE) {
            if (k2 < kPruneLo) pd[60 * k2] = pw;
            else pm[-60 * k2] = pw;
          } else if (k2 < 16) {
            if (direct) pd[60 * k2] = pw;
          } else if (k2 == 16) {
            if (lane == 0) s_ex[960] = pw;          // the Nyquist bin belongs to row 0
            else if (mirror) pm[-60 * 16] = pw;
          } else {
            if (mirror) pm[-60 * k2] = pw;
          }
        }
      }
      __syncwarp();

      // ---- mel: per-lane segment sums (four products per step: 16-byte loads of weights and bins), then the <= 4 segments of a filter ----
      for (int r = 0; r < rounds; ++r) {
        const int len4 = int(s_mel[2 + r]);
        const float4* __restrict__ w = reinterpret_cast<const float4*>(s_mel + s_mel[6 + r]) + lane;
        const float4* __restrict__ pp = reinterpret_cast<const float4*>(s_ex + s_start[r * 32 + lane]);
        float a0 = 0.0f, a1 = 0.0f, a2 = 0.0f, a3 = 0.0f;
        // (four steps per trip with all eight loads ahead of the products measured slower: 0.3745 vs 0.3679 ms)
#pragma unroll 2
        for (int g = 0; g < len4; ++g) {
          const float4 wv = w[g * 32], pv = pp[g];
          a0 = fmaf(wv.x, pv.x, a0);
          a1 = fmaf(wv.y, pv.y, a1);
          a2 = fmaf(wv.z, pv.z, a2);
          a3 = fmaf(wv.w, pv.w, a3);
        }
        s_part[r * 32 + lane] = (a0 + a1) + (a2 + a3);
      }
      __syncwarp();
      for (int q = 0; q < kWpfRounds; ++q) {
        const int m = lane + 32 * q;
        if (m < M) {
          const int4 fi = s_fin[m];
          float v = ((s_part[fi.x] + s_part[fi.y]) + s_part[fi.z]) + s_part[fi.w];
          if (log_mode == LOG_LN) v = wpf_lg2(fmaxf(v, log_floor)) * 0.69314718055994531f;
          else if (log_mode == LOG_LOG10) v = wpf_lg2(fmaxf(v, log_floor)) * 0.30102999566398120f;
          else if (log_mode == LOG_DB20) v = wpf_lg2(fmaxf(v, log_floor)) * 6.0205999132796239f;
          if (affine) v = (v - prm.post_sub) / prm.post_div;
          // (M, T'): 4-byte stores; the eight frames of a strip fill whole 32-byte sectors in the L2 within microseconds (staging the
          // strip's (M, 8) block in shared memory and writing 32-byte segments measured the same: 0.3716 vs 0.3694 ms)
          oc[(long long)m * n_frames + t0 + f] = v;
        }
      }
      // (the next frame's first shared-memory write -- exchange rows or an edge frame -- follows the barrier above: every lane is
      // done reading the spectrum; s_part is rewritten only after the next frame's own barriers)
    }
  }
}

int cuda_fail(cudaError_t e, const char* what, std::string* err) {
  if (err) *err = std::string(what) + ": " + cudaGetErrorString(e);
  return B2A_E_CUDA;
}

template <int NW, bool MAG, bool RAGGED, bool PRUNE>
int launch_t(const WpfParams& prm, cudaStream_t st, std::string* err) {
  constexpr size_t smem = sizeof(float) * size_t(kTableWords + NW * kWarpWords);
  struct DevInfo { std::atomic<int> ready{0}; int n_sm = 0; };
  static DevInfo infos[64];
  static std::mutex mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) {
    if (err) *err = "device index out of range";
    return B2A_E_CUDA;
  }
  DevInfo& di = infos[dev];
  cudaError_t e;
  if (!di.ready.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(mu);
    if (!di.ready.load(std::memory_order_relaxed)) {
      if ((e = cudaFuncSetAttribute(wpf1920_kernel<NW, MAG, RAGGED, PRUNE>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) != cudaSuccess)
        return cuda_fail(e, "cudaFuncSetAttribute", err);
      int n_sm = 148;
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      di.n_sm = n_sm;
      di.ready.store(1, std::memory_order_release);
    }
  }
  const long long blocks = std::min<long long>((prm.total_strips + NW - 1) / NW, di.n_sm);   // one persistent CTA per SM
  wpf1920_kernel<NW, MAG, RAGGED, PRUNE><<<unsigned(blocks), NW * 32, smem, st>>>(prm);
  if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "wpf1920_kernel launch", err);
  return B2A_OK;
}

}  // namespace

int init_wpf1920_tables(std::string* err) {
  std::vector<float2> t(30 * 32);
  for (int k1 = 1; k1 <= 30; ++k1)
    for (int n2 = 0; n2 < 32; ++n2) {
      const double a = -2.0 * M_PI * double((n2 * k1) % kN) / double(kN);
      t[size_t(k1 - 1) * 32 + n2] = make_float2(float(cos(a)), float(sin(a)));
    }
  const cudaError_t e = cudaMemcpyToSymbol(c_wpf_tw, t.data(), t.size() * sizeof(float2));
  if (e != cudaSuccess) return cuda_fail(e, "twiddle upload", err);
  return B2A_OK;
}

// On by default; B2A_WPF1920=0 in the environment or b2a_debug_wpf1920(0) keeps the tiled lane == frame kernel of frontend.cu (A/B switch)
static int g_wpf_enabled = -1;
static int g_wpf_prune = 1;   // b2a_debug_wpf1920(2) = warp-per-frame kernel without the pruned stage B (A/B and the parity test of the unpruned path)
void wpf1920_enable(int on) {
  g_wpf_enabled = on ? 1 : 0;
  g_wpf_prune = on == 2 ? 0 : 1;
}

bool wpf1920_applicable(const FrontendArgs& a) {
  if (g_wpf_enabled < 0) {
    const char* v = getenv("B2A_WPF1920");
    g_wpf_enabled = (v != nullptr && v[0] == '0') ? 0 : 1;
    if (v != nullptr && v[0] == '2') g_wpf_prune = 0;
  }
  return g_wpf_enabled == 1 && a.n_fft == kN && a.hop == kHop && a.win_len == kN && a.pre_mode == PRE_NONE && a.bank.wpf_mel != nullptr && (a.clip_tab == nullptr || (a.tile_tab != nullptr && a.total_tiles > 0)) &&
         a.bank.n_mels <= kWpfRounds * 32 && !a.whisper_norm && !a.out_f16 && a.out_mode == OUT_MT &&
         a.n_frames > 0 && a.n_frames <= 0x7fffffffLL && a.batch > 0;
}

int launch_wpf1920(const FrontendArgs& a, void* stream, int* launches, std::string* err) {
  WpfParams prm;
  prm.x = a.x;
  prm.out = a.out;
  prm.mel = a.bank.wpf_mel;
  prm.mel_words = a.bank.wpf_words;
  prm.clip_stride = a.n_samples;
  prm.n_samples = a.n_samples;
  prm.zero_tail = a.zero_tail;
  prm.pad_left = a.pad_left;
  prm.n_frames = a.n_frames;
  prm.out_clip_stride = a.n_frames * (long long)a.bank.n_mels;
  prm.strips_per_clip = int((a.n_frames + kStrip - 1) / kStrip);
  prm.total_strips = a.clip_tab ? (long long)a.total_tiles : (long long)prm.strips_per_clip * a.batch;
  prm.tile_tab = static_cast<const int4*>(a.tile_tab);
  prm.pad_mode = a.pad_mode;
  prm.log_mode = a.log_mode;
  prm.post_affine = a.post_affine;
  prm.n_mels = a.bank.n_mels;
  prm.rot = int(((a.pad_left % 32) + 32) % 32);
  prm.log_floor = a.log_floor;
  prm.post_sub = a.post_sub;
  prm.post_div = a.post_div;
  for (int o = 0; o < kN; ++o) prm.window[o] = a.window[o];
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool mag = a.spec_mode == SPEC_MAGNITUDE;
  int rc;
  // banks that end at or below bin 640 (every S3Gen configuration of the reference: fmax 8000 Hz at 24 kHz) run the pruned stage B
  const bool prune = g_wpf_prune != 0 && a.bank.n_bins_used > 0 && a.bank.n_bins_used <= kPruneBins;
  constexpr int W = B2A_WPF_WARPS;
  if (prune) {
    if (a.clip_tab) rc = mag ? launch_t<W, true, true, true>(prm, st, err) : launch_t<W, false, true, true>(prm, st, err);
    else rc = mag ? launch_t<W, true, false, true>(prm, st, err) : launch_t<W, false, false, true>(prm, st, err);
  } else {
    if (a.clip_tab) rc = mag ? launch_t<W, true, true, false>(prm, st, err) : launch_t<W, false, true, false>(prm, st, err);
    else rc = mag ? launch_t<W, true, false, false>(prm, st, err) : launch_t<W, false, false, false>(prm, st, err);
  }
  if (rc == B2A_OK) *launches += 1;
  return rc;
}

}  // namespace b2a
