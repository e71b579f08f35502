// Fused STFT -> |X|^2 / |X| -> sparse mel projection -> log / clamp / normalise for sm_100a.
//
// One kernel replaces the reference's chain of ~10 lazily evaluated MLX ops
// (reflectPad -> asStrided -> * window -> rfft -> abs -> pow -> matmul -> log10 -> max -> scale;
// Codec/S3Tokenizer/S3TokenizerUtils.swift:224-263, STT/Whisper/WhisperAudio.swift:78-137 and the
// other front ends listed in include/b200audio.h).  Nothing here is derived from MLX source.
//
// Design:
//   * persistent CTAs; a tile is FT = 32 (or 16) consecutive frames of one clip; LANE == FRAME in every stage, so
//     twiddles / window values / filter weights are warp-uniform operands and every shared-memory
//     access is stride-1 across lanes (conflict-free by construction);
//   * PCM for the tile is copied once from HBM with cp.async into shared memory (row pitch HOP+1 so that the
//     stride-HOP frame starts fall in distinct banks) -- the 2.5x (or 4x) frame overlap is served from shared
//     memory, never from HBM; the next tile's copy is issued as soon as stage A has consumed the current one;
//   * real FFT of size N = N1*N2 as two register-resident stages of generated straight-line
//     codelets (tools/gen_codelets.py): stage A = N2 real DFTs of size N1 (+ inter-stage twiddle),
//     one shared-memory exchange, stage B = N1/2 items of size N2 (complex; k1 = 0 and N1/2 share one item);
//     warps split the items of each stage;
//   * shared memory is used three times per tile: stage B writes |X|^2 back into its own rows of the exchange
//     buffer, the mel stage leaves its sums in the rows the spectrum does not use;
//   * mel projection is sparse (a bin feeds at most two adjacent triangular filters): a step program interpreted
//     in a loop for arbitrary banks, build-time generated straight-line code (mel_baked.h) for the standard ones;
//     log / floor / scale fused into the coalesced store loop;
//   * the per-clip max-8 clamp of Whisper / S3Tokenizer needs a clip-global maximum: the main
//     kernel writes normalised values, tracks per-clip max and per-tile min, and a second small
//     kernel rewrites only tiles whose minimum is below the clamp threshold;
//   * the tile loop's instruction footprint is kept near 32 KB (what the SM's instruction cache serves at full rate).
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <utility>
#include <string>
#include <vector>

#include "../../include/b200audio.h"
#include "codelets.h"
#include "internal.h"
#include "mel_baked.h"
#include "pad_index.cuh"

// 1 = interior tiles of the standard 7/6 LFR stacking take the division-free store (0 keeps the generic segment loop: A/B switch)
#ifndef B2A_LFR_FAST
#define B2A_LFR_FAST 1
#endif
// 1 = the (M, T') store of the baked banks walks out_base_words() incrementally (0 keeps the per-row evaluation: A/B switch)
#ifndef B2A_MT_WALK
#define B2A_MT_WALK 1
#endif
// 1 = Whisper-style clamp bookkeeping on the RAW mel sums (A/B switch; 0 keeps the round-1 form: floor and min / max of the normalised
// values inside the store loop).  The store phase is the latency-critical stretch of a tile (removing its 177 tracking instructions,
// 1.6 % of the tile, measured -4.8 %), so: the log floor moves into the clamp kernel (max(max(L, floor), Lmax - 8) = max(L, max(floor,
// Lmax - 8)): one FMNMX per value less, and nothing in front of the MUFU), the minimum / maximum are taken over the raw sums (log is
// monotone: max f(v) = f(max v)) in trees of three-input FMNMX that do not wait for the logarithms, and the two per-warp atomics are
// plain RED instructions instead of the compiler's warp-aggregated sequence.
#ifndef B2A_RAW_TRACK
#define B2A_RAW_TRACK 1
#endif
// 1 = tiles that lie entirely in the caller's zero tail (Whisper `padding`: WhisperSTT.swift:139-144 appends 30 s of zeros to every clip, so
// half of the frames of a 30 s call are silence) are not transformed at all: the main kernel only marks them (kTileFill in tile_min) and the
// clamp kernel, which would have lifted their log floor to max(floor, Lmax - 8) anyway, fills them without reading (A/B switch)
#ifndef B2A_SKIP_ZERO_TILES
#define B2A_SKIP_ZERO_TILES 1
#endif
constexpr int kTileFill = 0x7fffffff;   // tile_min entry of a skipped tile (no float's ordered-int encoding: enc_ordered(+inf) = 0x7f800000)
// 1 = the (M, T') store of the baked 32-frame banks reads its staging offsets from a per-warp shared-memory table built in the prologue
// (0: the incremental out_base_words() walk, ~8 integer instructions per stored row in the unrolled loop; A/B switch)
#ifndef B2A_MT_TABLE
#define B2A_MT_TABLE 1
#endif


namespace b2a {

// ------------------------------------------------------------------------------------------------
// codelet dispatch by array size
// ------------------------------------------------------------------------------------------------
B2A_DEV void rdft(const float (&x)[16], float (&yr)[9], float (&yi)[9]) { b2a_rdft16(x, yr, yi); }
B2A_DEV void rdft(const float (&x)[20], float (&yr)[11], float (&yi)[11]) { b2a_rdft20(x, yr, yi); }
B2A_DEV void rdft(const float (&x)[32], float (&yr)[17], float (&yi)[17]) { b2a_rdft32(x, yr, yi); }
B2A_DEV void rdft(const float (&x)[60], float (&yr)[31], float (&yi)[31]) { b2a_rdft60(x, yr, yi); }
B2A_DEV void rdftodd(const float (&x)[16], float (&yr)[8], float (&yi)[8]) { b2a_rdftodd16(x, yr, yi); }
B2A_DEV void rdftodd(const float (&x)[20], float (&yr)[10], float (&yi)[10]) { b2a_rdftodd20(x, yr, yi); }
B2A_DEV void rdftodd(const float (&x)[32], float (&yr)[16], float (&yi)[16]) { b2a_rdftodd32(x, yr, yi); }
B2A_DEV void cdft(const float (&xr)[16], const float (&xi)[16], float (&yr)[16], float (&yi)[16]) { b2a_cdft16(xr, xi, yr, yi); }
B2A_DEV void cdft(const float (&xr)[20], const float (&xi)[20], float (&yr)[20], float (&yi)[20]) { b2a_cdft20(xr, xi, yr, yi); }
B2A_DEV void cdft(const float (&xr)[32], const float (&xi)[32], float (&yr)[32], float (&yi)[32]) { b2a_cdft32(xr, xi, yr, yi); }

// inter-stage twiddles W_N^{n2*k1}, k1 = 1..N1/2-1, as (cos, -sin); filled once per device
__constant__ float2 c_tw400[20 * 9];
__constant__ float2 c_tw512[16 * 15];
__constant__ float2 c_tw1920[32 * 29];

template <int N_, int WIN_, int N1_, int N2_, int HOP_, int FT_, int NWARPS_, int MINB_>
struct Plan {
  static constexpr int N = N_, WIN = WIN_, N1 = N1_, N2 = N2_, HOP = HOP_, FT = FT_, NWARPS = NWARPS_, MINB = MINB_;
  static constexpr int H1 = N1 / 2;
  static constexpr int NBINS = N / 2 + 1;
  static constexpr int NTHREADS = NWARPS * 32;
  static constexpr int TS = (FT - 1) * HOP + WIN;  // samples per tile
  static constexpr int PITCH = HOP + 1;            // skewed row pitch
  static constexpr int NROWS = (TS + HOP - 1) / HOP;  // rows of HOP samples staged per tile (the last one may be partial)
  static constexpr int PCM_WORDS = (NROWS * PITCH + 3) & ~3;
  static constexpr int Y_WORDS = N * FT;           // H1 float2 slots x N2 x FT; stage B leaves the real spectrum tile here too
  static constexpr int P_PITCH = FT + 1;
  static constexpr bool CPLX_DIRECT = (N / 2 + 1) * (FT + 1) * 2 * 4 > 64 * 1024;  // complex tile would not fit: stft() stores directly
  static constexpr int P_WORDS_CPLX = NBINS * P_PITCH * 2;
  // region 0: the PCM tile, later the [m][frame] output staging tile (plain stft(): the complex spectrum tile)
  static constexpr int R0_WORDS_REAL = PCM_WORDS;
  static constexpr int R0_WORDS_CPLX = CPLX_DIRECT ? ((PCM_WORDS + 3) & ~3) : (((PCM_WORDS > P_WORDS_CPLX ? PCM_WORDS : P_WORDS_CPLX) + 3) & ~3);
  static constexpr int TW_ROW = (2 * (N1 / 2 - 1) + 3) & ~3;         // twiddles of one stage-A item, padded to whole float4s
  static constexpr int TW_WORDS = N2 * TW_ROW;
  static constexpr int SUB = 32 / FT;               // a warp covers FT frames x SUB items (lane = sub * FT + frame)
  static constexpr int NCHUNK = NWARPS * SUB;       // mel-program chunks
  static_assert(N1 * N2 == N, "N = N1*N2");
  static_assert(FT == 32 || FT == 16, "lane == (item, frame)");
};
// 400 = 20 x 20: 20 stage-A items and 9 complex + 1 (real + odd-real) stage-B items split evenly over 10 warps; 74.4 KB of
// shared memory per CTA -> three CTAs (30 warps) per SM, which hides the barrier / shared-memory latencies better than
// prefetching the next tile into a second buffer with two CTAs per SM
using Plan400 = Plan<400, 400, 20, 20, 160, 32, 10, 3>;
// 512 = 32 x 16: 16 stage-A items (real DFTs of size 32) and 15 complex + 1 (real + odd-real) stage-B items of size 16, one per warp
using Plan512 = Plan<512, 400, 32, 16, 160, 32, 16, 2>;
// n_fft 1920 (S3Gen 24 kHz mel): the exchange buffer only fits 16 frames, so half-warps take different items
using Plan1920 = Plan<1920, 1920, 60, 32, 480, 16, 16, 1>;
template <class P> constexpr bool plan_matches(const PlanShape& s) {
  return s.n_fft == P::N && s.n1 == P::N1 && s.n2 == P::N2 && s.frame_tile == P::FT && s.n_warps == P::NWARPS && s.n_chunks == P::NCHUNK;
}
static_assert(plan_matches<Plan400>(kPlanShapes[0]) && plan_matches<Plan512>(kPlanShapes[1]) && plan_matches<Plan1920>(kPlanShapes[2]),
              "internal.h kPlanShapes (host tables, baked mel generator) must describe these plans");

template <class P> struct TwTable;
template <> struct TwTable<Plan400> { static B2A_DEV const float2* get() { return c_tw400; } };
template <> struct TwTable<Plan512> { static B2A_DEV const float2* get() { return c_tw512; } };
template <> struct TwTable<Plan1920> { static B2A_DEV const float2* get() { return c_tw1920; } };

enum SpecKind { SK_POWER = 0, SK_MAG = 1, SK_CPLX = 2 };
// mel-program steps a kernel keeps in shared memory when its plan leaves the room: 32 KB with one CTA per SM (1920), 16 KB with two (512)
template <class P> constexpr int steps_smem_max() { return P::MINB == 1 ? 2048 : (P::MINB == 2 ? 1024 : 0); }
// Post-processing of a finished mel value.  POST_RUNTIME: log mode / Whisper normalisation are kernel parameters and the
// filterbank is the interpreted step program (any bank); the other kinds belong to the baked banks of mel_baked.h.
enum PostKind { POST_RUNTIME = 0, POST_WNORM = 1, POST_LN = 2, POST_NONE = 3 /* the mel value itself (voice-encoder "amp" mel) */ };

template <int MEL> struct MelTraits { static constexpr int M = 0; };
template <> struct MelTraits<1> { static constexpr int M = 128; };
template <> struct MelTraits<2> { static constexpr int M = 80; };
template <> struct MelTraits<3> { static constexpr int M = 80; };
template <> struct MelTraits<4> { static constexpr int M = 80; };
template <> struct MelTraits<5> { static constexpr int M = 40; };
// (M, T') store of a baked bank: rows per warp, padded to whole 16-byte table reads
template <class P, int MEL> constexpr int mt_slots() { return ((MelTraits<MEL>::M + P::NWARPS * P::SUB - 1) / (P::NWARPS * P::SUB) + 3) & ~3; }
template <class P, int MEL, int OUT> constexpr bool mt_table() { return B2A_MT_TABLE && MEL > 0 && OUT == OUT_MT && P::FT == 32; }
template <int MEL, class Emit>
__device__ __forceinline__ void mel_baked(int chunk, const float* __restrict__ p, Emit&& emit) {
  if constexpr (MEL == 1) mel_baked_1(chunk, p, emit);
  else if constexpr (MEL == 2) mel_baked_2(chunk, p, emit);
  else if constexpr (MEL == 3) mel_baked_3(chunk, p, emit);
  else if constexpr (MEL == 4) mel_baked_4(chunk, p, emit);
  else if constexpr (MEL == 5) mel_baked_5(chunk, p, emit);
}

template <class P>
struct FrontendParams {
  const float* x;
  long long clip_stride, n_samples, n_eff, pad_left, n_frames, out_clip_stride, lfr_rows, total_tiles;
  int pad_mode, log_mode, whisper_norm, post_affine, out_mode, tiles_per_clip, lfr_m, lfr_n, n_clips;
  float log_floor, post_sub, post_div;
  const int4* fb_desc;    // per filter: (first bin, number of bins, offset into fb_w, 0)
  const float* fb_w;
  const float4* fb_steps; // mel step program (see mel_steps); null = generic per-filter path
  int n_mels, n_steps;
  int chunk_m[P::NCHUNK + 1];  // filters [chunk_m[c], chunk_m[c+1]) belong to chunk c = warp * SUB + sub
  int chunk_s[P::NCHUNK + 1];  // steps   [chunk_s[c], chunk_s[c+1]) of the program belong to chunk c
  float* out;
  int* clip_max;
  int* tile_min;   // per tile: ordered-int encoding of MINUS the minimum normalised value (atomicMax, same 0x80.. initial pattern as clip_max)
  int* tile_max;   // per tile: the maximum normalised value (floor applied), or null: a tile that lies entirely at or below the clip's clamp
                   // threshold -- silence -- becomes that constant, which the clamp kernel then writes without reading the tile (RAWT kernels)
  // Ragged batches (per-clip lengths; RAGGED kernels only): clip b = (n_samples, n_frames, lfr_rows, index of its first tile);
  // tile g of the launch = tile_tab[g] = (clip, tile within the clip, the clip's n_samples, n_frames), built on the device from
  // clip_tab (tile_table_kernel): ONE load per tile, issued a tile ahead -- a second, dependent load of the clip's row cost ~1000
  // cycles per tile (ncu launch list: 524 vs 486 us for 512 x 20 s through the ragged entry with equal lengths).
  // Strides (clip_stride, out_clip_stride) stay those of the longest clip.
  const int4* clip_tab;
  const int4* tile_tab;
  // Dynamic tile walk (equal-length launches of more than two rounds): after its first tile (blockIdx.x) a CTA takes tile
  // gridDim.x + (atomicAdd(tile_ctr, 1) - tile_ctr_init) instead of the static blockIdx.x + k * gridDim.x -- CTAs that run ahead (an SM with
  // 7 instead of 8 warps on a scheduler, fewer edge tiles) take more tiles and the launch ends without a tail.  null = static walk.
  int* tile_ctr;
  int tile_ctr_init;
  // Tiles [walk_tpc, tiles_per_clip) of every clip are tiles of silence (the caller's zero tail, Whisper's `padding`): the walk runs over the
  // first walk_tpc tiles of every clip only, and the kernel's prologue marks the rest for the clamp kernel's fill (equal-length launches;
  // ragged ones and tails the right-edge reflection reaches back out of keep the ZS instantiation's test per tile).
  int walk_tpc;
  float window[P::WIN];
};

B2A_DEV int enc_ordered(float f) {
  const int b = __float_as_int(f);
  return b >= 0 ? b : b ^ 0x7fffffff;
}
B2A_DEV float dec_ordered(int e) { return __int_as_float(e >= 0 ? e : e ^ 0x7fffffff); }
// one thread's atomic max without a return value (atomicMax() under `if (lane == 0)` compiles to a warp-aggregated sequence: vote, leader
// election and a second reduction, ~20 instructions)
B2A_DEV void red_max_s32(int* p, int v) { asm volatile("red.relaxed.gpu.global.max.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory"); }

// Shared-memory word offset (relative to the lane's frame start) of sample o = N2*n1 + n2 in the skewed
// PCM tile: o + o / HOP.  With n1 a compile-time constant this is (immediate) + n2 + carry, and the carry
// can only be non-zero for the few n1 whose row remainder is within N2 of the row end.
template <class P, int n1>
B2A_DEV int pcm_off(int n2) {
  constexpr int base = P::N2 * n1;
  constexpr int c = base / P::HOP, r = base % P::HOP;
  if (r + P::N2 - 1 >= P::HOP) return base + c + n2 + (n2 >= P::HOP - r ? 1 : 0);
  return base + c + n2;
}

template <class P, int PRE, int n1>
B2A_DEV float load_sample(const float* lane_pcm, int n2, float mu) {
  constexpr int base = P::N2 * n1;
  if (base >= P::WIN) return 0.0f;
  const int off = pcm_off<P, n1>(n2);
  float v = lane_pcm[off];
  if (PRE == PRE_KALDI) {
    // (x[o]-mu) - 0.97*(x[o-1]-mu), first sample only DC-removed (CAMPPlus.swift:66-72)
    v = v - mu;
    const int o = base + n2;
    if (o > 0) {
      const int poff = off - 1 - ((o % P::HOP) == 0 ? 1 : 0);  // previous sample may sit before the row's pad word
      v = v - 0.97f * (lane_pcm[poff] - mu);
    }
  }
  if (base + P::N2 - 1 >= P::WIN && base + n2 >= P::WIN) v = 0.0f;  // zero-extended window (win_length < n_fft)
  return v;
}

// Kaldi pre-processing in two steps around the per-frame mean: the raw samples of the item (zero beyond the window) ...
template <class P, int n1>
B2A_DEV float raw_sample(const float* lane_pcm, int n2) {
  constexpr int base = P::N2 * n1;
  if (base >= P::WIN) return 0.0f;
  const float v = lane_pcm[pcm_off<P, n1>(n2)];
  return (base + P::N2 - 1 >= P::WIN && base + n2 >= P::WIN) ? 0.0f : v;
}
// ... and (x[o]-mu) - 0.97*(x[o-1]-mu), first sample only DC-removed (CAMPPlus.swift:66-72), times the window
template <class P, int n1>
B2A_DEV float kaldi_sample(float raw, const float* lane_pcm, int n2, float mu, float w) {
  constexpr int base = P::N2 * n1;
  if (base >= P::WIN) return 0.0f;
  const int o = base + n2;
  float v = raw - mu;
  if (o > 0) {
    const int poff = pcm_off<P, n1>(n2) - 1 - ((o % P::HOP) == 0 ? 1 : 0);  // previous sample may sit before the row's pad word
    v = v - 0.97f * (lane_pcm[poff] - mu);
  }
  if (base + P::N2 - 1 >= P::WIN && o >= P::WIN) v = 0.0f;
  return v * w;
}
template <class P, int... I>
B2A_DEV void load_raw_item(const float* lane_pcm, int n2, float (&raw)[P::N1], std::integer_sequence<int, I...>) {
  ((raw[I] = raw_sample<P, I>(lane_pcm, n2)), ...);
}
template <class P, int... I>
B2A_DEV void kaldi_item(const float* lane_pcm, int n2, float mu, const float* wrow, float (&v)[P::N1], std::integer_sequence<int, I...>) {
  ((v[I] = kaldi_sample<P, I>(v[I], lane_pcm, n2, mu, wrow[I])), ...);
}

template <class P, int PRE, int... I>
B2A_DEV void load_item(const float* lane_pcm, int n2, float mu, const float* wrow, float (&in)[P::N1], std::integer_sequence<int, I...>) {
  ((in[I] = load_sample<P, PRE, I>(lane_pcm, n2, mu) * wrow[I]), ...);
}

B2A_DEV float sqrt_approx(float x) {   // MUFU-based square root (~2 ulp): |X| for the magnitude front ends, 1e-4 tolerance downstream
  float y;
  asm("sqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

B2A_DEV float lg2_ftz(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Sparse mel projection for one warp's chunk of filters.  The filterbank is compiled on the host into a "step
// program": step = (w_lo, w_hi, byte offset of the spectrum bin's row, emit word).  Each step adds w_lo*P to the open filter
// and w_hi*P to the next one; a non-zero emit word (byte offset of the filter's staging position, see output_words() in
// host_tables.cpp) stores the open filter's sum and shifts the accumulators.  A bin touches at most two adjacent triangular
// filters; filters without bins get a zero-weight step.  LANE == FRAME.  The program is read straight from global memory:
// every lane reads the same 16 bytes (one L1 sector per step), which keeps shared memory for a third CTA per SM.
B2A_DEV void mel_step_apply(const float4& t, float pk, float& acc0, float& acc1, char* so_lane) {
  const int w = __float_as_int(t.w);
  acc0 = fmaf(t.x, pk, acc0);
  acc1 = fmaf(t.y, pk, acc1);
  if (w != 0) {   // warp-uniform
    *reinterpret_cast<float*>(so_lane + w) = acc0;
    acc0 = acc1;
    acc1 = 0.0f;
  }
}

// SMEM: the program was copied into shared memory (plans with one CTA per SM have the room; global-memory reads of the program are
// mostly L2 hits there, because the cp.async traffic of the PCM staging sweeps the L1)
template <bool SMEM>
B2A_DEV float4 load_step(const float4* steps, int s) {
  return SMEM ? steps[s] : __ldg(steps + s);   // (an L1::evict_last hint on the global loads changed nothing: measured)
}

template <bool SMEM>
B2A_DEV void mel_steps(const float* __restrict__ p_lane, const float4* steps, int s0, int s1, float* so_lane_f) {
  char* so_lane = reinterpret_cast<char*>(so_lane_f);
  float acc0 = 0.0f, acc1 = 0.0f;
  const char* pb = reinterpret_cast<const char*>(p_lane);
  int s = s0;
  // four steps per iteration, all loads issued before the (serial) accumulator updates
  for (; s + 4 <= s1; s += 4) {
    const float4 t0 = load_step<SMEM>(steps, s), t1 = load_step<SMEM>(steps, s + 1), t2 = load_step<SMEM>(steps, s + 2),
                 t3 = load_step<SMEM>(steps, s + 3);
    const float p0 = *reinterpret_cast<const float*>(pb + __float_as_int(t0.z));
    const float p1 = *reinterpret_cast<const float*>(pb + __float_as_int(t1.z));
    const float p2 = *reinterpret_cast<const float*>(pb + __float_as_int(t2.z));
    const float p3 = *reinterpret_cast<const float*>(pb + __float_as_int(t3.z));
    mel_step_apply(t0, p0, acc0, acc1, so_lane);
    mel_step_apply(t1, p1, acc0, acc1, so_lane);
    mel_step_apply(t2, p2, acc0, acc1, so_lane);
    mel_step_apply(t3, p3, acc0, acc1, so_lane);
  }
  for (; s < s1; ++s) {
    const float4 t = load_step<SMEM>(steps, s);
    mel_step_apply(t, *reinterpret_cast<const float*>(pb + __float_as_int(t.z)), acc0, acc1, so_lane);
  }
}

template <int LOGM, bool WNORM>
B2A_DEV float mel_post(float v, float log_floor, float& lmax, float& vmin) {
  // log2-based logs: MUFU.LG2 is accurate to ~1e-7 absolute on the log value, far inside the 1e-4 tolerance
  constexpr float kLog = LOGM == LOG_LOG10 ? 0.30102999566398120f : (LOGM == LOG_LN ? 0.69314718055994531f : 6.0205999132796239f);
  if (LOGM != LOG_NONE) v = lg2_ftz(fmaxf(v, log_floor)) * kLog;
  if (WNORM) {
    v = (v + 4.0f) * 0.25f;
    lmax = fmaxf(lmax, v);
    vmin = fminf(vmin, v);
  }
  return v;
}

B2A_DEV void cp_async4(float* smem_dst, const float* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<unsigned>(__cvta_generic_to_shared(smem_dst))), "l"(gmem_src));
}
B2A_DEV void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// Row of the exchange buffer that holds bin k after stage B (same map as spectrum_slots() in host_tables.cpp)
template <class P>
B2A_DEV int spectrum_slot(int k) {
  constexpr int N1 = P::N1, N2 = P::N2, H1 = P::H1, N = P::N;
  const int r = k % N1;
  if (r == 0) return k / N1;
  if (r == H1) return N2 / 2 + 1 + (k - H1) / N1;
  if (r < H1) return r * 2 * N2 + k / N1;
  return (N1 - r) * 2 * N2 + (N - (N1 - r) - k) / N1;
}

// Staging position of mel value (filter m, frame f) = out_base_words(m) + f: word offset into the exchange buffer
// (same map as output_words() in host_tables.cpp; bank = (m + f) mod 32)
template <class P>
B2A_DEV constexpr int out_base_words(int m) {
  constexpr int CAP = ((P::N2 - 1) * P::FT - (P::FT - 1)) / (P::FT + 1);
  return ((m / CAP) * 2 * P::N2 + P::N2 + 1) * P::FT + (CAP * (m / CAP)) % P::FT + (m % CAP) * (P::FT + 1);
}

// out_base_words<P>(m) for m = m0, m0 + STEP, ... without a division per row: the (M, T') store loops walk the filters with a
// run-time m, and the compiler's code for the two constant divisions and the modulo was ~35 instructions per stored row (45 % of
// the Chatterbox kernel).  q = m / CAP and r = m % CAP advance incrementally.
template <class P, int STEP>
struct OutBaseWalk {
  static constexpr int CAP = ((P::N2 - 1) * P::FT - (P::FT - 1)) / (P::FT + 1);
  static_assert((P::FT & (P::FT - 1)) == 0, "frame tile is a power of two");
  int q, r;
  B2A_DEV explicit OutBaseWalk(int m0) : q(m0 / CAP), r(m0 - (m0 / CAP) * CAP) {}
  B2A_DEV int words() const { return (q * 2 * P::N2 + P::N2 + 1) * P::FT + ((CAP * q) & (P::FT - 1)) + r * (P::FT + 1); }
  B2A_DEV void next() {
    r += STEP % CAP;
    q += STEP / CAP;
    if (r >= CAP) {
      r -= CAP;
      ++q;
    }
  }
};

// Edge tiles (2 of 94 for a 30 s clip, 2 of 32 for a 10 s one): kept out of line so that the 64-bit index arithmetic of the padding
// map does not sit in the instruction stream of the interior tiles (the hot loop has to stay inside the 32 KB the instruction
// cache serves at full rate: tools/microbench/icache.cu).  Like the interior tiles they are ASYNCHRONOUS: every sample is a
// 4-byte cp.async from its mapped source index (`padded_index`; zero padding is a plain shared-memory store), so an
// edge tile prefetched behind stage A lands during stage B / mel / store instead of stalling the CTA on ~17 rounds of global
// loads; rows that lie inside the clip take the interior tiles' row copy, only the rows that touch the padding go through the
// map, and the modulo of the reference's repeated reflection only runs for clips shorter than the pad.
template <class P>
__device__ __noinline__ void stage_pcm_edge(const float* __restrict__ xc, float* __restrict__ buf, long long p0, long long pad_left,
                                            long long n_samples, long long n_eff, int pad_mode, int tid) {
  const int lane = tid & 31, warp = tid >> 5;
  const long long j0 = p0 - pad_left;
  for (int row = warp; row < P::NROWS; row += P::NWARPS) {
    const long long jr = j0 + (long long)row * P::HOP;
    float* d = buf + row * P::PITCH + lane;
    if (jr >= 0 && jr + P::HOP <= n_samples) {   // (warp-uniform) the row lies inside the clip: the interior tiles' row copy
#pragma unroll
      for (int q = 0; q < P::HOP / 32; ++q) cp_async4(d + q * 32, xc + jr + lane + q * 32);
      continue;
    }
    for (int q = 0; q < P::HOP / 32; ++q) {      // rows that touch the padding: sample by sample through the index map
      const long long j = padded_index(p0 + (long long)row * P::HOP + lane + q * 32, pad_left, n_eff, pad_mode);
      if (j >= 0 && j < n_samples) cp_async4(d + q * 32, xc + j);
      else d[q * 32] = 0.0f;
    }
  }
}

// Stages the PCM of tile (clip, f0) into `buf` (skewed rows, pitch HOP+1).  Interior tiles use cp.async so that
// the copy overlaps with the previous tile's FFT stages: whole rows of HOP samples, warp w takes rows w, w+NW, ...,
// every copy an immediate offset from two per-warp base pointers.  Edge tiles (reflect / zero padding, clip end) go
// through the index map.
template <class P>
B2A_DEV void stage_pcm(const FrontendParams<P>& prm, float* __restrict__ buf, int clip, int f0, long long n_samples, long long n_eff,
                       int tid, int lane, int warp) {
  constexpr int HOP = P::HOP, NW = P::NWARPS, NROWS = P::NROWS;
  static_assert(HOP % 32 == 0, "rows are copied as whole 32-lane chunks");
  const float* __restrict__ xc = prm.x + (long long)clip * prm.clip_stride;
  const long long p0 = (long long)f0 * HOP;
  const long long j0 = p0 - prm.pad_left;
  if (j0 >= 0 && j0 + NROWS * HOP <= n_samples) {
    const float* __restrict__ src = xc + j0 + warp * HOP + lane;
    float* dst = buf + warp * P::PITCH + lane;
#pragma unroll
    for (int rr = 0; rr < (NROWS + NW - 1) / NW; ++rr) {
      if (rr * NW + warp < NROWS) {
#pragma unroll
        for (int j = 0; j < HOP / 32; ++j) cp_async4(dst + rr * NW * P::PITCH + j * 32, src + rr * NW * HOP + j * 32);
      }
    }
  } else {
    stage_pcm_edge<P>(xc, buf, p0, prm.pad_left, n_samples, n_eff, prm.pad_mode, tid);
  }
}

// Persistent kernel: grid = MINB CTAs per SM; every CTA loads its tables once and walks tiles
// blockIdx.x, blockIdx.x + gridDim.x, ...  The next tile's PCM is prefetched (cp.async) into the PCM region as soon as
// stage A has consumed the current one.
// MEL == 0: any bank -- mel step program interpreted in a loop, run-time log / output modes (POST_RUNTIME, OUT = -1).
// MEL > 0: bank MEL of mel_baked.h as straight-line code, with compile-time post-processing POST (applied in the store
// loop) and output layout OUT (OUT_TM / OUT_MT / OUT_LFR).
// RAGGED: per-clip lengths -- the tile walk and the clip geometry come from prm.tile_tab / prm.clip_tab (one uniform load each
// per tile, issued one tile ahead) instead of the launch-wide constants.
// F16: the (T', M) store of a baked bank writes __half (round to nearest even of the fp32 value: bit-identical to the reference's
// asType(.float16) of the fp32 feature, WhisperSTT.swift:156-157,181-182); clamp bookkeeping stays in fp32.
// ZS: the launch has a zero tail (Whisper `padding` > 0) -- the instantiation that skips tiles of silence (B2A_SKIP_ZERO_TILES); launches
// without one keep the kernel without that code (measured: the test alone cost 0.35 - 0.5 % of the unpadded 1024 x 30 s step).
template <class P, int PRE, int SPEC, int MEL, int POST, int OUT, bool RAGGED = false, bool F16 = false, bool ZS = false, bool DYN = false>
__global__ void __launch_bounds__(P::NTHREADS, P::MINB) frontend_kernel(const __grid_constant__ FrontendParams<P> prm) {
  constexpr int N1 = P::N1, N2 = P::N2, H1 = P::H1, FT = P::FT, HOP = P::HOP, NW = P::NWARPS, N = P::N, WIN = P::WIN;
  constexpr bool cplx = SPEC == SK_CPLX;
  constexpr bool BAKED = MEL > 0;                 // straight-line mel code
  constexpr bool FIXED = POST != POST_RUNTIME;    // compile-time post-processing and output layout: short store code
  static_assert((!BAKED || FIXED) && FIXED == (OUT >= 0), "known banks come with compile-time POST and OUT");
  static_assert(BAKED || !FIXED || OUT == OUT_MT, "interpreted bank with compile-time POST: built for the (M, T') layout");
  constexpr bool EARLY_PREFETCH = !cplx;   // the PCM region is dead after stage A: refill it during stage B / mel / store
  // LATE_TOP: the barrier that separates a tile from the previous one (every warp done reading the staged mel values out of the
  // exchange buffer) sits in front of the tile's first exchange-buffer WRITE instead of at the loop top, so a warp that leaves
  // the store phase early already loads and transforms its first stage-A item while the others finish.  The prefetched PCM is
  // published by the post-mel barrier of the previous tile (cp.async wait in front of it).  Kaldi keeps the barrier at the top
  // (its stage A has a barrier of its own for the frame means; letting that one double as the tile separator measured 2 % slower).
  constexpr bool LATE_TOP = EARLY_PREFETCH && PRE != PRE_KALDI;
  constexpr int R0W = cplx ? P::R0_WORDS_CPLX : P::R0_WORDS_REAL;
  static_assert(MEL == 0 || (SPEC == SK_POWER && FT == 32), "known banks are power-spectrum banks of the 32-frame plans");
  static_assert(!F16 || (BAKED && OUT == OUT_TM), "fp16 output is built for the (T', M) store of the baked banks");
  extern __shared__ __align__(16) float smem[];
  // the clamp / statistics kernel behind this one may be scheduled as soon as every CTA here has started (it waits for this grid
  // to complete before it reads anything): hides its launch latency, which matters for the reference's one-clip calls
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  float2* s_y = reinterpret_cast<float2*>(smem + R0W);
  float* s_p = reinterpret_cast<float*>(s_y);               // stage B leaves the spectrum tile in the exchange buffer (spectrum_slots)
  float* s_wt = reinterpret_cast<float*>(s_y) + P::Y_WORDS;  // window, item-major [n2][n1]
  float2* s_tw = reinterpret_cast<float2*>(s_wt + N);       // inter-stage twiddles [n2][k1-1]
  constexpr int SUB = P::SUB, NIT = NW * SUB;
  // plans with one or two CTAs per SM: the interpreted mel program lives in shared memory behind the tables
  constexpr int kStepsSmemMax = steps_smem_max<P>();
  constexpr bool STEPS_SMEM = kStepsSmemMax > 0 && MEL == 0 && !cplx;
  float4* s_steps = reinterpret_cast<float4*>(reinterpret_cast<float*>(s_tw) + P::TW_WORDS);
  int* s_mt = reinterpret_cast<int*>(reinterpret_cast<float*>(s_tw) + P::TW_WORDS);   // baked (M, T') kernels: staging word offset of row warp + j * NW, [warp][j]
  const bool steps_in_smem = STEPS_SMEM && prm.fb_steps != nullptr && prm.n_steps <= kStepsSmemMax;
  if (steps_in_smem)
    for (int i = threadIdx.x; i < prm.n_steps; i += P::NTHREADS) s_steps[i] = __ldg(prm.fb_steps + i);
  __shared__ float s_part[PRE == PRE_KALDI ? NW * FT : 1];   // Kaldi: per-warp partial sums of the frame mean
  // dynamic tile walk (DYN: a compile-time variant of the equal-length kernels -- as a run-time switch its branches and the per-thread
  // division cost 660 instructions per tile, 6 %): the (clip, tile) the counter handed to this CTA for its next round, worked out by
  // thread 0 alone and published by the post-A barrier
  __shared__ int2 s_next2;
  __shared__ int s_next;   // (ragged batches: the index into the tile table)
  static_assert(!DYN || (!RAGGED && !ZS), "the compile-time dynamic walk is the equal-length kernels'");
  constexpr bool dyn = DYN;
  // ragged batches: the counter's answer is an index into tile_tab, whose entry has to be loaded a tile ahead of its use -- so the walk is
  // requested TWO tiles ahead (the CTA's first two tiles are static: blockIdx.x and blockIdx.x + gridDim.x)
  const bool dyn_r = RAGGED && !ZS && prm.tile_ctr != nullptr;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  // frame lane; item / chunk slot of this (half-)warp.  The two half-warps of a 16-frame plan take slots NW apart (not adjacent):
  // their stage-A items then start 16 samples apart, so the 32 lanes' PCM loads (row pitch HOP + 1) fall into 32 different banks
  const int fl = lane % FT, wsub = warp + NW * (lane / FT);

  // ---- 0. per-CTA tables (once): window in item-major order, twiddles (the loop-top barrier publishes them) ----
  for (int i = tid; i < N; i += P::NTHREADS) {
    const int n2 = i / N1, n1 = i - n2 * N1;
    const int o = N2 * n1 + n2;
    s_wt[i] = o < WIN ? prm.window[o] : 0.0f;
  }
  {
    const float2* __restrict__ tw = TwTable<P>::get();
    for (int i = tid; i < N2 * (H1 - 1); i += P::NTHREADS) {
      const int n2 = i / (H1 - 1), k = i - n2 * (H1 - 1);
      *reinterpret_cast<float2*>(reinterpret_cast<float*>(s_tw) + n2 * P::TW_ROW + 2 * k) = tw[i];
    }
  }

  if constexpr (mt_table<P, MEL, OUT>()) {
    constexpr int SL = mt_slots<P, MEL>();
    for (int i = tid; i < NW * SL; i += P::NTHREADS) {
      const int w = i / SL, m = w + (i - w * SL) * NIT;
      s_mt[i] = out_base_words<P>(m < MelTraits<MEL>::M ? m : MelTraits<MEL>::M - 1);
    }
  }

  // tile walk.  Static (launches of at most two rounds, B2A_DYN_TILES=0): (clip, tile) advances by gridDim.x tiles per iteration without
  // divisions in the loop.  Dynamic (DYN / dyn_r): the first tile is blockIdx.x, every further one comes from the launch-wide counter.
  const int tpc = prm.tiles_per_clip, n_clips = prm.n_clips;
  const int wtpc = RAGGED ? tpc : prm.walk_tpc;   // tiles per clip the walk visits
  const int step_clip = int(gridDim.x) / wtpc, step_tile = int(gridDim.x) - step_clip * wtpc;
  int clip = int(blockIdx.x) / wtpc, tile = int(blockIdx.x) - clip * wtpc;
  if (POST == POST_WNORM && !RAGGED && wtpc < tpc) {   // the tiles of silence behind the walk: filled by the clamp kernel
    const int nz = tpc - wtpc;
    for (long long i = (long long)blockIdx.x * P::NTHREADS + threadIdx.x; i < (long long)n_clips * nz; i += (long long)gridDim.x * P::NTHREADS) {
      const long long c = i / nz;
      prm.tile_min[c * tpc + wtpc + (i - c * nz)] = kTileFill;
    }
  }
  // geometry of the current clip: launch-wide constants, or (RAGGED) the clip's own row of clip_tab
  long long n_samples = prm.n_samples, n_frames = prm.n_frames, lfr_rows = prm.lfr_rows;
  const long long zero_tail = prm.n_eff - prm.n_samples;
  long long g = blockIdx.x;    // RAGGED: index of the tile within the launch
  long long g_next = (long long)blockIdx.x + gridDim.x, g_next2 = 0;   // dyn_r: the next tile's index and the one after it
  int first_tile = 0;
  const int lfr_n = prm.lfr_n > 0 ? prm.lfr_n : 1;
  auto set_clip = [&](const int4& t, long long gg) {   // tile_tab entry -> geometry of its clip
    n_samples = t.z;
    n_frames = t.w;
    lfr_rows = OUT == OUT_LFR || OUT < 0 ? (t.w + lfr_n - 1) / lfr_n : 0;   // ceil(T' / n) (FunASRAudio.swift:114)
    first_tile = int(gg) - t.y;
  };
  if (RAGGED) {
    clip = n_clips;
    if (g < prm.total_tiles) {
      const int4 t = __ldg(prm.tile_tab + g);
      clip = t.x;
      tile = t.y;
      set_clip(t, g);
    }
  }
  // ZSKIP: is every sample of this tile a zero of the caller's zero tail (directly, or reflected from it at the right edge)?
  constexpr bool ZSKIP = ZS && B2A_SKIP_ZERO_TILES && B2A_RAW_TRACK && POST == POST_WNORM && LATE_TOP;
  auto tile_zero = [&](int tl, long long ns) {
    if (!ZSKIP || zero_tail <= 0) return false;
    const long long j0 = (long long)tl * (FT * HOP) - prm.pad_left;
    if (j0 < ns) return false;
    const long long over = j0 + (P::TS - 1) - (ns + zero_tail);   // how far the tile's last sample lies beyond the padded signal
    return over < 0 || prm.pad_mode != PAD_REFLECT || over <= zero_tail - 2;
  };
  if (ZSKIP) {   // the CTA's first tiles may be tiles of silence too
    while (clip < n_clips && tile_zero(tile, n_samples)) {
      if (tid == 0) prm.tile_min[RAGGED ? g : (long long)clip * tpc + tile] = kTileFill;
      if (RAGGED) {
        g += gridDim.x;
        clip = n_clips;
        if (g < prm.total_tiles) {
          const int4 t = __ldg(prm.tile_tab + g);
          clip = t.x;
          tile = t.y;
          set_clip(t, g);
        }
      } else {
        clip += step_clip;
        tile += step_tile;
        if (tile >= wtpc) {
          tile -= wtpc;
          ++clip;
        }
      }
    }
  }
  if (clip < n_clips) stage_pcm<P>(prm, smem, clip, tile * FT, n_samples, n_samples + zero_tail, tid, lane, warp);
  if (LATE_TOP) {   // first tile's PCM and the tables
    cp_async_commit_wait_all();
    __syncthreads();
  }

  while (clip < n_clips) {
    const int f0 = tile * FT;
    float* s_r0 = smem;    // PCM tile (plain stft(): later the complex spectrum tile)
    int nclip = clip + step_clip, ntile = tile + step_tile;
    if (ntile >= wtpc) {
      ntile -= wtpc;
      ++nclip;
    }
    long long nn_samples = n_samples;   // next tile's clip length
    int4 nt = make_int4(0, 0, 0, 0);    // RAGGED: the next tile's table entry
    if (RAGGED) {
      nclip = n_clips;
      const long long gn = dyn_r ? g_next : g + gridDim.x;
      if (gn < prm.total_tiles) {
        nt = __ldg(prm.tile_tab + gn);
        nclip = nt.x;
        ntile = nt.y;
        nn_samples = nt.z;
      }
    }
    // ZSKIP: the walk passes over tiles of silence (it only marks them for the clamp kernel's fill), so the tile staged behind stage A is
    // the next tile that is actually transformed -- no barrier, no exposed staging latency for a skipped tile
    long long ng = dyn_r ? g_next : g + gridDim.x;   // (RAGGED: launch-wide index of the next tile)
    if (ZSKIP) {
      while (nclip < n_clips && tile_zero(ntile, nn_samples)) {
        if (tid == 0) prm.tile_min[RAGGED ? ng : (long long)nclip * tpc + ntile] = kTileFill;
        if (RAGGED) {
          ng += gridDim.x;
          nclip = n_clips;
          if (ng < prm.total_tiles) {
            nt = __ldg(prm.tile_tab + ng);
            nclip = nt.x;
            ntile = nt.y;
            nn_samples = nt.z;
          }
        } else {
          nclip += step_clip;
          ntile += step_tile;
          if (ntile >= wtpc) {
            ntile -= wtpc;
            ++nclip;
          }
        }
      }
    }

    int next_req = 0;   // dynamic walk, thread 0: the CTA's next tile (the answer is needed behind stage A)
    if (dyn && tid == 0) next_req = int(gridDim.x) + (atomicAdd(prm.tile_ctr, 1) - prm.tile_ctr_init);
    if (dyn_r && tid == 0) next_req = 2 * int(gridDim.x) + (atomicAdd(prm.tile_ctr, 1) - prm.tile_ctr_init);

    // ---- 1. this tile's PCM has landed (and every warp is done with the previous tile's staging rows) -------------
    if (!LATE_TOP) {
      cp_async_commit_wait_all();
      __syncthreads();
    }

    // Lanes past the clip's last frame recompute the last valid frame (same shared-memory words: a broadcast,
    // not a conflict), so that per-tile max / min need no lane masking.
    const int rows = int(n_frames - f0 < FT ? n_frames - f0 : FT);
    const int flane = fl < rows ? fl : rows - 1;

    // ---- 2. stage A: N2 real DFTs of size N1 over samples o = N2*n1 + n2, twiddle, exchange --------
    {
      const float* lane_pcm = s_r0 + flane * P::PITCH;
      for (int n2 = wsub; n2 < N2; n2 += NIT) {
        float wrow[N1];
#pragma unroll
        for (int q = 0; q < N1 / 4; ++q) {
          const float4 w4 = reinterpret_cast<const float4*>(s_wt + n2 * N1)[q];
          wrow[4 * q] = w4.x; wrow[4 * q + 1] = w4.y; wrow[4 * q + 2] = w4.z; wrow[4 * q + 3] = w4.w;
        }
        float in[N1];
        if (PRE == PRE_KALDI) {
          // Per-frame mean (CAMPPlus.swift:66) without a pass of its own: every warp owns exactly one item, sums the raw
          // samples it has just loaded, and the warps' partial sums are combined in fixed order.
          static_assert(PRE != PRE_KALDI || (SUB == 1 && N2 == NIT), "Kaldi pre-processing: 32-frame tiles, one stage-A item per warp");
          load_raw_item<P>(lane_pcm, n2, in, std::make_integer_sequence<int, N1>{});
          float part = 0.0f;
#pragma unroll
          for (int i = 0; i < N1; ++i) part += in[i];
          s_part[warp * FT + fl] = part;
          __syncthreads();
          float tot = 0.0f;
#pragma unroll
          for (int w = 0; w < NW; ++w) tot += s_part[w * FT + fl];
          const float mu = tot / float(WIN);
          kaldi_item<P>(lane_pcm, n2, mu, wrow, in, std::make_integer_sequence<int, N1>{});
        } else {
          load_item<P, PRE>(lane_pcm, n2, 0.0f, wrow, in, std::make_integer_sequence<int, N1>{});
        }
        float yr[H1 + 1], yi[H1 + 1];
        rdft(in, yr, yi);
        if (LATE_TOP && n2 == wsub) __syncthreads();   // (every warp owns at least one stage-A item: all threads get here once per tile)
        float2* yb = s_y + n2 * FT + fl;
        yb[0] = make_float2(yr[0], yr[H1]);
        // the item's twiddles as whole float4 loads, all in flight before the first complex multiply
        float twv[P::TW_ROW];
#pragma unroll
        for (int q = 0; q < P::TW_ROW / 4; ++q) {
          const float4 t4 = reinterpret_cast<const float4*>(reinterpret_cast<const float*>(s_tw) + n2 * P::TW_ROW)[q];
          twv[4 * q] = t4.x; twv[4 * q + 1] = t4.y; twv[4 * q + 2] = t4.z; twv[4 * q + 3] = t4.w;
        }
#pragma unroll
        for (int k1 = 1; k1 < H1; ++k1) {
          const float tx = twv[2 * (k1 - 1)], ty = twv[2 * (k1 - 1) + 1];
          yb[k1 * N2 * FT] = make_float2(yr[k1] * tx - yi[k1] * ty, yr[k1] * ty + yi[k1] * tx);
        }
      }
    }
    if (dyn_r && tid == 0) s_next = next_req;
    if (dyn && tid == 0) {
      const int nc = next_req / wtpc;
      s_next2 = make_int2(next_req >= prm.total_tiles ? n_clips : nc, next_req - nc * wtpc);
    }
    __syncthreads();
    if (dyn_r) g_next2 = s_next;
    if (dyn) {
      const int2 v = s_next2;
      nclip = v.x;
      ntile = v.y;
    }
    if (EARLY_PREFETCH && nclip < n_clips) stage_pcm<P>(prm, smem, nclip, ntile * FT, nn_samples, nn_samples + zero_tail, tid, lane, warp);

    // ---- 3. stage B: DFTs of size N2 over n2; bins k = k1 + N1*k2 (mirrored above N/2) -------------
    // The power / magnitude of every bin goes back into the item's own rows of the exchange buffer (row = slot of
    // spectrum_slots(), FT floats per row): no separate spectrum tile.  An item's rows are touched by its (half-)warp only,
    // and every lane has its inputs in registers before any lane stores (__syncwarp).
    {
      const unsigned item_mask = FT == 32 ? 0xffffffffu : (0xffffu << (lane & 16));
      auto put = [&](int it, int slot, int k, float re, float im) {
        if (cplx) {
          if (P::CPLX_DIRECT) {
            // tile too large for a staged complex spectrum: store straight to (T', F) global memory
            if (fl < rows)
              reinterpret_cast<float2*>(prm.out + (long long)clip * prm.out_clip_stride)[(long long)(f0 + fl) * P::NBINS + k] = make_float2(re, im);
          } else {
            reinterpret_cast<float2*>(s_r0)[k * P::P_PITCH + fl] = make_float2(re, im);
          }
        } else {
          const float pw = re * re + im * im;
          s_p[(it * 2 * N2 + slot) * FT + fl] = SPEC == SK_POWER ? pw : sqrt_approx(pw);
        }
      };
      for (int it = wsub; it < H1; it += NIT) {
        if (it == 0) {
          // k1 = 0 (real DFT of the DC row) and k1 = N1/2 (odd-frequency real DFT of the Nyquist row) share the row block
          float xa[N2], xb[N2];
#pragma unroll
          for (int n2 = 0; n2 < N2; ++n2) {
            const float2 v = s_y[n2 * FT + fl];
            xa[n2] = v.x;
            xb[n2] = v.y;
          }
          __syncwarp(item_mask);
          {
            float ur[N2 / 2 + 1], ui[N2 / 2 + 1];
            rdft(xa, ur, ui);
#pragma unroll
            for (int k2 = 0; k2 <= N2 / 2; ++k2) put(0, k2, N1 * k2, ur[k2], ui[k2]);
          }
          {
            float ur[(N2 - 1) / 2 + 1], ui[(N2 - 1) / 2 + 1];
            rdftodd(xb, ur, ui);
#pragma unroll
            for (int k2 = 0; k2 <= (N2 - 1) / 2; ++k2) put(0, N2 / 2 + 1 + k2, H1 + N1 * k2, ur[k2], ui[k2]);
          }
        } else {
          float xr[N2], xi[N2], ur[N2], ui[N2];
          const float2* yb = s_y + it * N2 * FT + fl;
#pragma unroll
          for (int n2 = 0; n2 < N2; ++n2) {
            const float2 v = yb[n2 * FT];
            xr[n2] = v.x;
            xi[n2] = v.y;
          }
          __syncwarp(item_mask);
          cdft(xr, xi, ur, ui);
#pragma unroll
          for (int k2 = 0; k2 < N2; ++k2) {
            const int kc = N1 * k2;  // k = it + kc
            if (kc + H1 <= N / 2) put(it, k2, it + kc, ur[k2], ui[k2]);       // it < H1  =>  it + kc <= N/2
            else put(it, k2, N - kc - it, ur[k2], -ui[k2]);                    // conjugate mirror
          }
        }
      }
    }
    __syncthreads();

    const bool frame_ok = fl < rows;

    // ---- 4a. plain stft(): write the complex spectrum tile ------------------------------------------
    if (cplx) {
      if (!P::CPLX_DIRECT) {
        const int nb = P::NBINS;
        float2* __restrict__ dst = reinterpret_cast<float2*>(prm.out + (long long)clip * prm.out_clip_stride) + (long long)f0 * nb;
        const float2* sp = reinterpret_cast<const float2*>(s_r0);
        for (int e = tid; e < rows * nb; e += P::NTHREADS) {
          const int r = e / nb, k = e - r * nb;
          dst[e] = sp[k * P::P_PITCH + r];
        }
      }
    } else {   // (the loop-top barrier orders a complex tile's reads before the next tile's writes)

    // ---- 4b. sparse mel projection; finished values wait in the free rows of the exchange buffer (output_words()) ----
    const int M = MEL > 0 ? MelTraits<MEL>::M : prm.n_mels;
    const int out_mode = OUT >= 0 ? OUT : prm.out_mode;
    constexpr bool RAWT = B2A_RAW_TRACK && POST == POST_WNORM;   // bookkeeping on the raw mel sums (see B2A_RAW_TRACK)
    float lmax = RAWT ? 0.0f : -3.0e38f, vmin = 3.0e38f;   // of the normalised values, or (RAWT) of the raw mel sums (>= 0)
    const float log_floor = prm.log_floor;
    // Known banks: post-processing sits in the store loop, which keeps the unrolled mel code a third shorter (the tile loop
    // has to fit the instruction cache).
    auto post_pure = [&](float v) {
      // (log10(max(v, floor)) + 4) / 4 as one FMA on the MUFU log2; RAWT: no floor here (a zero sum becomes -inf, its tile's minimum
      // is below any threshold, and whisper_clamp_kernel lifts it to max(floor, Lmax - 8))
      if (POST == POST_WNORM) v = fmaf(lg2_ftz(RAWT ? v : fmaxf(v, log_floor)), 0.25f * 0.30102999566398120f, 1.0f);
      else if (POST == POST_LN) v = lg2_ftz(fmaxf(v, log_floor)) * 0.69314718055994531f;
      return v;
    };
    auto track2 = [&](float a, float b) {   // two values per three-input min / max (FMNMX3)
      if (POST == POST_WNORM) {
        lmax = fmaxf(lmax, fmaxf(a, b));
        vmin = fminf(vmin, fminf(a, b));
      }
    };
    auto post = [&](float v) {
      if (RAWT) {
        track2(v, v);
        return post_pure(v);
      }
      v = post_pure(v);
      track2(v, v);
      return v;
    };
    {
      if (BAKED) {
        mel_baked<MEL>(wsub, s_p + fl, [&](int m, float v) { s_p[out_base_words<P>(m) + fl] = v; });
      } else if (STEPS_SMEM && steps_in_smem) {
        mel_steps<true>(s_p + fl, s_steps, prm.chunk_s[wsub], prm.chunk_s[wsub + 1], s_p + fl);
      } else if (prm.fb_steps != nullptr) {
        mel_steps<false>(s_p + fl, prm.fb_steps, prm.chunk_s[wsub], prm.chunk_s[wsub + 1], s_p + fl);
      } else {
        // generic path: arbitrary filterbank, one short loop per filter
        const int ma = prm.chunk_m[wsub], mb = prm.chunk_m[wsub + 1];
        const int4* __restrict__ fdesc = prm.fb_desc;
        const float* __restrict__ fw = prm.fb_w;
        for (int m = ma; m < mb; ++m) {
          const int4 d = __ldg(fdesc + m);
          const float* __restrict__ w = fw + d.z;
          float v = 0.0f;
          for (int i = 0; i < d.y; ++i) v = fmaf(__ldg(w + i), s_p[spectrum_slot<P>(d.x + i) * FT + fl], v);
          s_p[out_base_words<P>(m) + fl] = v;
        }
      }
    }
    if (LATE_TOP) cp_async_commit_wait_all();   // the next tile's PCM (issued after stage A) becomes visible with this barrier
    __syncthreads();

    // ---- 5. store of the staged tile (baked: plain transposing copy; otherwise log / floor / scale fused in) ----
    float* __restrict__ dst = prm.out + (long long)clip * prm.out_clip_stride;
    if (FIXED) {
      constexpr int MB = MelTraits<MEL>::M;
      constexpr int NC = MB > 0 ? (MB + 31) / 32 : 1;   // (MEL == 0: the row-major branches below are dead, MB == 0)
      if (out_mode == OUT_MT) {
        // (M, T') rows: lanes run over frames
        const long long nfr = prm.n_frames;
        float* d = dst + f0 + fl;
        if constexpr (mt_table<P, MEL, OUT>()) {
          // staging offsets from the per-warp table: whole 16-byte (warp-uniform) reads, then one LDS / post / STG per row
          constexpr int SL = mt_slots<P, MEL>(), MB = MelTraits<MEL>::M;
          int off[SL];
#pragma unroll
          for (int q = 0; q < SL / 4; ++q) {
            const int4 o4 = reinterpret_cast<const int4*>(s_mt + wsub * SL)[q];
            off[4 * q] = o4.x; off[4 * q + 1] = o4.y; off[4 * q + 2] = o4.z; off[4 * q + 3] = o4.w;
          }
          if (frame_ok) {
            float* dm = d + wsub * nfr;
#pragma unroll
            for (int j = 0; j < SL; ++j) {
              if (j * NIT < MB && (j * NIT + NIT <= MB || wsub + j * NIT < MB)) *dm = post(s_p[off[j] + fl]);
              dm += NIT * nfr;
            }
          }
        } else if (frame_ok)
          if (B2A_MT_WALK) {
            OutBaseWalk<P, NIT> ob(wsub);
            float* dm = d + wsub * nfr;
            for (int m = wsub; m < M; m += NIT, ob.next(), dm += NIT * nfr) *dm = post(s_p[ob.words() + fl]);
          } else {
            for (int m = wsub; m < M; m += NIT) d[m * nfr] = post(s_p[out_base_words<P>(m) + fl]);
          }
      } else {
        // lanes run over m: one staging pointer per 32-filter chunk (bank = (m + frame) mod 32: conflict free)
        const float* srow[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) srow[c] = s_p + out_base_words<P>(c * 32 + lane < MB ? c * 32 + lane : MB - 1);
        if (out_mode == OUT_TM) {
          // (T', M) rows, every row a run of coalesced 128-byte segments; rows warp, warp + NW, ... as immediates
          float* d = dst + ((long long)f0 + warp) * MB + lane;
          __half* dh = reinterpret_cast<__half*>(prm.out) + (long long)clip * prm.out_clip_stride + ((long long)f0 + warp) * MB + lane;
          auto put = [&](int off, float v) {
            if (F16) dh[off] = __float2half_rn(v);
            else d[off] = v;
          };
          // Row slots warp, warp + NW, ...: NF of them exist for every warp (branch-free: loads unconditional -- staging
          // rows past `rows` hold the recomputed last frame -- only the stores are guarded); the remaining FT - NF*NW rows
          // go to the first warps under a warp-uniform branch.
          constexpr int NF = FT / NW;
          bool ok[NF + 1];
#pragma unroll
          for (int i = 0; i <= NF; ++i) ok[i] = warp + i * NW < rows;
          const bool extra = FT % NW != 0 && warp + NF * NW < FT;   // warp-uniform
#pragma unroll
          for (int c = 0; c < NC; ++c) {
            const bool col_ok = c < MB / 32 || lane < MB % 32;
            const float* sr = srow[c] + warp;
            float v[NF + 1];
            if (RAWT) {
              // raw sums: their minimum / maximum (one three-input FMNMX each per two values, independent of the logarithms below)
#pragma unroll
              for (int i = 0; i < NF; ++i) v[i] = sr[i * NW];
              float cmax = v[0], cmin = v[0];
#pragma unroll
              for (int i = 1; i + 1 < NF; i += 2) {
                cmax = fmaxf(cmax, fmaxf(v[i], v[i + 1]));
                cmin = fminf(cmin, fminf(v[i], v[i + 1]));
              }
              if (NF % 2 == 0) {
                cmax = fmaxf(cmax, v[NF - 1]);
                cmin = fminf(cmin, v[NF - 1]);
              }
              track2(cmax, cmin);   // (max3(lmax, cmax, cmin) = max(lmax, cmax); min3 alike)
#pragma unroll
              for (int i = 0; i < NF; ++i) v[i] = post_pure(v[i]);
            } else {
#pragma unroll
              for (int i = 0; i < NF; ++i) v[i] = post_pure(sr[i * NW]);
#pragma unroll
              for (int i = 0; i + 1 < NF; i += 2) track2(v[i], v[i + 1]);
              if (NF % 2) track2(v[NF - 1], v[NF - 1]);
            }
#pragma unroll
            for (int i = 0; i < NF; ++i)
              if (ok[i] && col_ok) put(i * NW * MB + c * 32, v[i]);
            if (extra) {
              v[NF] = post(sr[NF * NW]);
              if (ok[NF] && col_ok) put(NF * NW * MB + c * 32, v[NF]);
            }
          }
        } else {  // OUT_LFR: out[i][j*M + m] = feat[clamp(i*n + j - left, 0, T'-1)][m]   (FunASRAudio.swift:108-154)
          const int lm = prm.lfr_m, ln = prm.lfr_n, left = (lm - 1) / 2;
          const int T = int(n_frames), last_row = int(lfr_rows) - 1;
          if (B2A_LFR_FAST && lm == 7 && ln == 6 && f0 > 0 && f0 + FT < T) {
            // Interior tile of the standard 7/6 stacking (no replicated edge frame, all FT rows valid).  The output of a clip is a
            // stream of M-float slots, slot(i, j) = 7 i + j; frame t (u = t + 3) is element j = u - 6 i of row i = u / 6, i.e. slot
            // u + u / 6, and when u is a multiple of 6 also element 6 of row i - 1: the slot just before.  Rows warp, warp + NW, ...
            // of the tile as in the (T', M) store; no per-segment division, no skipped segments.
            constexpr int NF = FT / NW;
            const bool extra = FT % NW != 0 && warp + NF * NW < FT;   // warp-uniform
            auto put_row = [&](int x) {
              const int u = f0 + x + 3, i1 = u / 6;
              const bool own = i1 <= last_row, dup = u == 6 * i1;     // (i1 - 1 <= last_row always: u <= T + 1)
              float* d = dst + (long long)(u + i1) * MB + lane;
#pragma unroll
              for (int c = 0; c < NC; ++c) {
                if (c < MB / 32 || lane < MB % 32) {
                  const float v = post(srow[c][x]);
                  if (own) d[c * 32] = v;
                  if (dup) d[c * 32 - MB] = v;
                }
              }
            };
#pragma unroll
            for (int i = 0; i < NF; ++i) put_row(warp + i * NW);
            if (extra) put_row(warp + NF * NW);
          } else {
          int i_lo = f0 + left - (lm - 1) < 0 ? 0 : (f0 + left - (lm - 1)) / ln;
          int i_hi = (f0 + rows - 1 + left) / ln;
          if (f0 + rows >= T || i_hi > last_row) i_hi = last_row;
          const int nseg = (i_hi - i_lo + 1) * lm;
          for (int sg = warp; sg < nseg; sg += NW) {
            const int i = i_lo + sg / lm;
            const int j = sg - (sg / lm) * lm;
            int t = i * ln + j - left;
            t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
            if (t < f0 || t >= f0 + rows) continue;
            const int x = t - f0;
            float* d = dst + ((long long)i * lm + j) * MB + lane;
#pragma unroll
            for (int c = 0; c < MB / 32; ++c) d[c * 32] = post(srow[c][x]);
            if (MB % 32 != 0 && lane < MB % 32) d[(MB / 32) * 32] = post(srow[NC - 1][x]);
          }
          }
        }
      }
      if (POST == POST_WNORM) {
        int* tmin_p = prm.tile_min + (RAGGED ? first_tile : clip * tpc) + tile;
        if (RAWT) {
          // raw sums are >= +0: their bit patterns order like integers.  Lane 0 turns the two extremes into normalised values (the
          // maximum with the floor, as every stored value will have it after the clamp kernel) and publishes them with plain REDs.
          const int wmax = __reduce_max_sync(0xffffffffu, __float_as_int(lmax));
          const int wmin = __reduce_min_sync(0xffffffffu, __float_as_int(vmin));
          if (lane == 0) {
            const float nmax = fmaf(lg2_ftz(fmaxf(__int_as_float(wmax), log_floor)), 0.25f * 0.30102999566398120f, 1.0f);
            const float nmin = fmaf(lg2_ftz(__int_as_float(wmin)), 0.25f * 0.30102999566398120f, 1.0f);
            red_max_s32(prm.clip_max + clip, enc_ordered(nmax));
            if (prm.tile_max != nullptr) red_max_s32(prm.tile_max + (tmin_p - prm.tile_min), enc_ordered(nmax));
            red_max_s32(tmin_p, enc_ordered(-nmin));   // the NEGATED minimum, ordered-int encoded: one 0x80 memset initialises clip_max and tile_min alike
          }
        } else {
          const int wmax = __reduce_max_sync(0xffffffffu, enc_ordered(lmax));
          const int wmin = __reduce_min_sync(0xffffffffu, enc_ordered(vmin));
          if (lane == 0) {
            atomicMax(prm.clip_max + clip, wmax);
            atomicMax(tmin_p, enc_ordered(-dec_ordered(wmin)));  // the NEGATED minimum, ordered-int encoded: one 0x80 memset initialises clip_max and tile_min alike
          }
        }
      }
    } else {
      const int log_mode = prm.log_mode;
      const float log_floor = prm.log_floor;
      const bool wnorm = prm.whisper_norm != 0;
      if (out_mode == OUT_TM) {
        // (T', M) rows: lanes run over m
        dst += (long long)f0 * M;
        auto store_tm = [&](auto post) {
          for (int c = lane; c < M; c += 32) {
            const float* sr = s_p + out_base_words<P>(c);
            for (int r = warp; r < rows; r += NW) dst[r * M + c] = post(sr[r]);
          }
        };
        if (wnorm) store_tm([&](float v) { return mel_post<LOG_LOG10, true>(v, log_floor, lmax, vmin); });
        else if (log_mode == LOG_LN) store_tm([&](float v) { return mel_post<LOG_LN, false>(v, log_floor, lmax, vmin); });
        else if (log_mode == LOG_LOG10) store_tm([&](float v) { return mel_post<LOG_LOG10, false>(v, log_floor, lmax, vmin); });
        else if (log_mode == LOG_DB20) store_tm([&](float v) { return mel_post<LOG_DB20, false>(v, log_floor, lmax, vmin); });
        else store_tm([&](float v) { return v; });
      } else if (out_mode == OUT_MT) {
        // (M, T') rows: lanes run over frames
        dst += f0 + fl;
        const long long nfr = prm.n_frames;
        const bool post = prm.post_affine != 0;
        auto store_mt = [&](auto fn) {
          OutBaseWalk<P, NIT> ob(wsub);
          for (int m = wsub; m < M; m += NIT, ob.next()) {
            float v = fn(s_p[(B2A_MT_WALK ? ob.words() : out_base_words<P>(m)) + fl]);
            if (post) v = (v - prm.post_sub) / prm.post_div;
            if (frame_ok) dst[m * nfr] = v;
          }
        };
        if (wnorm) store_mt([&](float v) { return mel_post<LOG_LOG10, true>(v, log_floor, lmax, vmin); });
        else if (log_mode == LOG_LN) store_mt([&](float v) { return mel_post<LOG_LN, false>(v, log_floor, lmax, vmin); });
        else if (log_mode == LOG_LOG10) store_mt([&](float v) { return mel_post<LOG_LOG10, false>(v, log_floor, lmax, vmin); });
        else if (log_mode == LOG_DB20) store_mt([&](float v) { return mel_post<LOG_DB20, false>(v, log_floor, lmax, vmin); });
        else store_mt([&](float v) { return v; });
      } else {  // OUT_LFR: out[i][j*M + m] = ln(feat)[clamp(i*n + j - left, 0, T'-1)][m]   (FunASRAudio.swift:108-154)
        const int lm = prm.lfr_m, ln = prm.lfr_n, left = (lm - 1) / 2;
        const long long T = n_frames;
        long long i_lo = (f0 + left - (lm - 1)) / ln;
        if (f0 + left - (lm - 1) < 0) i_lo = 0;
        long long i_hi = (f0 + rows - 1 + left) / ln;
        if (f0 + rows >= T) i_hi = lfr_rows - 1;
        if (i_hi > lfr_rows - 1) i_hi = lfr_rows - 1;
        const int nseg = int(i_hi - i_lo + 1) * lm;
        for (int sg = warp; sg < nseg; sg += NW) {
          const long long i = i_lo + sg / lm;
          const int j = sg % lm;
          long long t = i * ln + j - left;
          t = t < 0 ? 0 : (t > T - 1 ? T - 1 : t);
          if (t < f0 || t >= f0 + rows) continue;
          const int x = int(t - f0);
          float* d = dst + (i * lm + j) * (long long)M;
          for (int c = lane; c < M; c += 32) {
            float v = s_p[out_base_words<P>(c) + x];
            if (log_mode == LOG_LN) v = lg2_ftz(fmaxf(v, log_floor)) * 0.69314718055994531f;
            else if (log_mode == LOG_LOG10) v = lg2_ftz(fmaxf(v, log_floor)) * 0.30102999566398120f;
            d[c] = v;
          }
        }
      }
    }
    if (!FIXED && prm.whisper_norm) {
      const int wmax = __reduce_max_sync(0xffffffffu, enc_ordered(lmax));
      const int wmin = __reduce_min_sync(0xffffffffu, enc_ordered(vmin));
      if (lane == 0) {
        atomicMax(prm.clip_max + clip, wmax);
        atomicMax(prm.tile_min + (RAGGED ? first_tile : clip * tpc) + tile, enc_ordered(-dec_ordered(wmin)));  // the NEGATED minimum, ordered-int encoded: one 0x80 memset initialises clip_max and tile_min alike
      }
    }
    }  // !cplx
    if (!EARLY_PREFETCH) {
      // plain stft(), single buffer: the next tile's PCM can only be staged once every warp is done with the complex tile
      __syncthreads();
      if (nclip < n_clips) stage_pcm<P>(prm, smem, nclip, ntile * FT, nn_samples, nn_samples + zero_tail, tid, lane, warp);
    }
    clip = nclip;
    tile = ntile;
    if (RAGGED) {
      g = ng;
      g_next = g_next2;
      if (clip < n_clips) set_clip(nt, g);
    }
  }
}

// Rewrites only the tiles whose minimum lies below the clip's clamp threshold:
//   (max(L, Lmax - 8) + 4) / 4 == max((L + 4) / 4, (Lmax + 4) / 4 - 2)   (x -> (x+4)/4 is monotone)
// WhisperAudio.swift:130-134, S3TokenizerUtils.swift:203-205.
// One CTA per (clip, group of `group` <= kClampTilesPerCta tiles; small launches take smaller groups, down to one tile per CTA: for one 30 s
// clip three CTAs walking 32 tiles each took 18 us, more than the front-end kernel itself): the group's tile minima are tested in parallel (one round trip to
// memory), the tiles that need it are rewritten with 16-byte accesses.
constexpr int kClampTilesPerCta = 32;
// tiles per CTA of the clamp kernel: enough CTAs to fill the GPU for small launches (one tile per CTA up to ~2000 tiles), kClampTilesPerCta for large ones
static inline int clamp_group(long long total_tiles) {
  int g = 1;
  while (g < kClampTilesPerCta && total_tiles / g > 2048) g *= 2;
  return g;
}
constexpr int kFillFlag = 0x40000000;   // s_list entry: the tile was skipped by the main kernel (kTileFill), write the threshold without reading
// clip_tab != null (ragged batch): the clip's own frame count and first tile; `n_frames` stays the (M, T') row stride.
__global__ void __launch_bounds__(256) whisper_clamp_kernel(float* out, const int* clip_max, const int* tile_min, int tiles_per_clip,
                                                            long long n_frames, int n_mels, long long out_clip_stride, int out_mode, int ft,
                                                            const int4* __restrict__ clip_tab, int f16, float log_floor, int group,
                                                            const int* __restrict__ tile_max) {
  __shared__ int s_list[kClampTilesPerCta];
  __shared__ int s_count;
  // launched with programmatic stream serialisation behind the front-end kernel (launch_plan): the blocks may already be resident
  // while that kernel drains; nothing it wrote is read before this point (a no-op for a plain launch)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  const long long clip = blockIdx.y;
  const long long mt_stride = n_frames;
  long long tile_base = clip * tiles_per_clip;
  if (clip_tab != nullptr) {
    const int4 ci = clip_tab[clip];
    n_frames = ci.y;
    tiles_per_clip = int((n_frames + ft - 1) / ft);
    tile_base = ci.w;
  }
  // clip_max holds the maximum of the normalised values: ((Lmax - 8) + 4) / 4 = (Lmax + 4) / 4 - 2.  log_floor > 0 (B2A_RAW_TRACK): the main
  // kernel stored un-floored values -- max(max(L, floor), Lmax - 8) = max(L, max(floor, Lmax - 8)) -- and the normalised floor is formed here
  // with the expression (and the MUFU) the store loop used for floored values in round 1
  const float norm_floor = log_floor > 0.0f ? fmaf(lg2_ftz(log_floor), 0.25f * 0.30102999566398120f, 1.0f) : -3.0e38f;
  const float thr = fmaxf(dec_ordered(clip_max[clip]) - 2.0f, norm_floor);
  const int t0 = blockIdx.x * group;
  if (threadIdx.x == 0) s_count = 0;
  __syncthreads();
  if (int(threadIdx.x) < group && t0 + int(threadIdx.x) < tiles_per_clip) {
    const int tm = tile_min[tile_base + t0 + threadIdx.x];   // the negated minimum, or kTileFill: a tile of silence the main kernel skipped
    // (tile_max: a tile whose largest value does not exceed the threshold becomes the threshold everywhere -- written without reading)
    if (tm == kTileFill || (tile_max != nullptr && dec_ordered(tile_max[tile_base + t0 + threadIdx.x]) <= thr))
      s_list[atomicAdd(&s_count, 1)] = (t0 + threadIdx.x) | kFillFlag;
    else if (-dec_ordered(tm) < thr) s_list[atomicAdd(&s_count, 1)] = t0 + threadIdx.x;   // (order within the list is irrelevant)
  }
  __syncthreads();
  const int count = s_count;
  float* o = out + clip * out_clip_stride;
  for (int i = 0; i < count; ++i) {
    const bool fill = (s_list[i] & kFillFlag) != 0;   // nothing valid was written there: every value is the floor, i.e. becomes thr
    const int t = s_list[i] & ~kFillFlag;
    const long long f0 = (long long)t * ft;
    const int rows = int(n_frames - f0 < ft ? n_frames - f0 : ft);
    if (out_mode == OUT_TM && f16) {
      // fp16 features: rounding is monotone, so max(half(v), half(thr)) == half(max(v, thr))
      __half* d = reinterpret_cast<__half*>(out) + clip * out_clip_stride + f0 * n_mels;
      const int n = rows * n_mels;
      const __half th = __float2half_rn(thr);
      if ((reinterpret_cast<uintptr_t>(d) & 15) == 0 && (n & 7) == 0) {
        const __half2 th2 = __half2half2(th);
        uint4* d4 = reinterpret_cast<uint4*>(d);
        for (int e = threadIdx.x; e < n / 8; e += blockDim.x) {
          uint4 v;
          __half2* h = reinterpret_cast<__half2*>(&v);
          if (fill) {
            h[0] = h[1] = h[2] = h[3] = th2;
          } else {
            v = d4[e];
            h[0] = __hmax2(h[0], th2); h[1] = __hmax2(h[1], th2); h[2] = __hmax2(h[2], th2); h[3] = __hmax2(h[3], th2);
          }
          d4[e] = v;
        }
      } else {
        for (int e = threadIdx.x; e < n; e += blockDim.x) d[e] = fill ? th : __hmax(d[e], th);
      }
    } else if (out_mode == OUT_TM) {
      float* d = o + f0 * n_mels;
      const int n = rows * n_mels;
      if ((reinterpret_cast<uintptr_t>(d) & 15) == 0 && (n & 3) == 0) {
        float4* d4 = reinterpret_cast<float4*>(d);
        for (int e = threadIdx.x; e < n / 4; e += blockDim.x) {
          float4 v = make_float4(thr, thr, thr, thr);
          if (!fill) {
            v = d4[e];
            v.x = fmaxf(v.x, thr); v.y = fmaxf(v.y, thr); v.z = fmaxf(v.z, thr); v.w = fmaxf(v.w, thr);
          }
          d4[e] = v;
        }
      } else {
        for (int e = threadIdx.x; e < n; e += blockDim.x) d[e] = fill ? thr : fmaxf(d[e], thr);
      }
    } else {  // OUT_MT: lanes over the tile's (<= 32) frames, warps over the filters -- no per-element division
      const int r = threadIdx.x & 31;
      if (r < rows)
        for (int m = threadIdx.x >> 5; m < n_mels; m += int(blockDim.x >> 5)) {
          float* d = o + (long long)m * mt_stride + f0 + r;
          *d = fill ? thr : fmaxf(*d, thr);
        }
    }
  }
}

// ------------------------------------------------------------------------------------------------
// per-clip column statistics: CMVN (FunASRAudio.swift:165-180) and time-mean removal (CAMPPlus.swift:800)
// block = 32 columns x 8 row groups; grid = (column chunks, batch)
// ------------------------------------------------------------------------------------------------
// `in` and `out` may be the same buffer (CMVN / mean-norm run in place on the front end's output): no __restrict__, no
// non-coherent loads on them; every thread re-reads only elements it has not yet written.
// clip_tab != null (ragged batch): the statistics run over the clip's own rows (component `tab_rows` of its clip_tab entry:
// 1 = frames, 2 = LFR rows); the clip stride stays `rows` (the longest clip's) x dim.
__global__ void __launch_bounds__(256) colstat_kernel(const float* in, float* out, long long rows,
                                                      int dim, const float* __restrict__ gmean, const float* __restrict__ gistd,
                                                      int do_var, const int4* __restrict__ clip_tab, int tab_rows) {
  __shared__ float s_red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const long long clip = blockIdx.y;
  const float* src = in + clip * rows * dim;
  float* dst = out + clip * rows * dim;
  if (clip_tab != nullptr) {
    const int4 ci = clip_tab[clip];
    rows = tab_rows == 2 ? ci.z : ci.y;
  }
  const bool ok = col < dim;
  if (gmean != nullptr) {  // (x + mean) * istd with precomputed statistics
    if (ok) {
      const float a = gmean[col], b = gistd[col];
      for (long long r = ry; r < rows; r += 8) dst[r * dim + col] = (src[r * dim + col] + a) * b;
    }
    return;
  }
  float s = 0.0f;
  // (eight independent loads per thread in flight before the serial adds: the kernel is bound by memory latency)
  if (ok) {
    long long r = ry;
    for (; r + 56 < rows; r += 64) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(r + 8 * u) * dim + col];
#pragma unroll
      for (int u = 0; u < 8; ++u) s += v[u];
    }
    for (; r < rows; r += 8) s += src[r * dim + col];
  }
  s_red[ry][cx] = s;
  __syncthreads();
  float mean = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) mean += s_red[i][cx];
  mean = mean / float(rows);
  __syncthreads();
  float denom = 1.0f;
  if (do_var) {
    float q = 0.0f;
    if (ok) {
      long long r = ry;
      for (; r + 56 < rows; r += 64) {
        float v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[(r + 8 * u) * dim + col];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const float d = v[u] - mean; q = fmaf(d, d, q); }
      }
      for (; r < rows; r += 8) { const float d = src[r * dim + col] - mean; q = fmaf(d, d, q); }
    }
    s_red[ry][cx] = q;
    __syncthreads();
    float var = 0.0f;
#pragma unroll
    for (int i = 0; i < 8; ++i) var += s_red[i][cx];
    var = var / float(rows);
    denom = sqrtf(var) + 1e-6f;
  }
  if (ok) {
    long long r = ry;
    for (; r + 56 < rows; r += 64) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(r + 8 * u) * dim + col];
#pragma unroll
      for (int u = 0; u < 8; ++u) dst[(r + 8 * u) * dim + col] = do_var ? (v[u] - mean) / denom : v[u] - mean;
    }
    for (; r < rows; r += 8) dst[r * dim + col] = do_var ? (src[r * dim + col] - mean) / denom : src[r * dim + col] - mean;
  }
}

// The same statistics for clips of at most 8 * RPT rows (Fun-ASR CMVN: ceil(2001 / 6) = 334 rows of 560): a thread keeps its
// <= RPT rows of the column in REGISTERS, so the slab is read once (all loads in flight together) and the variance and the
// normalisation run out of registers -- no second and third pass over L2, a third of the instructions of colstat_kernel (which
// is bound by instruction issue, not by memory: ~39 instructions per element, ten of them the IEEE division).  Same layout
// (32 columns x 8 row groups), same summation order per column.  The quotient (x - mean) / denom is one reciprocal per column
// and a Newton correction per element (the fast path of the IEEE division without its range check; the denominators are
// >= 1e-6 and the quotients far from the fp32 range limits).
template <int RPT>
__global__ void __launch_bounds__(256) colstat_reg_kernel(const float* in, float* out, long long rows, int dim,
                                                          int do_var, const int4* __restrict__ clip_tab, int tab_rows) {
  __shared__ float s_red[8][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const long long clip = blockIdx.y;
  int nrows = int(rows);
  if (clip_tab != nullptr) {
    const int4 ci = clip_tab[clip];
    nrows = tab_rows == 2 ? ci.z : ci.y;
  }
  const bool ok = col < dim;
  const int nr = ok && nrows > ry ? (nrows - ry + 7) >> 3 : 0;   // rows ry, ry + 8, ... of this thread
  // rows of the thread: one 64-bit pointer walked by a 32-bit byte step
  const long long off0 = clip * rows * dim + (long long)ry * dim + col;
  const char* __restrict__ src = reinterpret_cast<const char*>(in + off0);
  char* __restrict__ dst = reinterpret_cast<char*>(out + off0);
  const unsigned step = 32u * unsigned(dim);   // bytes between rows r and r + 8
  float v[RPT];
#pragma unroll
  for (int u = 0; u < RPT; ++u) {
    v[u] = u < nr ? *reinterpret_cast<const float*>(src) : 0.0f;
    src += step;
  }
  float s = 0.0f;
#pragma unroll
  for (int u = 0; u < RPT; ++u) s += v[u];
  s_red[ry][cx] = s;
  __syncthreads();
  float mean = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) mean += s_red[i][cx];
  mean = mean / float(nrows);
  if (!do_var) {
#pragma unroll
    for (int u = 0; u < RPT; ++u) {
      const float r = v[u] - mean;
      if (u < nr) *reinterpret_cast<float*>(dst) = r;
      dst += step;
    }
    return;
  }
  __syncthreads();
  float q = 0.0f;
#pragma unroll
  for (int u = 0; u < RPT; ++u) {
    v[u] -= mean;
    q = fmaf(u < nr ? v[u] : 0.0f, v[u], q);
  }
  s_red[ry][cx] = q;
  __syncthreads();
  float var = 0.0f;
#pragma unroll
  for (int i = 0; i < 8; ++i) var += s_red[i][cx];
  var = var / float(nrows);
  const float denom = sqrtf(var) + 1e-6f;
  const float inv = 1.0f / denom;
#pragma unroll
  for (int u = 0; u < RPT; ++u) {
    const float q0 = v[u] * inv;
    const float r = fmaf(fmaf(-q0, denom, v[u]), inv, q0);
    if (u < nr) *reinterpret_cast<float*>(dst) = r;
    dst += step;
  }
}

// The same statistics for clips whose CW-column slab (rows x CW floats) fits shared memory (CAM++ mean-norm: 1998 rows of 80):
// the slab is copied in ONCE with 16-byte cp.async (all copies of the block in flight together), reduced and normalised out of
// shared memory, and written back -- one DRAM read and one write instead of the two reads of the streaming kernels, whose
// slabs (512 resident clips x 639 KB) do not survive in the L2.  Three blocks per SM overlap each other's load / store phases.
// Column sums run over rows g, g + G, ... (G = 256 / CW row groups) and are combined in fixed order: deterministic.
template <int CW>
__global__ void __launch_bounds__(256) colstat_smem_kernel(const float* in, float* out, long long rows, int dim,
                                                           int do_var, const int4* __restrict__ clip_tab, int tab_rows) {
  extern __shared__ __align__(16) float s_slab[];   // [row][CW]
  __shared__ float s_red[256];
  __shared__ float s_stat[2 * CW];
  constexpr int Q = CW / 4, G = 256 / CW;
  const int tid = threadIdx.x;
  const long long clip = blockIdx.y;
  const int c0 = int(blockIdx.x) * CW;
  int nrows = int(rows);
  if (clip_tab != nullptr) {
    const int4 ci = clip_tab[clip];
    nrows = tab_rows == 2 ? ci.z : ci.y;
  }
  const float* __restrict__ src = in + clip * rows * dim + c0;
  float* __restrict__ dst = out + clip * rows * dim + c0;
  const int cwq = (dim - c0) / 4 < Q ? (dim - c0) / 4 : Q;   // float4 columns of this chunk (the last chunk may be narrower)
  const int q = tid % Q, r0 = tid / Q;
  if (q < cwq) {
    const unsigned sbase = static_cast<unsigned>(__cvta_generic_to_shared(s_slab + 4 * q));
    for (int r = r0; r < nrows; r += 256 / Q)
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + unsigned(r) * (CW * 4)), "l"(src + (long long)r * dim + 4 * q));
  }
  cp_async_commit_wait_all();
  __syncthreads();
  const int c = tid % CW, g = tid / CW;   // (columns past a narrow last chunk hold stale shared memory: computed, never stored)
  {
    float a = 0.0f;
    for (int r = g; r < nrows; r += G) a += s_slab[r * CW + c];
    s_red[tid] = a;
  }
  __syncthreads();
  if (tid < CW) {
    float m = 0.0f;
#pragma unroll 8
    for (int i = 0; i < G; ++i) m += s_red[i * CW + tid];
    s_stat[tid] = m / float(nrows);
    s_stat[CW + tid] = 1.0f;
  }
  __syncthreads();
  if (do_var) {
    const float mean = s_stat[c];
    float a = 0.0f;
    for (int r = g; r < nrows; r += G) {
      const float d = s_slab[r * CW + c] - mean;
      a = fmaf(d, d, a);
    }
    s_red[tid] = a;
    __syncthreads();
    if (tid < CW) {
      float v = 0.0f;
#pragma unroll 8
      for (int i = 0; i < G; ++i) v += s_red[i * CW + tid];
      s_stat[CW + tid] = sqrtf(v / float(nrows)) + 1e-6f;
    }
    __syncthreads();
  }
  if (q < cwq) {
    const float4 mean = *reinterpret_cast<const float4*>(s_stat + 4 * q);
    const float4 den = *reinterpret_cast<const float4*>(s_stat + CW + 4 * q);
    for (int r = r0; r < nrows; r += 256 / Q) {
      float4 v = *reinterpret_cast<const float4*>(s_slab + r * CW + 4 * q);
      v.x -= mean.x; v.y -= mean.y; v.z -= mean.z; v.w -= mean.w;
      if (do_var) { v.x = v.x / den.x; v.y = v.y / den.y; v.z = v.z / den.z; v.w = v.w / den.w; }
      *reinterpret_cast<float4*>(dst + (long long)r * dim + 4 * q) = v;
    }
  }
}

// The same statistics with a thread owning FOUR adjacent columns (16-byte loads and stores: four times the bytes in flight per
// thread; the kernel is bound by memory latency): used when dim % 4 == 0 and the buffers are 16-byte aligned.  Same two-pass
// arithmetic per column as colstat_kernel (RG = 8 also sums in the same order).
template <int RG>
__global__ void __launch_bounds__(32 * RG) colstat4_kernel(const float4* in, float4* out, long long rows, int dim4,
                                                          int do_var, const int4* __restrict__ clip_tab, int tab_rows) {
  __shared__ float4 s_red[RG][33];
  const int cx = threadIdx.x & 31, ry = threadIdx.x >> 5;
  const int col = blockIdx.x * 32 + cx;
  const long long clip = blockIdx.y;
  const float4* src = in + clip * rows * dim4;
  float4* dst = out + clip * rows * dim4;
  if (clip_tab != nullptr) {
    const int4 ci = clip_tab[clip];
    rows = tab_rows == 2 ? ci.z : ci.y;
  }
  const bool ok = col < dim4;
  float4 s = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
  if (ok) {
    long long r = ry;
    for (; r + 7 * RG < rows; r += 8 * RG) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(r + RG * u) * dim4 + col];
#pragma unroll
      for (int u = 0; u < 8; ++u) { s.x += v[u].x; s.y += v[u].y; s.z += v[u].z; s.w += v[u].w; }
    }
    for (; r < rows; r += RG) {
      const float4 v = src[r * dim4 + col];
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
  }
  s_red[ry][cx] = s;
  __syncthreads();
  float4 mean = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
  for (int i = 0; i < RG; ++i) { const float4 t = s_red[i][cx]; mean.x += t.x; mean.y += t.y; mean.z += t.z; mean.w += t.w; }
  const float fr = float(rows);
  mean.x = mean.x / fr; mean.y = mean.y / fr; mean.z = mean.z / fr; mean.w = mean.w / fr;
  __syncthreads();
  float4 denom = make_float4(1.0f, 1.0f, 1.0f, 1.0f);
  if (do_var) {
    float4 q = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    auto acc = [&](const float4& v) {
      float d = v.x - mean.x; q.x = fmaf(d, d, q.x);
      d = v.y - mean.y; q.y = fmaf(d, d, q.y);
      d = v.z - mean.z; q.z = fmaf(d, d, q.z);
      d = v.w - mean.w; q.w = fmaf(d, d, q.w);
    };
    if (ok) {
      long long r = ry;
      for (; r + 7 * RG < rows; r += 8 * RG) {
        float4 v[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) v[u] = src[(r + RG * u) * dim4 + col];
#pragma unroll
        for (int u = 0; u < 8; ++u) acc(v[u]);
      }
      for (; r < rows; r += RG) acc(src[r * dim4 + col]);
    }
    s_red[ry][cx] = q;
    __syncthreads();
    float4 var = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
#pragma unroll
    for (int i = 0; i < RG; ++i) { const float4 t = s_red[i][cx]; var.x += t.x; var.y += t.y; var.z += t.z; var.w += t.w; }
    denom = make_float4(sqrtf(var.x / fr) + 1e-6f, sqrtf(var.y / fr) + 1e-6f, sqrtf(var.z / fr) + 1e-6f, sqrtf(var.w / fr) + 1e-6f);
  }
  if (ok) {
    auto norm = [&](const float4& v) {
      return do_var ? make_float4((v.x - mean.x) / denom.x, (v.y - mean.y) / denom.y, (v.z - mean.z) / denom.z, (v.w - mean.w) / denom.w)
                    : make_float4(v.x - mean.x, v.y - mean.y, v.z - mean.z, v.w - mean.w);
    };
    long long r = ry;
    for (; r + 7 * RG < rows; r += 8 * RG) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = src[(r + RG * u) * dim4 + col];
#pragma unroll
      for (int u = 0; u < 8; ++u) dst[(r + RG * u) * dim4 + col] = norm(v[u]);
    }
    for (; r < rows; r += RG) dst[r * dim4 + col] = norm(src[r * dim4 + col]);
  }
}

// standalone applyLFR (FunASRAudio.swift:108-154): one warp per (row, slot) segment
__global__ void lfr_kernel(const float* __restrict__ in, float* __restrict__ out, long long n_frames, int n_mels, int lfr_m,
                           int lfr_n, long long lfr_rows) {
  const long long clip = blockIdx.y;
  const int left = (lfr_m - 1) / 2;
  const long long seg = (long long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (seg >= lfr_rows * lfr_m) return;
  const long long i = seg / lfr_m;
  const int j = int(seg - i * lfr_m);
  long long t = i * lfr_n + j - left;
  t = t < 0 ? 0 : (t > n_frames - 1 ? n_frames - 1 : t);
  const float* s = in + (clip * n_frames + t) * n_mels;
  float* d = out + (clip * lfr_rows * lfr_m + seg) * n_mels;
  for (int c = threadIdx.x & 31; c < n_mels; c += 32) d[c] = s[c];
}

// 16-bit PCM -> fp32 in [-1, 1): x / 32768, exact in fp32 (what AVAudioFile's float processing format hands the reference for a
// 16-bit file, STT/Whisper/WhisperEngine.swift:327-369).  One thread = 8 samples: one 16-byte load, two 16-byte stores.
__global__ void __launch_bounds__(256) pcm16_to_f32_kernel(const short* __restrict__ in, float* __restrict__ out, long long n) {
  const long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (e0 >= n) return;
  constexpr float k = 1.0f / 32768.0f;
  if (e0 + 8 <= n && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0) {
    const uint4 v = __ldg(reinterpret_cast<const uint4*>(in + e0));
    const short* h = reinterpret_cast<const short*>(&v);
    reinterpret_cast<float4*>(out + e0)[0] = make_float4(float(h[0]) * k, float(h[1]) * k, float(h[2]) * k, float(h[3]) * k);
    reinterpret_cast<float4*>(out + e0)[1] = make_float4(float(h[4]) * k, float(h[5]) * k, float(h[6]) * k, float(h[7]) * k);
  } else {
    for (long long e = e0; e < e0 + 8 && e < n; ++e) out[e] = float(in[e]) * k;
  }
}

int launch_pcm16_to_f32(const void* in_i16, float* out, int64_t n, void* stream, int* launches, std::string* err) {
  const long long blocks = (n + 2047) / 2048;
  if (blocks <= 0 || blocks > 0x7fffffffLL) {
    if (err) *err = "pcm16_to_f32: bad size";
    return B2A_E_BAD_ARG;
  }
  pcm16_to_f32_kernel<<<unsigned(blocks), 256, 0, static_cast<cudaStream_t>(stream)>>>(static_cast<const short*>(in_i16), out, n);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("pcm16_to_f32_kernel launch: ") + cudaGetErrorString(e);
    return B2A_E_CUDA;
  }
  *launches += 1;
  return B2A_OK;
}

// padOrTrim (WhisperAudio.swift:54-67)
__global__ void pad_or_trim_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, long long length) {
  const long long clip = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < length) out[clip * length + i] = i < n ? in[clip * n + i] : 0.0f;
}

// reflectPad / reflectPad1D (S3TokenizerUtils.swift:266-298, FunASRAudio.swift:280-310) as a stand-alone call: the index map the
// fused front ends apply on the fly (fetch_padded), including the reference's short-input loops
__global__ void reflect_pad_kernel(const float* __restrict__ in, float* __restrict__ out, long long n, long long pad) {
  const long long clip = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n + 2 * pad) out[clip * (n + 2 * pad) + i] = fetch_padded(in + clip * n, i, pad, n, n, PAD_REFLECT);
}

// Whisper seek window (STT/Whisper/WhisperSTT.swift:171-182,624-635): rows
// [seek, seek + min(length, content_frames - seek)) of a clip's (T', M) log-mel, zero-padded to `length` rows, cast to fp16
// (round to nearest even, like MLX asType(.float16)).  One thread = 8 consecutive values: two 16-byte loads, one 16-byte store.
// seek_content: per clip (seek, content_frames) as int64 pairs.
__global__ void __launch_bounds__(256) mel_segment_f16_kernel(const float* __restrict__ mel, __half* __restrict__ out, long long n_frames,
                                                              int n_mels, const long long* __restrict__ seek_content, int length) {
  const long long clip = blockIdx.y;
  const long long seek = seek_content[2 * clip], content = seek_content[2 * clip + 1];
  long long seg = content - seek;
  seg = seg < 0 ? 0 : (seg > length ? length : seg);
  if (seek + seg > n_frames) seg = n_frames - seek > 0 ? n_frames - seek : 0;   // never read past the mel
  const long long total = (long long)length * n_mels;
  const float* __restrict__ src = mel + (clip * n_frames + seek) * n_mels;
  __half* __restrict__ dst = out + clip * total;
  const long long valid = seg * n_mels;   // the segment's rows are contiguous in both tensors
  const long long e0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (e0 >= total) return;
  const bool vec = ((reinterpret_cast<uintptr_t>(src) | reinterpret_cast<uintptr_t>(dst)) & 15) == 0 && e0 + 8 <= total;
  if (vec && e0 + 8 <= valid) {
    const float4 a = __ldg(reinterpret_cast<const float4*>(src + e0)), b = __ldg(reinterpret_cast<const float4*>(src + e0 + 4));
    __half2 h[4] = {__floats2half2_rn(a.x, a.y), __floats2half2_rn(a.z, a.w), __floats2half2_rn(b.x, b.y), __floats2half2_rn(b.z, b.w)};
    *reinterpret_cast<uint4*>(dst + e0) = *reinterpret_cast<const uint4*>(h);
  } else if (vec && e0 >= valid) {
    *reinterpret_cast<uint4*>(dst + e0) = make_uint4(0u, 0u, 0u, 0u);
  } else {
    for (long long e = e0; e < e0 + 8 && e < total; ++e) dst[e] = e < valid ? __float2half_rn(__ldg(src + e)) : __float2half_rn(0.0f);
  }
}

int launch_mel_segment_f16(const float* mel, void* out_f16, int64_t batch, int64_t n_frames, int n_mels, const long long* d_seek_content,
                           int length, void* stream, int* launches, std::string* err) {
  const long long total = (long long)length * n_mels;
  dim3 grid(unsigned((total + 2047) / 2048), unsigned(batch));
  mel_segment_f16_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mel, static_cast<__half*>(out_f16), n_frames, n_mels,
                                                                              d_seek_content, length);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("mel_segment_f16_kernel launch: ") + cudaGetErrorString(e);
    return B2A_E_CUDA;
  }
  *launches += 1;
  return B2A_OK;
}

// S3Tokenizer long-audio windows (Codec/S3Tokenizer/S3Tokenizer.swift:499-571): segment s of the
// unified batch = mel[batch_idx[s]][:, start[s] ..< start[s] + len[s]] zero-padded to `window` frames.  (M, T) rows are
// contiguous along time: one warp-coalesced copy per row.  seg: per segment (batch index, start, length) as int32 triples.
__global__ void __launch_bounds__(256) mel_windows_kernel(const float* __restrict__ mel, float* __restrict__ out, int n_mels, long long t_max,
                                                          const int* __restrict__ seg, int window) {
  const int s = blockIdx.z, m = blockIdx.y;
  const int b = seg[3 * s], start = seg[3 * s + 1], len = seg[3 * s + 2];
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= window) return;
  const float* __restrict__ src = mel + ((long long)b * n_mels + m) * t_max + start;
  out[((long long)s * n_mels + m) * window + t] = t < len ? __ldg(src + t) : 0.0f;
}

int launch_mel_windows(const float* mel, float* out, int n_mels, int64_t t_max, const int* d_seg, int n_segments, int window, void* stream,
                       int* launches, std::string* err) {
  for (int s0 = 0; s0 < n_segments; s0 += 65535) {   // gridDim.z limit
    const int ns = std::min(65535, n_segments - s0);
    dim3 grid(unsigned((window + 255) / 256), unsigned(n_mels), unsigned(ns));
    mel_windows_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(mel, out + (long long)s0 * n_mels * window, n_mels, t_max,
                                                                            d_seg + 3 * s0, window);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
      if (err) *err = std::string("mel_windows_kernel launch: ") + cudaGetErrorString(e);
      return B2A_E_CUDA;
    }
    *launches += 1;
  }
  return B2A_OK;
}

// resampleAudio / linearInterpolate1d (TTS/CosyVoice2/CosyVoice2TTS.swift:733-744,
// TTS/CosyVoice2/HiFiGAN/CosyHiFTGenerator.swift:17-58): PyTorch-style align_corners=False linear interpolation, every
// step in fp32 in the reference's op order (the source index is computed in fp32, so the order matters for bit-exactness;
// __fmul_rn / __fadd_rn keep the compiler from contracting them into FMAs).
__global__ void __launch_bounds__(256) resample_linear_kernel(const float* __restrict__ x, float* __restrict__ out, long long T, long long new_t,
                                                              float step /* Float(T) / Float(newT) */, float hi_clip /* Float(T) - 1.001 */) {
  const long long clip = blockIdx.y;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= new_t) return;
  float idx = __fadd_rn(__fmul_rn(__fadd_rn(float(int(i)), 0.5f), step), -0.5f);   // (arange(int32).asType(float32) + 0.5) * step - 0.5
  idx = fminf(fmaxf(idx, 0.0f), hi_clip);
  const float fl = floorf(idx);
  // (T == 1: hi_clip = -0.001 < 0 leaves idx = -0.001 and floor -1; the reference's negative index wraps to x[T - 1] = x[0])
  const int lo = int(fl) < 0 ? int(fl) + int(T) : int(fl);
  const int hi = int(fl) + 1 < int(T - 1) ? int(fl) + 1 : int(T - 1);
  const float wh = __fadd_rn(idx, -fl);
  const float wl = __fadd_rn(1.0f, -wh);
  const float* __restrict__ xc = x + clip * T;
  out[clip * new_t + i] = __fadd_rn(__fmul_rn(__ldg(xc + lo), wl), __fmul_rn(__ldg(xc + hi), wh));
}

// Polyphase resampler (non-parity extension, see b2a_resample_poly_filter): y[j] = sum_i x[i] h[(j + pre) * down - i * up] over the taps
// that exist, zeros outside the clip.  One thread per output sample; the filter (a few hundred to a few thousand taps) is read
// through the read-only path (every thread of a warp walks it at stride `up` from a phase that repeats every `up` outputs).
__global__ void __launch_bounds__(256) resample_poly_kernel(const float* __restrict__ x, float* __restrict__ out, const float* __restrict__ h, long long T,
                                                            long long new_t, int up, int down, long long pre, int n_taps) {
  const long long clip = blockIdx.y;
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= new_t) return;
  const long long pos = (j + pre) * down;
  long long i = pos / up;                 // newest input sample under the filter
  int k = int(pos - i * up);              // its tap
  if (i >= T) {                           // past the end of the clip: skip the taps that would meet zeros
    const long long skip = i - (T - 1);
    i -= skip;
    k += int(skip) * up;
  }
  const float* __restrict__ xc = x + clip * T;
  float acc = 0.0f;
  for (; k < n_taps && i >= 0; k += up, --i) acc = fmaf(__ldg(h + k), __ldg(xc + i), acc);
  out[clip * new_t + j] = acc;
}

int launch_resample_poly(const float* x, float* out, const float* h_dev, int64_t batch, int64_t T, int64_t new_t, int up, int down, int64_t pre, int n_taps,
                         void* stream, int* launches, std::string* err) {
  dim3 grid(unsigned((new_t + 255) / 256), unsigned(batch));
  resample_poly_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, h_dev, T, new_t, up, down, pre, n_taps);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("resample_poly_kernel launch: ") + cudaGetErrorString(e);
    return B2A_E_CUDA;
  }
  *launches += 1;
  return B2A_OK;
}

int launch_resample_linear(const float* x, float* out, int64_t batch, int64_t T, int64_t new_t, float step, float hi_clip, void* stream,
                           int* launches, std::string* err) {
  dim3 grid(unsigned((new_t + 255) / 256), unsigned(batch));
  resample_linear_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(x, out, T, new_t, step, hi_clip);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("resample_linear_kernel launch: ") + cudaGetErrorString(e);
    return B2A_E_CUDA;
  }
  *launches += 1;
  return B2A_OK;
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static int cuda_fail(cudaError_t e, const char* what, std::string* err) {
  if (err) *err = std::string(what) + ": " + cudaGetErrorString(e);
  return B2A_E_CUDA;
}

static void fill_tw(std::vector<float2>& v, int N, int N1, int N2) {
  const int H1 = N1 / 2;
  v.resize(size_t(N2) * (H1 - 1));
  for (int n2 = 0; n2 < N2; ++n2)
    for (int k1 = 1; k1 < H1; ++k1) {
      const double a = -2.0 * M_PI * double((long long)n2 * k1 % N) / double(N);
      v[size_t(n2) * (H1 - 1) + (k1 - 1)] = make_float2(float(cos(a)), float(sin(a)));
    }
}

int init_frontend_tables(std::string* err) {
  std::vector<float2> t;
  cudaError_t e;
  fill_tw(t, 400, 20, 20);
  if ((e = cudaMemcpyToSymbol(c_tw400, t.data(), t.size() * sizeof(float2))) != cudaSuccess) return cuda_fail(e, "twiddle upload", err);
  fill_tw(t, 512, 32, 16);
  if ((e = cudaMemcpyToSymbol(c_tw512, t.data(), t.size() * sizeof(float2))) != cudaSuccess) return cuda_fail(e, "twiddle upload", err);
  fill_tw(t, 1920, 60, 32);
  if ((e = cudaMemcpyToSymbol(c_tw1920, t.data(), t.size() * sizeof(float2))) != cudaSuccess) return cuda_fail(e, "twiddle upload", err);
  return init_wpf1920_tables(err);
}

bool frontend_plan_exists(int n_fft, int hop, int win_len) {
  return (n_fft == 400 && hop == 160 && win_len == 400) || (n_fft == 512 && hop == 160 && win_len == 400) ||
         (n_fft == 1920 && hop == 480 && win_len == 1920);
}

bool frontend_dyn_tiles();
static int sm_count() {   // SMs of the current device (cached per device)
  static std::atomic<int> cache[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return 148;
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    n = 148;
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
int frontend_tiles_per_clip(int n_fft, int64_t n_frames) {
  const PlanShape* ps = plan_shape(n_fft);
  const int ft = ps ? ps->frame_tile : 32;
  return int((n_frames + ft - 1) / ft);
}

template <class P, int PRE, int SPEC, int MEL = 0, int POST = POST_RUNTIME, int OUT = -1, bool RAGGED = false, bool F16 = false, bool ZS = false, bool DYN = false>
static int launch_plan(const FrontendArgs& a, cudaStream_t st, int* launches, std::string* err) {
  int walk_tpc = frontend_tiles_per_clip(P::N, a.n_frames);
  if constexpr (!ZS && B2A_SKIP_ZERO_TILES && B2A_RAW_TRACK && POST == POST_WNORM && MEL > 0 && PRE != PRE_KALDI) {
    // A zero tail (Whisper `padding`).  Equal lengths: the tiles of silence are the last tiles of every clip unless the right-edge
    // reflection reaches back into the content -- then the walk simply ends in front of them (walk_tpc; the same test as the kernel's
    // tile_zero).  Otherwise: the instantiation that tests every tile of its walk.
    if (a.zero_tail > 0) {
      const int tpc_all = walk_tpc;
      auto tile_zero = [&](long long tl) {
        const long long j0 = tl * (P::FT * P::HOP) - a.pad_left;
        if (j0 < a.n_samples) return false;
        const long long over = j0 + (P::TS - 1) - (a.n_samples + a.zero_tail);
        return over < 0 || a.pad_mode != PAD_REFLECT || over <= a.zero_tail - 2;
      };
      int z0 = 0;
      while (z0 < tpc_all && !tile_zero(z0)) ++z0;
      bool tail_only = !RAGGED && a.clip_tab == nullptr && z0 > 0 && frontend_dyn_tiles();
      for (int t = z0; t < tpc_all && tail_only; ++t) tail_only = tile_zero(t);
      if (!tail_only) return launch_plan<P, PRE, SPEC, MEL, POST, OUT, RAGGED, F16, true>(a, st, launches, err);
      walk_tpc = z0;
    }
  }
  if (!RAGGED && a.clip_tab != nullptr) {
    // per-clip lengths: the RAGGED instantiation of the run-time-configured kernel of the same plan (and of the Whisper
    // 128-mel kernel, dispatched in launch_frontend), reading the clip table
    if constexpr (MEL == 0 && POST == POST_RUNTIME && OUT == -1 && SPEC != SK_CPLX) return launch_plan<P, PRE, SPEC, 0, POST_RUNTIME, -1, true>(a, st, launches, err);
    if (err) *err = "ragged batches are built for the mel front ends only";
    return B2A_E_UNSUPPORTED;
  }
  if constexpr (!DYN && !RAGGED && !ZS && SPEC != SK_CPLX) {
    // equal-length mel launches of more than two rounds of persistent CTAs: the instantiation with the dynamic tile walk
    if (a.tile_ctr != nullptr && frontend_dyn_tiles() && (long long)walk_tpc * a.batch > 2LL * sm_count() * P::MINB)
      return launch_plan<P, PRE, SPEC, MEL, POST, OUT, false, F16, false, true>(a, st, launches, err);
  }
  // (by value on the stack: launches from different contexts / threads share nothing)
  FrontendParams<P> prm;
  prm.x = a.x;
  prm.clip_stride = a.n_samples;
  prm.n_samples = a.n_samples;
  prm.n_eff = a.n_samples + a.zero_tail;
  prm.pad_left = a.pad_left;
  prm.n_frames = a.n_frames;
  prm.pad_mode = a.pad_mode;
  prm.log_mode = a.log_mode;
  prm.whisper_norm = a.whisper_norm;
  prm.post_affine = a.post_affine;
  prm.post_sub = a.post_sub;
  prm.post_div = a.post_div;
  prm.out_mode = a.out_mode;
  prm.log_floor = a.log_floor;
  prm.lfr_m = a.lfr_m;
  prm.lfr_n = a.lfr_n;
  prm.lfr_rows = a.lfr_rows;
  prm.fb_desc = reinterpret_cast<const int4*>(a.bank.desc);
  prm.fb_w = a.bank.weights;
  prm.fb_steps = reinterpret_cast<const float4*>(a.bank.steps);
  prm.n_mels = a.bank.n_mels;
  prm.n_steps = a.bank.n_steps;
  for (int w = 0; w <= P::NCHUNK; ++w) prm.chunk_m[w] = prm.chunk_s[w] = 0;
  if (SPEC != SK_CPLX) {
    const int M = a.bank.n_mels;
    if (a.bank.steps != nullptr) {
      if (a.bank.n_chunks != P::NCHUNK || a.bank.frame_tile != P::FT) {
        if (err) *err = "mel program was compiled for another CTA shape";
        return B2A_E_BAD_ARG;
      }
      for (int w = 0; w <= P::NCHUNK; ++w) {
        prm.chunk_m[w] = a.bank.host_chunk_m[w];
        prm.chunk_s[w] = a.bank.host_chunk_s[w];
      }
    } else {
      for (int w = 0; w <= P::NCHUNK; ++w) prm.chunk_m[w] = int((long long)M * w / P::NCHUNK);
    }
  }
  prm.out = a.out;
  prm.clip_max = a.clip_max;
  prm.tile_min = reinterpret_cast<int*>(a.tile_min);
  prm.tile_max = nullptr;
  prm.tiles_per_clip = frontend_tiles_per_clip(P::N, a.n_frames);
  prm.walk_tpc = RAGGED ? prm.tiles_per_clip : walk_tpc;
  prm.n_clips = int(a.batch);
  prm.clip_tab = static_cast<const int4*>(a.clip_tab);
  prm.tile_tab = static_cast<const int4*>(a.tile_tab);
  switch (a.out_mode) {
    case OUT_TM: prm.out_clip_stride = a.n_frames * (long long)a.bank.n_mels; break;
    case OUT_MT: prm.out_clip_stride = a.n_frames * (long long)a.bank.n_mels; break;
    case OUT_LFR: prm.out_clip_stride = a.lfr_rows * (long long)a.lfr_m * a.bank.n_mels; break;
    default: prm.out_clip_stride = a.n_frames * (long long)P::NBINS * 2; break;
  }
  for (int o = 0; o < P::WIN; ++o) prm.window[o] = a.window[o];
  if (SPEC != SK_CPLX) {
    // the finished mel values are staged in the rows of the exchange buffer that the spectrum tile leaves free
    if (a.bank.n_mels <= 0 || a.bank.n_mels > (((P::N2 - 1) * P::FT - (P::FT - 1)) / (P::FT + 1)) * P::H1) {
      if (err) *err = "n_mels out of range for this plan";
      return B2A_E_BAD_ARG;
    }
  }
  const size_t smem = sizeof(float) * size_t((SPEC == SK_CPLX ? P::R0_WORDS_CPLX : P::R0_WORDS_REAL) + P::Y_WORDS +
                                             P::N + P::TW_WORDS) +
                      ((MEL == 0 && SPEC != SK_CPLX) ? sizeof(float4) * size_t(steps_smem_max<P>()) : 0) +
                      (mt_table<P, MEL, OUT>() ? sizeof(int) * size_t(P::NWARPS * mt_slots<P, MEL>()) : 0);
  static_assert((P::R0_WORDS_REAL % 4) == 0 && (P::R0_WORDS_CPLX % 4) == 0 && (P::Y_WORDS % 4) == 0 && (P::N % 4) == 0 && (P::TW_WORDS % 4) == 0,
                "shared-memory tables must stay 16-byte (window rows) / 8-byte (twiddles) aligned");
  prm.total_tiles = RAGGED ? (long long)a.total_tiles : (long long)prm.walk_tpc * a.batch;   // tiles of the walk
  const long long table_tiles = RAGGED ? (long long)a.total_tiles : (long long)prm.tiles_per_clip * a.batch;   // entries of tile_min
  if (prm.total_tiles <= 0 || prm.total_tiles > 0x7fffffffLL || a.batch > 0x7fffffffLL || a.n_frames > 0x7fffffffLL) {
    if (err) *err = "empty launch";
    return B2A_E_BAD_ARG;
  }
  // shared-memory opt-in and residency of this instantiation: queried once per device, then cached (a single 30 s clip is a
  // ~12 us kernel; two runtime queries per launch under a process-wide lock used to cost more than that on the host)
  struct DevInfo { std::atomic<int> ready{0}; int n_sm = 0, per_sm = 0; };
  static DevInfo infos[64];
  static std::mutex info_mu;
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) {
    if (err) *err = "device index out of range";
    return B2A_E_CUDA;
  }
  DevInfo& di = infos[dev];
  cudaError_t e;
  if (!di.ready.load(std::memory_order_acquire)) {
    std::lock_guard<std::mutex> lk(info_mu);
    if (!di.ready.load(std::memory_order_relaxed)) {
      int n_sm = 148, per_sm = 1;
      if ((e = cudaFuncSetAttribute(frontend_kernel<P, PRE, SPEC, MEL, POST, OUT, RAGGED, F16, ZS, DYN>, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) != cudaSuccess)
        return cuda_fail(e, "cudaFuncSetAttribute", err);
      cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
      if ((e = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, frontend_kernel<P, PRE, SPEC, MEL, POST, OUT, RAGGED, F16, ZS, DYN>, P::NTHREADS, smem)) != cudaSuccess)
        return cuda_fail(e, "occupancy query", err);
      if (per_sm < 1) {
        if (err) *err = "frontend kernel does not fit on this device";
        return B2A_E_CUDA;
      }
      di.n_sm = n_sm;
      di.per_sm = std::min(per_sm, P::MINB);
      di.ready.store(1, std::memory_order_release);
    }
  }
  const int n_sm = di.n_sm, per_sm = di.per_sm;
  const long long nblocks = std::min<long long>(prm.total_tiles, (long long)n_sm * per_sm);  // persistent CTAs
  // dynamic tile walk for launches of more than two rounds; its counter sits behind the clamp tables when there are any (their memset
  // initialises it too: 0x80808080), otherwise it gets a 4-byte memset of its own
  prm.tile_ctr = nullptr;
  prm.tile_ctr_init = 0;
  // per-tile maxima (kernels that track the raw sums): behind the tile minima, same initial pattern
  const bool contiguous = a.whisper_norm && reinterpret_cast<int*>(a.tile_min) == a.clip_max + a.batch;
  if (B2A_RAW_TRACK && POST == POST_WNORM && contiguous && a.tile_max == reinterpret_cast<int*>(a.tile_min) + table_tiles) prm.tile_max = a.tile_max;
  const long long table_words = a.batch + table_tiles * (prm.tile_max != nullptr ? 2 : 1);   // clip maxima, tile minima (, tile maxima)
  if (DYN || (RAGGED && !ZS && a.tile_ctr != nullptr && frontend_dyn_tiles() && prm.total_tiles > 3 * nblocks)) {
    prm.tile_ctr = a.tile_ctr;
    if (contiguous && a.tile_ctr == a.clip_max + table_words) {
      prm.tile_ctr_init = int(0x80808080u);
      if ((e = cudaMemsetAsync(a.clip_max, 0x80, sizeof(int) * size_t(table_words + 1), st)) != cudaSuccess) return cuda_fail(e, "memset", err);
    } else {
      if ((e = cudaMemsetAsync(a.tile_ctr, 0, sizeof(int), st)) != cudaSuccess) return cuda_fail(e, "memset", err);
    }
  }
  if (a.whisper_norm && prm.tile_ctr_init == 0) {
    // clip_max and the (negated) tile minima start from the same "very negative" pattern: one memset when the C ABI placed them
    // back to back
    if (reinterpret_cast<int*>(a.tile_min) == a.clip_max + a.batch) {
      if ((e = cudaMemsetAsync(a.clip_max, 0x80, sizeof(int) * size_t(table_words), st)) != cudaSuccess) return cuda_fail(e, "memset", err);
    } else {
      if ((e = cudaMemsetAsync(a.clip_max, 0x80, sizeof(int) * size_t(a.batch), st)) != cudaSuccess) return cuda_fail(e, "memset", err);
      if ((e = cudaMemsetAsync(a.tile_min, 0x80, sizeof(int) * size_t(table_tiles), st)) != cudaSuccess) return cuda_fail(e, "memset", err);
    }
  }
  frontend_kernel<P, PRE, SPEC, MEL, POST, OUT, RAGGED, F16, ZS, DYN><<<unsigned(nblocks), P::NTHREADS, smem, st>>>(prm);
  if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "frontend_kernel launch", err);
  *launches += 1;
  if (a.whisper_norm) {
    for (long long c0 = 0; c0 < a.batch; c0 += 65535) {   // gridDim.y limit
      const long long nb = std::min<long long>(65535, a.batch - c0);
      cudaLaunchConfig_t cfg = {};
      const int group = clamp_group(table_tiles);
      cfg.gridDim = dim3(unsigned((prm.tiles_per_clip + group - 1) / group), unsigned(nb));
      cfg.blockDim = dim3(256);
      cfg.dynamicSmemBytes = 0;
      cfg.stream = st;
      cudaLaunchAttribute attr[1];
      attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
      attr[0].val.programmaticStreamSerializationAllowed = 1;
      cfg.attrs = attr;
      cfg.numAttrs = 1;
      float* c_out = F16 ? reinterpret_cast<float*>(reinterpret_cast<__half*>(a.out) + c0 * prm.out_clip_stride) : a.out + c0 * prm.out_clip_stride;
      const float clamp_floor = (B2A_RAW_TRACK && POST == POST_WNORM) ? a.log_floor : -1.0f;   // > 0: the clamp kernel applies the log floor
      const int* c_max = a.clip_max + c0;
      const int* c_min = RAGGED ? prm.tile_min : prm.tile_min + c0 * prm.tiles_per_clip;
      const int4* c_tab = RAGGED ? prm.clip_tab + c0 : nullptr;
      const int* c_tmax = prm.tile_max == nullptr ? nullptr : (RAGGED ? prm.tile_max : prm.tile_max + c0 * prm.tiles_per_clip);
      if ((e = cudaLaunchKernelEx(&cfg, whisper_clamp_kernel, c_out, c_max, c_min, prm.tiles_per_clip, (long long)a.n_frames, a.bank.n_mels,
                                  prm.out_clip_stride, a.out_mode, int(P::FT), c_tab, F16 ? 1 : 0, clamp_floor, group, c_tmax)) != cudaSuccess)
        return cuda_fail(e, "whisper_clamp_kernel launch", err);
    }
    if ((e = cudaGetLastError()) != cudaSuccess) return cuda_fail(e, "whisper_clamp_kernel launch", err);
    *launches += 1;
  }
  return B2A_OK;
}

int launch_whisper_clamp(float* out, const int* clip_max, const int* tile_min, int64_t batch, int64_t n_frames, int n_mels, void* stream,
                         int* launches, std::string* err) {
  const int tiles = frontend_tiles_per_clip(400, n_frames);
  const long long stride = n_frames * (long long)n_mels;
  for (long long c0 = 0; c0 < batch; c0 += 65535) {   // gridDim.y limit
    const long long nb = std::min<long long>(65535, batch - c0);
    const int group = clamp_group((long long)tiles * batch);
    whisper_clamp_kernel<<<dim3(unsigned((tiles + group - 1) / group), unsigned(nb)), 256, 0, static_cast<cudaStream_t>(stream)>>>(
        out + c0 * stride, clip_max + c0, tile_min + c0 * tiles, tiles, n_frames, n_mels, stride, OUT_TM, 32, nullptr, 0, -1.0f, group, nullptr);
  }
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "whisper_clamp_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

// Id of the baked bank (mel_baked.h) whose step program equals this one word for word, or 0.
int frontend_match_baked(const float* steps, int n_steps, const int* chunk_m, const int* chunk_s, int n_chunks, int frame_tile, int n_mels) {
  for (int i = 0; i < kMelBakedCount; ++i) {
    const MelBakedInfo& b = mel_baked_infos[i];
    if (b.n_steps != n_steps || b.n_chunks != n_chunks || b.frame_tile != frame_tile || b.n_mels != n_mels) continue;
    if (memcmp(b.steps, steps, sizeof(float) * 4 * size_t(n_steps)) != 0) continue;
    if (memcmp(b.chunk_m, chunk_m, sizeof(int) * size_t(n_chunks + 1)) != 0) continue;
    if (memcmp(b.chunk_s, chunk_s, sizeof(int) * size_t(n_chunks + 1)) != 0) continue;
    return b.id;
  }
  return 0;
}

// B2A_DYN_TILES=0 in the environment or frontend_dyn_tiles_enable(0): static tile walk everywhere (A/B switch)
static int g_dyn_tiles = -1;
void frontend_dyn_tiles_enable(int on) { g_dyn_tiles = on ? 1 : 0; }
bool frontend_dyn_tiles() {
  if (g_dyn_tiles < 0) {
    const char* v = getenv("B2A_DYN_TILES");
    g_dyn_tiles = (v != nullptr && v[0] == '0') ? 0 : 1;
  }
  return g_dyn_tiles == 1;
}

int launch_frontend(const FrontendArgs& a, void* stream, int* launches, std::string* err) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int spec = a.out_mode == OUT_COMPLEX ? SK_CPLX : (a.spec_mode == SPEC_POWER ? SK_POWER : SK_MAG);
  // Known banks (mel_baked.h): the bank's step program matched one of them word for word.  Their kernels are specialised
  // at compile time on the post-processing and the output layout (small code: the whole tile loop stays inside the
  // instruction cache); everything else runs the run-time-configured kernel.
  int post = -1;
  const bool ragged = a.clip_tab != nullptr;   // per-clip lengths: RAGGED instantiations
  if (a.bank.baked_id > 0 && spec == SK_POWER && !a.post_affine && a.out_mode != OUT_COMPLEX) {
    if (a.whisper_norm && a.log_mode == LOG_LOG10 && a.out_mode != OUT_LFR) post = POST_WNORM;
    else if (!a.whisper_norm && a.log_mode == LOG_LN) post = POST_LN;
    else if (!a.whisper_norm && a.log_mode == LOG_NONE) post = POST_NONE;
  }
  const int id = a.bank.baked_id, om = a.out_mode;
  // B2A_WHISPER_TC=1: the tensor-core front end (tc_frontend.cu) instead of the FFT kernel, where it applies (A/B switch)
  if (tc_whisper_enabled() && tc_whisper_applicable(a)) return launch_tc_whisper(a, stream, launches, err);
  if (a.out_f16) {
    // fp16 features straight from the store loop: the Whisper front end's two standard banks, (T', M) layout
    if (a.n_fft == 400 && a.hop == 160 && a.win_len == 400 && a.pre_mode == PRE_NONE && post == POST_WNORM && om == OUT_TM) {
      if (id == 1) return ragged ? launch_plan<Plan400, PRE_NONE, SK_POWER, 1, POST_WNORM, OUT_TM, true, true>(a, st, launches, err)
                                 : launch_plan<Plan400, PRE_NONE, SK_POWER, 1, POST_WNORM, OUT_TM, false, true>(a, st, launches, err);
      if (id == 2 && !ragged) return launch_plan<Plan400, PRE_NONE, SK_POWER, 2, POST_WNORM, OUT_TM, false, true>(a, st, launches, err);
    }
    if (err) *err = "fp16 output is built for the Whisper log-mel front end with its standard 80 / 128-mel banks";
    return B2A_E_UNSUPPORTED;
  }
  if (ragged) {
    // the headline front ends keep their tuned kernels (Whisper 128 (T', M) and (M, T'), Fun-ASR LFR, CAM++ fbank); every other
    // ragged call runs the run-time-configured kernel of its plan (launch_plan forwards to the RAGGED instantiation)
    if (a.n_fft == 400 && a.hop == 160 && a.win_len == 400 && a.pre_mode == PRE_NONE) {
      if (post == POST_WNORM && id == 1 && om == OUT_TM) return launch_plan<Plan400, PRE_NONE, SK_POWER, 1, POST_WNORM, OUT_TM, true>(a, st, launches, err);
      if (post == POST_WNORM && id == 1 && om == OUT_MT) return launch_plan<Plan400, PRE_NONE, SK_POWER, 1, POST_WNORM, OUT_MT, true>(a, st, launches, err);
      if (post == POST_LN && id == 3 && om == OUT_LFR) return launch_plan<Plan400, PRE_NONE, SK_POWER, 3, POST_LN, OUT_LFR, true>(a, st, launches, err);
    }
    if (a.n_fft == 512 && a.hop == 160 && a.win_len == 400 && a.pre_mode == PRE_KALDI && post == POST_LN && id == 4 && om == OUT_TM)
      return launch_plan<Plan512, PRE_KALDI, SK_POWER, 4, POST_LN, OUT_TM, true>(a, st, launches, err);
    post = -1;
  }
  if (a.n_fft == 400 && a.hop == 160 && a.win_len == 400 && a.pre_mode == PRE_NONE) {
    if (post == POST_WNORM && id == 1 && om == OUT_TM) return launch_plan<Plan400, PRE_NONE, SK_POWER, 1, POST_WNORM, OUT_TM>(a, st, launches, err);
    if (post == POST_WNORM && id == 1 && om == OUT_MT) return launch_plan<Plan400, PRE_NONE, SK_POWER, 1, POST_WNORM, OUT_MT>(a, st, launches, err);
    if (post == POST_WNORM && id == 2 && om == OUT_TM) return launch_plan<Plan400, PRE_NONE, SK_POWER, 2, POST_WNORM, OUT_TM>(a, st, launches, err);
    if (post == POST_LN && id == 3 && om == OUT_LFR) return launch_plan<Plan400, PRE_NONE, SK_POWER, 3, POST_LN, OUT_LFR>(a, st, launches, err);
    if (post == POST_LN && id == 3 && om == OUT_TM) return launch_plan<Plan400, PRE_NONE, SK_POWER, 3, POST_LN, OUT_TM>(a, st, launches, err);
    // Chatterbox voice encoder (VoiceEncoderMelspec.swift:17-68, defaults): 40-mel power mel, no log, (M, T')
    if (post == POST_NONE && id == 5 && om == OUT_MT) return launch_plan<Plan400, PRE_NONE, SK_POWER, 5, POST_NONE, OUT_MT>(a, st, launches, err);
    if (spec == SK_POWER) return launch_plan<Plan400, PRE_NONE, SK_POWER>(a, st, launches, err);
    if (spec == SK_MAG) return launch_plan<Plan400, PRE_NONE, SK_MAG>(a, st, launches, err);
    return launch_plan<Plan400, PRE_NONE, SK_CPLX>(a, st, launches, err);
  }
  if (a.n_fft == 512 && a.hop == 160 && a.win_len == 400) {
    if (a.pre_mode == PRE_KALDI && post == POST_LN && id == 4 && om == OUT_TM)
      return launch_plan<Plan512, PRE_KALDI, SK_POWER, 4, POST_LN, OUT_TM>(a, st, launches, err);
    if (a.pre_mode == PRE_KALDI && spec == SK_POWER) return launch_plan<Plan512, PRE_KALDI, SK_POWER>(a, st, launches, err);
    if (a.pre_mode == PRE_NONE && spec == SK_CPLX) return launch_plan<Plan512, PRE_NONE, SK_CPLX>(a, st, launches, err);
  }
  if (a.n_fft == 1920 && a.hop == 480 && a.win_len == 1920 && a.pre_mode == PRE_NONE) {
    // mel batches (equal lengths or ragged): the warp-per-frame kernel (wpf1920.cu); B2A_WPF1920=0 / b2a_debug_wpf1920(0) keep the tiled kernel
    if (spec != SK_CPLX && wpf1920_applicable(a)) return launch_wpf1920(a, stream, launches, err);
    // S3Gen 24 kHz mel (S3GenMel.swift:43-102): magnitude spectrum, ln, (M, T') -- compile-time post-processing keeps the store code short
    if (!ragged && spec == SK_MAG && a.bank.steps != nullptr && a.log_mode == LOG_LN && !a.whisper_norm && !a.post_affine && a.out_mode == OUT_MT)
      return launch_plan<Plan1920, PRE_NONE, SK_MAG, 0, POST_LN, OUT_MT>(a, st, launches, err);
    if (spec == SK_POWER) return launch_plan<Plan1920, PRE_NONE, SK_POWER>(a, st, launches, err);
    if (spec == SK_MAG) return launch_plan<Plan1920, PRE_NONE, SK_MAG>(a, st, launches, err);
    return launch_plan<Plan1920, PRE_NONE, SK_CPLX>(a, st, launches, err);
  }
  if (err) *err = "no FFT plan built for this (n_fft, hop, win_length, mode)";
  return B2A_E_UNSUPPORTED;
}

static bool colstat_vec_ok(const float* in, const float* out, int dim) {
  return dim % 4 == 0 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15) == 0;
}

// Measured on B200 (512 clips): the scalar kernel wins while the slabs of all resident blocks (8 blocks per SM x 32 columns x rows
// x 4 B) fit the L2, because its second and third pass then hit the cache (Fun-ASR CMVN, 334 x 560: 0.20 ms vs 0.23 ms); beyond
// that the 16-byte version's larger number of bytes in flight wins (CAM++ mean-norm, 1998 x 80: 0.18 ms vs 0.20 ms).  Limiting
// the residency of the 16-byte version to make its slabs fit was slower than either.
static bool colstat_use_vec(const float* in, const float* out, int64_t rows, int dim) {
  if (!colstat_vec_ok(in, out, dim)) return false;
  int dev = 0, n_sm = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
  return double(n_sm) * 8.0 * 32.0 * double(rows) * 4.0 > 96.0e6;
}

static int launch_colstat4(const float* in, float* out, int64_t batch, int64_t rows, int dim, int do_var, const void* clip_tab, int tab_rows,
                           cudaStream_t st, int* launches, std::string* err) {
  const int dim4 = dim / 4;
  dim3 grid(unsigned((dim4 + 31) / 32), unsigned(batch));
  colstat4_kernel<8><<<grid, 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out), rows, dim4, do_var,
                                           static_cast<const int4*>(clip_tab), tab_rows);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "colstat4_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

// Register-resident statistics kernel for clips of at most 512 rows (-> true when launched)
#ifndef B2A_COLSTAT_REG
#define B2A_COLSTAT_REG 1
#endif
template <int RPT>
static void launch_colstat_reg_t(const float* in, float* out, int64_t batch, int64_t rows, int dim, int do_var, const void* clip_tab,
                                 int tab_rows, cudaStream_t st) {
  dim3 grid(unsigned((dim + 31) / 32), unsigned(batch));
  colstat_reg_kernel<RPT><<<grid, 256, 0, st>>>(in, out, rows, dim, do_var, static_cast<const int4*>(clip_tab), tab_rows);
}
static bool launch_colstat_reg(const float* in, float* out, int64_t batch, int64_t rows, int dim, int do_var, const void* clip_tab, int tab_rows,
                               cudaStream_t st, int* launches, std::string* err, int* rc) {
  if (!B2A_COLSTAT_REG || rows > 512 || dim > (1 << 20)) return false;
  if (rows <= 128) launch_colstat_reg_t<16>(in, out, batch, rows, dim, do_var, clip_tab, tab_rows, st);
  else if (rows <= 256) launch_colstat_reg_t<32>(in, out, batch, rows, dim, do_var, clip_tab, tab_rows, st);
  else if (rows <= 384) launch_colstat_reg_t<48>(in, out, batch, rows, dim, do_var, clip_tab, tab_rows, st);
  else launch_colstat_reg_t<64>(in, out, batch, rows, dim, do_var, clip_tab, tab_rows, st);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    *rc = cuda_fail(e, "colstat_reg_kernel launch", err);
    return true;
  }
  *launches += 1;
  *rc = B2A_OK;
  return true;
}

// Shared-memory-resident statistics kernel: 8-column slabs of at most 96 KB (-> true when launched)
#ifndef B2A_COLSTAT_SMEM
#define B2A_COLSTAT_SMEM 1
#endif
static bool launch_colstat_smem(const float* in, float* out, int64_t batch, int64_t rows, int dim, int do_var, const void* clip_tab, int tab_rows,
                                cudaStream_t st, int* launches, std::string* err, int* rc) {
  constexpr int CW = 8;
  const size_t bytes = size_t(rows) * CW * sizeof(float);
  if (!B2A_COLSTAT_SMEM || bytes > 96 * 1024 || !colstat_vec_ok(in, out, dim)) return false;
  cudaError_t e = cudaFuncSetAttribute(colstat_smem_kernel<CW>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);   // (per device: every launch)
  if (e != cudaSuccess) {
    *rc = cuda_fail(e, "cudaFuncSetAttribute", err);
    return true;
  }
  dim3 grid(unsigned((dim + CW - 1) / CW), unsigned(batch));
  colstat_smem_kernel<CW><<<grid, 256, bytes, st>>>(in, out, rows, dim, do_var, static_cast<const int4*>(clip_tab), tab_rows);
  e = cudaGetLastError();
  if (e != cudaSuccess) {
    *rc = cuda_fail(e, "colstat_smem_kernel launch", err);
    return true;
  }
  *launches += 1;
  *rc = B2A_OK;
  return true;
}

int launch_cmvn(const float* in, float* out, int64_t batch, int64_t rows, int dim, const float* mean, const float* istd,
                void* stream, int* launches, std::string* err, const void* clip_tab) {
  int rc = B2A_OK;
  if (mean == nullptr && launch_colstat_reg(in, out, batch, rows, dim, 1, clip_tab, 2, static_cast<cudaStream_t>(stream), launches, err, &rc))
    return rc;
  if (mean == nullptr && launch_colstat_smem(in, out, batch, rows, dim, 1, clip_tab, 2, static_cast<cudaStream_t>(stream), launches, err, &rc))
    return rc;
  if (mean == nullptr && colstat_use_vec(in, out, rows, dim))
    return launch_colstat4(in, out, batch, rows, dim, 1, clip_tab, 2, static_cast<cudaStream_t>(stream), launches, err);
  dim3 grid(unsigned((dim + 31) / 32), unsigned(batch));
  colstat_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, rows, dim, mean, istd, 1, static_cast<const int4*>(clip_tab), 2);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "colstat_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

int launch_mean_norm(float* inout, int64_t batch, int64_t rows, int dim, void* stream, int* launches, std::string* err, const void* clip_tab) {
  int rc = B2A_OK;
  if (launch_colstat_reg(inout, inout, batch, rows, dim, 0, clip_tab, 1, static_cast<cudaStream_t>(stream), launches, err, &rc)) return rc;
  if (launch_colstat_smem(inout, inout, batch, rows, dim, 0, clip_tab, 1, static_cast<cudaStream_t>(stream), launches, err, &rc)) return rc;
  if (colstat_use_vec(inout, inout, rows, dim))
    return launch_colstat4(inout, inout, batch, rows, dim, 0, clip_tab, 1, static_cast<cudaStream_t>(stream), launches, err);
  dim3 grid(unsigned((dim + 31) / 32), unsigned(batch));
  colstat_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(inout, inout, rows, dim, nullptr, nullptr, 0, static_cast<const int4*>(clip_tab), 1);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "colstat_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

int launch_lfr(const float* in, float* out, int64_t batch, int64_t n_frames, int n_mels, int lfr_m, int lfr_n, void* stream,
               int* launches, std::string* err) {
  const long long rows = (n_frames + lfr_n - 1) / lfr_n;
  const long long segs = rows * lfr_m;
  dim3 grid(unsigned((segs + 7) / 8), unsigned(batch));
  lfr_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n_frames, n_mels, lfr_m, lfr_n, rows);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "lfr_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

int launch_pad_or_trim(const float* in, float* out, int64_t batch, int64_t n, int64_t length, void* stream, int* launches,
                       std::string* err) {
  dim3 grid(unsigned((length + 255) / 256), unsigned(batch));
  pad_or_trim_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n, length);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "pad_or_trim_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

// Ragged batches: rows past a clip's own count are zero.  out is (batch, rows_max, row_len) ("time-major": whole rows) or
// (batch, n_rows, t_max) with the valid part the first `count` entries of every row ("mel-major").
__global__ void __launch_bounds__(256) zero_tail_kernel(float* __restrict__ out, const int4* __restrict__ clip_tab, int which, long long rows_max,
                                                        long long row_len, int mel_major) {
  const long long clip = blockIdx.y;
  const int4 ci = clip_tab[clip];
  const long long count = which == 2 ? ci.z : ci.y;
  float* o = out + clip * rows_max * row_len;
  const long long i0 = (long long)blockIdx.x * blockDim.x + threadIdx.x, stride = (long long)gridDim.x * blockDim.x;
  if (!mel_major) {
    for (long long i = count * row_len + i0; i < rows_max * row_len; i += stride) o[i] = 0.0f;
  } else {   // rows_max = number of mel rows, row_len = t_max
    const long long tail = row_len - count;
    for (long long i = i0; i < rows_max * tail; i += stride) o[(i / tail) * row_len + count + i % tail] = 0.0f;
  }
}

// tile_tab[g] = (clip, tile) for a ragged launch: the clip of tile g is the last one whose first tile is <= g (bisection over the
// first-tile column of clip_tab; every clip has at least one tile).  Built once per call on the device -- the main kernel would
// otherwise pay the ~10 dependent loads per tile itself, mostly as L2 hits (its cp.async traffic sweeps the L1).
__global__ void tile_table_kernel(const int4* __restrict__ clip_tab, int n_clips, int total_tiles, int4* __restrict__ tile_tab) {
  const int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= total_tiles) return;
  int lo = 0, hi = n_clips - 1;
  while (lo < hi) {
    const int mid = (lo + hi + 1) >> 1;
    if (clip_tab[mid].w <= g) lo = mid;
    else hi = mid - 1;
  }
  const int4 ci = clip_tab[lo];
  tile_tab[g] = make_int4(lo, g - ci.w, ci.x, ci.y);
}

int launch_tile_table(const void* clip_tab, int64_t n_clips, int64_t total_tiles, void* tile_tab, void* stream, int* launches, std::string* err) {
  tile_table_kernel<<<unsigned((total_tiles + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const int4*>(clip_tab), int(n_clips), int(total_tiles), static_cast<int4*>(tile_tab));
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "tile_table_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

// max_tail_rows: the largest number of rows any clip of the batch falls short of rows_max (mel-major: of t_max columns); the grid is
// sized for that tail and the launch is skipped when no clip has one (equal lengths through a ragged entry point)
int launch_zero_tails(float* out, const void* clip_tab, int which, int64_t batch, int64_t rows_max, int64_t row_len, int mel_major, int64_t max_tail,
                      void* stream, int* launches, std::string* err) {
  if (max_tail <= 0) return B2A_OK;
  const long long per_clip = mel_major ? rows_max * max_tail : max_tail * row_len;
  dim3 grid(unsigned(std::max<long long>(1, std::min<long long>(64, (per_clip + 4095) / 4096))), unsigned(batch));
  zero_tail_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(out, static_cast<const int4*>(clip_tab), which, rows_max, row_len, mel_major);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "zero_tail_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

int launch_reflect_pad(const float* in, float* out, int64_t batch, int64_t n, int64_t pad, void* stream, int* launches, std::string* err) {
  dim3 grid(unsigned((n + 2 * pad + 255) / 256), unsigned(batch));
  reflect_pad_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(in, out, n, pad);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) return cuda_fail(e, "reflect_pad_kernel launch", err);
  *launches += 1;
  return B2A_OK;
}

}  // namespace b2a
