// Internal declarations shared by the host tables, the CUDA kernels and the C ABI.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace b2a {

// ---- host tables (host_tables.cpp) ---------------------------------------------------------
struct SparseBank {
  int n_mels = 0, n_bins = 0, max_bin = -1;
  std::vector<int> start, count, offset;  // per filter: first bin, number of bins, offset into weights
  std::vector<float> weights;
  bool two_adjacent = false;       // every bin feeds at most two filters, and they are adjacent (m, m+1), m non-decreasing
  std::vector<int> bin_mlo;        // then: m_lo per bin
  std::vector<float> steps;        // mel step program for the frontend kernel (see build_mel_program), 4 floats per step
  std::vector<int> chunk_m, chunk_s;  // per warp: first filter / first step (n_chunks + 1 entries)
};
int make_window(int kind, int length, float* out);
void hann_periodic_via_hanning(int n, std::vector<float>& w);
int mel_filters_slaney(int sample_rate, int n_fft, int n_mels, float f_min, float f_max, float* out);
int mel_filters_funasr(int sample_rate, int n_fft, int n_mels, float* out);
int mel_filters_htk_int(int sample_rate, int n_fft, int n_mels, float f_min, float f_max, float* out);
int64_t reflect_pad_index(int64_t i, int64_t n, int64_t pad);
void build_sparse_bank(const float* bank, int n_mels, int n_bins, bool bin_major, SparseBank& sb);
// Shape of the frontend kernel's FFT plans (frontend.cu static_asserts that its Plan types agree): n_fft = n1 * n2, CTA tile
// of `frame_tile` frames, `n_chunks` mel chunks (one per warp, or per half-warp for 16-frame tiles).
struct PlanShape { int n_fft, n1, n2, frame_tile, n_warps, n_chunks; };
constexpr PlanShape kPlanShapes[3] = {{400, 20, 20, 32, 10, 10}, {512, 32, 16, 32, 16, 16}, {1920, 60, 32, 16, 16, 32}};
inline const PlanShape* plan_shape(int n_fft) {
  for (const PlanShape& p : kPlanShapes)
    if (p.n_fft == n_fft) return &p;
  return nullptr;
}
// Stage B leaves the spectrum tile in the exchange buffer, item by item: slot_of_bin[k] is the row (of frame_tile floats)
// that holds bin k.  Item k1 = k mod n1 (or n1 - k mod n1 for the conjugate-mirrored bins) owns rows [k1*2*n2, (k1+1)*2*n2).
void spectrum_slots(const PlanShape& ps, std::vector<int>& slot_of_bin);
// bin_slot: nullptr = identity (bin k in row k)
// emit_word[m]: what the step that finishes filter m carries (see output_words); false if n_mels does not fit the free rows
bool output_words(const PlanShape& ps, int n_mels, std::vector<int>& emit_word);
int output_block_capacity(const PlanShape& ps);
void build_mel_program(const float* bank, int n_mels, int n_bins, bool bin_major, int frame_tile, const int* emit_word, int n_chunks,
                       const int* bin_slot, SparseBank& sb);
// Mel schedule of the warp-per-frame front end (wpf1920.cu: one warp = one frame, the spectrum of a frame lies bin-major in the warp's
// shared memory).  The filters are cut into at most kWpfRounds * 32 segments of contiguous bins (the widest ones in halves / quarters), sorted
// by length and dealt round by round to the 32 lanes: in round r every lane accumulates len4[r] groups of FOUR products (two 16-byte
// loads: weights, bins) over its own segment, read from a 16-byte aligned start bin (weights zero outside the segment); the starts and
// the lane of every segment are chosen so that the eight 16-byte spectrum reads of a quarter-warp fall into different banks.  Then
// filter m is the sum of its <= 4 segment sums in ascending bin order.  32-bit words:
//   [0] n_mels  [1] rounds  [2..5] len4[r]  [6..9] word offset of round r's weights (float4 [group][lane])  [10] word offset of the
//   start table ([r][lane]: first bin read, a multiple of 4)  [11] word offset of the segment table ([m][4]: partial-sum slots
//   r * 32 + lane, or rounds * 32 = the zero slot)  [12] total words  [13] segments left with a two-way bank conflict.
// False when the bank does not fit (n_mels > kWpfRounds * 32, more than kWpfMelMaxWords words).
constexpr int kWpfRounds = 3, kWpfMelMaxWords = 3584, kWpfMelHeader = 16;
bool build_wpf_mel(const SparseBank& sb, int n_bins_spectrum, std::vector<uint32_t>& blob);

// ---- fused STFT -> (power | magnitude) -> mel -> log front-end (frontend.cu) -----------------
enum PadMode { PAD_NONE = 0, PAD_REFLECT = 1, PAD_ZERO = 2 };
enum PreMode { PRE_NONE = 0, PRE_KALDI = 1 };            // per-frame DC removal + 0.97 pre-emphasis
enum SpecMode { SPEC_POWER = 0, SPEC_MAGNITUDE = 1 };     // |X|^2 or |X|
enum LogMode { LOG_NONE = 0, LOG_LOG10 = 1, LOG_LN = 2, LOG_DB20 = 3 };
enum OutMode {
  OUT_TM = 0,       // (batch, T', M)            Whisper, Fun-ASR log-mel, Kaldi fbank
  OUT_MT = 1,       // (batch, M, T')            S3Tokenizer/Chatterbox 128-mel, S3Gen 80-mel, voice encoder
  OUT_LFR = 2,      // (batch, ceil(T'/n), m*M)  Fun-ASR preprocessAudio (LFR stacking fused into the store)
  OUT_COMPLEX = 3   // (batch, T', F) complex64  plain stft()
};

struct DeviceBank {  // sparse filterbank in device memory
  const int* desc = nullptr;      // 4 ints per filter: first bin, number of bins, offset into weights, 0
  const float* weights = nullptr;
  const float* steps = nullptr;   // mel step program, 4 floats per step (w_lo, w_hi, bits(bin*FT*4), bits(emit stride)); null = generic path
  int n_steps = 0;
  int n_chunks = 0, frame_tile = 0;  // CTA shape the program was compiled for
  const int* host_chunk_m = nullptr;  // HOST: per-warp first filter / first step of the program (n_chunks + 1 entries)
  const int* host_chunk_s = nullptr;
  int n_mels = 0;
  int n_bins_used = 0;  // bins [0, n_bins_used) are read by the mel stage
  int baked_id = 0;     // > 0: the step program equals baked bank `baked_id` of mel_baked.h (straight-line kernel available)
  const uint32_t* wpf_mel = nullptr;   // n_fft 1920: mel schedule of the warp-per-frame kernel (build_wpf_mel), device memory; null = none
  int wpf_words = 0;
};

struct FrontendArgs {
  // plan selection
  int n_fft = 0, hop = 0, win_len = 0;
  // input
  const float* x = nullptr;       // device, (batch, n_samples)
  int64_t batch = 0;
  int64_t n_samples = 0;          // real samples per clip
  int64_t zero_tail = 0;          // Whisper `padding`: virtual zeros appended before the reflect pad
  int pad_mode = PAD_NONE;
  int64_t pad_left = 0;           // padded coordinate p maps to signal index p - pad_left
  int pre_mode = PRE_NONE;
  const float* window = nullptr;  // HOST pointer, n_fft floats (already zero-extended)
  // spectrum -> mel -> log
  int spec_mode = SPEC_POWER;
  DeviceBank bank;
  int log_mode = LOG_NONE;
  float log_floor = 0.0f;
  int whisper_norm = 0;           // (x + 4) / 4 and per-clip max-8 clamp (needs clip_max / tile_min scratch)
  float post_sub = 0.0f, post_div = 1.0f;  // optional (x - post_sub) / post_div   (voice-encoder normalized_mels)
  int post_affine = 0;
  // output
  int out_mode = OUT_TM;
  int64_t n_frames = 0;           // frames to emit per clip
  float* out = nullptr;           // device
  int out_f16 = 0;                // `out` holds __half (Whisper (T', M) front end only): the store loop rounds, no second pass
  int lfr_m = 7, lfr_n = 6;       // OUT_LFR only
  int64_t lfr_rows = 0;
  // scratch for the Whisper clamp (device): clip_max (batch) ordered-int encoded, tile_min (batch * tiles)
  int* clip_max = nullptr;
  float* tile_min = nullptr;
  int* tile_max = nullptr;   // optional: batch * tiles more ints right behind tile_min (per-tile maxima: silent tiles are filled without being read)
  // ragged batch (per-clip lengths): device tables -- clip_tab[b] = int4(n_samples, n_frames, lfr_rows, first tile of the
  // clip), uploaded by the C ABI, and tile_tab[g] = int4(clip, tile, n_samples, n_frames), built from it on the device (launch_tile_table);
  // n_samples / n_frames / lfr_rows above are then those of the longest clip (they give the strides).
  // Null = every clip has n_samples samples.
  const void* clip_tab = nullptr;
  const void* tile_tab = nullptr;
  int64_t total_tiles = 0;
  int* tile_ctr = nullptr;   // one int of device scratch for the dynamic tile walk (behind the clamp tables when whisper_norm is set); null = static walk
};

// Returns 0 or a b2a_status; sets *launches to the number of kernels enqueued.
int launch_frontend(const FrontendArgs& a, void* stream, int* launches, std::string* err);
void frontend_dyn_tiles_enable(int on);
bool frontend_dyn_tiles();
int frontend_tiles_per_clip(int n_fft, int64_t n_frames);
int frontend_match_baked(const float* steps, int n_steps, const int* chunk_m, const int* chunk_s, int n_chunks, int frame_tile, int n_mels);
bool frontend_plan_exists(int n_fft, int hop, int win_len);
int init_frontend_tables(std::string* err);  // once per device: twiddle tables into __constant__ memory
// Warp-per-frame front end for n_fft 1920 / hop 480 (wpf1920.cu): |X| or |X|^2 -> any bank with a wpf_mel schedule -> log -> (M, T')
bool wpf1920_applicable(const FrontendArgs& a);
int launch_wpf1920(const FrontendArgs& a, void* stream, int* launches, std::string* err);
int init_wpf1920_tables(std::string* err);
void wpf1920_enable(int on);

// max - 8 clamp of the Whisper-style front ends over (batch, n_frames, n_mels) fp32 features in (T', M) layout, from the per-clip maxima and
// the (negated) per-32-frame-tile minima the main kernel left behind (frontend.cu: whisper_clamp_kernel)
int launch_whisper_clamp(float* out, const int* clip_max, const int* tile_min, int64_t batch, int64_t n_frames, int n_mels, void* stream,
                         int* launches, std::string* err);

// Tensor-core (tcgen05 + TMEM + TMA) Whisper front end (tc_frontend.cu): same inputs, outputs and clamp bookkeeping as the FFT kernel
bool tc_whisper_enabled();
void tc_whisper_enable(int on);
bool tc_whisper_applicable(const FrontendArgs& a);
void tc_debug_set_power_buffer(float* device_ptr);
int launch_tc_whisper(const FrontendArgs& a, void* stream, int* launches, std::string* err, float* dbg_power = nullptr);

// per-clip column statistics kernels (frontend.cu)
int launch_cmvn(const float* in, float* out, int64_t batch, int64_t rows, int dim, const float* mean, const float* istd,
                void* stream, int* launches, std::string* err, const void* clip_tab = nullptr);
int launch_mean_norm(float* inout, int64_t batch, int64_t rows, int dim, void* stream, int* launches, std::string* err,
                     const void* clip_tab = nullptr);
int launch_lfr(const float* in, float* out, int64_t batch, int64_t n_frames, int n_mels, int lfr_m, int lfr_n,
               void* stream, int* launches, std::string* err);
int launch_mel_segment_f16(const float* mel, void* out_f16, int64_t batch, int64_t n_frames, int n_mels, const long long* d_seek_content,
                           int length, void* stream, int* launches, std::string* err);
int launch_mel_windows(const float* mel, float* out, int n_mels, int64_t t_max, const int* d_seg, int n_segments, int window, void* stream,
                       int* launches, std::string* err);
int launch_resample_linear(const float* x, float* out, int64_t batch, int64_t T, int64_t new_t, float step, float hi_clip, void* stream,
                           int* launches, std::string* err);
int launch_resample_poly(const float* x, float* out, const float* h_dev, int64_t batch, int64_t T, int64_t new_t, int up, int down, int64_t pre, int n_taps,
                         void* stream, int* launches, std::string* err);
int launch_pcm16_to_f32(const void* in_i16, float* out, int64_t n, void* stream, int* launches, std::string* err);
int launch_pad_or_trim(const float* in, float* out, int64_t batch, int64_t n, int64_t length, void* stream,
                       int* launches, std::string* err);
int launch_tile_table(const void* clip_tab, int64_t n_clips, int64_t total_tiles, void* tile_tab, void* stream, int* launches, std::string* err);
int launch_zero_tails(float* out, const void* clip_tab, int which, int64_t batch, int64_t rows_max, int64_t row_len, int mel_major, int64_t max_tail,
                      void* stream, int* launches, std::string* err);
int launch_reflect_pad(const float* in, float* out, int64_t batch, int64_t n, int64_t pad, void* stream, int* launches, std::string* err);

// ---- generic (any n_fft / hop) STFT and spectrum -> mel kernels (generic_stft.cu): the sizes without a tuned plan -------------
bool generic_stft_supported(int n_fft, int hop);
int launch_generic_stft(const float* x, int64_t batch, int64_t n_samples, int64_t n_frames, int n_fft, int hop, int64_t pad_left, int pad_mode,
                        const float* window_dev, float* out_complex, void* stream, int* launches, std::string* err);
int launch_generic_mel(const float* spec_complex, float* out, int64_t batch, int64_t n_frames, int n_bins, const DeviceBank& bank, int spec_mode,
                       int log_mode, float log_floor, int post_affine, float post_sub, float post_div, int out_mode, void* stream, int* launches,
                       std::string* err);

// ---- vocoder STFT / iSTFT (vocoder.cu) -------------------------------------------------------
enum IstftNorm { NORM_WSQ_FLOOR = 0 /* HiFT, CosyVoice3: / max(sum w^2, 1e-8) */, NORM_WSUM_NONZERO = 1 /* Kokoro */ };
struct IstftArgs {
  int n_fft = 0, hop = 0;
  const float* mag = nullptr;     // device (batch, F, frames)
  const float* phase = nullptr;
  int64_t batch = 0, n_frames = 0;
  const float* window = nullptr;  // HOST, n_fft floats
  float clip_lo = 0.0f;
  int use_clip_lo = 0;            // CosyVoice3 clips below at 0; HiFT does not
  float clip_hi = 100.0f;
  int norm = NORM_WSQ_FLOOR;
  int unwrap = 0;                 // Kokoro: 0 none, 1 always unwrap (numpy-style, along time), 2 optimistic (flag + redo)
  int* d_flag = nullptr;          // unwrap == 2: device flag and pinned host mirror
  int* h_flag = nullptr;
  float* out = nullptr;           // device (batch, (frames-1)*hop)
  float* scratch_phase = nullptr; // device, same size as phase, when unwrap != 0
  int head = 0;                   // 1: `mag` is the vocoder's conv output (batch, 2F, frames); exp / sin formed in the kernel, `phase` unused
                                  // 2: `mag` is the complex64 spectrum (batch, F, frames) itself (mlxIstft); `phase` unused
  float out_limit = 0.0f;         // head: clip the waveform to +-out_limit (0 = none)
  const float* fade = nullptr;    // head: device, fade_len floats multiplying the start of every clip's waveform (null = none)
  int fade_len = 0;
};
int launch_istft(const IstftArgs& a, void* stream, int* launches, std::string* err);

enum SmallStftOut { SOUT_REAL_IMAG = 0, SOUT_MAG_PHASE = 1 };
struct SmallStftArgs {
  int n_fft = 0, hop = 0;
  const float* x = nullptr;       // device (batch, n_samples)
  int64_t batch = 0, n_samples = 0, n_frames = 0;
  int pad_mode = PAD_REFLECT;     // reflect (HiFT, Kokoro) or zero (CosyVoice3); pad = n_fft / 2
  const float* window = nullptr;  // HOST, n_fft floats
  int out_kind = SOUT_REAL_IMAG;
  float* out0 = nullptr;          // device (batch, F, frames): real or magnitude
  float* out1 = nullptr;          //                            imag or phase
};
int launch_small_stft(const SmallStftArgs& a, void* stream, int* launches, std::string* err);
int launch_unwrap(const float* phase, float* out, int64_t n_rows, int64_t n_frames, void* stream, int* launches, std::string* err);

}  // namespace b2a
