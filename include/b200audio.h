/*
 * b200audio.h -- C ABI of the B200-native STFT-family DSP path.
 *
 * Drop-in boundary for the DSP helpers of smdesai/mlx-swift-audio (the reference has no
 * FFI layer of its own: the seam is its set of Swift free functions, SURVEY.md section
 * 8b).  Every entry point below names the reference function it replaces (paths relative
 * to the reference's package/ directory).  Plain pointers and sizes only; no torch / MLX
 * types.  INTEGRATION.md shows the Swift-side binding.
 *
 * Conventions
 *  - All arithmetic is IEEE fp32 on the GPU (sm_100a CUDA kernels).  There is NO CPU
 *    fallback: every compute entry point fails with B2A_E_CUDA when no device is usable.
 *  - `space` says where the caller's buffers live: B2A_DEVICE (device pointers; work is
 *    enqueued on the context's stream and the call returns immediately) or B2A_HOST
 *    (host pointers, pinned preferred: the call stages clips through device memory in
 *    overlapped chunks and returns after the result is in the caller's buffer).
 *  - The reference helpers take ONE clip (T,).  Here every entry point takes `batch`
 *    independent clips of equal length laid out (batch, n_samples) row-major; batch = 1
 *    is exactly the reference call.  Outputs are (batch, ...reference shape...).
 *  - Output buffers are caller-allocated; the *_shape functions are pure integer code.
 *  - The reference calls fatalError on bad input; here a non-zero status is returned and
 *    b2a_last_error(ctx) describes it (the Swift shim turns it back into fatalError).
 *  - A context owns a stream, device tables (windows, sparse filterbanks, twiddles) and
 *    scratch.  One context per host thread / stream; contexts share no mutable state.
 */
#ifndef B200AUDIO_H_
#define B200AUDIO_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(_WIN32)
#define B2A_API
#else
#define B2A_API __attribute__((visibility("default")))
#endif

typedef struct b2a_ctx b2a_ctx;

enum b2a_status {
  B2A_OK = 0,
  B2A_E_BAD_ARG = 1,     /* null pointer, non-positive size, unsupported parameter value        */
  B2A_E_TOO_SHORT = 2,   /* reference: fatalError("Input is too short for STFT")                */
  B2A_E_CUDA = 3,        /* CUDA runtime / launch failure, or no sm_100 device                  */
  B2A_E_UNSUPPORTED = 4, /* valid in the reference but not built here (e.g. n_fft outside set)  */
  B2A_E_NOMEM = 5
};

enum b2a_space { B2A_HOST = 0, B2A_DEVICE = 1 };

/* window generators of the reference */
enum b2a_window_kind {
  B2A_WIN_WHISPER_HANN = 0,  /* whisperHannWindow        STT/Whisper/WhisperAudio.swift:32-44 (symmetric)         */
  B2A_WIN_HANNING = 1,       /* hanningWindow / hanning  Codec/S3Tokenizer/S3TokenizerUtils.swift:213-221,
                                                         TTS/Kokoro/Decoder/MLXSTFT.swift:12-20 (np.hanning)      */
  B2A_WIN_HAMMING = 2,       /* hammingWindow            STT/FunASR/FunASRAudio.swift:35-45                       */
  B2A_WIN_POVEY = 3,         /* poveyWindow              Codec/S3Gen/CAMPPlus.swift:15-19                         */
  B2A_WIN_HANN_PERIODIC = 4  /* hannWindowPeriodic       Codec/S3Gen/HiFiGAN.swift:15-20,
                                cosyVoice3HannWindowPeriodic TTS/CosyVoice3/HiFiGAN/CausalHiFTGenerator.swift:429-432 */
};

/* ---------------------------------------------------------------------------------------
 * context
 * ------------------------------------------------------------------------------------- */
B2A_API const char* b2a_version(void);
/* Creates a context on `device` with its own non-blocking stream. */
B2A_API int b2a_ctx_create(b2a_ctx** ctx, int device);
/* Same, but enqueues on the caller's stream (a cudaStream_t passed as void*; NULL = legacy default). */
B2A_API int b2a_ctx_create_on_stream(b2a_ctx** ctx, int device, void* cuda_stream);
B2A_API int b2a_ctx_destroy(b2a_ctx* ctx);
B2A_API int b2a_ctx_sync(b2a_ctx* ctx);
/* The cudaStream_t (as void*) the context enqueues on: callers that produce device inputs or consume device outputs on another
 * stream order the two with events (the Python mirror does this for torch tensors). */
B2A_API void* b2a_ctx_stream(const b2a_ctx* ctx);
B2A_API const char* b2a_last_error(const b2a_ctx* ctx);
/* Number of kernel launches issued through this context so far (bench.py's gpu_launches). */
B2A_API int64_t b2a_ctx_launch_count(const b2a_ctx* ctx);
/* Pinned host memory for B2A_HOST buffers (optional; pageable pointers also work, slower). */
B2A_API int b2a_host_alloc(void** ptr, uint64_t bytes);
B2A_API int b2a_host_free(void* ptr);

/* ---------------------------------------------------------------------------------------
 * multi-GPU: fused gather over peer memory
 *
 * One process per GPU.  The consumer rank allocates the full (all clips) feature buffer with
 * b2a_device_alloc and exports it; every producer rank opens the handle and passes
 * `peer_base + first_clip * clip_bytes` as the `out` pointer (space = B2A_DEVICE) of any compute
 * entry point: the kernels' own epilogue stores then travel over NVLink / NVSwitch into the
 * consumer's HBM while the producer is still computing -- no separate copy, no NCCL call on the
 * data path.  (The reference is single-device: MLX unified memory has no counterpart.)
 * The caller orders "all producers done" (b2a_ctx_sync on each rank + a host barrier) before the
 * consumer reads.  Handles are CUDA IPC handles (64 bytes) and are valid between processes of one
 * box.
 * ------------------------------------------------------------------------------------- */
#define B2A_IPC_HANDLE_BYTES 64
B2A_API int b2a_device_alloc(b2a_ctx* ctx, void** ptr, uint64_t bytes);   /* cudaMalloc on the context's device (IPC-exportable) */
B2A_API int b2a_device_free(b2a_ctx* ctx, void* ptr);
B2A_API int b2a_ipc_export(b2a_ctx* ctx, void* device_ptr, unsigned char handle[B2A_IPC_HANDLE_BYTES]);
B2A_API int b2a_ipc_open(b2a_ctx* ctx, const unsigned char handle[B2A_IPC_HANDLE_BYTES], void** peer_ptr);
B2A_API int b2a_ipc_close(b2a_ctx* ctx, void* peer_ptr);
/* Plain copies on the context's stream (device buffers obtained above have no torch / MLX owner to copy them). */
B2A_API int b2a_memcpy_d2h(b2a_ctx* ctx, void* host_dst, const void* device_src, uint64_t bytes);   /* synchronous */

/* ---------------------------------------------------------------------------------------
 * host-side pure functions (no GPU needed): windows, filterbanks, index and shape rules
 * ------------------------------------------------------------------------------------- */
B2A_API int b2a_window(int kind, int length, float* out);                       /* see enum b2a_window_kind */
/* melFilters (Slaney)      Codec/S3Tokenizer/S3TokenizerUtils.swift:301-375.  out (n_mels, n_fft/2+1).  f_max < 0 = nil (sr/2). */
B2A_API int b2a_mel_filters(int sample_rate, int n_fft, int n_mels, float f_min, float f_max, float* out);
/* funASRMelFilters (HTK)   STT/FunASR/FunASRAudio.swift:322-396.  out (n_mels, n_fft/2) -- 200-point linspace grid. */
B2A_API int b2a_funasr_mel_filters(int sample_rate, int n_fft, int n_mels, float* out);
/* melFiltersHTK            Codec/S3Gen/CAMPPlus.swift:134-175.  out (n_fft/2+1, n_mels) -- integer-bin triangles. */
B2A_API int b2a_mel_filters_htk(int sample_rate, int n_fft, int n_mels, float f_min, float f_max, float* out);
/* Source index of padded position i (0 <= i < n + 2*pad) under reflectPad / reflectPad1D
 * (S3TokenizerUtils.swift:266-298, FunASRAudio.swift:280-310), incl. the short-input loops. */
B2A_API int64_t b2a_reflect_pad_index(int64_t i, int64_t n, int64_t pad);
/* nextPowerOf2             Codec/S3Gen/CAMPPlus.swift:22-29 */
B2A_API int b2a_next_power_of_2(int n);
/* computeFeatureLength     STT/FunASR/FunASRAudio.swift:225-235 */
B2A_API int64_t b2a_funasr_compute_feature_length(int64_t audio_length, int hop_length, int lfr_n);

/* frame-count rules (all return < 0 for "too short") */
B2A_API int64_t b2a_stft_num_frames(int64_t n_samples, int n_fft, int hop, int center);  /* stft: 1 + (T + 2*(n_fft/2)*center - n_fft) / hop */
B2A_API int64_t b2a_whisper_num_frames(int64_t n_samples, int64_t padding);               /* stft frames - 1 */
B2A_API int64_t b2a_funasr_num_frames(int64_t n_samples);                                 /* 1 + T/160 */
B2A_API int64_t b2a_lfr_num_rows(int64_t n_frames, int lfr_n);                             /* ceil(T / lfr_n) */
B2A_API int64_t b2a_kaldi_num_frames(int64_t n_samples, int win_length, int hop);          /* max(1, (T - win)/hop + 1) */
B2A_API int64_t b2a_s3gen_num_frames(int64_t n_samples, int n_fft, int hop);               /* on reflectPad2D'ed signal, center=false */
B2A_API int64_t b2a_vocoder_stft_num_frames(int64_t n_samples, int n_fft, int hop);        /* (T + 2*(n_fft/2) - n_fft)/hop + 1 */
B2A_API int64_t b2a_istft_out_length(int64_t n_frames, int hop);                           /* (frames - 1) * hop */

/* ---------------------------------------------------------------------------------------
 * log-mel / fbank front ends
 * ------------------------------------------------------------------------------------- */

/* padOrTrim                STT/Whisper/WhisperAudio.swift:54-67.  out (batch, length). */
B2A_API int b2a_pad_or_trim(b2a_ctx* ctx, const float* x, int64_t batch, int64_t n_samples, int64_t length,
                            float* out, int space);
/* reflectPad / reflectPad1D   Codec/S3Tokenizer/S3TokenizerUtils.swift:266-298, STT/FunASR/FunASRAudio.swift:280-310.
 * x (batch, n_samples) -> out (batch, n_samples + 2 * padding), incl. the reference's loops for n_samples - 1 < padding and
 * the n_samples == 1 replicate case.  (The front ends never call this: they apply the same index map while staging frames.) */
B2A_API int b2a_reflect_pad(b2a_ctx* ctx, const float* x, int64_t batch, int64_t n_samples, int64_t padding, float* out, int space);

/* whisperLogMelSpectrogram STT/Whisper/WhisperAudio.swift:78-137.
 * audio (batch, n_samples) -> out (batch, T', n_mels), T' = b2a_whisper_num_frames(n_samples, padding).
 * The max-8 clamp uses each clip's own global maximum, as the single-clip reference does.
 * `padding` = zeros appended to every clip (the reference's own parameter; WhisperSTT.swift:139-144 concatenates 30 s of zeros before the
 * call -- pass padding = 480000 instead of concatenating: the zeros are neither stored nor copied, and the frames that lie entirely
 * inside them are filled with their final value (max(log floor, Lmax - 8)) instead of being transformed; the features are bit-identical
 * to those of explicit zeros). */
B2A_API int b2a_whisper_log_mel_spectrogram(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples,
                                            int n_mels, int64_t padding, float* out, int space);
/* The same features as IEEE fp16: whisperLogMelSpectrogram(...).asType(.float16), the only form the Whisper encoder ever consumes
 * (STT/Whisper/WhisperSTT.swift:156-157,181-182,649-655).  The cast is the last statement of the kernel's store loop (round to
 * nearest even of the fp32 value, so the result is bit-identical to casting b2a_whisper_log_mel_spectrogram's output): no fp32
 * feature tensor is written or re-read, and a B2A_HOST caller receives half the bytes.  n_mels = 80 or 128 (Whisper's banks).
 * out_f16 (batch, T', n_mels) of uint16 / __half. */
B2A_API int b2a_whisper_log_mel_spectrogram_f16(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples,
                                                int n_mels, int64_t padding, void* out_f16, int space);
/* 16-bit PCM in: sample = int16 / 32768 (exact in fp32), which is what AVAudioFile's float processing format hands the reference
 * for a 16-bit file (STT/Whisper/WhisperEngine.swift:327-369); then exactly the entry points above.  Halves the host->device
 * bytes of a B2A_HOST call.  out: float (out_is_f16 = 0) or fp16 (out_is_f16 = 1), (batch, T', n_mels). */
B2A_API int b2a_whisper_log_mel_spectrogram_pcm16(b2a_ctx* ctx, const int16_t* audio, int64_t batch, int64_t n_samples,
                                                  int n_mels, int64_t padding, int out_is_f16, void* out, int space);

/* logMelSpectrogramChatterbox Codec/S3Tokenizer/S3TokenizerUtils.swift:160-208 (and the wrapper
 * logMelSpectrogramCAMPPlus TTS/CosyVoice2/CosyVoice2TTS.swift:787-795).  out (batch, n_mels, T'). */
B2A_API int b2a_log_mel_spectrogram_chatterbox(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples,
                                               int n_mels, int64_t padding, float* out, int space);

/* funASRLogMelSpectrogram  STT/FunASR/FunASRAudio.swift:57-94 (n_fft 400, hop 160).  out (batch, T', n_mels). */
B2A_API int b2a_funasr_log_mel_spectrogram(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples,
                                           int n_mels, float* out, int space);
/* applyLFR                 STT/FunASR/FunASRAudio.swift:108-154.  features (batch, T, n_mels) -> (batch, ceil(T/lfr_n), lfr_m*n_mels). */
B2A_API int b2a_apply_lfr(b2a_ctx* ctx, const float* features, int64_t batch, int64_t n_frames, int n_mels,
                          int lfr_m, int lfr_n, float* out, int space);
/* applyCMVN                STT/FunASR/FunASRAudio.swift:165-180.  cmvn_mean/cmvn_istd (dim,) or both NULL (per-utterance). */
B2A_API int b2a_apply_cmvn(b2a_ctx* ctx, const float* features, int64_t batch, int64_t n_rows, int dim,
                           const float* cmvn_mean, const float* cmvn_istd, float* out, int space);
/* preprocessAudio          STT/FunASR/FunASRAudio.swift:197-216.  out (batch, ceil(T'/lfr_n), lfr_m*n_mels). */
B2A_API int b2a_funasr_preprocess_audio(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples,
                                        int n_mels, int lfr_m, int lfr_n, int apply_normalization,
                                        float* out, int space);
/* The same from 16-bit PCM (sample = int16 / 32768, as AVAudioFile decodes a 16-bit file; exact in fp32): bit-identical to the fp32
 * entry on the converted samples, half the bytes for a host caller.  Likewise for the two entries below. */
B2A_API int b2a_funasr_preprocess_audio_pcm16(b2a_ctx* ctx, const int16_t* audio, int64_t batch, int64_t n_samples,
                                              int n_mels, int lfr_m, int lfr_n, int apply_normalization,
                                              float* out, int space);

/* kaldiFbankCAMPPlus       Codec/S3Gen/CAMPPlus.swift:32-106.  out (batch, T', num_mel_bins).
 * mean_norm != 0 additionally applies the caller-side `fbank - mean(fbank, axis: 0)` of
 * CAMPPlus.inference (CAMPPlus.swift:797-802), per clip. */
B2A_API int b2a_kaldi_fbank_campplus(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples,
                                     int sample_rate, int num_mel_bins, float frame_length_ms, float frame_shift_ms,
                                     int mean_norm, float* out, int space);
B2A_API int b2a_kaldi_fbank_campplus_pcm16(b2a_ctx* ctx, const int16_t* audio, int64_t batch, int64_t n_samples,
                                           int sample_rate, int num_mel_bins, float frame_length_ms, float frame_shift_ms,
                                           int mean_norm, float* out, int space);

/* s3genMelSpectrogram      Codec/S3Gen/Mel/S3GenMel.swift:43-102 (wrappers computeMelSpectrogram80
 * CosyVoice2TTS.swift:754-770, CosyVoice3TTS.swift:770-787, melSpectrogramS3Gen ChatterboxTurboModel.swift:545-572).
 * y (batch, n_samples) -> out (batch, num_mels, T').  Built for n_fft=win_size=1920, hop=480. */
B2A_API int b2a_s3gen_mel_spectrogram(b2a_ctx* ctx, const float* y, int64_t batch, int64_t n_samples, int n_fft,
                                      int num_mels, int sampling_rate, int hop_size, int win_size, int fmin, int fmax,
                                      float* out, int space);
B2A_API int b2a_s3gen_mel_spectrogram_pcm16(b2a_ctx* ctx, const int16_t* y, int64_t batch, int64_t n_samples, int n_fft,
                                            int num_mels, int sampling_rate, int hop_size, int win_size, int fmin, int fmax,
                                            float* out, int space);

/* voiceEncoderMelspectrogram TTS/Chatterbox/VoiceEncoder/VoiceEncoderMelspec.swift:17-68 with VoiceEncConfig
 * (Config/ChatterboxConfig.swift:139-156).  out (batch, num_mels, T'). */
typedef struct b2a_voice_enc_config {
  int num_mels;            /* 40    */
  int sample_rate;         /* 16000 */
  int n_fft;               /* 400   */
  int hop_size;            /* 160   */
  int win_size;            /* 400   */
  int fmin;                /* 0     */
  int fmax;                /* 8000  */
  float mel_power;         /* 2.0 (1.0 and 2.0 are built) */
  int mel_type_db;         /* 0 = "amp" (default), 1 = "db" */
  int normalized_mels;     /* 0 */
  float stft_magnitude_min;/* 1e-4 */
} b2a_voice_enc_config;
B2A_API void b2a_voice_enc_config_default(b2a_voice_enc_config* cfg);
B2A_API int b2a_voice_encoder_melspectrogram(b2a_ctx* ctx, const float* wav, int64_t batch, int64_t n_samples,
                                             const b2a_voice_enc_config* cfg, float* out, int space);

/* ---------------------------------------------------------------------------------------
 * ragged batches: per-clip lengths ("optional per-clip lengths[B]")
 *
 * The reference helpers take one clip, so clips of different lengths are simply separate calls there.  These variants (every mel front end)
 * run a whole batch of unequal clips in ONE launch: clip b holds lengths[b] (host array, 1 <= lengths[b] <= n_samples) valid
 * samples at the start of row b of the (batch, n_samples) input, and every clip is processed exactly as the single-clip
 * reference call on audio[b, :lengths[b]] would process it (padding, frame count, per-clip max / CMVN / mean over ITS frames).
 * Output strides are those of an n_samples-long clip -- (batch, frames(n_samples), n_mels) etc. -- rows past a clip's own
 * count are zero; out_frames / out_rows (host, may be NULL) receive each clip's count.
 * ------------------------------------------------------------------------------------- */
B2A_API int b2a_whisper_log_mel_spectrogram_ragged(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                                   int n_mels, int64_t padding, float* out, int64_t* out_frames, int space);
B2A_API int b2a_whisper_log_mel_spectrogram_f16_ragged(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                                       int n_mels /* 128 */, int64_t padding, void* out_f16, int64_t* out_frames, int space);
B2A_API int b2a_log_mel_spectrogram_chatterbox_ragged(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                                      int n_mels, int64_t padding, float* out, int64_t* out_frames, int space);
B2A_API int b2a_funasr_log_mel_spectrogram_ragged(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                                  int n_mels, float* out, int64_t* out_frames, int space);
B2A_API int b2a_voice_encoder_melspectrogram_ragged(b2a_ctx* ctx, const float* wav, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                                    const b2a_voice_enc_config* cfg /* NULL = defaults */, float* out, int64_t* out_frames,
                                                    int space);
B2A_API int b2a_funasr_preprocess_audio_ragged(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                               int n_mels, int lfr_m, int lfr_n, int apply_normalization, float* out, int64_t* out_rows, int space);
B2A_API int b2a_kaldi_fbank_campplus_ragged(b2a_ctx* ctx, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                            int sample_rate, int num_mel_bins, float frame_length_ms, float frame_shift_ms, int mean_norm,
                                            float* out, int64_t* out_frames, int space);
B2A_API int b2a_s3gen_mel_spectrogram_ragged(b2a_ctx* ctx, const float* y, int64_t batch, int64_t n_samples, const int64_t* lengths, int n_fft,
                                             int num_mels, int sampling_rate, int hop_size, int win_size, int fmin, int fmax, float* out,
                                             int64_t* out_frames, int space);

/* stft                     Codec/S3Tokenizer/S3TokenizerUtils.swift:224-263 (== funASRSTFT FunASRAudio.swift:240-277).
 * window (win_len,) host pointer, zero-extended to n_fft; out (batch, T', n_fft/2+1) complex64 interleaved (re, im).
 * Built for n_fft in {400, 512, 1920} with hop {160, 160, 480}. */
B2A_API int b2a_stft(b2a_ctx* ctx, const float* x, int64_t batch, int64_t n_samples, const float* window, int win_len,
                     int n_fft, int hop, int center, float* out_complex, int space);

/* ---------------------------------------------------------------------------------------
 * vocoder STFT / iSTFT (n_fft 16 hop 4, n_fft 20 hop 5)
 * ------------------------------------------------------------------------------------- */

/* stftHiFiGAN              Codec/S3Gen/HiFiGAN.swift:257-295.  x (batch, T) -> real, imag (batch, n_fft/2+1, frames).
 * window (n_fft,) host pointer. */
B2A_API int b2a_stft_hifigan(b2a_ctx* ctx, const float* x, int64_t batch, int64_t n_samples, int n_fft, int hop,
                             const float* window, float* real_out, float* imag_out, int space);
/* istftHiFiGAN             Codec/S3Gen/HiFiGAN.swift:298-367.  magnitude, phase (batch, n_fft/2+1, frames)
 * -> out (batch, (frames-1)*hop). */
B2A_API int b2a_istft_hifigan(b2a_ctx* ctx, const float* magnitude, const float* phase, int64_t batch,
                              int64_t n_frames, int n_fft, int hop, const float* window, float* out, int space);
/* cosyVoice3Stft           TTS/CosyVoice3/HiFiGAN/CausalHiFTGenerator.swift:435-460 (zero padding). */
B2A_API int b2a_cosyvoice3_stft(b2a_ctx* ctx, const float* x, int64_t batch, int64_t n_samples, int n_fft, int hop,
                                const float* window, float* real_out, float* imag_out, int space);
/* cosyVoice3Istft          TTS/CosyVoice3/HiFiGAN/CausalHiFTGenerator.swift:463-514 (clip magnitude to [0, 100]). */
B2A_API int b2a_cosyvoice3_istft(b2a_ctx* ctx, const float* magnitude, const float* phase, int64_t batch,
                                 int64_t n_frames, int n_fft, int hop, const float* window, float* out, int space);
/* MLXSTFT.transform        TTS/Kokoro/Decoder/MLXSTFT.swift:181-209 (mlxStft :69-113, "hann", center, reflect).
 * x (batch, T) -> magnitude, phase (batch, filter_length/2+1, frames). */
B2A_API int b2a_kokoro_stft_transform(b2a_ctx* ctx, const float* x, int64_t batch, int64_t n_samples,
                                      int filter_length, int hop_length, int win_length,
                                      float* magnitude_out, float* phase_out, int space);
/* MLXSTFT.inverse          TTS/Kokoro/Decoder/MLXSTFT.swift:211-235 (unwrap :23-46, mlxIstft :115-163).
 * magnitude, phase (batch, F, frames) -> out (batch, 1, (frames-1)*hop). */
B2A_API int b2a_kokoro_stft_inverse(b2a_ctx* ctx, const float* magnitude, const float* phase, int64_t batch,
                                    int64_t n_frames, int filter_length, int hop_length, int win_length,
                                    float* out, int space);
/* mlxIstft TTS/Kokoro/Decoder/MLXSTFT.swift:115-163 (stand-alone: no unwrap, no polar conversion).  spec_complex: complex64
 * (batch, F, frames) as interleaved (re, im) floats, F = win_length / 2 + 1; "hann" window; out (batch, (frames-1)*hop).
 * Built for (win_length, hop_length) = (20, 5) and (16, 4). */
B2A_API int b2a_mlx_istft(b2a_ctx* ctx, const float* spec_complex, int64_t batch, int64_t n_frames, int win_length, int hop_length,
                          float* out, int space);
/* unwrap                   TTS/Kokoro/Decoder/MLXSTFT.swift:23-46: numpy-style phase unwrap along the last axis of (n_rows, n_frames)
 * (what b2a_kokoro_stft_inverse applies to the phase; the identity unless some |phase[t] - phase[t-1]| >= pi). */
B2A_API int b2a_unwrap(b2a_ctx* ctx, const float* phase, int64_t n_rows, int64_t n_frames, float* out, int space);

/* ---- adjacent rows: vocoder glue fused around the iSTFT ----------------------------------------
 * HiFT head: h = convPost output (batch, n_fft + 2, frames).  magnitude = exp(h[:, :F]), phase = sin(h[:, F:]),
 * istftHiFiGAN, clip(output, -audio_limit, audio_limit).  Codec/S3Gen/HiFiGAN.swift:577-589 (audio_limit 0.99). */
B2A_API int b2a_hift_head_istft(b2a_ctx* ctx, const float* conv_out, int64_t batch, int64_t n_frames, int n_fft, int hop,
                                const float* window, float audio_limit, float* out, int space);
/* The same followed by the 20 ms fade-in of S3Gen.callAsFunction (Codec/S3Gen/S3Gen.swift:259-262, 284-289), still one kernel:
 * the first fade_len samples of every clip are multiplied by fade[] (host, fade_len floats) when the clip has at least fade_len
 * samples.  b2a_s3gen_trim_fade writes the reference's window, zeros(sr/50) ++ (cos(linspace(pi, 0, sr/50)) + 1) / 2, into
 * out[2 * (sr / 50)] (pure host function). */
B2A_API int b2a_hift_head_istft_fade(b2a_ctx* ctx, const float* conv_out, int64_t batch, int64_t n_frames, int n_fft, int hop,
                                     const float* window, float audio_limit, const float* fade, int64_t fade_len, float* out, int space);
B2A_API int b2a_s3gen_trim_fade(int sampling_rate, float* out);
/* Kokoro head: x = conv_post output (batch, filter_length + 2, frames).  spec = exp(x[:, :F]), phase = sin(x[:, F:]),
 * MLXSTFT.inverse.  TTS/Kokoro/Decoder/Generator.swift:182-190.  out (batch, 1, (frames-1)*hop). */
B2A_API int b2a_kokoro_head_istft(b2a_ctx* ctx, const float* conv_out, int64_t batch, int64_t n_frames, int filter_length,
                                  int hop_length, int win_length, float* out, int space);

/* Whisper seek window: per clip, rows [seek, seek + min(length, content_frames - seek)) of the
 * (batch, n_frames, n_mels) fp32 log-mel, zero-padded to `length` rows and cast to fp16 -- melSegment / padOrTrimMel / asType(.float16),
 * STT/Whisper/WhisperSTT.swift:171-182,624-635 (and :156-157 with seek = 0).  seek / content_frames: HOST int64[batch].
 * out_f16: (batch, length, n_mels) IEEE half, in `space`. */
B2A_API int b2a_whisper_mel_segment_f16(b2a_ctx* ctx, const float* mel, int64_t batch, int64_t n_frames, int n_mels,
                                        const int64_t* seek, const int64_t* content_frames, int64_t length, void* out_f16, int space);

/* resampleAudio / linearInterpolate1d: align_corners=False linear interpolation in fp32, bit-exact with
 * TTS/CosyVoice2/CosyVoice2TTS.swift:733-744 + TTS/CosyVoice2/HiFiGAN/CosyHiFTGenerator.swift:17-58.  out (batch, new length).
 * (AVAudioConverter, Audio/AudioResampler.swift, is Apple's proprietary converter and has no reproducible definition.) */
B2A_API int64_t b2a_resample_linear_length(int64_t n_samples, int from_rate, int to_rate);
/* NON-PARITY EXTENSION -- anti-aliased polyphase resampler standing in for AudioResampler.resample (Audio/AudioResampler.swift:15-88,
 * call sites TTS/Chatterbox/ChatterboxModel.swift:445-462, STT/Whisper/WhisperEngine.swift:308-322), which is Apple's AVAudioConverter
 * and has no reproducible definition.  Defined as scipy.signal.resample_poly(x, up, down) with its default design (its documented
 * oracle, tests at 1e-5 absolute): rates reduced by their gcd, firwin(20 max(up, down) + 1, 1 / max(up, down), kaiser beta 5) * up,
 * zero phase, zeros outside the clip, out (batch, ceil(n up / down)).  b2a_resample_poly_filter (host) returns the padded filter. */
B2A_API int64_t b2a_resample_poly_length(int64_t n_samples, int from_rate, int to_rate);
B2A_API int64_t b2a_resample_poly_filter(int64_t n_samples, int from_rate, int to_rate, float* h_out, int64_t cap, int* up_out, int* down_out,
                                         int64_t* pre_remove_out);
B2A_API int b2a_resample_poly(b2a_ctx* ctx, const float* x, int64_t batch, int64_t n_samples, int from_rate, int to_rate, float* out, int space);
B2A_API int b2a_resample_linear(b2a_ctx* ctx, const float* x, int64_t batch, int64_t n_samples, int from_rate, int to_rate, float* out,
                                int space);

/* S3Tokenizer long-audio windows: the segment plan of quantize / quantizeMixedBatch
 * (Codec/S3Tokenizer/S3Tokenizer.swift:474-571; window 3000 frames, stride 2600) -- host only; pass null outputs to count -- and the
 * gather of the unified batch: out[s] = mel[batch_idx[s]][:, start[s] ..< start[s] + length[s]] zero-padded to `window` frames.
 * mel (batch, n_mels, t_max) fp32, out (n_segments, n_mels, window) fp32; batch_idx / start / length: HOST int32[n_segments]. */
B2A_API int64_t b2a_s3tokenizer_plan_segments(const int64_t* mel_len, int64_t batch, int64_t window, int64_t stride, int32_t* batch_idx,
                                              int32_t* start, int32_t* length, int64_t cap);
/* mergeTokenizedSegments Codec/S3Tokenizer/S3TokenizerUtils.swift:71-88 (pure host function): joins the token sequences of a long
 * clip's windows, dropping (overlap / 2) * token_rate tokens at every inner edge.  tokens = the segments back to back,
 * seg_len[n_segments]; returns the merged length (out may be NULL to query it), -1 on bad arguments or when `cap` is too small. */
B2A_API int64_t b2a_merge_tokenized_segments(const int32_t* tokens, const int64_t* seg_len, int64_t n_segments, int overlap, int token_rate,
                                             int32_t* out, int64_t cap);
B2A_API int b2a_s3tokenizer_gather_segments(b2a_ctx* ctx, const float* mel, int64_t batch, int n_mels, int64_t t_max, int64_t n_segments,
                                            const int32_t* batch_idx, const int32_t* start, const int32_t* length, int64_t window,
                                            float* out, int space);

/* Test hook (host only, no GPU): compiles a dense filterbank ((n_mels, n_bins), or (n_bins, n_mels) when
 * bin_major) into the kernel's sparse mel "step program" and interprets it on the host for one spectrum p.
 * Returns the number of steps, -1 if the bank is not of the <=2-adjacent-filters-per-bin form. */
B2A_API int b2a_debug_mel_program_apply(const float* bank, int n_mels, int n_bins, int bin_major, const float* p, float* out);
/* Test hook (host only): builds the mel schedule of the warp-per-frame n_fft 1920 front end (balanced per-lane segments) for a dense
 * filterbank and interprets it on the host for one spectrum p, exactly as the kernel does.  Returns the schedule's size in 32-bit
 * words (+ 65536 x the number of segments left with a shared-memory bank conflict), -1 if the bank does not fit the schedule, -2 if a
 * lane would read past the spectrum row, -3 if a start bin is not 16-byte aligned. */
B2A_API int b2a_debug_wpf_mel_apply(const float* bank, int n_mels, int n_bins, int bin_major, const float* p, float* out);
/* Test hook (host only): shared-memory layout of the FFT plan of n_fft (spectrum rows and mel staging words in the exchange buffer). */
B2A_API int b2a_debug_plan_layout(int n_fft, int n_mels, int* slots_out, int* words_out);
/* Build hook (host only): raw mel step program for a CTA shape (see tools/gen_mel_baked.py). */
B2A_API int b2a_debug_mel_program_dump(const float* bank, int n_mels, int n_bins, int bin_major, int n_fft,
                                       unsigned* steps_out, int cap_steps, int* chunk_m, int* chunk_s, int* n_chunks_out,
                                       int* frame_tile_out);

/* ---------------------------------------------------------------------------------------
 * instrumentation used by bench.py (device-side timing of the last call's kernels)
 * ------------------------------------------------------------------------------------- */
/* When enabled, every compute entry point brackets its kernels with CUDA events on the
 * context's stream; b2a_ctx_last_kernel_ms returns the elapsed time of the last call
 * (synchronises the two events).  Off by default. */
/* Switches the experimental tensor-core Whisper front end (tcgen05 split-precision DFT, csrc/tc_frontend.cu) on / off for the
 * process; the environment variable B2A_WHISPER_TC=1 sets the initial state.  Off by default: DESIGN.md section 6. */
B2A_API int b2a_debug_whisper_tc(int on);
/* Switches the warp-per-frame n_fft 1920 front end (csrc/wpf1920.cu; on by default, B2A_WPF1920=0 sets the initial state to off) on / off
 * for the process: off = the tiled lane == frame kernel of csrc/frontend.cu runs the S3Gen mel (A/B switch, both are parity-tested).
 * on == 2: the warp-per-frame kernel without its pruned stage B (banks that end at or below bin 640, e.g. S3Gen's fmax 8000 Hz at
 * 24 kHz, normally skip the ten dead outputs of every 32-point transform; results are bit-identical either way). */
B2A_API int b2a_debug_wpf1920(int on);
/* Switches the dynamic tile walk of the tiled front-end kernels on / off for the process (on by default; B2A_DYN_TILES=0 sets the initial
 * state to off): on = after its first tile a persistent CTA takes its next tile from a launch-wide atomic counter, off = the static stride
 * blockIdx.x + k * gridDim.x.  Results are bit-identical (every tile is computed the same way whoever computes it); A/B switch. */
B2A_API int b2a_debug_dyn_tiles(int on);
/* Bring-up hook of the tensor-core Whisper front end (B2A_WHISPER_TC=1): when a device buffer of (batch, T', 201) floats is set, the
 * kernel also leaves the power spectrum |X[k]|^2 of every frame there.  NULL switches it off. */
B2A_API int b2a_debug_tc_power_buffer(void* device_ptr);
B2A_API int b2a_ctx_enable_timing(b2a_ctx* ctx, int on);
B2A_API int b2a_ctx_last_kernel_ms(b2a_ctx* ctx, float* ms);

#ifdef __cplusplus
}
#endif
#endif /* B200AUDIO_H_ */
