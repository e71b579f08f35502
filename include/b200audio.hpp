// b200audio.hpp -- header-only C++ mirror of the reference's Swift DSP helpers over the C ABI.
//
// The reference is Swift (a compiled language) and no Swift toolchain exists in the build image, so the
// compiled-language host layer is C++: same function names, argument meaning, defaults and return shapes
// as the Swift helpers; a non-zero status throws b2a::Error where the reference
// calls fatalError.  Host buffers (std::vector<float>) go through B2A_HOST; device pointers can be passed
// to the C ABI directly with B2A_DEVICE.
#pragma once
#include <cstdint>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

#include "b200audio.h"

namespace b2a {

struct Error : std::runtime_error {
  int status;
  Error(int s, const std::string& m) : std::runtime_error(m), status(s) {}
};

struct Array {  // contiguous fp32, row-major
  std::vector<int64_t> shape;
  std::vector<float> data;
  Array() = default;
  explicit Array(std::vector<int64_t> s) : shape(std::move(s)) {
    int64_t n = 1;
    for (auto d : shape) n *= d;
    data.assign(size_t(n), 0.0f);
  }
};

class Context {
 public:
  explicit Context(int device = 0) {
    if (b2a_ctx_create(&c_, device) != B2A_OK) throw Error(B2A_E_CUDA, "b200audio: no sm_100 CUDA device");
  }
  ~Context() { b2a_ctx_destroy(c_); }
  Context(const Context&) = delete;
  Context& operator=(const Context&) = delete;
  b2a_ctx* raw() const { return c_; }
  void check(int rc) const {
    if (rc != B2A_OK) throw Error(rc, b2a_last_error(c_));
  }

  // STT/Whisper/WhisperAudio.swift:54-67
  Array padOrTrim(const Array& array, int64_t length = 480000) const {
    Array out({length});
    check(b2a_pad_or_trim(c_, array.data.data(), 1, array.shape[0], length, out.data.data(), B2A_HOST));
    return out;
  }
  // STT/Whisper/WhisperAudio.swift:78-137 -> (T', nMels); a leading batch axis is accepted
  Array whisperLogMelSpectrogram(const Array& audio, int nMels, int64_t padding = 0) const {
    const int64_t b = audio.shape.size() == 2 ? audio.shape[0] : 1, n = audio.shape.back();
    const int64_t frames = b2a_whisper_num_frames(n, padding);
    if (frames <= 0) throw Error(B2A_E_TOO_SHORT, "Input is too short for STFT");
    Array out(audio.shape.size() == 2 ? std::vector<int64_t>{b, frames, nMels} : std::vector<int64_t>{frames, nMels});
    check(b2a_whisper_log_mel_spectrogram(c_, audio.data.data(), b, n, nMels, padding, out.data.data(), B2A_HOST));
    return out;
  }
  // Codec/S3Tokenizer/S3TokenizerUtils.swift:160-208 -> (nMels, T')
  Array logMelSpectrogramChatterbox(const Array& audio, int nMels = 128, int64_t padding = 0) const {
    const int64_t n = audio.shape.back(), frames = b2a_whisper_num_frames(n, padding);
    if (frames <= 0) throw Error(B2A_E_TOO_SHORT, "Input is too short for STFT");
    Array out({nMels, frames});
    check(b2a_log_mel_spectrogram_chatterbox(c_, audio.data.data(), 1, n, nMels, padding, out.data.data(), B2A_HOST));
    return out;
  }
  // STT/FunASR/FunASRAudio.swift:197-216
  Array preprocessAudio(const Array& audio, int nMels = 80, int lfrM = 7, int lfrN = 6, bool applyNormalization = true) const {
    const int64_t n = audio.shape.back(), frames = b2a_funasr_num_frames(n);
    if (frames <= 0) throw Error(B2A_E_TOO_SHORT, "Input is too short for STFT");
    Array out({b2a_lfr_num_rows(frames, lfrN), int64_t(nMels) * lfrM});
    check(b2a_funasr_preprocess_audio(c_, audio.data.data(), 1, n, nMels, lfrM, lfrN, applyNormalization, out.data.data(), B2A_HOST));
    return out;
  }
  // Codec/S3Gen/CAMPPlus.swift:32-106
  Array kaldiFbankCAMPPlus(const Array& audio, int sampleRate = 16000, int numMelBins = 80, float frameLength = 25.0f,
                           float frameShift = 10.0f) const {
    const int64_t n = audio.shape.back();
    const int win = int(float(sampleRate) * frameLength / 1000), hop = int(float(sampleRate) * frameShift / 1000);
    const int64_t frames = b2a_kaldi_num_frames(n, win, hop);
    if (frames <= 0) throw Error(B2A_E_TOO_SHORT, "signal shorter than one analysis window");
    Array out({frames, numMelBins});
    check(b2a_kaldi_fbank_campplus(c_, audio.data.data(), 1, n, sampleRate, numMelBins, frameLength, frameShift, 0, out.data.data(), B2A_HOST));
    return out;
  }
  // Codec/S3Gen/Mel/S3GenMel.swift:43-102: (B, T) -> (B, numMels, T')
  Array s3genMelSpectrogram(const Array& y, int nFft = 1920, int numMels = 80, int samplingRate = 24000, int hopSize = 480,
                            int winSize = 1920, int fmin = 0, int fmax = 8000) const {
    const bool was1d = y.shape.size() == 1;
    const int64_t b = was1d ? 1 : y.shape[0], n = y.shape.back(), frames = b2a_s3gen_num_frames(n, nFft, hopSize);
    if (frames <= 0) throw Error(B2A_E_TOO_SHORT, "Input is too short for STFT");
    Array out(was1d ? std::vector<int64_t>{numMels, frames} : std::vector<int64_t>{b, numMels, frames});
    check(b2a_s3gen_mel_spectrogram(c_, y.data.data(), b, n, nFft, numMels, samplingRate, hopSize, winSize, fmin, fmax, out.data.data(), B2A_HOST));
    return out;
  }
  // Codec/S3Gen/HiFiGAN.swift:298-367
  Array istftHiFiGAN(const Array& magnitude, const Array& phase, int nFft, int hopLength, const std::vector<float>& window) const {
    const int64_t b = magnitude.shape[0], frames = magnitude.shape[2];
    Array out({b, (frames - 1) * hopLength});
    check(b2a_istft_hifigan(c_, magnitude.data.data(), phase.data.data(), b, frames, nFft, hopLength, window.data(), out.data.data(), B2A_HOST));
    return out;
  }
  // Codec/S3Gen/HiFiGAN.swift:257-295
  std::pair<Array, Array> stftHiFiGAN(const Array& x, int nFft, int hopLength, const std::vector<float>& window) const {
    const int64_t b = x.shape[0], n = x.shape[1], frames = b2a_vocoder_stft_num_frames(n, nFft, hopLength);
    if (frames <= 0) throw Error(B2A_E_TOO_SHORT, "Input is too short");
    Array re({b, nFft / 2 + 1, frames}), im({b, nFft / 2 + 1, frames});
    check(b2a_stft_hifigan(c_, x.data.data(), b, n, nFft, hopLength, window.data(), re.data.data(), im.data.data(), B2A_HOST));
    return {std::move(re), std::move(im)};
  }
  // TTS/Kokoro/Decoder/MLXSTFT.swift:211-235
  Array mlxStftInverse(const Array& magnitude, const Array& phase, int filterLength = 20, int hopLength = 5, int winLength = 20) const {
    const int64_t b = magnitude.shape[0], frames = magnitude.shape[2];
    Array out({b, 1, (frames - 1) * hopLength});
    check(b2a_kokoro_stft_inverse(c_, magnitude.data.data(), phase.data.data(), b, frames, filterLength, hopLength, winLength, out.data.data(), B2A_HOST));
    return out;
  }

  // ---- adjacent rows ----
  // Tail of HiFTGenerator.decode, Codec/S3Gen/HiFiGAN.swift:577-589: convPost output (B, nFft+2, frames) -> waveform, one kernel
  Array hiftHeadIstft(const Array& convOut, int nFft, int hopLength, const std::vector<float>& window, float audioLimit = 0.99f) const {
    const int64_t b = convOut.shape[0], frames = convOut.shape[2];
    Array out({b, (frames - 1) * hopLength});
    check(b2a_hift_head_istft(c_, convOut.data.data(), b, frames, nFft, hopLength, window.data(), audioLimit, out.data.data(), B2A_HOST));
    return out;
  }
  // ... followed by the fade-in of S3Token2Wav.callAsFunction (Codec/S3Gen/S3Gen.swift:284-289); trimFade from s3genTrimFade()
  Array hiftHeadIstftFade(const Array& convOut, int nFft, int hopLength, const std::vector<float>& window, const std::vector<float>& trimFade,
                          float audioLimit = 0.99f) const {
    const int64_t b = convOut.shape[0], frames = convOut.shape[2];
    Array out({b, (frames - 1) * hopLength});
    check(b2a_hift_head_istft_fade(c_, convOut.data.data(), b, frames, nFft, hopLength, window.data(), audioLimit, trimFade.data(),
                                   int64_t(trimFade.size()), out.data.data(), B2A_HOST));
    return out;
  }
  // Tail of the Kokoro generator, TTS/Kokoro/Decoder/Generator.swift:182-190
  Array kokoroHeadIstft(const Array& convOut, int filterLength = 20, int hopLength = 5, int winLength = 20) const {
    const int64_t b = convOut.shape[0], frames = convOut.shape[2];
    Array out({b, 1, (frames - 1) * hopLength});
    check(b2a_kokoro_head_istft(c_, convOut.data.data(), b, frames, filterLength, hopLength, winLength, out.data.data(), B2A_HOST));
    return out;
  }
  // Seek window of the Whisper decode loop, STT/Whisper/WhisperSTT.swift:171-182,624-635: (T', M) fp32 -> (length, M) IEEE half bits
  std::vector<uint16_t> whisperMelSegment(const Array& mel, int64_t seek, int64_t contentFrames, int64_t length = 3000) const {
    const int64_t t = mel.shape[0], m = mel.shape[1];
    std::vector<uint16_t> out(size_t(length * m));
    check(b2a_whisper_mel_segment_f16(c_, mel.data.data(), 1, t, int(m), &seek, &contentFrames, length, out.data(), B2A_HOST));
    return out;
  }

  // ---- ragged batches: clips of different lengths in one launch (b200audio.h, "ragged batches") ----
  // audio (B, T_max) with lengths[b] valid samples per row -> (B, T'max, nMels) (rows past a clip's frame count are zero) and the
  // frame count of every clip; each clip is processed exactly like whisperLogMelSpectrogram(audio[b, :lengths[b]])
  std::pair<Array, std::vector<int64_t>> whisperLogMelSpectrogramRagged(const Array& audio, const std::vector<int64_t>& lengths, int nMels,
                                                                       int64_t padding = 0) const {
    const int64_t b = audio.shape.at(0), n = audio.shape.at(1), frames = b2a_whisper_num_frames(n, padding);
    if (frames <= 0 || int64_t(lengths.size()) != b) throw Error(B2A_E_BAD_ARG, "bad ragged batch");
    Array out({b, frames, nMels});
    std::vector<int64_t> rows(size_t(b), 0);
    check(b2a_whisper_log_mel_spectrogram_ragged(c_, audio.data.data(), b, n, lengths.data(), nMels, padding, out.data.data(), rows.data(), B2A_HOST));
    return {std::move(out), std::move(rows)};
  }
  // preprocessAudio per clip -> (B, rows_max, nMels * lfrM) and the LFR row count of every clip
  std::pair<Array, std::vector<int64_t>> preprocessAudioRagged(const Array& audio, const std::vector<int64_t>& lengths, int nMels = 80, int lfrM = 7,
                                                              int lfrN = 6, bool applyNormalization = true) const {
    const int64_t b = audio.shape.at(0), n = audio.shape.at(1), frames = b2a_funasr_num_frames(n);
    if (frames <= 0 || int64_t(lengths.size()) != b) throw Error(B2A_E_BAD_ARG, "bad ragged batch");
    Array out({b, b2a_lfr_num_rows(frames, lfrN), int64_t(nMels) * lfrM});
    std::vector<int64_t> rows(size_t(b), 0);
    check(b2a_funasr_preprocess_audio_ragged(c_, audio.data.data(), b, n, lengths.data(), nMels, lfrM, lfrN, applyNormalization, out.data.data(),
                                             rows.data(), B2A_HOST));
    return {std::move(out), std::move(rows)};
  }

  // ---- multi-GPU: the consumer rank's feature buffer, mapped into every producer process (b200audio.h, "fused gather") ----
  void* deviceAlloc(uint64_t bytes) const {
    void* p = nullptr;
    check(b2a_device_alloc(c_, &p, bytes));
    return p;
  }
  void deviceFree(void* p) const { check(b2a_device_free(c_, p)); }
  std::vector<unsigned char> ipcExport(void* devicePtr) const {
    std::vector<unsigned char> h(B2A_IPC_HANDLE_BYTES);
    check(b2a_ipc_export(c_, devicePtr, h.data()));
    return h;
  }
  void* ipcOpen(const std::vector<unsigned char>& handle) const {
    void* p = nullptr;
    check(b2a_ipc_open(c_, handle.data(), &p));
    return p;
  }
  void ipcClose(void* peerPtr) const { check(b2a_ipc_close(c_, peerPtr)); }

 private:
  b2a_ctx* c_ = nullptr;
};

// host-only helpers
inline std::vector<float> s3genTrimFade(int samplingRate = 24000) {  // Codec/S3Gen/S3Gen.swift:259-262
  std::vector<float> f(size_t(2 * (samplingRate / 50)));
  if (b2a_s3gen_trim_fade(samplingRate, f.data()) != B2A_OK) throw Error(B2A_E_BAD_ARG, "bad sampling rate");
  return f;
}
inline std::vector<float> hannWindowPeriodic(int size) {  // Codec/S3Gen/HiFiGAN.swift:15-20
  std::vector<float> w(size);
  b2a_window(B2A_WIN_HANN_PERIODIC, size, w.data());
  return w;
}
inline std::vector<float> melFilters(int sampleRate, int nFft, int nMels, float fMin = 0.0f, float fMax = -1.0f) {  // S3TokenizerUtils.swift:301-375
  std::vector<float> f(size_t(nMels) * (nFft / 2 + 1));
  if (b2a_mel_filters(sampleRate, nFft, nMels, fMin, fMax, f.data()) != B2A_OK) throw Error(B2A_E_BAD_ARG, "bad filterbank parameters");
  return f;
}

}  // namespace b2a
