// B200AudioShim.swift -- Linux-host replacement for the reference's DSP helpers.
//
// SOURCE ONLY: there is no Swift toolchain in the build image, so this file is not compiled or run here;
// its correctness rests on the C ABI (include/b200audio.h), which is what the tests exercise.  It shows
// exactly what a maintainer of smdesai/mlx-swift-audio adds to call the CUDA path from package/STT and
// package/TTS: same function names, labels, defaults and return shapes as the reference (file:line cited
// per function), `fatalError` on a non-zero status as the reference does on bad input.
//
// `Tensor` stands for whatever host-visible float buffer the Linux port uses in place of MLXArray
// (MLX has no CUDA/Linux array type in this package); it only needs `shape` and contiguous fp32 storage.

import CB200Audio
import Foundation

public struct Tensor {
  public var shape: [Int]
  public var data: [Float]
  public init(shape: [Int], data: [Float]) { self.shape = shape; self.data = data }
  public init(zeros shape: [Int]) { self.shape = shape; self.data = [Float](repeating: 0, count: shape.reduce(1, *)) }
}

/// One context per thread / actor (the reference's helpers are re-entrant free functions called from actors).
public final class B200Audio {
  let ctx: OpaquePointer
  public init(device: Int32 = 0) {
    var c: OpaquePointer?
    guard b2a_ctx_create(&c, device) == B2A_OK.rawValue, let cc = c else { fatalError("b200audio: no sm_100 CUDA device") }
    ctx = cc
  }
  deinit { b2a_ctx_destroy(ctx) }

  @inline(__always) func check(_ rc: Int32) {
    if rc != B2A_OK.rawValue { fatalError(String(cString: b2a_last_error(ctx))) }  // reference: fatalError("Input is too short for STFT")
  }

  /// STT/Whisper/WhisperAudio.swift:54-67
  public func padOrTrim(_ array: Tensor, length: Int = 480_000) -> Tensor {
    var out = Tensor(zeros: [length])
    array.data.withUnsafeBufferPointer { x in out.data.withUnsafeMutableBufferPointer { o in
      check(b2a_pad_or_trim(ctx, x.baseAddress, 1, Int64(array.shape[0]), Int64(length), o.baseAddress, Int32(B2A_HOST.rawValue)))
    } }
    return out
  }

  /// STT/Whisper/WhisperAudio.swift:78-137 -> (n_frames, n_mels)
  public func whisperLogMelSpectrogram(audio: Tensor, nMels: Int, padding: Int = 0) -> Tensor {
    let frames = Int(b2a_whisper_num_frames(Int64(audio.shape[0]), Int64(padding)))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    var out = Tensor(zeros: [frames, nMels])
    audio.data.withUnsafeBufferPointer { x in out.data.withUnsafeMutableBufferPointer { o in
      check(b2a_whisper_log_mel_spectrogram(ctx, x.baseAddress, 1, Int64(audio.shape[0]), Int32(nMels), Int64(padding), o.baseAddress,
                                            Int32(B2A_HOST.rawValue)))
    } }
    return out
  }

  /// Codec/S3Tokenizer/S3TokenizerUtils.swift:160-208 -> (n_mels, T')
  public func logMelSpectrogramChatterbox(audio: Tensor, nMels: Int = 128, padding: Int = 0) -> Tensor {
    let frames = Int(b2a_whisper_num_frames(Int64(audio.shape[0]), Int64(padding)))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    var out = Tensor(zeros: [nMels, frames])
    audio.data.withUnsafeBufferPointer { x in out.data.withUnsafeMutableBufferPointer { o in
      check(b2a_log_mel_spectrogram_chatterbox(ctx, x.baseAddress, 1, Int64(audio.shape[0]), Int32(nMels), Int64(padding), o.baseAddress,
                                               Int32(B2A_HOST.rawValue)))
    } }
    return out
  }

  /// STT/FunASR/FunASRAudio.swift:197-216 -> (ceil(T'/lfrN), nMels*lfrM)
  public func preprocessAudio(_ audio: Tensor, nMels: Int = 80, lfrM: Int = 7, lfrN: Int = 6, applyNormalization: Bool = true) -> Tensor {
    let frames = b2a_funasr_num_frames(Int64(audio.shape[0]))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    let rows = Int(b2a_lfr_num_rows(frames, Int32(lfrN)))
    var out = Tensor(zeros: [rows, nMels * lfrM])
    audio.data.withUnsafeBufferPointer { x in out.data.withUnsafeMutableBufferPointer { o in
      check(b2a_funasr_preprocess_audio(ctx, x.baseAddress, 1, Int64(audio.shape[0]), Int32(nMels), Int32(lfrM), Int32(lfrN),
                                        applyNormalization ? 1 : 0, o.baseAddress, Int32(B2A_HOST.rawValue)))
    } }
    return out
  }

  /// Codec/S3Gen/CAMPPlus.swift:32-106 -> (T', numMelBins)
  public func kaldiFbankCAMPPlus(audio: Tensor, sampleRate: Int = 16000, numMelBins: Int = 80, frameLength: Float = 25.0,
                                 frameShift: Float = 10.0) -> Tensor {
    let win = Int32(Float(sampleRate) * frameLength / 1000), hop = Int32(Float(sampleRate) * frameShift / 1000)
    let frames = Int(b2a_kaldi_num_frames(Int64(audio.shape[0]), win, hop))
    if frames <= 0 { fatalError("signal shorter than one analysis window") }
    var out = Tensor(zeros: [frames, numMelBins])
    audio.data.withUnsafeBufferPointer { x in out.data.withUnsafeMutableBufferPointer { o in
      check(b2a_kaldi_fbank_campplus(ctx, x.baseAddress, 1, Int64(audio.shape[0]), Int32(sampleRate), Int32(numMelBins), frameLength, frameShift,
                                     0, o.baseAddress, Int32(B2A_HOST.rawValue)))
    } }
    return out
  }

  /// Codec/S3Gen/Mel/S3GenMel.swift:43-102: y (B, T) -> (B, numMels, T')
  public func s3genMelSpectrogram(y: Tensor, nFft: Int = 1920, numMels: Int = 80, samplingRate: Int = 24000, hopSize: Int = 480,
                                  winSize: Int = 1920, fmin: Int = 0, fmax: Int = 8000, center: Bool = false) -> Tensor {
    let was1D = y.shape.count == 1
    let b = was1D ? 1 : y.shape[0], t = y.shape.last!
    let frames = Int(b2a_s3gen_num_frames(Int64(t), Int32(nFft), Int32(hopSize)))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    var out = Tensor(zeros: was1D ? [numMels, frames] : [b, numMels, frames])
    y.data.withUnsafeBufferPointer { x in out.data.withUnsafeMutableBufferPointer { o in
      check(b2a_s3gen_mel_spectrogram(ctx, x.baseAddress, Int64(b), Int64(t), Int32(nFft), Int32(numMels), Int32(samplingRate), Int32(hopSize),
                                      Int32(winSize), Int32(fmin), Int32(fmax), o.baseAddress, Int32(B2A_HOST.rawValue)))
    } }
    return out
  }

  /// Codec/S3Gen/HiFiGAN.swift:298-367: magnitude, phase (B, nFft/2+1, frames) -> (B, (frames-1)*hop)
  public func istftHiFiGAN(magnitude: Tensor, phase: Tensor, nFft: Int, hopLength: Int, window: Tensor) -> Tensor {
    let b = magnitude.shape[0], frames = magnitude.shape[2]
    var out = Tensor(zeros: [b, (frames - 1) * hopLength])
    magnitude.data.withUnsafeBufferPointer { m in phase.data.withUnsafeBufferPointer { p in window.data.withUnsafeBufferPointer { w in
      out.data.withUnsafeMutableBufferPointer { o in
        check(b2a_istft_hifigan(ctx, m.baseAddress, p.baseAddress, Int64(b), Int64(frames), Int32(nFft), Int32(hopLength), w.baseAddress,
                                o.baseAddress, Int32(B2A_HOST.rawValue)))
      } } } }
    return out
  }

  /// Codec/S3Gen/HiFiGAN.swift:257-295: x (B, T) -> (real, imag) each (B, nFft/2+1, frames)
  public func stftHiFiGAN(x: Tensor, nFft: Int, hopLength: Int, window: Tensor) -> (Tensor, Tensor) {
    let b = x.shape[0], t = x.shape[1]
    let frames = Int(b2a_vocoder_stft_num_frames(Int64(t), Int32(nFft), Int32(hopLength)))
    if frames <= 0 { fatalError("Input is too short") }
    var re = Tensor(zeros: [b, nFft / 2 + 1, frames]), im = Tensor(zeros: [b, nFft / 2 + 1, frames])
    x.data.withUnsafeBufferPointer { xp in window.data.withUnsafeBufferPointer { w in
      re.data.withUnsafeMutableBufferPointer { r in im.data.withUnsafeMutableBufferPointer { i in
        check(b2a_stft_hifigan(ctx, xp.baseAddress, Int64(b), Int64(t), Int32(nFft), Int32(hopLength), w.baseAddress, r.baseAddress, i.baseAddress,
                               Int32(B2A_HOST.rawValue)))
      } } } }
    return (re, im)
  }
  /// Tail of HiFTGenerator.decode (Codec/S3Gen/HiFiGAN.swift:577-589) as one kernel: exp / sin split of the convPost output
  /// (B, nFft+2, frames), istftHiFiGAN, clip to +-audioLimit.
  public func hiftHeadIstft(convOut: Tensor, nFft: Int, hopLength: Int, window: Tensor, audioLimit: Float = 0.99) -> Tensor {
    let b = convOut.shape[0], frames = convOut.shape[2]
    var out = Tensor(zeros: [b, (frames - 1) * hopLength])
    convOut.data.withUnsafeBufferPointer { h in window.data.withUnsafeBufferPointer { w in
      out.data.withUnsafeMutableBufferPointer { o in
        check(b2a_hift_head_istft(ctx, h.baseAddress, Int64(b), Int64(frames), Int32(nFft), Int32(hopLength), w.baseAddress, audioLimit,
                                  o.baseAddress, Int32(B2A_HOST.rawValue)))
      } } }
    return out
  }

  /// Seek window of the Whisper decode loop (STT/Whisper/WhisperSTT.swift:171-182,624-635): (T', M) fp32 -> (length, M) Float16
  public func whisperMelSegment(mel: Tensor, seek: Int, contentFrames: Int, length: Int = 3000) -> [Float16] {
    var out = [Float16](repeating: 0, count: length * mel.shape[1])
    var s = Int64(seek), c = Int64(contentFrames)
    mel.data.withUnsafeBufferPointer { m in out.withUnsafeMutableBytes { o in
      check(b2a_whisper_mel_segment_f16(ctx, m.baseAddress, 1, Int64(mel.shape[0]), Int32(mel.shape[1]), &s, &c, Int64(length),
                                        o.baseAddress, Int32(B2A_HOST.rawValue)))
    } }
    return out
  }
  /// Ragged batch: clips of different lengths in one launch (include/b200audio.h, "ragged batches").  `audio` is (B, T_max),
  /// `lengths[b]` the valid samples of row b; every clip comes out as whisperLogMelSpectrogram(audio[b, ..<lengths[b]]) would,
  /// rows past a clip's frame count are zero.  Replaces the per-clip Swift loops of the callers (e.g. S3Tokenizer.swift:474-571).
  public func whisperLogMelSpectrogramRagged(audio: Tensor, lengths: [Int], nMels: Int, padding: Int = 0) -> (Tensor, [Int]) {
    let b = audio.shape[0], n = audio.shape[1]
    let frames = Int(b2a_whisper_num_frames(Int64(n), Int64(padding)))
    if frames <= 0 || lengths.count != b { fatalError("Input is too short for STFT") }
    var out = Tensor(zeros: [b, frames, nMels])
    var rows = [Int64](repeating: 0, count: b)
    let len64 = lengths.map { Int64($0) }
    audio.data.withUnsafeBufferPointer { x in out.data.withUnsafeMutableBufferPointer { o in
      check(b2a_whisper_log_mel_spectrogram_ragged(ctx, x.baseAddress, Int64(b), Int64(n), len64, Int32(nMels), Int64(padding), o.baseAddress,
                                                   &rows, Int32(B2A_HOST.rawValue)))
    } }
    return (out, rows.map { Int($0) })
  }
  // logMelSpectrogramChatterboxRagged, preprocessAudioRagged, kaldiFbankCAMPPlusRagged, funASRLogMelSpectrogramRagged,
  // voiceEncoderMelspectrogramRagged and s3genMelSpectrogramRagged bind the other b2a_*_ragged entry points in the same way.
  //
  // Multi-GPU (one process per GPU): the consumer rank calls b2a_device_alloc + b2a_ipc_export and ships the 64-byte handle to the
  // producer processes (any host channel); each producer calls b2a_ipc_open and passes `peer + firstClip * clipBytes` as the `out`
  // pointer of a B2A_DEVICE call -- the kernel's stores then land in the consumer's HBM over NVLink (INTEGRATION.md section 6).

  // ---- helpers shared by the bindings below -------------------------------------------------------------------------------
  @inline(__always) func unary(_ x: Tensor, _ outShape: [Int], _ body: (UnsafePointer<Float>?, UnsafeMutablePointer<Float>?) -> Int32) -> Tensor {
    var out = Tensor(zeros: outShape)
    x.data.withUnsafeBufferPointer { xp in out.data.withUnsafeMutableBufferPointer { op in check(body(xp.baseAddress, op.baseAddress)) } }
    return out
  }
  @inline(__always) static func table(_ n: Int, _ body: (UnsafeMutablePointer<Float>?) -> Int32) -> Tensor {
    var out = Tensor(zeros: [n])
    out.data.withUnsafeMutableBufferPointer { op in if body(op.baseAddress) != B2A_OK.rawValue { fatalError("b200audio: bad table parameters") } }
    return out
  }
  var host: Int32 { Int32(B2A_HOST.rawValue) }

  // ---- window generators (host-side tables) -------------------------------------------------------------------------------
  /// STT/Whisper/WhisperAudio.swift:32-44
  public static func whisperHannWindow(length: Int) -> Tensor { table(length) { b2a_window(Int32(B2A_WIN_WHISPER_HANN.rawValue), Int32(length), $0) } }
  /// Codec/S3Tokenizer/S3TokenizerUtils.swift:213-221 (== Kokoro `hanning`, TTS/Kokoro/Decoder/MLXSTFT.swift:12-20)
  public static func hanningWindow(length: Int) -> Tensor { table(length) { b2a_window(Int32(B2A_WIN_HANNING.rawValue), Int32(length), $0) } }
  /// STT/FunASR/FunASRAudio.swift:35-45
  public static func hammingWindow(length: Int) -> Tensor { table(length) { b2a_window(Int32(B2A_WIN_HAMMING.rawValue), Int32(length), $0) } }
  /// Codec/S3Gen/CAMPPlus.swift:15-19
  public static func poveyWindow(size: Int) -> Tensor { table(size) { b2a_window(Int32(B2A_WIN_POVEY.rawValue), Int32(size), $0) } }
  /// Codec/S3Gen/HiFiGAN.swift:15-20 and cosyVoice3HannWindowPeriodic, TTS/CosyVoice3/HiFiGAN/CausalHiFTGenerator.swift:429-432
  public static func hannWindowPeriodic(size: Int) -> Tensor { table(size) { b2a_window(Int32(B2A_WIN_HANN_PERIODIC.rawValue), Int32(size), $0) } }
  public static func cosyVoice3HannWindowPeriodic(size: Int) -> Tensor { hannWindowPeriodic(size: size) }

  // ---- filterbanks and integer rules (host-side) ---------------------------------------------------------------------------
  /// Codec/S3Tokenizer/S3TokenizerUtils.swift:301-375 -> (nMels, nFft/2+1)
  public static func melFilters(sampleRate: Int, nFft: Int, nMels: Int, fMin: Float = 0, fMax: Float? = nil) -> Tensor {
    var t = table(nMels * (nFft / 2 + 1)) { b2a_mel_filters(Int32(sampleRate), Int32(nFft), Int32(nMels), fMin, fMax ?? -1, $0) }
    t.shape = [nMels, nFft / 2 + 1]
    return t
  }
  /// STT/FunASR/FunASRAudio.swift:322-396 -> (nMels, nFft/2)
  public static func funASRMelFilters(sampleRate: Int = 16000, nFft: Int = 400, nMels: Int = 80) -> Tensor {
    var t = table(nMels * (nFft / 2)) { b2a_funasr_mel_filters(Int32(sampleRate), Int32(nFft), Int32(nMels), $0) }
    t.shape = [nMels, nFft / 2]
    return t
  }
  /// computeMelFiltersHTK, Codec/S3Gen/CAMPPlus.swift:134-175 -> (nFft/2+1, nMels)
  public static func computeMelFiltersHTK(sampleRate: Int, nFft: Int, nMels: Int, fMin: Float, fMax: Float) -> Tensor {
    var t = table((nFft / 2 + 1) * nMels) { b2a_mel_filters_htk(Int32(sampleRate), Int32(nFft), Int32(nMels), fMin, fMax, $0) }
    t.shape = [nFft / 2 + 1, nMels]
    return t
  }
  /// Codec/S3Gen/CAMPPlus.swift:22-29
  public static func nextPowerOf2(_ n: Int) -> Int { Int(b2a_next_power_of_2(Int32(n))) }
  /// STT/FunASR/FunASRAudio.swift:225-235
  public static func computeFeatureLength(audioLength: Int, hopLength: Int = 160, lfrN: Int = 6) -> Int {
    Int(b2a_funasr_compute_feature_length(Int64(audioLength), Int32(hopLength), Int32(lfrN)))
  }
  /// Codec/S3Tokenizer/S3TokenizerUtils.swift:71-88
  public static func mergeTokenizedSegments(_ tokenizedSegments: [[Int]], overlap: Int, tokenRate: Int) -> [Int] {
    let flat = tokenizedSegments.flatMap { $0.map { Int32($0) } }
    let lens = tokenizedSegments.map { Int64($0.count) }
    var out = [Int32](repeating: 0, count: max(1, flat.count))
    let n = b2a_merge_tokenized_segments(flat, lens, Int64(lens.count), Int32(overlap), Int32(tokenRate), &out, Int64(out.count))
    if n < 0 { fatalError("mergeTokenizedSegments: bad segment lengths") }
    return out[0 ..< Int(n)].map { Int($0) }
  }

  // ---- padding / plain STFT -------------------------------------------------------------------------------------------------
  /// reflectPad, Codec/S3Tokenizer/S3TokenizerUtils.swift:266-298 (== reflectPad1D, STT/FunASR/FunASRAudio.swift:280-310)
  public func reflectPad(_ x: Tensor, padding: Int) -> Tensor {
    unary(x, [x.shape[0] + 2 * padding]) { b2a_reflect_pad(ctx, $0, 1, Int64(x.shape[0]), Int64(padding), $1, host) }
  }
  public func reflectPad1D(_ x: Tensor, padding: Int) -> Tensor { reflectPad(x, padding: padding) }
  /// stft, Codec/S3Tokenizer/S3TokenizerUtils.swift:224-263 (== funASRSTFT, STT/FunASR/FunASRAudio.swift:240-277)
  /// -> complex (T', nFft/2+1) as (T', nFft/2+1, 2) floats.  `winLength` and `padMode` are ignored, as in the reference (:229,231).
  public func stft(_ x: Tensor, window: Tensor, nFft: Int, hopLength: Int, winLength: Int? = nil, center: Bool = true, padMode: String = "reflect") -> Tensor {
    let frames = Int(b2a_stft_num_frames(Int64(x.shape[0]), Int32(nFft), Int32(hopLength), center ? 1 : 0))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    return window.data.withUnsafeBufferPointer { w in
      unary(x, [frames, nFft / 2 + 1, 2]) { b2a_stft(ctx, $0, 1, Int64(x.shape[0]), w.baseAddress, Int32(window.shape[0]), Int32(nFft), Int32(hopLength), center ? 1 : 0, $1, host) }
    }
  }

  // ---- Fun-ASR ---------------------------------------------------------------------------------------------------------------
  /// STT/FunASR/FunASRAudio.swift:57-94 -> (T', nMels)
  public func funASRLogMelSpectrogram(audio: Tensor, nMels: Int = 80, nFft: Int = 400, hopLength: Int = 160) -> Tensor {
    precondition(nFft == 400 && hopLength == 160, "Fun-ASR front end: n_fft 400 / hop 160")
    let frames = Int(b2a_funasr_num_frames(Int64(audio.shape[0])))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    return unary(audio, [frames, nMels]) { b2a_funasr_log_mel_spectrogram(ctx, $0, 1, Int64(audio.shape[0]), Int32(nMels), $1, host) }
  }
  /// STT/FunASR/FunASRAudio.swift:108-154: (T, M) -> (ceil(T / lfrN), M * lfrM)
  public func applyLFR(_ features: Tensor, lfrM: Int = 7, lfrN: Int = 6) -> Tensor {
    let t = features.shape[0], m = features.shape[1]
    let rows = Int(b2a_lfr_num_rows(Int64(t), Int32(lfrN)))
    return unary(features, [rows, m * lfrM]) { b2a_apply_lfr(ctx, $0, 1, Int64(t), Int32(m), Int32(lfrM), Int32(lfrN), $1, host) }
  }
  /// STT/FunASR/FunASRAudio.swift:165-180: (x + mean) * istd with statistics, per-utterance (x - mean) / (std + 1e-6) without
  public func applyCMVN(_ features: Tensor, cmvnMean: Tensor? = nil, cmvnIstd: Tensor? = nil) -> Tensor {
    let t = features.shape[0], d = features.shape[1]
    guard let mean = cmvnMean, let istd = cmvnIstd else {
      return unary(features, [t, d]) { b2a_apply_cmvn(ctx, $0, 1, Int64(t), Int32(d), nil, nil, $1, host) }
    }
    return mean.data.withUnsafeBufferPointer { mp in istd.data.withUnsafeBufferPointer { ip in
      unary(features, [t, d]) { b2a_apply_cmvn(ctx, $0, 1, Int64(t), Int32(d), mp.baseAddress, ip.baseAddress, $1, host) }
    } }
  }

  // ---- Chatterbox voice encoder / CAM++ wrappers ---------------------------------------------------------------------------------
  /// Config/ChatterboxConfig.swift:139-156 (the fields the mel front end reads)
  public struct VoiceEncConfig {
    public var numMels = 40, sampleRate = 16000, nFft = 400, hopSize = 160, winSize = 400, fmin = 0, fmax = 8000
    public var melPower: Float = 2.0, melType = "amp", normalizedMels = false, stftMagnitudeMin: Float = 1e-4
    public init() {}
  }
  /// TTS/Chatterbox/VoiceEncoder/VoiceEncoderMelspec.swift:17-68 -> (numMels, T')
  public func voiceEncoderMelspectrogram(wav: Tensor, config: VoiceEncConfig, pad: Bool = true) -> Tensor {
    var c = b2a_voice_enc_config()
    b2a_voice_enc_config_default(&c)
    c.num_mels = Int32(config.numMels); c.sample_rate = Int32(config.sampleRate); c.n_fft = Int32(config.nFft)
    c.hop_size = Int32(config.hopSize); c.win_size = Int32(config.winSize); c.fmin = Int32(config.fmin); c.fmax = Int32(config.fmax)
    c.mel_power = config.melPower; c.mel_type_db = config.melType == "db" ? 1 : 0
    c.normalized_mels = config.normalizedMels ? 1 : 0; c.stft_magnitude_min = config.stftMagnitudeMin
    let frames = Int(b2a_stft_num_frames(Int64(wav.shape[0]), c.n_fft, c.hop_size, 1))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    return unary(wav, [config.numMels, frames]) { b2a_voice_encoder_melspectrogram(ctx, $0, 1, Int64(wav.shape[0]), &c, $1, host) }
  }
  /// TTS/CosyVoice2/CosyVoice2TTS.swift:787-795
  public func logMelSpectrogramCAMPPlus(audio: Tensor, sampleRate: Int = 16000, numMelBins: Int = 128) -> Tensor {
    logMelSpectrogramChatterbox(audio: audio, nMels: numMelBins, padding: 0)
  }
  /// kaldiFbankCAMPPlus followed by the caller's time-mean removal (Codec/S3Gen/CAMPPlus.swift:797-802) in one call
  public func kaldiFbankCAMPPlusMeanNorm(audio: Tensor, sampleRate: Int = 16000, numMelBins: Int = 80, frameLength: Float = 25.0,
                                         frameShift: Float = 10.0) -> Tensor {
    let win = Int32(Float(sampleRate) * frameLength / 1000), hop = Int32(Float(sampleRate) * frameShift / 1000)
    let frames = Int(b2a_kaldi_num_frames(Int64(audio.shape[0]), win, hop))
    if frames <= 0 { fatalError("signal shorter than one analysis window") }
    return unary(audio, [frames, numMelBins]) {
      b2a_kaldi_fbank_campplus(ctx, $0, 1, Int64(audio.shape[0]), Int32(sampleRate), Int32(numMelBins), frameLength, frameShift, 1, $1, host)
    }
  }

  // ---- Whisper: the forms the encoder consumes ---------------------------------------------------------------------------------
  /// whisperLogMelSpectrogram(...).asType(.float16) in one kernel (STT/Whisper/WhisperSTT.swift:156-157,181-182)
  public func whisperLogMelSpectrogramF16(audio: Tensor, nMels: Int, padding: Int = 0) -> [Float16] {
    let frames = Int(b2a_whisper_num_frames(Int64(audio.shape[0]), Int64(padding)))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    var out = [Float16](repeating: 0, count: frames * nMels)
    audio.data.withUnsafeBufferPointer { x in out.withUnsafeMutableBytes { o in
      check(b2a_whisper_log_mel_spectrogram_f16(ctx, x.baseAddress, 1, Int64(audio.shape[0]), Int32(nMels), Int64(padding), o.baseAddress, host))
    } }
    return out
  }
  /// 16-bit PCM in (sample = int16 / 32768, what AVAudioFile hands WhisperEngine.swift:327-369 for a 16-bit file), fp16 features out
  public func whisperLogMelSpectrogramPCM16(audio: [Int16], nMels: Int, padding: Int = 0) -> [Float16] {
    let frames = Int(b2a_whisper_num_frames(Int64(audio.count), Int64(padding)))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    var out = [Float16](repeating: 0, count: frames * nMels)
    audio.withUnsafeBufferPointer { x in out.withUnsafeMutableBytes { o in
      check(b2a_whisper_log_mel_spectrogram_pcm16(ctx, x.baseAddress, 1, Int64(audio.count), Int32(nMels), Int64(padding), 1, o.baseAddress, host))
    } }
    return out
  }

  // ---- 16-bit PCM in (sample = int16 / 32768, as AVAudioFile decodes a 16-bit file): half the bytes for a host caller, the same features --
  /// preprocessAudio (STT/FunASR/FunASRAudio.swift:197-216) of 16-bit PCM
  public func preprocessAudioPCM16(_ audio: [Int16], nMels: Int = 80, lfrM: Int = 7, lfrN: Int = 6, applyNormalization: Bool = true) -> Tensor {
    let frames = b2a_funasr_num_frames(Int64(audio.count))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    var out = Tensor(zeros: [Int(b2a_lfr_num_rows(frames, Int32(lfrN))), nMels * lfrM])
    audio.withUnsafeBufferPointer { x in out.data.withUnsafeMutableBufferPointer { o in
      check(b2a_funasr_preprocess_audio_pcm16(ctx, x.baseAddress, 1, Int64(audio.count), Int32(nMels), Int32(lfrM), Int32(lfrN), applyNormalization ? 1 : 0, o.baseAddress, host))
    } }
    return out
  }
  /// kaldiFbankCAMPPlus (Codec/S3Gen/CAMPPlus.swift:32-106) of 16-bit PCM; meanNorm adds CAMPPlus.swift:797-802
  public func kaldiFbankCAMPPlusPCM16(audio: [Int16], sampleRate: Int = 16000, numMelBins: Int = 80, frameLength: Float = 25.0,
                                      frameShift: Float = 10.0, meanNorm: Bool = false) -> Tensor {
    let win = Int32(Float(sampleRate) * frameLength / 1000), hop = Int32(Float(sampleRate) * frameShift / 1000)
    let frames = Int(b2a_kaldi_num_frames(Int64(audio.count), win, hop))
    if frames <= 0 { fatalError("signal shorter than one analysis window") }
    var out = Tensor(zeros: [frames, numMelBins])
    audio.withUnsafeBufferPointer { x in out.data.withUnsafeMutableBufferPointer { o in
      check(b2a_kaldi_fbank_campplus_pcm16(ctx, x.baseAddress, 1, Int64(audio.count), Int32(sampleRate), Int32(numMelBins), frameLength, frameShift, meanNorm ? 1 : 0, o.baseAddress, host))
    } }
    return out
  }
  /// s3genMelSpectrogram (Codec/S3Gen/Mel/S3GenMel.swift:43-102) of one 16-bit PCM clip -> (numMels, T')
  public func s3genMelSpectrogramPCM16(y: [Int16], nFft: Int = 1920, numMels: Int = 80, samplingRate: Int = 24000, hopSize: Int = 480,
                                       winSize: Int = 1920, fmin: Int = 0, fmax: Int = 8000) -> Tensor {
    let frames = Int(b2a_s3gen_num_frames(Int64(y.count), Int32(nFft), Int32(hopSize)))
    if frames <= 0 { fatalError("Input is too short for STFT") }
    var out = Tensor(zeros: [numMels, frames])
    y.withUnsafeBufferPointer { x in out.data.withUnsafeMutableBufferPointer { o in
      check(b2a_s3gen_mel_spectrogram_pcm16(ctx, x.baseAddress, 1, Int64(y.count), Int32(nFft), Int32(numMels), Int32(samplingRate), Int32(hopSize), Int32(winSize), Int32(fmin), Int32(fmax), o.baseAddress, host))
    } }
    return out
  }

  // ---- CosyVoice3 / Kokoro vocoder transforms -----------------------------------------------------------------------------------
  /// TTS/CosyVoice3/HiFiGAN/CausalHiFTGenerator.swift:435-460: x (B, T) -> (real, imag) each (B, nFft/2+1, frames), zero padding
  public func cosyVoice3Stft(x: Tensor, nFft: Int, hopLength: Int, window: Tensor) -> (Tensor, Tensor) {
    let b = x.shape[0], t = x.shape[1]
    let frames = Int(b2a_vocoder_stft_num_frames(Int64(t), Int32(nFft), Int32(hopLength)))
    if frames <= 0 { fatalError("Input is too short") }
    var re = Tensor(zeros: [b, nFft / 2 + 1, frames]), im = Tensor(zeros: [b, nFft / 2 + 1, frames])
    x.data.withUnsafeBufferPointer { xp in window.data.withUnsafeBufferPointer { w in
      re.data.withUnsafeMutableBufferPointer { r in im.data.withUnsafeMutableBufferPointer { i in
        check(b2a_cosyvoice3_stft(ctx, xp.baseAddress, Int64(b), Int64(t), Int32(nFft), Int32(hopLength), w.baseAddress, r.baseAddress, i.baseAddress, host))
      } } } }
    return (re, im)
  }
  /// TTS/CosyVoice3/HiFiGAN/CausalHiFTGenerator.swift:463-514: magnitude clipped to [0, 100]
  public func cosyVoice3Istft(magnitude: Tensor, phase: Tensor, nFft: Int, hopLength: Int, window: Tensor) -> Tensor {
    let b = magnitude.shape[0], frames = magnitude.shape[2]
    var out = Tensor(zeros: [b, (frames - 1) * hopLength])
    magnitude.data.withUnsafeBufferPointer { m in phase.data.withUnsafeBufferPointer { p in window.data.withUnsafeBufferPointer { w in
      out.data.withUnsafeMutableBufferPointer { o in
        check(b2a_cosyvoice3_istft(ctx, m.baseAddress, p.baseAddress, Int64(b), Int64(frames), Int32(nFft), Int32(hopLength), w.baseAddress, o.baseAddress, host))
      } } } }
    return out
  }
  /// unwrap, TTS/Kokoro/Decoder/MLXSTFT.swift:23-46: (rows, frames) along the last axis
  public func unwrap(_ p: Tensor) -> Tensor {
    let frames = p.shape.last!, rows = p.data.count / max(1, frames)
    return unary(p, p.shape) { b2a_unwrap(ctx, $0, Int64(rows), Int64(frames), $1, host) }
  }
  /// mlxStft, TTS/Kokoro/Decoder/MLXSTFT.swift:69-113 in the package's configuration (periodic Hann of nFft taps, centre reflect
  /// padding): x (T,) -> complex (F, frames) as (real, imag)
  public func mlxStft(x: Tensor, nFft: Int = 20, hopLength: Int = 5) -> (Tensor, Tensor) {
    var w = B200Audio.hanningWindow(length: nFft + 1)
    w.data.removeLast(); w.shape = [nFft]
    let (re, im) = stftHiFiGAN(x: Tensor(shape: [1, x.shape[0]], data: x.data), nFft: nFft, hopLength: hopLength, window: w)
    return (Tensor(shape: Array(re.shape.dropFirst()), data: re.data), Tensor(shape: Array(im.shape.dropFirst()), data: im.data))
  }
  /// mlxIstft, TTS/Kokoro/Decoder/MLXSTFT.swift:115-163: complex (F, frames) as (F, frames, 2) floats -> ((frames - 1) * hop,)
  public func mlxIstft(x: Tensor, hopLength: Int? = nil, winLength: Int? = nil) -> Tensor {
    let f = x.shape[0], frames = x.shape[1]
    let win = winLength ?? (f - 1) * 2, hop = hopLength ?? win / 4
    return unary(x, [(frames - 1) * hop]) { b2a_mlx_istft(ctx, $0, 1, Int64(frames), Int32(win), Int32(hop), $1, host) }
  }
  /// Kokoro head (TTS/Kokoro/Decoder/Generator.swift:182-190): exp / sin split of conv_post's output + MLXSTFT.inverse
  public func kokoroHeadIstft(convOut: Tensor, filterLength: Int = 20, hopLength: Int = 5, winLength: Int = 20) -> Tensor {
    let b = convOut.shape[0], frames = convOut.shape[2]
    return unary(convOut, [b, 1, (frames - 1) * hopLength]) {
      b2a_kokoro_head_istft(ctx, $0, Int64(b), Int64(frames), Int32(filterLength), Int32(hopLength), Int32(winLength), $1, host)
    }
  }
  /// hiftHeadIstft followed by `result[0..., 0 ..< fadeLen] *= trimFade` of S3Token2Wav (Codec/S3Gen/S3Gen.swift:284-289), one kernel
  public func hiftHeadIstftFade(convOut: Tensor, nFft: Int, hopLength: Int, window: Tensor, trimFade: Tensor, audioLimit: Float = 0.99) -> Tensor {
    let b = convOut.shape[0], frames = convOut.shape[2]
    return window.data.withUnsafeBufferPointer { w in trimFade.data.withUnsafeBufferPointer { fd in
      unary(convOut, [b, (frames - 1) * hopLength]) {
        b2a_hift_head_istft_fade(ctx, $0, Int64(b), Int64(frames), Int32(nFft), Int32(hopLength), w.baseAddress, audioLimit, fd.baseAddress, Int64(trimFade.data.count), $1, host)
      }
    } }
  }
  /// `_trimFade` of S3Token2Wav (Codec/S3Gen/S3Gen.swift:259-262)
  public static func s3genTrimFade(samplingRate: Int = 24000) -> Tensor {
    table(2 * (samplingRate / 50)) { b2a_s3gen_trim_fade(Int32(samplingRate), $0) }
  }

  // ---- resampling ------------------------------------------------------------------------------------------------------------
  /// resampleAudio / linearInterpolate1d (TTS/CosyVoice2/CosyVoice2TTS.swift:733-744, CosyHiFTGenerator.swift:17-58), bit-exact
  public func resampleAudio(_ audio: Tensor, fromRate: Int, toRate: Int) -> Tensor {
    let n = Int(b2a_resample_linear_length(Int64(audio.shape[0]), Int32(fromRate), Int32(toRate)))
    return unary(audio, [n]) { b2a_resample_linear(ctx, $0, 1, Int64(audio.shape[0]), Int32(fromRate), Int32(toRate), $1, host) }
  }
  /// Stand-in for AudioResampler.resample (Audio/AudioResampler.swift:15-88: AVAudioConverter is not available on Linux): the
  /// polyphase design of scipy.signal.resample_poly -- a non-parity extension (include/b200audio.h)
  public func resample(_ audio: Tensor, from sourceSampleRate: Int, to targetSampleRate: Int) -> Tensor {
    let n = Int(b2a_resample_poly_length(Int64(audio.shape[0]), Int32(sourceSampleRate), Int32(targetSampleRate)))
    return unary(audio, [n]) { b2a_resample_poly(ctx, $0, 1, Int64(audio.shape[0]), Int32(sourceSampleRate), Int32(targetSampleRate), $1, host) }
  }
}

/// TTS/Kokoro/Decoder/MLXSTFT.swift:165-235
public final class MLXSTFT {
  let dsp: B200Audio
  public let filterLength: Int, hopLength: Int, winLength: Int
  public init(filterLength: Int = 800, hopLength: Int = 200, winLength: Int = 800, window: String = "hann", dsp: B200Audio) {
    if window.lowercased() != "hann" { fatalError("Only hanning is supported for window, not \(window)") }   // MLXSTFT.swift:54
    self.filterLength = filterLength; self.hopLength = hopLength; self.winLength = winLength; self.dsp = dsp
  }
  /// inputData (B, T) -> (magnitude, phase) each (B, F, frames)
  public func transform(inputData: Tensor) -> (Tensor, Tensor) {
    let b = inputData.shape[0], t = inputData.shape[1]
    let frames = Int(b2a_vocoder_stft_num_frames(Int64(t), Int32(filterLength), Int32(hopLength)))
    if frames <= 0 { fatalError("Input is too short") }
    let f = filterLength / 2 + 1
    var mag = Tensor(zeros: [b, f, frames]), ph = Tensor(zeros: [b, f, frames])
    inputData.data.withUnsafeBufferPointer { x in mag.data.withUnsafeMutableBufferPointer { m in ph.data.withUnsafeMutableBufferPointer { p in
      dsp.check(b2a_kokoro_stft_transform(dsp.ctx, x.baseAddress, Int64(b), Int64(t), Int32(filterLength), Int32(hopLength), Int32(winLength), m.baseAddress, p.baseAddress, dsp.host))
    } } }
    return (mag, ph)
  }
  /// magnitude, phase (B, F, frames) -> (B, 1, (frames - 1) * hop); unwrap, irfft, window, overlap-add, window-sum normalisation
  public func inverse(magnitude: Tensor, phase: Tensor) -> Tensor {
    let b = magnitude.shape[0], frames = magnitude.shape[2]
    var out = Tensor(zeros: [b, 1, (frames - 1) * hopLength])
    magnitude.data.withUnsafeBufferPointer { m in phase.data.withUnsafeBufferPointer { p in out.data.withUnsafeMutableBufferPointer { o in
      dsp.check(b2a_kokoro_stft_inverse(dsp.ctx, m.baseAddress, p.baseAddress, Int64(b), Int64(frames), Int32(filterLength), Int32(hopLength), Int32(winLength), o.baseAddress, dsp.host))
    } } }
    return out
  }
  public func callAsFunction(_ inputData: Tensor) -> Tensor {
    let (m, p) = transform(inputData: inputData)
    return inverse(magnitude: m, phase: p)
  }
}
