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
  // logMelSpectrogramChatterboxRagged, preprocessAudioRagged, kaldiFbankCAMPPlusRagged and s3genMelSpectrogramRagged bind the other
  // b2a_*_ragged entry points in the same way.
  //
  // Multi-GPU (one process per GPU): the consumer rank calls b2a_device_alloc + b2a_ipc_export and ships the 64-byte handle to the
  // producer processes (any host channel); each producer calls b2a_ipc_open and passes `peer + firstClip * clipBytes` as the `out`
  // pointer of a B2A_DEVICE call -- the kernel's stores then land in the consumer's HBM over NVLink (INTEGRATION.md section 6).

  // hiftHeadIstftFade(convOut:..., trimFade:) binds b2a_hift_head_istft_fade (the head plus `result[0..., 0 ..< fadeLen] *= trimFade`
  // of S3Token2Wav.callAsFunction, S3Gen.swift:284-289); b2a_s3gen_trim_fade builds `_trimFade` (S3Gen.swift:259-262).
  // kokoroHeadIstft binds b2a_kokoro_head_istft like hiftHeadIstft.
  // cosyVoice3Stft / cosyVoice3Istft, MLXSTFT.transform / .inverse, funASRLogMelSpectrogram, applyLFR, applyCMVN,
  // voiceEncoderMelspectrogram and stft bind b2a_cosyvoice3_*, b2a_kokoro_stft_*, b2a_funasr_log_mel_spectrogram,
  // b2a_apply_lfr, b2a_apply_cmvn, b2a_voice_encoder_melspectrogram and b2a_stft in exactly the same way.
}
