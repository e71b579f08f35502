"""CPU restatement (NumPy/SciPy) of the reference's STFT-family DSP helpers.

TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py.  PARITY UNPINNED (no golden vectors
exist in the reference; the reference cannot run here).

Every function cites the reference file:line it follows (paths relative to the
reference's ``package/`` directory).  The arithmetic lives in MLX (mlx-swift @ b1e80760,
Package.resolved:13-20), which is not vendored; its NumPy-convention ops are restated
with NumPy/SciPy in the same order: ``rfft`` unnormalised, ``irfft``/``ifft`` scaled
1/n, ``irfft`` ignoring the imaginary part of DC/Nyquist, ``variance`` = population
variance, ``padded`` = zero padding.  SciPy's FFT is pocketfft, the family MLX's CPU
backend uses.

All functions take ``dt`` (np.float32 = the reference's precision, np.float64 = error
attribution).  Single-clip signatures mirror the Swift helpers; ``*_batch`` wrappers
just loop (the reference has no batched path either).
"""
from __future__ import annotations

import functools
import math

import numpy as np
import scipy.fft as sfft

F32 = np.float32
F64 = np.float64


def _c(dt):
    return np.complex64 if dt == np.float32 else np.complex128


def _pi(dt):
    return dt(np.pi)


# --------------------------------------------------------------------------------------
# windows (a1-a5)
# --------------------------------------------------------------------------------------

def whisper_hann_window(length: int, dt=F32) -> np.ndarray:
    """STT/Whisper/WhisperAudio.swift:32-44 -- symmetric Hann, 0.5*(1-cos(2*pi*n/(N-1)))."""
    if length == 1:
        return np.ones(1, dt)
    n = np.arange(length, dtype=dt)
    factor = dt(2.0) * _pi(dt) / dt(length - 1)
    return (dt(0.5) * (dt(1.0) - np.cos(n * factor))).astype(dt)


def hanning_window(length: int, dt=F32) -> np.ndarray:
    """Codec/S3Tokenizer/S3TokenizerUtils.swift:213-221 and TTS/Kokoro/Decoder/MLXSTFT.swift:12-20.

    np.hanning(length): 0.5 + 0.5*cos(pi*n/(L-1)), n = 1-L, 3-L, ..., L-1.
    """
    if length == 1:
        return np.ones(1, dt)
    n = np.arange(1 - length, length, 2).astype(dt)
    factor = _pi(dt) / dt(length - 1)
    return (dt(0.5) + dt(0.5) * np.cos(n * factor)).astype(dt)


def hann_periodic_via_hanning(n_fft: int, dt=F32) -> np.ndarray:
    """``hanningWindow(length: nFft + 1)[0 ..< nFft]`` (S3TokenizerUtils.swift:172,
    S3GenMel.swift:66, VoiceEncoderMelspec.swift:23, MLXSTFT.swift:52)."""
    return hanning_window(n_fft + 1, dt)[:n_fft].copy()


def hamming_window(length: int, dt=F32) -> np.ndarray:
    """STT/FunASR/FunASRAudio.swift:35-45 -- symmetric Hamming."""
    if length == 1:
        return np.ones(1, dt)
    n = np.arange(length, dtype=dt)
    factor = dt(2.0) * _pi(dt) / dt(length - 1)
    return (dt(0.54) - dt(0.46) * np.cos(n * factor)).astype(dt)


def povey_window(size: int, dt=F32) -> np.ndarray:
    """Codec/S3Gen/CAMPPlus.swift:15-19 -- (0.5-0.5cos(2*pi*n/(N-1)))**0.85."""
    n = np.arange(size, dtype=dt)
    hann = dt(0.5) - dt(0.5) * np.cos(dt(2.0) * _pi(dt) * n / dt(size - 1))
    return np.power(hann.astype(dt), dt(0.85)).astype(dt)


def hann_window_periodic(size: int, dt=F32) -> np.ndarray:
    """Codec/S3Gen/HiFiGAN.swift:15-20 and CosyVoice3 CausalHiFTGenerator.swift:429-432."""
    n = np.arange(size, dtype=dt)
    return (dt(0.5) * (dt(1.0) - np.cos(dt(2.0) * _pi(dt) * n / dt(size)))).astype(dt)


# --------------------------------------------------------------------------------------
# padding / framing (a6-a9, a29)
# --------------------------------------------------------------------------------------

def pad_or_trim(x: np.ndarray, length: int = 480000) -> np.ndarray:
    """WhisperAudio.swift:54-67."""
    n = x.shape[0]
    if n > length:
        return x[:length]
    if n < length:
        return np.concatenate([x, np.zeros(length - n, x.dtype)])
    return x


def reflect_pad(x: np.ndarray, padding: int) -> np.ndarray:
    """S3TokenizerUtils.swift:266-298 == FunASRAudio.swift:280-310 (incl. the non-standard
    short-input ``while`` loops)."""
    if padding == 0:
        return x
    n = x.shape[0]
    if n == 1:
        return np.concatenate([np.full(padding, x[0], x.dtype), x, np.full(padding, x[0], x.dtype)])
    prefix = x[1:min(padding + 1, n)][::-1]
    suffix = x[max(0, n - padding - 1):n - 1][::-1]
    while prefix.shape[0] < padding:
        additional = min(padding - prefix.shape[0], n - 1)
        prefix = np.concatenate([x[1:additional + 1][::-1], prefix])
    while suffix.shape[0] < padding:
        additional = min(padding - suffix.shape[0], n - 1)
        suffix = np.concatenate([suffix, x[n - additional - 1:n - 1][::-1]])
    return np.concatenate([prefix[:padding], x, suffix[:padding]])


def reflect_pad_index(n: int, padding: int) -> np.ndarray:
    """Source index of every padded position (integer restatement of reflect_pad; used to
    check the device index function bit-exactly)."""
    idx = np.arange(n, dtype=np.int64)
    return reflect_pad(idx, padding)


def stft(x: np.ndarray, window: np.ndarray, n_fft: int, hop: int, center: bool = True, dt=F32) -> np.ndarray:
    """S3TokenizerUtils.swift:224-263 == FunASRAudio.swift:240-277.  Returns complex (T', F).
    Raises ValueError where the reference calls fatalError("Input is too short for STFT")."""
    x = np.asarray(x, dt)
    w = np.asarray(window, dt)
    if w.shape[0] < n_fft:
        w = np.concatenate([w, np.zeros(n_fft - w.shape[0], dt)])
    if center:
        x = reflect_pad(x, n_fft // 2)
    # Swift Int division truncates toward zero, so for -hop < num < 0 the reference gets
    # numFrames == 1 and asStrided reads past the buffer (undefined).  Only reachable with
    # center=false; this restatement (and the CUDA path: B2A_E_TOO_SHORT) rejects num < 0.
    num = x.shape[0] - n_fft
    if num < 0:
        raise ValueError("Input is too short for STFT")
    n_frames = 1 + num // hop
    frames = np.lib.stride_tricks.as_strided(x, (n_frames, n_fft), (hop * x.itemsize, x.itemsize))
    return sfft.rfft((frames * w).astype(dt), axis=-1).astype(_c(dt))


# --------------------------------------------------------------------------------------
# filterbanks (a10-a12)
# --------------------------------------------------------------------------------------

@functools.lru_cache(maxsize=None)
def _mel_filters_cached(sample_rate, n_fft, n_mels, f_min, f_max, dtname):
    dt = np.float32 if dtname == "f32" else np.float64
    lg = (lambda v: dt(math.log(v))) if dt == np.float64 else (lambda v: np.log(dt(v)))
    ex = (lambda v: dt(math.exp(v))) if dt == np.float64 else (lambda v: np.exp(dt(v)))
    f_sp = dt(200.0) / dt(3.0)
    min_log_hz = dt(1000.0)
    min_log_mel = min_log_hz / f_sp
    logstep = lg(6.4) / dt(27.0)

    def hz_to_mel(hz):
        hz = dt(hz)
        if hz >= min_log_hz:
            return dt(min_log_mel + lg(hz / min_log_hz) / logstep)
        return dt(hz / f_sp)

    def mel_to_hz(mel):
        mel = dt(mel)
        if mel >= min_log_mel:
            return dt(min_log_hz * ex(logstep * (mel - min_log_mel)))
        return dt(f_sp * mel)

    actual_fmax = dt(f_max) if f_max is not None else dt(sample_rate) / dt(2.0)
    mel_min = hz_to_mel(f_min)
    mel_max = hz_to_mel(actual_fmax)
    pts = [mel_to_hz(dt(mel_min + dt(dt(i) * dt(mel_max - mel_min)) / dt(n_mels + 1))) for i in range(n_mels + 2)]
    nb = n_fft // 2 + 1
    freqs = [dt(dt(i) * dt(sample_rate)) / dt(n_fft) for i in range(nb)]
    fb = np.zeros((n_mels, nb), dt)
    for m in range(n_mels):
        fl, fc, fr = pts[m], pts[m + 1], pts[m + 2]
        for k in range(nb):
            f = freqs[k]
            if fl <= f <= fc:
                fb[m, k] = dt(f - fl) / dt(fc - fl)
            elif fc < f <= fr:
                fb[m, k] = dt(fr - f) / dt(fr - fc)
        enorm = dt(2.0) / dt(pts[m + 2] - pts[m])
        fb[m, :] *= enorm
    fb.setflags(write=False)
    return fb


def mel_filters(sample_rate: int, n_fft: int, n_mels: int, f_min: float = 0.0, f_max=None, dt=F32) -> np.ndarray:
    """S3TokenizerUtils.swift:301-375 -- Slaney scale + Slaney area norm, scalar loops.  (M, F).

    A degenerate 0/0 (fCenter == fLeft at a bin) would be NaN in Swift as well; it does not
    occur for any configuration the reference uses."""
    with np.errstate(invalid="ignore", divide="ignore"):
        return _mel_filters_cached(sample_rate, n_fft, n_mels, float(f_min), None if f_max is None else float(f_max),
                                   "f32" if dt == np.float32 else "f64")


@functools.lru_cache(maxsize=None)
def _funasr_filters_cached(sample_rate, n_fft, n_mels, dtname):
    dt = np.float32 if dtname == "f32" else np.float64

    def lin(a, b, num):
        # MLX linspace: (1-t)*start + t*stop with t = arange(num)/(num-1), evaluated in fp32
        t = np.arange(num, dtype=dt) / dt(num - 1)
        return ((dt(1.0) - t) * dt(a) + t * dt(b)).astype(dt)

    def hz_to_mel(hz):
        return dt(dt(2595.0) * np.log10(dt(1.0) + dt(hz) / dt(700.0)))

    n_freqs = n_fft // 2
    all_freqs = lin(0.0, dt(sample_rate) / dt(2.0), n_freqs)
    m_min = hz_to_mel(0.0)
    m_max = hz_to_mel(dt(sample_rate) / dt(2.0))
    m_pts = lin(m_min, m_max, n_mels + 2)
    f_pts = (dt(700.0) * (np.power(dt(10.0), m_pts / dt(2595.0)).astype(dt) - dt(1.0))).astype(dt)
    f_diff = f_pts[1:] - f_pts[:-1]
    slopes = f_pts[None, :] - all_freqs[:, None]
    down = -slopes[:, :-2] / f_diff[:-1]
    up = slopes[:, 2:] / f_diff[1:]
    fb = np.maximum(dt(0.0), np.minimum(down, up)).astype(dt)
    enorm = dt(2.0) / (f_pts[2:n_mels + 2] - f_pts[:n_mels])
    fb = (fb * enorm[None, :]).astype(dt)
    out = np.ascontiguousarray(fb.T)
    out.setflags(write=False)
    return out


def funasr_mel_filters(sample_rate: int = 16000, n_fft: int = 400, n_mels: int = 80, dt=F32) -> np.ndarray:
    """FunASRAudio.swift:322-396 (HTK branch).  (M, n_fft/2) -- note the 200-point linspace grid."""
    return _funasr_filters_cached(sample_rate, n_fft, n_mels, "f32" if dt == np.float32 else "f64")


def _c_round(v: float) -> int:
    return int(math.floor(abs(v) + 0.5)) * (1 if v >= 0 else -1)


@functools.lru_cache(maxsize=None)
def _htk_int_filters_cached(sample_rate, n_fft, n_mels, f_min, f_max, dtname):
    dt = np.float32 if dtname == "f32" else np.float64

    def hz_to_mel(hz):
        return dt(dt(2595.0) * np.log10(dt(1.0) + dt(hz) / dt(700.0)))

    def mel_to_hz(mel):
        return dt(dt(700.0) * (np.power(dt(10.0), dt(mel) / dt(2595.0)) - dt(1.0)))

    mel_min, mel_max = hz_to_mel(f_min), hz_to_mel(f_max)
    mel_pts = [dt(mel_min + dt(dt(i) * dt(mel_max - mel_min)) / dt(n_mels + 1)) for i in range(n_mels + 2)]
    hz_pts = [mel_to_hz(m) for m in mel_pts]
    bins = [_c_round(float(dt(dt(h) * dt(n_fft)) / dt(sample_rate))) for h in hz_pts]
    nb = n_fft // 2 + 1
    fb = np.zeros((nb, n_mels), dt)
    for m in range(1, n_mels + 1):
        lo, ce, hi = bins[m - 1], bins[m], bins[m + 1]
        for k in range(lo, ce):
            if 0 <= k < nb and ce != lo:
                fb[k, m - 1] = dt(k - lo) / dt(ce - lo)
        for k in range(ce, hi):
            if 0 <= k < nb and hi != ce:
                fb[k, m - 1] = dt(hi - k) / dt(hi - ce)
    fb.setflags(write=False)
    return fb


def mel_filters_htk(sample_rate: int, n_fft: int, n_mels: int, f_min: float, f_max: float, dt=F32) -> np.ndarray:
    """Codec/S3Gen/CAMPPlus.swift:134-175 -- integer-bin HTK triangles, (F, M), unnormalised."""
    return _htk_int_filters_cached(sample_rate, n_fft, n_mels, float(f_min), float(f_max),
                                   "f32" if dt == np.float32 else "f64")


# --------------------------------------------------------------------------------------
# log-mel front ends (a13-a21)
# --------------------------------------------------------------------------------------

def whisper_log_mel_spectrogram(audio, n_mels: int, padding: int = 0, dt=F32) -> np.ndarray:
    """WhisperAudio.swift:78-137.  (T', M)."""
    x = np.asarray(audio, dt)
    if padding > 0:
        x = np.concatenate([x, np.zeros(padding, dt)])
    spec = stft(x, whisper_hann_window(400, dt), 400, 160, dt=dt)
    freqs = spec[:-1, :]
    mag = np.power(np.abs(freqs).astype(dt), dt(2)).astype(dt)
    filt = mel_filters(16000, 400, n_mels, 0.0, 8000.0, dt)
    mel = np.matmul(mag, filt.T).astype(dt)
    log_spec = np.log10(np.maximum(mel, dt(1e-10))).astype(dt)
    log_spec = np.maximum(log_spec, log_spec.max() - dt(8.0))
    return ((log_spec + dt(4.0)) / dt(4.0)).astype(dt)


def log_mel_spectrogram_chatterbox(audio, n_mels: int = 128, padding: int = 0, dt=F32) -> np.ndarray:
    """S3TokenizerUtils.swift:160-208.  (M, T')."""
    x = np.asarray(audio, dt)
    if padding > 0:
        x = np.concatenate([x, np.zeros(padding, dt)])
    spec = stft(x, hann_periodic_via_hanning(400, dt), 400, 160, dt=dt)[:-1, :]
    mag = np.power(np.abs(spec).astype(dt), dt(2)).astype(dt)
    filt = mel_filters(16000, 400, n_mels, dt=dt)
    mel_t = np.matmul(mag, filt.T).astype(dt).T
    log_spec = np.log10(np.maximum(mel_t, dt(1e-10))).astype(dt)
    log_spec = np.maximum(log_spec, log_spec.max() - dt(8.0))
    return np.ascontiguousarray(((log_spec + dt(4.0)) / dt(4.0)).astype(dt))


def funasr_log_mel_spectrogram(audio, n_mels: int = 80, n_fft: int = 400, hop: int = 160, dt=F32) -> np.ndarray:
    """FunASRAudio.swift:57-94.  (T', M), natural log, Nyquist bin dropped."""
    x = np.asarray(audio, dt)
    spec = stft(x, hamming_window(n_fft, dt), n_fft, hop, dt=dt)
    freqs = spec[:, :n_fft // 2]
    mag = np.power(np.abs(freqs).astype(dt), dt(2)).astype(dt)
    filt = funasr_mel_filters(16000, n_fft, n_mels, dt)
    mel = np.matmul(mag, filt.T).astype(dt)
    return np.log(np.maximum(mel, dt(1e-10))).astype(dt)


def apply_lfr(features: np.ndarray, lfr_m: int = 7, lfr_n: int = 6) -> np.ndarray:
    """FunASRAudio.swift:108-154."""
    t, n_mels = features.shape
    t_lfr = int(math.ceil(t / lfr_n))
    left = (lfr_m - 1) // 2
    padded = features
    if left > 0:
        padded = np.concatenate([np.broadcast_to(features[0:1], (left, n_mels)), padded], axis=0)
    t_padded = padded.shape[0]
    total_needed = (t_lfr - 1) * lfr_n + lfr_m
    if total_needed > t_padded:
        padded = np.concatenate([padded, np.broadcast_to(padded[t_padded - 1:t_padded], (total_needed - t_padded, n_mels))], axis=0)
    idx = (np.arange(t_lfr) * lfr_n)[:, None] + np.arange(lfr_m)[None, :]
    return np.ascontiguousarray(padded[idx].reshape(t_lfr, lfr_m * n_mels))


def apply_cmvn(features: np.ndarray, cmvn_mean=None, cmvn_istd=None) -> np.ndarray:
    """FunASRAudio.swift:165-180."""
    dt = features.dtype.type
    if cmvn_mean is not None and cmvn_istd is not None:
        return ((features + np.asarray(cmvn_mean, dt)) * np.asarray(cmvn_istd, dt)).astype(dt)
    mean = features.mean(axis=0, keepdims=True, dtype=dt)
    std = np.sqrt(features.var(axis=0, keepdims=True, dtype=dt)).astype(dt) + dt(1e-6)
    return ((features - mean) / std).astype(dt)


def preprocess_audio(audio, n_mels: int = 80, lfr_m: int = 7, lfr_n: int = 6, apply_normalization: bool = True, dt=F32):
    """FunASRAudio.swift:197-216."""
    f = funasr_log_mel_spectrogram(audio, n_mels, dt=dt)
    f = apply_lfr(f, lfr_m, lfr_n)
    if apply_normalization:
        f = apply_cmvn(f)
    return f


def compute_feature_length(audio_length: int, hop: int = 160, lfr_n: int = 6) -> int:
    """FunASRAudio.swift:225-235."""
    return (audio_length // hop + lfr_n - 1) // lfr_n


def next_power_of_2(n: int) -> int:
    """CAMPPlus.swift:22-29."""
    if n <= 1:
        return 1
    p = 1
    while p < n:
        p *= 2
    return p


def kaldi_fbank_camp_plus(audio, sample_rate: int = 16000, num_mel_bins: int = 80, frame_length: float = 25.0,
                          frame_shift: float = 10.0, dt=F32) -> np.ndarray:
    """CAMPPlus.swift:32-106.  (T', M).  The caller's mean-normalisation (:800) is
    ``kaldi_fbank_mean_norm``."""
    win_length = int(np.float32(sample_rate) * np.float32(frame_length) / np.float32(1000))
    hop = int(np.float32(sample_rate) * np.float32(frame_shift) / np.float32(1000))
    n_fft = next_power_of_2(win_length)
    x = np.asarray(audio, dt).reshape(-1)
    num = x.shape[0] - win_length
    n_frames = (int(num / hop) if num < 0 else num // hop) + 1
    if n_frames < 1:
        n_frames = 1
    window = povey_window(win_length, dt)
    idx = (np.arange(n_frames) * hop)[:, None] + np.arange(win_length)[None, :]
    frames = x[idx]  # IndexError where MLX.take would read out of bounds (signal shorter than one window)
    frames = (frames - frames.mean(axis=1, keepdims=True, dtype=dt)).astype(dt)
    first = frames[:, 0:1]
    rest = frames[:, 1:] - dt(0.97) * frames[:, :win_length - 1]
    frames = np.concatenate([first, rest], axis=1).astype(dt)
    frames = frames * window
    if win_length < n_fft:
        frames = np.concatenate([frames, np.zeros((n_frames, n_fft - win_length), dt)], axis=1)
    spec = sfft.rfft(frames.astype(dt), axis=1).astype(_c(dt))
    power = np.power(np.abs(spec).astype(dt), dt(2)).astype(dt)
    filt = mel_filters_htk(sample_rate, n_fft, num_mel_bins, 20.0, float(sample_rate) / 2, dt)
    mel = np.matmul(power, filt).astype(dt)
    return np.log(np.maximum(mel, dt(1.1920929e-07))).astype(dt)


def kaldi_fbank_mean_norm(fbank: np.ndarray) -> np.ndarray:
    """CAMPPlus.swift:797-802 (``fbank - mean(fbank, axis=0)``)."""
    dt = fbank.dtype.type
    return (fbank - fbank.mean(axis=0, keepdims=True, dtype=dt)).astype(dt)


def reflect_pad_2d(x: np.ndarray, pad: int) -> np.ndarray:
    """S3GenMel.swift:10-28 (no short-input loop)."""
    if pad == 0:
        return x
    t = x.shape[1]
    prefix = x[:, 1:min(pad + 1, t)][:, ::-1]
    suffix = x[:, max(0, t - pad - 1):t - 1][:, ::-1]
    return np.concatenate([prefix, x, suffix], axis=1)


def s3gen_mel_spectrogram(y, n_fft=1920, num_mels=80, sampling_rate=24000, hop_size=480, win_size=1920,
                          fmin=0, fmax=8000, dt=F32) -> np.ndarray:
    """S3GenMel.swift:43-102.  (B, M, T') (or (M, T') for 1-D input)."""
    y = np.asarray(y, dt)
    was_1d = y.ndim == 1
    if was_1d:
        y = y[None, :]
    y = reflect_pad_2d(y, (n_fft - hop_size) // 2)
    window = hann_periodic_via_hanning(win_size, dt)
    specs = [stft(y[i], window, n_fft, hop_size, center=False, dt=dt) for i in range(y.shape[0])]
    spec = np.stack(specs, axis=0)
    mag = np.abs(spec).astype(dt)
    filt = mel_filters(sampling_rate, n_fft, num_mels, float(fmin), float(fmax), dt)
    mel = np.matmul(mag, filt.T).astype(dt).transpose(0, 2, 1)
    mel = np.log(np.maximum(mel, dt(1e-5))).astype(dt)
    mel = np.ascontiguousarray(mel)
    return mel[0] if was_1d else mel


def voice_encoder_melspectrogram(wav, num_mels=40, sample_rate=16000, n_fft=400, hop_size=160, win_size=400,
                                 fmin=0, fmax=8000, mel_power=2.0, mel_type="amp", normalized_mels=False,
                                 stft_magnitude_min=1e-4, dt=F32) -> np.ndarray:
    """VoiceEncoderMelspec.swift:17-68 with defaults from Config/ChatterboxConfig.swift:139-156.  (M, T')."""
    x = np.asarray(wav, dt)
    spec = stft(x, hann_periodic_via_hanning(win_size, dt), n_fft, hop_size, dt=dt)
    mag = np.abs(spec).astype(dt)
    if mel_power != 1.0:
        mag = np.power(mag, dt(mel_power)).astype(dt)
    filt = mel_filters(sample_rate, n_fft, num_mels, float(fmin), float(fmax), dt)
    mel = np.matmul(mag, filt.T).astype(dt).T
    if mel_type == "db":
        mel = (dt(20) * np.log10(np.maximum(mel, dt(stft_magnitude_min)))).astype(dt)
    if normalized_mels:
        min_level_db = dt(20) * np.log10(dt(stft_magnitude_min))
        mel = ((mel - min_level_db) / (-min_level_db + dt(15))).astype(dt)
    return np.ascontiguousarray(mel)


# --------------------------------------------------------------------------------------
# vocoder STFT / iSTFT (a22-a28)
# --------------------------------------------------------------------------------------

def stft_hifigan(x, n_fft: int, hop: int, window, dt=F32):
    """Codec/S3Gen/HiFiGAN.swift:257-295.  x (B, T) -> (real, imag) each (B, F, frames)."""
    x = np.asarray(x, dt)
    w = np.asarray(window, dt)
    b, t = x.shape
    p = n_fft // 2
    left = x[:, 1:p + 1][:, ::-1]
    right = x[:, t - p - 1:t - 1][:, ::-1]
    xp = np.concatenate([left, x, right], axis=1)
    n_frames = (xp.shape[1] - n_fft) // hop + 1
    idx = (np.arange(n_frames) * hop)[:, None] + np.arange(n_fft)[None, :]
    frames = xp[:, idx].transpose(0, 2, 1)  # (B, n_fft, frames)
    frames = (frames * w.reshape(1, -1, 1)).astype(dt)
    spec = sfft.fft(frames.astype(_c(dt)), axis=1)[:, :n_fft // 2 + 1, :]
    return np.ascontiguousarray(spec.real.astype(dt)), np.ascontiguousarray(spec.imag.astype(dt))


def _ola(frames: np.ndarray, hop: int) -> np.ndarray:
    """frames (B, n_frames, n_fft) -> (B, L) scatter-add in frame order (deterministic)."""
    b, nf, n = frames.shape
    out = np.zeros((b, (nf - 1) * hop + n), frames.dtype)
    for j in range(0, n, hop):
        # all frames contribute samples [j, j+hop) of themselves to disjoint output ranges
        seg = frames[:, :, j:j + hop].reshape(b, -1)
        out[:, j:j + nf * hop] += seg
    return out


def _ola_1d(w: np.ndarray, n_frames: int, hop: int) -> np.ndarray:
    return _ola(np.broadcast_to(w, (1, n_frames, w.shape[0])).copy(), hop)[0]


def istft_hifigan(magnitude, phase, n_fft: int, hop: int, window, dt=F32) -> np.ndarray:
    """Codec/S3Gen/HiFiGAN.swift:298-367.  (B, F, frames) x2 -> (B, (frames-1)*hop).
    Requires n_fft % hop == 0 in this restatement (16/4 in the reference)."""
    mag = np.minimum(np.asarray(magnitude, dt), dt(1e2))
    ph = np.asarray(phase, dt)
    w = np.asarray(window, dt)
    real = (mag * np.cos(ph)).astype(dt)
    imag = (mag * np.sin(ph)).astype(dt)
    f = real.shape[1]
    real_full = np.concatenate([real, real[:, 1:f - 1, :][:, ::-1, :]], axis=1)
    imag_full = np.concatenate([imag, -imag[:, 1:f - 1, :][:, ::-1, :]], axis=1)
    spectrum = (real_full + 1j * imag_full).astype(_c(dt))
    frames = sfft.ifft(spectrum, axis=1).real.astype(dt)  # (B, n_fft, frames)
    frames = (frames * w.reshape(1, -1, 1)).astype(dt)
    n_frames = frames.shape[2]
    out_len = (n_frames - 1) * hop + n_fft
    window_sum = np.maximum(_ola_1d((w ** 2).astype(dt), n_frames, hop), dt(1e-8))
    out = _ola(np.ascontiguousarray(frames.transpose(0, 2, 1)), hop)
    out = (out / window_sum).astype(dt)
    p = n_fft // 2
    return np.ascontiguousarray(out[:, p:out_len - p])


def cosyvoice3_stft(x, n_fft: int, hop: int, window, dt=F32):
    """CausalHiFTGenerator.swift:435-460 -- zero pad n_fft/2, rfft.  -> (real, imag) (B, F, frames)."""
    x = np.asarray(x, dt)
    w = np.asarray(window, dt)
    p = n_fft // 2
    xp = np.pad(x, ((0, 0), (p, p)))
    n_frames = (xp.shape[1] - n_fft) // hop + 1
    idx = (np.arange(n_frames) * hop)[:, None] + np.arange(n_fft)[None, :]
    frames = (xp[:, idx] * w).astype(dt)
    spec = sfft.rfft(frames, axis=-1).astype(_c(dt))
    return (np.ascontiguousarray(spec.real.transpose(0, 2, 1).astype(dt)),
            np.ascontiguousarray(spec.imag.transpose(0, 2, 1).astype(dt)))


def cosyvoice3_istft(magnitude, phase, n_fft: int, hop: int, window, dt=F32) -> np.ndarray:
    """CausalHiFTGenerator.swift:463-514."""
    mag = np.clip(np.asarray(magnitude, dt), dt(0.0), dt(1e2))
    ph = np.asarray(phase, dt)
    w = np.asarray(window, dt)
    real = (mag * np.cos(ph)).astype(dt).transpose(0, 2, 1)
    imag = (mag * np.sin(ph)).astype(dt).transpose(0, 2, 1)
    spec = (real + 1j * imag).astype(_c(dt))
    frames = sfft.irfft(spec, n=n_fft, axis=-1).astype(dt)
    frames = (frames * w).astype(dt)
    n_frames = frames.shape[1]
    out_len = n_fft + (n_frames - 1) * hop
    window_sum = np.maximum(_ola_1d((w * w).astype(dt), n_frames, hop), dt(1e-8))
    out = (_ola(np.ascontiguousarray(frames), hop) / window_sum[None, :]).astype(dt)
    p = n_fft // 2
    return np.ascontiguousarray(out[:, p:out_len - p])


def kokoro_get_window(win_len: int, n_fft: int, dt=F32) -> np.ndarray:
    """MLXSTFT.swift:48-67 ("hann" branch)."""
    w = hanning_window(win_len + 1, dt)[:win_len]
    if w.shape[0] < n_fft:
        w = np.concatenate([w, np.zeros(n_fft - w.shape[0], dt)])
    return w


def mlx_stft(x, n_fft=800, hop=None, win_length=None, center=True, pad_mode="reflect", dt=F32) -> np.ndarray:
    """MLXSTFT.swift:69-113.  1-D x -> complex (F, frames)."""
    hop = hop if hop is not None else n_fft // 4
    win_length = win_length if win_length is not None else n_fft
    w = kokoro_get_window(win_length, n_fft, dt)
    x = np.asarray(x, dt)
    if center:
        p = n_fft // 2
        if pad_mode == "constant":
            x = np.pad(x, (p, p))
        elif pad_mode == "reflect":
            x = np.concatenate([x[1:p + 1][::-1], x, x[-(p + 1):-1][::-1]])
        else:
            raise ValueError(f"Invalid pad mode {pad_mode}")
    num = x.shape[0] - n_fft
    if num < 0:
        raise ValueError("Input is too short")
    n_frames = 1 + num // hop
    frames = np.lib.stride_tricks.as_strided(x, (n_frames, n_fft), (hop * x.itemsize, x.itemsize))
    return np.ascontiguousarray(sfft.rfft((frames * w).astype(dt), axis=-1).astype(_c(dt)).T)


def kokoro_transform(x, n_fft=20, hop=5, win_length=20, dt=F32):
    """MLXSTFT.transform, MLXSTFT.swift:181-209.  (B, T) -> (|X|, atan2(Im, Re)) each (B, F, frames)."""
    x = np.asarray(x, dt)
    if x.ndim == 1:
        x = x[None]
    mags, phases = [], []
    for b in range(x.shape[0]):
        s = mlx_stft(x[b], n_fft, hop, win_length, dt=dt)
        mags.append(np.abs(s).astype(dt))
        phases.append(np.arctan2(s.imag.astype(dt), s.real.astype(dt)).astype(dt))
    return np.stack(mags), np.stack(phases)


def unwrap(p: np.ndarray) -> np.ndarray:
    """MLXSTFT.swift:23-46 (axis 1 of a 2-D array)."""
    dt = p.dtype.type
    period = dt(2.0) * dt(np.pi)
    discont = period / dt(2.0)
    d = p[:, 1:] - p[:, :-1]
    hi = period / dt(2.0)
    lo = -hi
    dm = d - lo
    dm = (np.fmod(np.fmod(dm, period) + period, period) + lo).astype(dt)
    # MLX remainder follows Python/NumPy sign conventions; the double-mod makes either convention agree
    dd_sign = np.where(d > 0, hi, dm).astype(dt)
    dm = np.where(dm == lo, dd_sign, dm).astype(dt)
    corr = (dm - d).astype(dt)
    corr = np.where(np.abs(d) < discont, dt(0.0), corr).astype(dt)
    return np.concatenate([p[:, :1], p[:, 1:] + np.cumsum(corr, axis=1, dtype=dt)], axis=1).astype(dt)


def mlx_istft(x: np.ndarray, hop=None, win_length=None, center=True, dt=F32) -> np.ndarray:
    """MLXSTFT.swift:115-163.  complex (F, frames) -> (L,).  Normalises by the plain window sum."""
    win_length = win_length if win_length is not None else (x.shape[0] - 1) * 2
    hop = hop if hop is not None else win_length // 4
    w = kokoro_get_window(win_length, win_length, dt)
    xt = x.T
    n_frames = xt.shape[0]
    frames = sfft.irfft(xt.astype(_c(dt)), axis=1).astype(dt)
    recon = _ola(np.ascontiguousarray((frames * w).astype(dt))[None], hop)[0]
    wsum = _ola_1d(w, n_frames, hop)
    with np.errstate(divide="ignore", invalid="ignore"):
        recon = np.where(wsum != 0, recon / wsum, recon).astype(dt)
    if center:
        recon = recon[win_length // 2:recon.shape[0] - win_length // 2]
    return np.ascontiguousarray(recon)


def kokoro_inverse(magnitude, phase, n_fft=20, hop=5, win_length=20, dt=F32) -> np.ndarray:
    """MLXSTFT.inverse, MLXSTFT.swift:211-235.  (B, F, frames) x2 -> (B, 1, L)."""
    mag = np.asarray(magnitude, dt)
    ph = np.asarray(phase, dt)
    outs = []
    for b in range(mag.shape[0]):
        pc = unwrap(ph[b])
        s = (mag[b] * np.exp(1j * pc.astype(dt))).astype(_c(dt))
        outs.append(mlx_istft(s, hop, win_length, True, dt))
    return np.stack(outs)[:, None, :]


# --------------------------------------------------------------------------------------
# batch helpers used by tests / bench (loop over clips; the reference has no batch path)
# --------------------------------------------------------------------------------------

def batch(fn, clips, *a, **k):
    return np.stack([fn(c, *a, **k) for c in clips])


# --------------------------------------------------------------------------------------
# adjacent rows: vocoder glue around the iSTFT
# --------------------------------------------------------------------------------------

def hift_head_istft(conv_out, n_fft: int, hop: int, window, audio_limit: float = 0.99, dt=F32) -> np.ndarray:
    """Tail of HiFTGenerator.decode, Codec/S3Gen/HiFiGAN.swift:577-589: h (B, n_fft+2, frames) -> magnitude = exp(h[:, :F]),
    phase = sin(h[:, F:]), istftHiFiGAN, clip(output, -audio_limit, audio_limit)."""
    h = np.asarray(conv_out, dt)
    f = n_fft // 2 + 1
    mag = np.exp(h[:, :f, :]).astype(dt)
    ph = np.sin(h[:, f:, :]).astype(dt)
    y = istft_hifigan(mag, ph, n_fft, hop, window, dt)
    return np.clip(y, dt(-audio_limit), dt(audio_limit)).astype(dt)


def s3gen_trim_fade(sampling_rate: int = 24000, dt=F32) -> np.ndarray:
    """Codec/S3Gen/S3Gen.swift:259-262: nTrim = sr / 50; zeros(nTrim) ++ (cos(linspace(pi, 0, nTrim)) + 1) / 2"""
    n = sampling_rate // 50
    ramp = (np.cos(np.linspace(np.pi, 0.0, n).astype(dt)).astype(dt) + dt(1.0)) / dt(2.0)
    return np.concatenate([np.zeros(n, dt), ramp.astype(dt)])


def apply_trim_fade(wav, fade) -> np.ndarray:
    """Codec/S3Gen/S3Gen.swift:284-289: result[..., 0 ..< fadeLen] *= trimFade when the waveform has at least fadeLen samples"""
    out = np.array(wav, copy=True)
    n = len(fade)
    if out.shape[1] >= n:
        out[:, :n] = out[:, :n] * np.asarray(fade, out.dtype)
    return out


def kokoro_head_istft(conv_out, n_fft: int = 20, hop: int = 5, win_length: int = 20, dt=F32) -> np.ndarray:
    """Tail of the Kokoro generator, TTS/Kokoro/Decoder/Generator.swift:182-190: x (B, n_fft+2, frames) ->
    spec = exp(x[:, :F]), phase = sin(x[:, F:]), MLXSTFT.inverse.  -> (B, 1, L)."""
    x = np.asarray(conv_out, dt)
    f = n_fft // 2 + 1
    return kokoro_inverse(np.exp(x[:, :f, :]).astype(dt), np.sin(x[:, f:, :]).astype(dt), n_fft, hop, win_length, dt)



def whisper_mel_segment(mel, seek: int, content_frames: int, length: int = 3000) -> np.ndarray:
    """melSegment of the seek loop, STT/Whisper/WhisperSTT.swift:171-182 with padOrTrimMel (:624-635):
    segmentSize = min(nFrames, contentFrames - seek); fullMel[seek ..< seek + segmentSize] zero-padded to nFrames rows,
    cast to float16.  mel (T', M) fp32 -> (length, M) float16."""
    mel = np.asarray(mel, F32)
    seg = max(0, min(length, content_frames - seek))
    rows = mel[seek:seek + seg]
    out = np.zeros((length, mel.shape[1]), F32)
    out[:rows.shape[0]] = rows
    return out.astype(np.float16)



def linear_interpolate_1d(x, scale_factor) -> np.ndarray:
    """linearInterpolate1d, TTS/CosyVoice2/HiFiGAN/CosyHiFTGenerator.swift:17-58, on (T,) or (B, T): PyTorch
    align_corners=False linear interpolation, every step in fp32 in the reference's op order."""
    x = np.asarray(x, F32)
    t = x.shape[-1]
    new_t = int(F32(t) * F32(scale_factor))
    if new_t == 0:
        new_t = 1
    step = F32(t) / F32(new_t)
    idx = ((np.arange(new_t, dtype=np.int32).astype(F32) + F32(0.5)) * step - F32(0.5)).astype(F32)
    idx = np.clip(idx, F32(0), F32(t) - F32(1.001)).astype(F32)
    lo = np.floor(idx).astype(np.int32)
    hi = np.minimum(lo + 1, np.int32(t - 1))
    wh = (idx - lo.astype(F32)).astype(F32)
    wl = (F32(1.0) - wh).astype(F32)
    return ((x[..., lo] * wl).astype(F32) + (x[..., hi] * wh).astype(F32)).astype(F32)


def resample_audio(audio, from_rate: int, to_rate: int) -> np.ndarray:
    """resampleAudio, TTS/CosyVoice2/CosyVoice2TTS.swift:733-744."""
    if from_rate == to_rate:
        return np.asarray(audio, F32)
    return linear_interpolate_1d(audio, F32(to_rate) / F32(from_rate))



def merge_tokenized_segments(tokenized_segments, overlap: int, token_rate: int):
    """mergeTokenizedSegments, Codec/S3Tokenizer/S3TokenizerUtils.swift:71-88."""
    merged = []
    overlap_tokens = (overlap // 2) * token_rate
    n = len(tokenized_segments)
    for i, tokens in enumerate(tokenized_segments):
        tokens = list(tokens)
        left = 0 if i == 0 else overlap_tokens
        right = len(tokens) - overlap_tokens if i != n - 1 else len(tokens)
        if left < right:
            merged.extend(tokens[left:right])
    return merged


def s3tokenizer_segments(mel, mel_len, window: int = 3000, stride: int = 2600):
    """Segment plan + unified batch of S3Tokenizer.quantize / quantizeMixedBatch, Codec/S3Tokenizer/S3Tokenizer.swift:474-571.
    mel (B, M, Tmax), mel_len (B,) -> (segments (S, M, window) fp32, lengths (S,), info [(batch_idx, segment_idx), ...]).
    (The all-short path stacks mel as it is; this is the mixed-batch path, whose short clips are padded to the window.)"""
    mel = np.asarray(mel, F32)
    segs, lens, info = [], [], []
    for b in range(mel.shape[0]):
        n = int(mel_len[b])
        if n <= window:
            seg = mel[b][:, :n]
            segs.append(np.pad(seg, ((0, 0), (0, window - n))))
            lens.append(n)
            info.append((b, 0))
        else:
            start, k = 0, 0
            while start < n:
                end = min(start + window, n)
                seg = mel[b][:, start:end]
                segs.append(np.pad(seg, ((0, 0), (0, window - seg.shape[1]))))
                lens.append(seg.shape[1])
                info.append((b, k))
                k += 1
                start += stride
    return np.stack(segs).astype(F32), np.asarray(lens, np.int32), info
