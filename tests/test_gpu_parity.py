"""GPU parity: CUDA path (through the C ABI) vs the CPU oracle on the same seeded inputs.

Tolerances are BASELINE.json's: log-mel / fbank within 1e-4 relative, evaluated as
|a - b| <= 1e-4 * max(1, |b|) (the normalised log-mel crosses zero); iSTFT waveforms within 1e-5
absolute; frame counts / shapes exact.
"""
import numpy as np
import pytest

from oracle import reference_dsp as R
from tests import synth

pytestmark = pytest.mark.gpu

RTOL = 1e-4
ISTFT_ATOL = 1e-5


def assert_feat_close(got, want, tol=RTOL, what=""):
    got = np.asarray(got)
    want = np.asarray(want)
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    err = np.abs(got.astype(np.float64) - want.astype(np.float64)) / np.maximum(1.0, np.abs(want.astype(np.float64)))
    i = np.unravel_index(np.argmax(err), err.shape)
    assert err.max() <= tol, f"{what}: max scaled err {err.max():.3e} at {i}: got {got[i]} want {want[i]}"


@pytest.fixture(scope="module")
def api(ctx):
    from mlx_swift_audio_b200 import api as A
    return A


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("n", [480000 // 10, 16000 * 3 + 37, 5000])
def test_whisper_log_mel(api, ctx, n_mels, n):
    x = synth.pcm(3, n, seed=1001)
    got = api.whisperLogMelSpectrogram(x, nMels=n_mels, ctx=ctx)
    want = np.stack([R.whisper_log_mel_spectrogram(c, n_mels) for c in x])
    assert_feat_close(got, want, what="whisper")


def test_random_lengths_frontends(api, ctx):
    # seeded random clip lengths around the tile (32 frames = 5120 samples) and row (160 samples) boundaries of the kernel
    rng = np.random.default_rng(77)
    lengths = sorted(set(int(v) for v in rng.integers(401, 40000, 10)) | {5120, 5119, 5121, 10240 + 200, 160 * 33})
    for n in lengths:
        x = synth.pcm(2, n, seed=n)
        assert_feat_close(api.whisperLogMelSpectrogram(x, nMels=80, ctx=ctx), np.stack([R.whisper_log_mel_spectrogram(c, 80) for c in x]),
                          what=f"whisper n={n}")
        assert_feat_close(api.kaldiFbankCAMPPlus(x, ctx=ctx), np.stack([R.kaldi_fbank_camp_plus(c) for c in x]), what=f"kaldi n={n}")
        assert_feat_close(api.funASRLogMelSpectrogram(x, ctx=ctx), np.stack([R.funasr_log_mel_spectrogram(c) for c in x]),
                          what=f"funasr n={n}")


def test_whisper_padding_and_single_clip(api, ctx):
    x = synth.pcm(1, 20000, seed=7)[0]
    got = api.whisperLogMelSpectrogram(x, nMels=80, padding=4000, ctx=ctx)
    want = R.whisper_log_mel_spectrogram(x, 80, padding=4000)
    assert_feat_close(got, want, what="whisper padding")


@pytest.mark.parametrize("n", [96000 // 4, 16000])
def test_chatterbox_log_mel(api, ctx, n):
    x = synth.pcm(2, n, seed=1002)
    got = api.logMelSpectrogramChatterbox(x, nMels=128, ctx=ctx)
    want = np.stack([R.log_mel_spectrogram_chatterbox(c, 128) for c in x])
    assert_feat_close(got, want, what="chatterbox")


def test_short_inputs_reflect_loops(api, ctx):
    # clips shorter than the reflect pad exercise the reference's while-loops (S3TokenizerUtils.swift:287-295)
    for n in (161, 170, 250, 399, 777):
        x = synth.pcm(1, n, seed=n, zero_tail_frac=0.0)[0]
        got = api.funASRLogMelSpectrogram(x, ctx=ctx)
        want = R.funasr_log_mel_spectrogram(x)
        assert_feat_close(got, want, what=f"funasr short n={n}")


def test_funasr_log_mel_and_pipeline(api, ctx):
    x = synth.pcm(3, 32000 + 55, seed=1003)
    got = api.funASRLogMelSpectrogram(x, ctx=ctx)
    want = np.stack([R.funasr_log_mel_spectrogram(c) for c in x])
    assert_feat_close(got, want, what="funasr logmel")
    lfr = api.applyLFR(want, ctx=ctx)
    want_lfr = np.stack([R.apply_lfr(f) for f in want])
    assert np.array_equal(lfr, want_lfr), "applyLFR is pure indexing: must be bit-exact"
    cm = api.applyCMVN(want_lfr, ctx=ctx)
    want_cm = np.stack([R.apply_cmvn(f) for f in want_lfr])
    assert_feat_close(cm, want_cm, what="cmvn")
    got_pp = api.preprocessAudio(x, ctx=ctx)
    want_pp = np.stack([R.preprocess_audio(c) for c in x])
    # CMVN divides by the per-column std: tolerance applies to the normalised features
    assert_feat_close(got_pp, want_pp, tol=2e-4, what="preprocessAudio")
    got_nonorm = api.preprocessAudio(x, applyNormalization=False, ctx=ctx)
    want_nonorm = np.stack([R.preprocess_audio(c, apply_normalization=False) for c in x])
    assert_feat_close(got_nonorm, want_nonorm, what="preprocessAudio no norm")


def test_cmvn_with_stats(api, ctx):
    rng = np.random.default_rng(3)
    f = rng.standard_normal((2, 50, 560)).astype(np.float32)
    mean = rng.standard_normal(560).astype(np.float32)
    istd = rng.uniform(0.5, 2, 560).astype(np.float32)
    got = api.applyCMVN(f, mean, istd, ctx=ctx)
    want = np.stack([R.apply_cmvn(a, mean, istd) for a in f])
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)


@pytest.mark.parametrize("dim", [560, 80, 100, 33])
def test_cmvn_row_counts(api, ctx, dim):
    # every bracket of the register-resident statistics kernel (<= 128 / 256 / 384 / 512 rows), its edges, and the streaming
    # kernels beyond it; rows = 1 has zero variance (x - mean = 0 over 1e-6)
    rng = np.random.default_rng(dim)
    for rows in (1, 2, 7, 8, 9, 100, 128, 129, 256, 257, 334, 384, 385, 511, 512, 513, 700):
        f = (rng.standard_normal((3, rows, dim)) * rng.uniform(0.1, 5.0, (1, 1, dim)) + rng.uniform(-8, 2, (1, 1, dim))).astype(np.float32)
        got = api.applyCMVN(f, ctx=ctx)
        want = np.stack([R.apply_cmvn(a) for a in f])
        assert_feat_close(got, want, tol=2e-4, what=f"cmvn rows={rows} dim={dim}")


@pytest.mark.parametrize("n", [32000, 400, 16000 + 123])
def test_kaldi_fbank(api, ctx, n):
    x = synth.pcm(2, n, seed=1004)
    got = api.kaldiFbankCAMPPlus(x, ctx=ctx)
    want = np.stack([R.kaldi_fbank_camp_plus(c) for c in x])
    assert_feat_close(got, want, what="kaldi")
    got_n = api.kaldiFbankCAMPPlus(x, meanNorm=True, ctx=ctx)
    want_n = np.stack([R.kaldi_fbank_mean_norm(f) for f in want])
    assert_feat_close(got_n, want_n, tol=2e-4, what="kaldi mean norm")


@pytest.mark.parametrize("n", [24000 * 2, 24000 + 333, 1000, 700])
def test_s3gen_mel(api, ctx, n):
    # n = 700 < pad + 1 exercises reflectPad2D's truncated reflection (S3GenMel.swift:17-25)
    x = synth.pcm(3, n, sample_rate=24000, seed=1008)
    got = api.s3genMelSpectrogram(x, ctx=ctx)
    want = R.s3gen_mel_spectrogram(x)
    assert_feat_close(got, want, what="s3gen mel")
    got1 = api.s3genMelSpectrogram(x[0], ctx=ctx)
    assert got1.shape == want[0].shape
    assert_feat_close(got1, want[0], what="s3gen mel 1-D")


def test_s3gen_too_short(api, ctx):
    from mlx_swift_audio_b200.api import B2ATooShort
    with pytest.raises(B2ATooShort):
        api.s3genMelSpectrogram(np.zeros(600, np.float32), ctx=ctx)


def test_stft_complex_1920(api, ctx):
    x = synth.pcm(2, 24000, sample_rate=24000, seed=1009)
    w = R.hann_periodic_via_hanning(1920)
    got = api.stft(x, w, 1920, 480, center=False, ctx=ctx)
    want = np.stack([R.stft(c, w, 1920, 480, center=False) for c in x])
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
    got_c = api.stft(x[0], w, 1920, 480, ctx=ctx)
    want_c = R.stft(x[0], w, 1920, 480)
    assert np.abs(got_c - want_c).max() <= 1e-4 * np.abs(want_c).max()


def test_voice_encoder_mel(api, ctx):
    x = synth.pcm(2, 24000, seed=1006)
    got = api.voiceEncoderMelspectrogram(x, ctx=ctx)
    want = np.stack([R.voice_encoder_melspectrogram(c) for c in x])
    # amplitude (not log) features: relative to the clip's scale
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-4 * scale


def test_stft_complex(api, ctx):
    x = synth.pcm(2, 8000, seed=1007)
    w = R.whisper_hann_window(400)
    got = api.stft(x, w, 400, 160, ctx=ctx)
    want = np.stack([R.stft(c, w, 400, 160) for c in x])
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
    got_nc = api.stft(x[0], w, 400, 160, center=False, ctx=ctx)
    want_nc = R.stft(x[0], w, 400, 160, center=False)
    assert np.abs(got_nc - want_nc).max() <= 1e-4 * np.abs(want_nc).max()


@pytest.mark.parametrize("frames", [2, 5, 253, 254, 1001, 4097])
def test_istft_hifigan(api, ctx, frames):
    mag, ph = synth.mag_phase(3, 9, frames, seed=frames)
    w = R.hann_window_periodic(16)
    got = api.istftHiFiGAN(mag, ph, 16, 4, w, ctx=ctx)
    want = R.istft_hifigan(mag, ph, 16, 4, w)
    assert got.shape == want.shape == (3, (frames - 1) * 4)
    assert np.abs(got - want).max() <= ISTFT_ATOL, np.abs(got - want).max()


def test_istft_cosyvoice3_negative_magnitudes(api, ctx):
    mag, ph = synth.mag_phase(2, 9, 777, seed=5)
    mag[:, :, ::7] *= -1.0  # CosyVoice3 clips below at 0, HiFT does not
    w = R.hann_window_periodic(16)
    got = api.cosyVoice3Istft(mag, ph, 16, 4, w, ctx=ctx)
    want = R.cosyvoice3_istft(mag, ph, 16, 4, w)
    assert np.abs(got - want).max() <= ISTFT_ATOL
    got_h = api.istftHiFiGAN(mag, ph, 16, 4, w, ctx=ctx)
    want_h = R.istft_hifigan(mag, ph, 16, 4, w)
    assert np.abs(got_h - want_h).max() <= ISTFT_ATOL


@pytest.mark.parametrize("frames", [2, 6, 253, 255, 1201])
def test_kokoro_inverse(api, ctx, frames):
    mag, ph = synth.mag_phase(2, 11, frames, seed=100 + frames)
    st = api.MLXSTFT(20, 5, 20, ctx=ctx)
    got = st.inverse(mag, ph)
    want = R.kokoro_inverse(mag, ph)
    assert got.shape == want.shape == (2, 1, (frames - 1) * 5)
    assert np.abs(got - want).max() <= ISTFT_ATOL


def test_kokoro_inverse_unwrap_path(api, ctx):
    # phase ~ U(-pi, pi) makes unwrap non-trivial: the unwrapped phase grows to hundreds of radians, where
    # fp32 spacing is ~3e-5 rad, so the result depends on cumsum order; compare at the resolution the
    # reference itself has there (mag <= 7.4 typical) -- documented in DESIGN.md.
    mag, ph = synth.mag_phase(2, 11, 600, seed=9, phase_mode="uniform")
    mag = np.minimum(mag, 1.0)
    st = api.MLXSTFT(20, 5, 20, ctx=ctx)
    got = st.inverse(mag, ph)
    want = R.kokoro_inverse(mag, ph)
    want64 = R.kokoro_inverse(mag, ph, dt=np.float64)
    ref_err = np.abs(want - want64).max()
    assert np.abs(got - want64).max() <= max(4 * ref_err, 1e-4)


@pytest.mark.parametrize("t_jump", [28, 29, 60, 61, 252, 253, 300])
def test_kokoro_inverse_single_phase_jump(api, ctx, t_jump):
    # All phases small (the kernel's fast path: no unwrap test, no range reduction) except ONE frame with a large phase:
    # the step of >= pi into and out of it must still be found, also when the neighbouring frame sits in another warp or
    # block of the kernel (frame f is lane (f + 3) % 32 of its warp, 253 segments per block).
    mag, ph = synth.mag_phase(2, 11, 400, seed=21)
    mag = np.minimum(mag, 1.0)
    ph = (0.3 * ph).astype(np.float32)
    ph[:, :, t_jump] = 3.0
    ph[1, 4, t_jump] = -3.1
    st = api.MLXSTFT(20, 5, 20, ctx=ctx)
    got = st.inverse(mag, ph)
    want = R.kokoro_inverse(mag, ph)
    # the reference's unwrap changes the phases after the jump by multiples of 2 pi (visible in fp32 as ~1e-6 rad)
    assert np.abs(R.unwrap(ph[0]) - ph[0]).max() > 3.0
    assert np.abs(got - want).max() <= 2e-5


@pytest.mark.parametrize("rates", [(24000, 16000), (16000, 24000), (44100, 16000), (16000, 16000), (48000, 7)])
def test_resample_audio_bit_exact(api, ctx, rates):
    x = synth.pcm(2, 24000 * 3 + 17, sample_rate=rates[0], seed=41, zero_tail_frac=0.0)
    got = api.resampleAudio(x, rates[0], rates[1], ctx=ctx)
    want = R.resample_audio(x, rates[0], rates[1])
    assert got.shape == want.shape
    assert np.array_equal(got, want)
    one = api.resampleAudio(x[0, :5], rates[0], rates[1], ctx=ctx)
    assert np.array_equal(one, R.resample_audio(x[0, :5], rates[0], rates[1]))


def test_s3tokenizer_segments(api, ctx):
    rng = np.random.default_rng(51)
    mel = rng.standard_normal((4, 8, 700)).astype(np.float32)
    lens = np.array([700, 120, 300, 301])               # window 300, stride 260: 3 windows, 1 short, exactly one window, 2 windows
    got, got_len, got_info = api.s3TokenizerSegments(mel, lens, window=300, stride=260, ctx=ctx)
    want, want_len, want_info = R.s3tokenizer_segments(mel, lens, window=300, stride=260)
    assert got.shape == want.shape and np.array_equal(got, want)
    assert np.array_equal(got_len, want_len) and got_info == want_info
    import torch
    got_d, _, _ = api.s3TokenizerSegments(torch.from_numpy(mel).cuda(), lens, window=300, stride=260)
    torch.cuda.synchronize()
    assert np.array_equal(got_d.cpu().numpy(), want)


def test_whisper_mel_segment_f16(api, ctx):
    # transcribe(): mel of the audio + 30 s of padding, content frames = len(audio) // 160, windows at arbitrary seeks
    x = synth.pcm(2, 16000 * 7 + 123, seed=31)
    mel = api.whisperLogMelSpectrogram(x, nMels=80, padding=480000, ctx=ctx)          # (2, 3700, 80)
    content = x.shape[1] // 160
    for seek in (0, 137, content - 5, content):
        got = api.whisperMelSegment(mel, seek, content, ctx=ctx)
        want = np.stack([R.whisper_mel_segment(m, seek, content) for m in mel])
        assert got.dtype == np.float16 and got.shape == (2, 3000, 80)
        assert np.array_equal(got.view(np.uint16), want.view(np.uint16)), seek     # a cast and a copy: bit exact
    # per-clip seeks, 128 mels, a window that runs into the end of the mel, device tensors
    import torch
    mel128 = api.whisperLogMelSpectrogram(torch.from_numpy(x).cuda(), nMels=128)
    seeks = np.array([10, mel128.shape[1] - 100])
    got = api.whisperMelSegment(mel128, seeks, mel128.shape[1], length=3000)
    torch.cuda.synchronize()
    for b in range(2):
        want = R.whisper_mel_segment(mel128[b].cpu().numpy(), int(seeks[b]), mel128.shape[1])
        assert np.array_equal(got[b].cpu().numpy().view(np.uint16), want.view(np.uint16))


@pytest.mark.parametrize("frames", [2, 254, 3001])
def test_hift_head_istft(api, ctx, frames):
    # convPost output: log-magnitudes N(-2, 1) (a few large ones: exp > 100 exercises the clip at 100), phase arguments N(0, 2^2)
    rng = np.random.default_rng(300 + frames)
    h = np.concatenate([rng.normal(-2.0, 1.0, (2, 9, frames)), rng.normal(0.0, 2.0, (2, 9, frames))], axis=1).astype(np.float32)
    h[0, 3, frames // 2] = 5.5          # exp -> 244 -> clipped to 100 -> output beyond the +-0.99 limiter
    h[1, 12, 0] = 4.0e4                 # large sine argument: range reduction path
    w = R.hann_window_periodic(16)
    got = api.hiftHeadIstft(h, 16, 4, w, 0.99, ctx=ctx)
    want = R.hift_head_istft(h, 16, 4, w, 0.99)
    assert got.shape == want.shape == (2, (frames - 1) * 4)
    assert np.abs(got - want).max() <= 2e-5   # fp32 sin(4e4) itself is only good to ~4e-3 * 2^-23 * 4e4 ... see DESIGN.md
    assert np.abs(got).max() <= 0.99



@pytest.mark.parametrize("frames", [2001, 300, 200])
def test_hift_head_istft_with_s3gen_fade(api, ctx, frames):
    # S3Token2Wav: vocoder head, then result[..., :960] *= trimFade when the waveform has >= 960 samples (S3Gen.swift:259-262, 284-289)
    rng = np.random.default_rng(frames)
    h = rng.standard_normal((2, 18, frames)).astype(np.float32)
    h[:, :9] -= 2.0
    h[:, 9:] *= 2.0
    w = R.hann_window_periodic(16)
    fade = api.s3genTrimFade(24000)
    assert np.abs(fade - R.s3gen_trim_fade(24000)).max() <= 2e-7
    got = api.hiftHeadIstftFade(h, 16, 4, w, fade, ctx=ctx)
    want = R.apply_trim_fade(R.hift_head_istft(h, 16, 4, w), fade)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= ISTFT_ATOL
    if want.shape[1] >= len(fade):
        assert not np.any(got[:, :480])                      # the first 20 ms are silenced
    else:
        assert np.abs(got - R.hift_head_istft(h, 16, 4, w)).max() <= ISTFT_ATOL   # too short: no fade at all
    # a different window on the same context replaces the cached table
    fade2 = np.linspace(0.0, 1.0, 100).astype(np.float32)
    got2 = api.hiftHeadIstftFade(h, 16, 4, w, fade2, ctx=ctx)
    assert np.abs(got2 - R.apply_trim_fade(R.hift_head_istft(h, 16, 4, w), fade2)).max() <= ISTFT_ATOL

@pytest.mark.parametrize("frames", [2, 253, 1201])
def test_kokoro_head_istft(api, ctx, frames):
    rng = np.random.default_rng(400 + frames)
    x = np.concatenate([rng.normal(-2.0, 1.0, (2, 11, frames)), rng.normal(0.0, 2.0, (2, 11, frames))], axis=1).astype(np.float32)
    got = api.kokoroHeadIstft(x, 20, 5, 20, ctx=ctx)
    want = R.kokoro_head_istft(x)
    assert got.shape == want.shape == (2, 1, (frames - 1) * 5)
    assert np.abs(got - want).max() <= ISTFT_ATOL


def test_forward_vocoder_stfts(api, ctx):
    x = synth.pcm(2, 4000, sample_rate=24000, seed=11, zero_tail_frac=0.0)
    w = R.hann_window_periodic(16)
    re, im = api.stftHiFiGAN(x, 16, 4, w, ctx=ctx)
    wr, wi = R.stft_hifigan(x, 16, 4, w)
    assert re.shape == wr.shape
    assert max(np.abs(re - wr).max(), np.abs(im - wi).max()) <= 1e-5
    re, im = api.cosyVoice3Stft(x, 16, 4, w, ctx=ctx)
    wr, wi = R.cosyvoice3_stft(x, 16, 4, w)
    assert max(np.abs(re - wr).max(), np.abs(im - wi).max()) <= 1e-5
    st = api.MLXSTFT(20, 5, 20, ctx=ctx)
    mag, ph = st.transform(x)
    wm, wp = R.kokoro_transform(x)
    assert np.abs(mag - wm).max() <= 1e-5
    # phase is ill-conditioned where the magnitude vanishes; compare the reconstructed complex value
    assert np.abs(mag * np.exp(1j * ph) - wm * np.exp(1j * wp)).max() <= 2e-5


def test_pad_or_trim(api, ctx):
    for n in (1000, 600000):
        x = np.full(n, 0.5, np.float32)  # the reference's own test input (Tests/WhisperTests.swift:85-96)
        y = api.padOrTrim(x, ctx=ctx)
        assert y.shape == (480000,)
        assert np.array_equal(y, R.pad_or_trim(x))


def test_too_short_is_an_error(api, ctx):
    from mlx_swift_audio_b200.api import B2ATooShort
    with pytest.raises(B2ATooShort):
        api.kaldiFbankCAMPPlus(np.zeros(100, np.float32), ctx=ctx)
    with pytest.raises(B2ATooShort):
        api.stftHiFiGAN(np.zeros((1, 8), np.float32), 16, 4, R.hann_window_periodic(16), ctx=ctx)


def test_device_tensors_roundtrip(api, ctx):
    import torch
    x = synth.pcm(4, 48000, seed=21)
    want = np.stack([R.whisper_log_mel_spectrogram(c, 128) for c in x])
    xt = torch.from_numpy(x).cuda()
    got = api.whisperLogMelSpectrogram(xt, nMels=128)
    torch.cuda.synchronize()
    assert_feat_close(got.cpu().numpy(), want, what="whisper device")
    mag, ph = synth.mag_phase(2, 9, 3000, seed=4)
    w = R.hann_window_periodic(16)
    y = api.istftHiFiGAN(torch.from_numpy(mag).cuda(), torch.from_numpy(ph).cuda(), 16, 4, w)
    torch.cuda.synchronize()
    assert np.abs(y.cpu().numpy() - R.istft_hifigan(mag, ph, 16, 4, w)).max() <= ISTFT_ATOL


def test_contexts_on_concurrent_host_threads(api):
    """The reference's helpers are re-entrant free functions called from several Swift actors ("Threading"): one
    context per host thread, no shared mutable state.  Four threads with their own contexts hammer different front ends at
    once; every result must equal the single-threaded one bit for bit."""
    import threading
    x = synth.pcm(4, 16000 * 4 + 321, seed=61)
    mag, ph = synth.mag_phase(2, 9, 3001, seed=62)
    w16 = R.hann_window_periodic(16)
    jobs = [
        lambda c: api.whisperLogMelSpectrogram(x, nMels=128, ctx=c),
        lambda c: api.preprocessAudio(x, ctx=c),
        lambda c: api.kaldiFbankCAMPPlus(x, meanNorm=True, ctx=c),
        lambda c: api.istftHiFiGAN(mag, ph, 16, 4, w16, ctx=c),
    ]
    ref_ctx = api.Context(0)
    want = [np.array(j(ref_ctx)) for j in jobs]
    ref_ctx.close()
    errors = []

    def worker(i):
        try:
            c = api.Context(0)
            for rep in range(15):
                k = (i + rep) % len(jobs)
                got = np.array(jobs[k](c))
                if not np.array_equal(got, want[k]):
                    errors.append((i, rep, k))
            c.close()
        except Exception as e:   # noqa: BLE001
            errors.append((i, repr(e)))

    threads = [threading.Thread(target=worker, args=(i,)) for i in range(4)]
    for t in threads:
        t.start()
    for t in threads:
        t.join(120)
    assert not errors, errors


def test_unwrap_and_mlx_stft(api, ctx):
    # unwrap (MLXSTFT.swift:23-46) as a stand-alone call: rows with phase jumps >= pi, rows without, 3-D input
    rng = np.random.default_rng(5)
    p = np.cumsum(rng.uniform(-2.5, 2.5, (3, 11, 700)), axis=-1)
    p = ((p + np.pi) % (2 * np.pi) - np.pi).astype(np.float32)          # wrapped random walk: many 2 pi jumps
    p[1] = np.sin(rng.standard_normal((11, 700))).astype(np.float32)    # the vocoders' case: no jump at all -> identity
    got = api.unwrap(p, ctx=ctx)
    want = np.stack([R.unwrap(q) for q in p])
    assert got.shape == want.shape
    assert np.array_equal(got[1], p[1])
    # fp32 cumulative sums over 700 steps: a few ulp of the running phase (hundreds of radians)
    assert np.abs(got - want).max() <= 2e-4 * max(1.0, np.abs(want).max())
    assert np.abs(got - np.unwrap(p.astype(np.float64), axis=-1)).max() <= 1e-3
    # mlxStft (MLXSTFT.swift:69-113) in the package's configuration: complex (F, frames)
    x = synth.pcm(2, 24000 // 4 + 3, sample_rate=24000, seed=6)
    z = api.mlxStft(x, 20, 5, ctx=ctx)
    ref = np.stack([R.mlx_stft(c, 20, 5, 20) for c in x])
    assert z.shape == ref.shape and z.dtype == np.complex64
    assert np.abs(z - ref).max() <= 1e-5 * max(1.0, np.abs(ref).max())
    z1 = api.mlxStft(x[0], 20, 5, ctx=ctx)
    assert np.array_equal(z1, z[0])


def test_reflect_pad_stand_alone(api, ctx):
    # reflectPad (S3TokenizerUtils.swift:266-298): ordinary inputs, the short-input loops (n - 1 < padding), n == 1, padding 0
    rng = np.random.default_rng(9)
    for n, pad in ((1000, 200), (201, 200), (150, 200), (37, 200), (2, 5), (1, 4), (64, 0), (1920, 720)):
        x = rng.standard_normal((3, n)).astype(np.float32)
        got = api.reflectPad(x, pad, ctx=ctx)
        want = np.stack([R.reflect_pad(c, pad) for c in x])
        assert got.shape == want.shape == (3, n + 2 * pad)
        assert np.array_equal(got, want), (n, pad)
    assert np.array_equal(api.reflectPad1D(x[0], 720, ctx=ctx), R.reflect_pad(x[0], 720))
