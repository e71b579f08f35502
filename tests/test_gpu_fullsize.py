"""GPU tests at BASELINE.json's full clip lengths (30 s clips, large batches), through size-independent properties --
the oracle only sees a handful of clips:

* batch invariance: clip b of a big batch == the same clip run alone, bit for bit (per-clip statistics stay per clip);
* spot checks of a few clips against the CPU oracle at the stated tolerance;
* the Whisper clamp invariant (min >= max - 2 in the normalised domain);
* forward STFT -> inverse STFT round trip (HiFT 16/4 is COLA with the envelope normalisation; Kokoro's pair is not), with phases from atan2, i.e.
  outside (-pi/2, pi/2): exercises the range-reduced sincos path of the iSTFT kernel;
* linearity of the iSTFT in the magnitudes.
"""
import numpy as np
import pytest

from oracle import reference_dsp as R
from tests import synth

pytestmark = pytest.mark.gpu


def test_whisper_128_full_length_batch(ctx):
    import torch
    from mlx_swift_audio_b200 import api
    B, n = 96, 480000                       # 96 x 30 s: 9024 tiles, every persistent CTA walks several clips
    g = torch.Generator(device="cuda").manual_seed(7)
    x = 0.1 * torch.randn((B, n), generator=g, device="cuda")
    t = torch.arange(n, device="cuda", dtype=torch.float32) / 16000.0
    x += 0.3 * torch.sin(2 * np.pi * 440.0 * t)[None, :] * torch.rand((B, 1), generator=g, device="cuda")
    x[:, n - n // 10:] = 0.0                # silent tail: the max-8 clamp is active
    x[5] *= 1e-3                            # a quiet clip: its own maximum, not the batch's
    got = api.whisperLogMelSpectrogram(x, nMels=128)
    torch.cuda.synchronize()
    assert got.shape == (B, 3000, 128)
    # clamp invariant per clip
    mx = got.amax(dim=(1, 2))
    mn = got.amin(dim=(1, 2))
    assert bool(torch.all(mn >= mx - 2.0 - 1e-6))
    assert bool(torch.all(torch.isfinite(got)))
    # batch invariance, bit for bit
    for b in (0, 5, 95):
        alone = api.whisperLogMelSpectrogram(x[b:b + 1].contiguous(), nMels=128)
        torch.cuda.synchronize()
        assert torch.equal(alone[0], got[b]), f"clip {b} differs between batch and single run"
    # oracle on a few clips
    for b in (3, 5):
        want = R.whisper_log_mel_spectrogram(x[b].cpu().numpy(), 128)
        err = np.abs(got[b].cpu().numpy().astype(np.float64) - want) / np.maximum(1.0, np.abs(want))
        assert err.max() <= 1e-4, (b, err.max())


def test_hift_round_trip_and_linearity_full_length(ctx):
    import torch
    from mlx_swift_audio_b200 import api
    B, n = 24, 720000                       # 24 x 30 s at 24 kHz: 180001 frames per clip
    g = torch.Generator(device="cuda").manual_seed(11)
    x = 0.2 * torch.randn((B, n), generator=g, device="cuda")
    w = R.hann_window_periodic(16)
    re, im = api.stftHiFiGAN(x, 16, 4, w)
    assert re.shape == (B, 9, n // 4 + 1)
    mag = torch.sqrt(re * re + im * im)
    ph = torch.atan2(im, re)                # in (-pi, pi]: the iSTFT kernel's range-reduced sincos path
    y = api.istftHiFiGAN(mag, ph, 16, 4, w)
    torch.cuda.synchronize()
    assert y.shape == (B, n)
    # COLA + envelope normalisation: exact reconstruction up to fp32 rounding (the reflect-padded edges included)
    assert float((y - x).abs().max()) <= 2e-6 * 16
    # oracle on one clip, at the iSTFT tolerance
    want = R.istft_hifigan(mag[1:2].cpu().numpy(), ph[1:2].cpu().numpy(), 16, 4, w)
    assert np.abs(y[1:2].cpu().numpy() - want).max() <= 1e-5
    # linearity in the magnitudes (no clipping active: mag << 100)
    y3 = api.istftHiFiGAN(3.0 * mag, ph, 16, 4, w)
    torch.cuda.synchronize()
    assert float((y3 - 3.0 * y).abs().max()) <= 1e-5


def test_kokoro_full_length(ctx):
    import torch
    from mlx_swift_audio_b200 import api
    B, n = 8, 720000
    g = torch.Generator(device="cuda").manual_seed(13)
    x = 0.2 * torch.randn((B, n), generator=g, device="cuda")
    st = api.MLXSTFT(20, 5, 20, ctx=None)
    mag, ph = st.transform(x)
    assert mag.shape == (B, 11, n // 5 + 1)
    y = st.inverse(mag, ph)                 # phases from atan2: unwrap is active, the flag path runs
    torch.cuda.synchronize()
    assert y.shape == (B, 1, n)
    assert bool(torch.all(torch.isfinite(y)))
    # (the reference's pair is not a perfect-reconstruction pair: the inverse divides by sum w, not sum w^2)
    # unwrap is causal along time, so the head of a clip only depends on the head of its spectrum: oracle on 3000 frames
    # (the unwrapped phase random-walks to hundreds of radians, where the fp32 cumsum order matters: compare with fp64 truth at
    # the resolution the fp32 reference itself has there, as tests/test_gpu_parity.py::test_kokoro_inverse_unwrap_path does)
    m_np, p_np = mag[2:3, :, :3000].cpu().numpy(), ph[2:3, :, :3000].cpu().numpy()
    want = R.kokoro_inverse(m_np, p_np)
    want64 = R.kokoro_inverse(m_np, p_np, dt=np.float64)
    k = (3000 - 4) * 5
    ref_err = np.abs(want[:, :, :k] - want64[:, :, :k]).max()
    assert np.abs(y[2:3, :, :k].cpu().numpy() - want64[:, :, :k]).max() <= max(4 * ref_err, 1e-4)
    # batch invariance, bit for bit
    alone = st.inverse(mag[5:6].contiguous(), ph[5:6].contiguous())
    torch.cuda.synchronize()
    assert torch.equal(alone[0], y[5])


def test_one_very_long_clip(ctx):
    """Maximum sizes along the time axis: ONE 30-minute clip (28.8 M samples, 180 000 frames, 5625 tiles) through the Whisper,
    Fun-ASR and Kaldi front ends -- the clip-global maximum, the tile walk and the 64-bit output offsets of a clip that is 60x
    longer than the benchmark's -- and a 12.5-minute HiFT iSTFT (4.5 M frames).  Checked against the oracle on the whole clip."""
    from mlx_swift_audio_b200 import api
    n = 16000 * 60 * 30 + 77
    x = synth.pcm(1, n, seed=404)[0]
    got = api.whisperLogMelSpectrogram(x, nMels=80, ctx=ctx)
    want = R.whisper_log_mel_spectrogram(x, 80)
    assert got.shape == want.shape == (n // 160, 80)
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) <= 1e-4
    # Un-clamped natural logs over 14 M values: the rare bins that lie ~100 dB below the frame's peak hold fp32 rounding noise in
    # BOTH fp32 implementations, so the yardstick is fp64 truth and the fp32 oracle's own error against it.
    def close_to_truth(got, want32, want64):
        assert got.shape == want32.shape
        scale = np.maximum(1.0, np.abs(want64))
        e_oracle = float(np.max(np.abs(want32 - want64) / scale))
        e_gpu = float(np.max(np.abs(got - want64) / scale))
        assert e_gpu <= max(1e-4, 4.0 * e_oracle), (e_gpu, e_oracle)
        assert float(np.quantile(np.abs(got - want32) / np.maximum(1.0, np.abs(want32)), 0.99999)) <= 1e-4

    close_to_truth(api.kaldiFbankCAMPPlus(x, ctx=ctx), R.kaldi_fbank_camp_plus(x), R.kaldi_fbank_camp_plus(x, dt=np.float64))
    close_to_truth(api.preprocessAudio(x, applyNormalization=False, ctx=ctx), R.preprocess_audio(x, apply_normalization=False),
                   R.preprocess_audio(x, apply_normalization=False, dt=np.float64))
    frames = 4_500_001
    mag, ph = synth.mag_phase(1, 9, frames, seed=405)
    w = R.hann_window_periodic(16)
    y = api.istftHiFiGAN(mag, ph, 16, 4, w, ctx=ctx)
    assert y.shape == (1, (frames - 1) * 4)
    assert np.abs(y - R.istft_hifigan(mag, ph, 16, 4, w)).max() <= 1e-5


def _pcm_batch(B, n, sr, seed):
    import torch
    g = torch.Generator(device="cuda").manual_seed(seed)
    x = 0.1 * torch.randn((B, n), generator=g, device="cuda")
    t = torch.arange(n, device="cuda", dtype=torch.float32) / float(sr)
    for f in (220.0, 1000.0, 3300.0):
        x += 0.2 * torch.sin(2 * np.pi * f * t[None, :] + 2 * np.pi * torch.rand((B, 1), generator=g, device="cuda"))
    x.clamp_(-1.0, 1.0)
    x[:, n - n // 10:] = 0.0
    x[3] *= 1e-3   # a quiet clip: per-clip statistics must stay per clip
    return x


def test_funasr_kaldi_s3gen_full_length_batches(ctx):
    """BASELINE configs 3a / 3b / 4 at their full clip lengths (20 s @16 kHz, 10 s @24 kHz) in batches large enough that every
    persistent CTA walks several clips: batch invariance bit for bit (CMVN / mean-norm statistics are per clip), finiteness, the
    defining property of each normalisation (zero column mean; unit column variance for CMVN), spot checks against the oracle."""
    import torch
    from mlx_swift_audio_b200 import api
    B, n = 64, 320000
    x = _pcm_batch(B, n, 16000, 11)
    # 3a: Fun-ASR preprocessAudio (log-mel + LFR 7/6 + per-utterance CMVN)
    got = api.preprocessAudio(x)
    torch.cuda.synchronize()
    assert got.shape == (B, 334, 560) and bool(torch.all(torch.isfinite(got)))
    assert float(got.mean(dim=1).abs().max()) <= 2e-4
    assert float((got.var(dim=1, unbiased=False).sqrt() - 1.0).abs().max()) <= 2e-3
    for b in (0, 3, B - 1):
        alone = api.preprocessAudio(x[b:b + 1].contiguous())
        torch.cuda.synchronize()
        assert torch.equal(alone[0], got[b]), f"funasr clip {b} differs between batch and single run"
    want = R.preprocess_audio(x[5].cpu().numpy())
    assert np.max(np.abs(got[5].cpu().numpy() - want) / np.maximum(1.0, np.abs(want))) <= 2e-4
    # 3b: Kaldi fbank + CAM++ mean normalisation
    got = api.kaldiFbankCAMPPlus(x, meanNorm=True)
    torch.cuda.synchronize()
    assert got.shape == (B, 1998, 80) and bool(torch.all(torch.isfinite(got)))
    assert float(got.mean(dim=1).abs().max()) <= 2e-4
    for b in (0, 3, B - 1):
        alone = api.kaldiFbankCAMPPlus(x[b:b + 1].contiguous(), meanNorm=True)
        torch.cuda.synchronize()
        assert torch.equal(alone[0], got[b]), f"kaldi clip {b} differs between batch and single run"
    # (un-clamped logs: fp64 truth and the fp32 oracle's own error are the yardstick)
    x5 = x[5].cpu().numpy()
    want = R.kaldi_fbank_mean_norm(R.kaldi_fbank_camp_plus(x5))
    truth = R.kaldi_fbank_mean_norm(R.kaldi_fbank_camp_plus(x5, dt=np.float64))
    scale = np.maximum(1.0, np.abs(truth))
    e_oracle = float(np.max(np.abs(want - truth) / scale))
    e_gpu = float(np.max(np.abs(got[5].cpu().numpy() - truth) / scale))
    assert e_gpu <= max(2e-4, 4.0 * e_oracle), (e_gpu, e_oracle)
    # 4: S3Gen 24 kHz mel
    B, n = 64, 240000
    y = _pcm_batch(B, n, 24000, 12)
    got = api.s3genMelSpectrogram(y)
    torch.cuda.synchronize()
    assert got.shape == (B, 80, 500) and bool(torch.all(torch.isfinite(got)))
    assert float(got.min()) >= np.log(1e-5) - 1e-5          # ln(max(., 1e-5)) floor
    for b in (0, 3, B - 1):
        alone = api.s3genMelSpectrogram(y[b:b + 1].contiguous())
        torch.cuda.synchronize()
        assert torch.equal(alone[0], got[b]), f"s3gen clip {b} differs between batch and single run"
    want = R.s3gen_mel_spectrogram(y[4:6].cpu().numpy())
    assert np.max(np.abs(got[4:6].cpu().numpy() - want) / np.maximum(1.0, np.abs(want))) <= 1e-4


def test_repeated_runs_are_bit_identical(ctx):
    """Determinism under load: the front-end kernels reuse shared memory several times per tile (PCM tile, exchange buffer, spectrum,
    staged mel values) behind four CTA barriers, the iSTFT overlaps frames through a shared tile; a missing barrier shows up as
    run-to-run differences long before it shows up as a tolerance failure.  20 back-to-back runs of the large batches must agree
    bit for bit (compute-sanitizer is not available on the GPU pool)."""
    import torch
    from mlx_swift_audio_b200 import api
    x = _pcm_batch(48, 480000, 16000, 21)
    first = api.whisperLogMelSpectrogram(x, nMels=128)
    firstk = api.kaldiFbankCAMPPlus(x[:, :320000].contiguous(), meanNorm=True)
    firstf = api.preprocessAudio(x[:, :320000].contiguous())
    y = _pcm_batch(16, 240000, 24000, 22)
    firsts = api.s3genMelSpectrogram(y)
    g = torch.Generator(device="cuda").manual_seed(23)
    mag = torch.exp(torch.randn((8, 9, 180001), generator=g, device="cuda") - 2.0)
    ph = torch.sin(2.0 * torch.randn((8, 9, 180001), generator=g, device="cuda"))
    w16 = R.hann_window_periodic(16)
    firsti = api.istftHiFiGAN(mag, ph, 16, 4, w16)
    torch.cuda.synchronize()
    for rep in range(20):
        assert torch.equal(api.whisperLogMelSpectrogram(x, nMels=128), first), f"whisper run {rep}"
        assert torch.equal(api.kaldiFbankCAMPPlus(x[:, :320000].contiguous(), meanNorm=True), firstk), f"kaldi run {rep}"
        assert torch.equal(api.preprocessAudio(x[:, :320000].contiguous()), firstf), f"funasr run {rep}"
        assert torch.equal(api.s3genMelSpectrogram(y), firsts), f"s3gen run {rep}"
        assert torch.equal(api.istftHiFiGAN(mag, ph, 16, 4, w16), firsti), f"istft run {rep}"


def test_whisper_f16_and_pcm16_full_length_batch(ctx):
    """Round-2 entries at BASELINE's clip length: the fp16 features are the cast of the fp32 ones bit for bit, for every clip of a
    96 x 30 s batch (device-resident and through the chunked host pipeline), and 16-bit PCM gives exactly what the scaled floats give."""
    import torch
    from mlx_swift_audio_b200 import api
    B, n = 96, 480000
    g = torch.Generator(device="cuda").manual_seed(13)
    xi = torch.randint(-20000, 20000, (B, n), generator=g, device="cuda", dtype=torch.int32).to(torch.int16)
    xi[:, n - n // 10:] = 0
    xf = xi.to(torch.float32) / 32768.0
    f32 = api.whisperLogMelSpectrogram(xf, nMels=128)
    f16 = api.whisperLogMelSpectrogramF16(xf, nMels=128)
    p16 = api.whisperLogMelSpectrogramF16(xi, nMels=128)
    torch.cuda.synchronize()
    assert f16.dtype == torch.float16 and f16.shape == (B, 3000, 128)
    assert torch.equal(f16, f32.to(torch.float16))
    assert torch.equal(p16, f16)
    # host pipeline (3 x 128 MB chunks at this size): same bits
    h16 = api.whisperLogMelSpectrogramF16(xi.cpu().numpy(), nMels=128, ctx=ctx)
    assert np.array_equal(h16.view(np.uint16), f16.cpu().numpy().view(np.uint16))
