"""GPU: the CUDA path against the frozen golden fixtures (tests/golden/, made by tests/make_golden.py)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_fp32_v1.npz")


def close(a, b, tol=1e-4):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape
    err = np.abs(a - b) / np.maximum(1.0, np.abs(b))
    assert err.max() <= tol, err.max()


def test_cuda_path_matches_golden_vectors(ctx):
    from mlx_swift_audio_b200 import api as A
    g = np.load(GOLD)
    x16, x24, sine = g["in_x16"], g["in_x24"], g["in_sine"]
    close(A.whisperLogMelSpectrogram(x16, nMels=80, ctx=ctx), g["whisper80"])
    close(A.whisperLogMelSpectrogram(x16, nMels=128, ctx=ctx), g["whisper128"])
    close(A.whisperLogMelSpectrogram(sine, nMels=80, ctx=ctx), g["sine_whisper80"])
    close(A.logMelSpectrogramChatterbox(x16, ctx=ctx), g["chatterbox128"])
    # pure tone, no clamp: compare against fp64 truth with the error the fp32 oracle itself has there as the yardstick
    t64 = g["sine_funasr_logmel_f64"]
    e32 = np.max(np.abs(g["sine_funasr_logmel"] - t64) / np.maximum(1.0, np.abs(t64)))
    close(A.funASRLogMelSpectrogram(sine, ctx=ctx), t64, max(1e-4, 4.0 * e32))
    assert np.array_equal(A.applyLFR(g["sine_funasr_logmel"], ctx=ctx), g["sine_funasr_lfr"])
    close(A.preprocessAudio(x16, ctx=ctx), g["funasr_preprocess"], 2e-4)
    close(A.kaldiFbankCAMPPlus(x16, ctx=ctx), g["kaldi_fbank"])
    close(A.kaldiFbankCAMPPlus(x16, meanNorm=True, ctx=ctx), g["kaldi_fbank_meannorm"], 2e-4)
    close(A.s3genMelSpectrogram(x24, ctx=ctx), g["s3gen_mel"])
    ve = A.voiceEncoderMelspectrogram(x16, ctx=ctx)
    assert np.abs(ve - g["voice_encoder_mel"]).max() <= 1e-4 * np.abs(g["voice_encoder_mel"]).max()
    z = A.stft(x16, A.whisperHannWindow(400), 400, 160, ctx=ctx)
    want = g["stft400"].view(np.complex64)
    assert np.abs(z - want).max() <= 1e-4 * np.abs(want).max()
    w16 = A.hannWindowPeriodic(16)
    assert np.abs(A.istftHiFiGAN(g["in_mag16"], g["in_ph16"], 16, 4, w16, ctx=ctx) - g["istft_hifigan"]).max() <= 1e-5
    assert np.abs(A.cosyVoice3Istft(g["in_mag16"], g["in_ph16"], 16, 4, w16, ctx=ctx) - g["cv3_istft"]).max() <= 1e-5
    st = A.MLXSTFT(20, 5, 20, ctx=ctx)
    assert np.abs(st.inverse(g["in_mag20"], g["in_ph20"]) - g["kokoro_inverse"]).max() <= 1e-5
    re, im = A.stftHiFiGAN(x24[:, :2000], 16, 4, w16, ctx=ctx)
    assert max(np.abs(re - g["stft_hifigan_re"]).max(), np.abs(im - g["stft_hifigan_im"]).max()) <= 1e-5
    re, im = A.cosyVoice3Stft(x24[:, :2000], 16, 4, w16, ctx=ctx)
    assert max(np.abs(re - g["cv3_stft_re"]).max(), np.abs(im - g["cv3_stft_im"]).max()) <= 1e-5
    mag, ph = st.transform(x24[:, :2000])
    assert np.abs(mag - g["kokoro_mag"]).max() <= 1e-5


def test_cuda_adjacent_rows_match_golden_vectors(ctx):
    from mlx_swift_audio_b200 import api as A
    a = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_fp32_v2_adjacent.npz"))
    g = np.load(GOLD)
    w16 = A.hannWindowPeriodic(16)
    assert np.abs(A.hiftHeadIstft(a["in_h16"], 16, 4, w16, 0.99, ctx=ctx) - a["hift_head_istft"]).max() <= 2e-5
    assert np.abs(A.kokoroHeadIstft(a["in_h20"], ctx=ctx) - a["kokoro_head_istft"]).max() <= 1e-5
    for seek, key in ((0, "mel_segment_seek0"), (37, "mel_segment_seek37")):
        got = A.whisperMelSegment(a["in_mel"], seek, 50, length=64, ctx=ctx)
        assert np.array_equal(got.view(np.uint16), a[key].view(np.uint16))
    assert np.array_equal(A.resampleAudio(g["in_x24"], 24000, 16000, ctx=ctx), a["resample_24k_16k"])
    assert np.array_equal(A.resampleAudio(g["in_x16"], 16000, 24000, ctx=ctx), a["resample_16k_24k"])
