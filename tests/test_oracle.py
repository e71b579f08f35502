"""CPU: the oracle against its frozen golden vectors, against independent implementations
(torch.stft / torch.istft) and against the only assertions the reference's own tests make for this
path (shapes: Tests/FunASRTests.swift:141-186, Tests/WhisperTests.swift:85-96)."""
import os

import numpy as np
import pytest

from oracle import reference_dsp as R
from tests import synth

GOLD = os.path.join(os.path.dirname(__file__), "golden", "oracle_fp32_v1.npz")


@pytest.fixture(scope="module")
def gold():
    return np.load(GOLD)


def close(a, b, tol=2e-5):
    a = np.asarray(a, np.float64)
    b = np.asarray(b, np.float64)
    assert a.shape == b.shape
    assert np.max(np.abs(a - b) / np.maximum(1.0, np.abs(b))) <= tol


def test_reference_unit_test_shapes(gold):
    # FunASRTests.funASRAudioPreprocessing: 1 s 440 Hz sine
    sine = gold["in_sine"]
    mel = R.funasr_log_mel_spectrogram(sine)
    assert mel.ndim == 2 and mel.shape[1] == 80
    lfr = R.apply_lfr(mel)
    assert lfr.shape[1] == 560 and lfr.shape[0] == (mel.shape[0] + 5) // 6
    # WhisperTests.whisperAudioPreprocessing: constant 0.5 arrays
    assert R.pad_or_trim(np.full(1000, 0.5, np.float32)).shape == (480000,)
    assert R.pad_or_trim(np.full(600000, 0.5, np.float32)).shape == (480000,)
    assert R.compute_feature_length(16000) == (100 + 5) // 6


def test_frame_counts():
    assert R.whisper_log_mel_spectrogram(np.zeros(480000, np.float32) + 1e-3, 80).shape == (3000, 80)
    assert R.funasr_log_mel_spectrogram(np.ones(320000, np.float32)).shape == (2001, 80)
    assert R.apply_lfr(np.zeros((2001, 80), np.float32)).shape == (334, 560)
    assert R.kaldi_fbank_camp_plus(np.ones(320000, np.float32)).shape == (1998, 80)
    assert R.s3gen_mel_spectrogram(np.ones((1, 240000), np.float32)).shape == (1, 80, 500)
    assert R.voice_encoder_melspectrogram(np.ones(96000, np.float32)).shape == (40, 601)


def test_golden_frontends(gold):
    x16, x24 = gold["in_x16"], gold["in_x24"]
    close(np.stack([R.whisper_log_mel_spectrogram(c, 80) for c in x16]), gold["whisper80"])
    close(np.stack([R.whisper_log_mel_spectrogram(c, 128) for c in x16]), gold["whisper128"])
    close(np.stack([R.log_mel_spectrogram_chatterbox(c, 128) for c in x16]), gold["chatterbox128"])
    close(np.stack([R.preprocess_audio(c) for c in x16]), gold["funasr_preprocess"], 1e-4)
    close(np.stack([R.kaldi_fbank_camp_plus(c) for c in x16]), gold["kaldi_fbank"])
    close(R.s3gen_mel_spectrogram(x24), gold["s3gen_mel"])
    close(R.preprocess_audio(gold["in_sine"]), gold["sine_funasr_preprocess"], 1e-3)  # pure tone: CMVN divides by tiny std
    close(R.whisper_log_mel_spectrogram(gold["in_sine"], 80), gold["sine_whisper80"])


def test_golden_vocoder(gold):
    w16 = R.hann_window_periodic(16)
    assert np.abs(R.istft_hifigan(gold["in_mag16"], gold["in_ph16"], 16, 4, w16) - gold["istft_hifigan"]).max() <= 2e-6
    assert np.abs(R.cosyvoice3_istft(gold["in_mag16"], gold["in_ph16"], 16, 4, w16) - gold["cv3_istft"]).max() <= 2e-6
    assert np.abs(R.kokoro_inverse(gold["in_mag20"], gold["in_ph20"]) - gold["kokoro_inverse"]).max() <= 2e-6
    re, im = R.stft_hifigan(gold["in_x24"][:, :2000], 16, 4, w16)
    assert np.abs(re - gold["stft_hifigan_re"]).max() <= 2e-6 and np.abs(im - gold["stft_hifigan_im"]).max() <= 2e-6


def test_fp32_oracle_close_to_fp64(gold):
    x = gold["in_x16"][0]
    a = R.whisper_log_mel_spectrogram(x, 128)
    b = R.whisper_log_mel_spectrogram(x, 128, dt=np.float64)
    assert np.abs(a - b).max() <= 2e-5
    w = R.hann_window_periodic(16)
    y32 = R.istft_hifigan(gold["in_mag16"], gold["in_ph16"], 16, 4, w)
    y64 = R.istft_hifigan(gold["in_mag16"], gold["in_ph16"], 16, 4, w, dt=np.float64)
    assert np.abs(y32 - y64).max() <= 5e-6


def test_against_torch(gold):
    torch = pytest.importorskip("torch")
    x = gold["in_x16"][0]
    st = torch.stft(torch.from_numpy(x), 400, 160, window=torch.hann_window(400, periodic=False), return_complex=True).numpy().T
    mine = R.stft(x, R.whisper_hann_window(400), 400, 160)
    assert np.abs(st - mine).max() <= 2e-4 * np.abs(st).max()
    # iSTFT: inverse of the forward transform is the identity (window-sum-square normalisation)
    w16 = R.hann_window_periodic(16)
    sig = gold["in_x24"][:, :2000]
    re, im = R.stft_hifigan(sig, 16, 4, w16)
    rec = R.istft_hifigan(np.sqrt(re ** 2 + im ** 2), np.arctan2(im, re), 16, 4, w16)
    assert np.abs(rec - sig).max() <= 1e-6
    ti = torch.istft(torch.complex(torch.from_numpy(re), torch.from_numpy(im)), 16, 4, window=torch.from_numpy(w16)).numpy()
    assert np.abs(ti - rec).max() <= 1e-6
    # Kokoro normalises by sum(w) instead of sum(w^2): interior gain 1.5 / 2.0
    m, p = R.kokoro_transform(sig)
    rk = R.kokoro_inverse(m, p)[:, 0]
    assert np.abs(rk[:, 40:-40] - 0.75 * sig[:, 40:-40]).max() <= 1e-4


def test_unwrap_matches_numpy():
    rng = np.random.default_rng(0)
    p = rng.uniform(-np.pi, np.pi, (7, 300)).astype(np.float32)
    assert np.abs(R.unwrap(p) - np.unwrap(p.astype(np.float64), axis=1)).max() <= 1e-4
    q = np.sin(rng.normal(0, 2, (7, 300))).astype(np.float32)  # vocoder phases: unwrap is the identity
    assert np.array_equal(R.unwrap(q), q)


def test_reflect_pad_short_inputs():
    # numpy-style reflect where it is defined ...
    x = np.arange(10, dtype=np.float32)
    assert np.array_equal(R.reflect_pad(x, 4), np.pad(x, 4, mode="reflect"))
    # ... and the reference's own looped variant below that (S3TokenizerUtils.swift:287-295)
    assert R.reflect_pad_index(5, 8).tolist() == [4, 3, 2, 1, 4, 3, 2, 1, 0, 1, 2, 3, 4, 3, 2, 1, 0, 3, 2, 1, 0]
    assert R.reflect_pad(np.array([7.0], np.float32), 3).tolist() == [7.0] * 7


def test_too_short_raises():
    with pytest.raises(ValueError):
        R.stft(np.zeros(100, np.float32), R.whisper_hann_window(400), 400, 160, center=False)
    with pytest.raises(ValueError):
        R.s3gen_mel_spectrogram(np.zeros(600, np.float32))


def test_htk_bank_has_the_all_zero_filter():
    # one integer-bin HTK triangle is identically zero; reproduce, do not fix
    fb = R.mel_filters_htk(16000, 512, 80, 20.0, 8000.0)
    assert (np.abs(fb).sum(axis=0) == 0).sum() == 1
    for nm in (80, 128, 40):
        f = R.mel_filters(16000, 400, nm, 0.0, 8000.0)
        assert ((f != 0).sum(axis=0) <= 2).all()


def test_vocoder_heads_compose_the_pieces():
    rng = np.random.default_rng(5)
    h = rng.normal(0.0, 1.5, (2, 18, 40)).astype(np.float32)
    w = R.hann_window_periodic(16)
    y = R.hift_head_istft(h, 16, 4, w, 0.5)
    ref = np.clip(R.istft_hifigan(np.exp(h[:, :9]), np.sin(h[:, 9:]), 16, 4, w), -0.5, 0.5)
    assert np.array_equal(y, ref.astype(np.float32)) and np.abs(y).max() <= 0.5
    x = rng.normal(0.0, 1.5, (1, 22, 30)).astype(np.float32)
    assert np.array_equal(R.kokoro_head_istft(x), R.kokoro_inverse(np.exp(x[:, :11]), np.sin(x[:, 11:])))



def test_whisper_mel_segment_shapes_and_padding():
    mel = np.arange(50 * 4, dtype=np.float32).reshape(50, 4) / 7
    seg = R.whisper_mel_segment(mel, 45, 48, length=10)
    assert seg.dtype == np.float16 and seg.shape == (10, 4)
    assert np.array_equal(seg[:3], mel[45:48].astype(np.float16)) and not seg[3:].any()
    assert not R.whisper_mel_segment(mel, 48, 48, length=10).any()
    assert np.array_equal(R.whisper_mel_segment(mel, 0, 48, length=10), mel[:10].astype(np.float16))



def test_resample_audio_matches_torch_interpolate():
    import torch
    rng = np.random.default_rng(9)
    x = rng.standard_normal(4801).astype(np.float32)
    for fr, to in ((24000, 16000), (16000, 24000), (22050, 16000)):
        got = R.resample_audio(x, fr, to)
        ref = torch.nn.functional.interpolate(torch.from_numpy(x)[None, None], size=got.shape[-1], mode="linear", align_corners=False)[0, 0].numpy()
        assert got.shape == (int(np.float32(4801) * (np.float32(to) / np.float32(fr))),)
        # same rule as torch (the reference's comment cites it) except at the very end, where the reference clips the source
        # index to T - 1.001 instead of T - 1; elsewhere only the fp32 rounding of the source index differs
        # (near sample 4800 an fp32 index has a spacing of 4.9e-4, i.e. the interpolation weight is only good to ~2.4e-4)
        assert np.abs(got[:-2] - ref[:-2]).max() <= 2e-3
    assert R.resample_audio(x, 16000, 16000) is not None and np.array_equal(R.resample_audio(x, 16000, 16000), x)


def test_golden_adjacent_rows():
    a = np.load(os.path.join(os.path.dirname(__file__), "golden", "oracle_fp32_v2_adjacent.npz"))
    g = np.load(GOLD)
    w16 = R.hann_window_periodic(16)
    assert np.abs(R.hift_head_istft(a["in_h16"], 16, 4, w16, 0.99) - a["hift_head_istft"]).max() <= 1e-6
    assert np.abs(R.kokoro_head_istft(a["in_h20"]) - a["kokoro_head_istft"]).max() <= 1e-6
    assert np.array_equal(R.whisper_mel_segment(a["in_mel"], 0, 50, length=64).view(np.uint16), a["mel_segment_seek0"].view(np.uint16))
    assert np.array_equal(R.whisper_mel_segment(a["in_mel"], 37, 50, length=64).view(np.uint16), a["mel_segment_seek37"].view(np.uint16))
    assert np.array_equal(R.resample_audio(g["in_x24"], 24000, 16000), a["resample_24k_16k"])
    assert np.array_equal(R.resample_audio(g["in_x16"], 16000, 24000), a["resample_16k_24k"])


# ---- independent pins: public implementations of the published algorithms the reference restates -------------------
# The reference ships no golden vectors, so the oracle is additionally pinned against two libraries that implement
# the same published definitions independently of it: torchaudio (melscale_fbanks, Kaldi-compatible framing) and Hugging Face
# transformers (the OpenAI Whisper log-mel pipeline).

def test_filterbanks_match_torchaudio_and_transformers():
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    # funASRMelFilters (FunASRAudio.swift:322-396) is torchaudio's melscale_fbanks evaluated on nFft/2 = 200 grid points
    fb = ta.functional.melscale_fbanks(200, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="htk").numpy().T
    assert np.abs(R.funasr_mel_filters(16000, 400, 80) - fb).max() <= 5e-7
    # melFilters (S3TokenizerUtils.swift:301-375) is the Slaney bank of librosa / OpenAI Whisper
    for n_mels in (80, 128):
        fb = ta.functional.melscale_fbanks(201, 0.0, 8000.0, n_mels, 16000, norm="slaney", mel_scale="slaney").numpy().T
        assert np.abs(R.mel_filters(16000, 400, n_mels, 0.0, 8000.0) - fb).max() <= 1e-6
    au = pytest.importorskip("transformers.audio_utils")
    fb = au.mel_filter_bank(num_frequency_bins=961, num_mel_filters=80, min_frequency=0.0, max_frequency=8000.0, sampling_rate=24000,
                            norm="slaney", mel_scale="slaney").T
    assert np.abs(R.mel_filters(24000, 1920, 80, 0.0, 8000.0) - fb).max() <= 1e-6


@pytest.mark.parametrize("which", ["chatterbox", "whisper"])
def test_log_mel_matches_the_openai_whisper_pipeline(which):
    au = pytest.importorskip("transformers.audio_utils")
    x = synth.pcm(1, 16000 * 6 + 123, seed=31)[0]
    bank = au.mel_filter_bank(num_frequency_bins=201, num_mel_filters=128, min_frequency=0.0, max_frequency=8000.0, sampling_rate=16000,
                              norm="slaney", mel_scale="slaney")
    # S3Tokenizer / Chatterbox use the periodic Hann window of OpenAI Whisper; the reference's own Whisper front end deviates
    # from OpenAI only in the window (symmetric Hann, WhisperAudio.swift:32-44), so the same pipeline with that window pins it too
    n = np.arange(400)
    w = au.window_function(400, "hann") if which == "chatterbox" else 0.5 * (1.0 - np.cos(2.0 * np.pi * n / 399.0))
    ls = au.spectrogram(x.astype(np.float64), w, frame_length=400, hop_length=160, power=2.0, mel_filters=bank, log_mel="log10")[:, :-1]
    ls = (np.maximum(ls, ls.max() - 8.0) + 4.0) / 4.0
    if which == "chatterbox":
        o32, o64 = R.log_mel_spectrogram_chatterbox(x, 128), R.log_mel_spectrogram_chatterbox(x, 128, dt=np.float64)
    else:
        o32, o64 = R.whisper_log_mel_spectrogram(x, 128).T, R.whisper_log_mel_spectrogram(x, 128, dt=np.float64).T
    assert o64.shape == ls.shape
    assert np.abs(o64 - ls).max() <= 1e-6      # same algorithm in fp64: only the fp32 filterbank tables differ
    assert np.abs(o32 - ls).max() <= 2e-4      # the fp32 restatement (what the GPU path is compared with)


def test_kaldi_framing_matches_torchaudio():
    torch = pytest.importorskip("torch")
    kaldi = pytest.importorskip("torchaudio.compliance.kaldi")
    x = synth.pcm(1, 16000 + 77, seed=33)[0]
    # snip-edges framing, per-frame DC removal, 0.97 pre-emphasis and the Povey window of kaldiFbankCAMPPlus
    # (CAMPPlus.swift:32-106) are Kaldi's; torchaudio's Kaldi-compatible front end gives the same windowed frames
    # (the two treat the first sample of a frame differently, but the Povey window is zero there)
    frames, _ = kaldi._get_window(torch.from_numpy(x), 512, 400, 160, "povey", 0.42, True, True, 0.0, 0.0, True, 0.97)
    n_frames = (len(x) - 400) // 160 + 1
    assert frames.shape == (n_frames, 512)
    fr = np.stack([x[f * 160:f * 160 + 400] for f in range(n_frames)]).astype(np.float32)
    fr = fr - fr.mean(axis=1, keepdims=True)
    pre = fr.copy()
    pre[:, 1:] = fr[:, 1:] - np.float32(0.97) * fr[:, :-1]
    mine = np.pad(pre * R.povey_window(400), ((0, 0), (0, 112)))
    assert np.abs(frames.numpy() - mine).max() <= 1e-6
    # ... and the oracle's fbank is the log of (|rfft|^2 of exactly these frames) through the reference's integer-bin HTK bank
    power = np.abs(np.fft.rfft(mine.astype(np.float64), axis=1)) ** 2
    want = np.log(np.maximum(power @ R.mel_filters_htk(16000, 512, 80, 20.0, 8000.0).astype(np.float64), 1.1920929e-07))
    assert np.abs(R.kaldi_fbank_camp_plus(x) - want).max() <= 2e-3   # fp32 restatement vs fp64 on un-clamped logs


def test_s3gen_trim_fade_window():
    # S3Gen.swift:259-262: 20 ms of zeros, then a raised-cosine ramp from 0 to 1 over 20 ms
    f = R.s3gen_trim_fade(24000)
    assert f.shape == (960,) and not np.any(f[:480])
    assert f[480] == 0.0 and f[-1] == 1.0 and np.all(np.diff(f[480:]) >= 0)
    assert abs(float(f[480 + 240]) - 0.5) < 5e-3
    y = np.ones((2, 1000), np.float32)
    assert np.array_equal(R.apply_trim_fade(y, f)[:, :960], np.broadcast_to(f, (2, 960)))
    assert np.array_equal(R.apply_trim_fade(y[:, :900], f), y[:, :900])      # shorter than the window: untouched


def test_s3gen_and_voice_encoder_mels_match_torchaudio():
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    # s3genMelSpectrogram (S3GenMel.swift:43-102) is the Matcha / HiFi-GAN mel: reflect pad (n_fft - hop) / 2, periodic Hann,
    # magnitude spectrum, Slaney bank 0..8000 Hz, ln(max(., 1e-5)) -- torchaudio's MelSpectrogram with power 1 on the padded signal
    y = synth.pcm(2, 24000 + 480 * 3, sample_rate=24000, seed=41)
    t = torch.from_numpy(y)
    padded = torch.nn.functional.pad(t[:, None, :], (720, 720), mode="reflect")[:, 0]
    ms = ta.transforms.MelSpectrogram(sample_rate=24000, n_fft=1920, win_length=1920, hop_length=480, f_min=0.0, f_max=8000.0, n_mels=80,
                                      power=1.0, center=False, norm="slaney", mel_scale="slaney")
    want = torch.log(torch.clamp(ms(padded.double().float()), min=1e-5)).numpy()
    got = R.s3gen_mel_spectrogram(y)
    assert got.shape == want.shape == (2, 80, 53)
    assert np.abs(got - want).max() <= 2e-4 * max(1.0, np.abs(want).max())
    # voiceEncoderMelspectrogram (VoiceEncoderMelspec.swift:17-68): centred reflect-padded power mel, 40 Slaney filters, no log
    x = synth.pcm(1, 16000 + 55, seed=42)[0]
    ve = ta.transforms.MelSpectrogram(sample_rate=16000, n_fft=400, win_length=400, hop_length=160, f_min=0.0, f_max=8000.0, n_mels=40,
                                      power=2.0, center=True, pad_mode="reflect", norm="slaney", mel_scale="slaney")
    want = ve(torch.from_numpy(x)).numpy()
    got = R.voice_encoder_melspectrogram(x)
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


def test_funasr_log_mel_and_vocoder_stfts_match_torch():
    torch = pytest.importorskip("torch")
    ta = pytest.importorskip("torchaudio")
    # funASRLogMelSpectrogram (FunASRAudio.swift:57-94): centred STFT with a symmetric Hamming window, |X|^2 of bins 0..199, the
    # 200-point torchaudio-style HTK bank, natural log floored at 1e-10 -- composed here from torch.stft and torchaudio's bank
    x = synth.pcm(1, 16000 * 2 + 9, seed=51)[0]
    st = torch.stft(torch.from_numpy(x).double(), 400, 160, window=torch.hamming_window(400, periodic=False, dtype=torch.float64),
                    center=True, pad_mode="reflect", return_complex=True)
    bank = ta.functional.melscale_fbanks(200, 0.0, 8000.0, 80, 16000, norm="slaney", mel_scale="htk").double()
    want = torch.log(torch.clamp((st.abs() ** 2)[:200].T @ bank, min=1e-10)).numpy()
    got64 = R.funasr_log_mel_spectrogram(x, dt=np.float64)
    assert got64.shape == want.shape == (1 + len(x) // 160, 80)
    assert np.abs(got64 - want).max() <= 2e-4            # (fp32 filterbank tables on the torchaudio side)
    # fp32 restatement: its fp32 bank differs from torchaudio's by ~1e-7 absolute, which a strong tone sitting on a filter's edge
    # (weight ~1e-6, power ~1e3) turns into ~1e-3 of an un-clamped log: table rounding, not the transform
    assert np.abs(R.funasr_log_mel_spectrogram(x) - want).max() <= 2e-3
    # stftHiFiGAN (reflect), cosyVoice3Stft (zero pad) and MLXSTFT.transform: torch.stft with n_fft 16 / 20 and periodic Hann
    sig = synth.pcm(2, 2403, sample_rate=24000, seed=52)
    ts = torch.from_numpy(sig).double()
    w16 = torch.hann_window(16, periodic=True, dtype=torch.float64)
    for fn, mode in ((R.stft_hifigan, "reflect"), (R.cosyvoice3_stft, "constant")):
        re, im = fn(sig, 16, 4, R.hann_window_periodic(16))
        ref = torch.stft(ts, 16, 4, window=w16, center=True, pad_mode=mode, return_complex=True).numpy()
        assert re.shape == ref.shape
        assert np.abs((re + 1j * im) - ref).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    mag, ph = R.kokoro_transform(sig)
    ref = torch.stft(ts, 20, 5, window=torch.hann_window(20, periodic=True, dtype=torch.float64), center=True, pad_mode="reflect",
                     return_complex=True).numpy()
    assert np.abs(mag - np.abs(ref)).max() <= 2e-6 * max(1.0, np.abs(ref).max())
    strong = np.abs(ref) > 1e-2          # the phase of a near-zero bin is ill-conditioned
    d = np.angle(np.exp(1j * (ph - np.angle(ref))))
    assert np.abs(d[strong]).max() <= 1e-4
