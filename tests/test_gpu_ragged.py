"""GPU parity of the ragged-batch entry points (per-clip lengths, include/b200audio.h "ragged batches"): every clip of a batch
of unequal clips must come out exactly as the single-clip reference call on audio[b, :lengths[b]] (oracle), rows past a clip's
own frame count must be zero, and the reported counts must follow the reference's frame-count rules."""
import os

import numpy as np
import pytest

from oracle import reference_dsp as R
from tests import synth
from tests.test_gpu_parity import assert_feat_close

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(ctx):
    from mlx_swift_audio_b200 import api as A
    return A


def _batch(lengths, sr=16000, seed=5):
    n_max = max(lengths)
    x = np.zeros((len(lengths), n_max), np.float32)
    for b, n in enumerate(lengths):
        x[b, :n] = synth.pcm(1, n, sample_rate=sr, seed=seed + b, zero_tail_frac=0.1 if b % 2 else 0.0)[0]
        x[b, n:] = 7.0   # whatever lies past a clip's length must never be read
    return x


# around the kernel's tile (32 frames = 5120 samples) and row (160 samples) boundaries, a clip shorter than the reflect pad,
# one of a single frame, and the longest clip in the middle of the batch
LENGTHS = [5120, 163, 160 * 33 + 1, 48000 + 11, 401, 5119, 16000, 250, 31999]


def _check(got, rows, want_list, what, time_axis, tol=1e-4):
    got = np.asarray(got.cpu() if hasattr(got, "cpu") else got)
    for b, want in enumerate(want_list):
        t = want.shape[time_axis]
        assert rows[b] == t, f"{what}: clip {b} reports {rows[b]} rows, the reference rule gives {t}"
        sl = [b, slice(None), slice(None)]
        sl[1 + time_axis] = slice(0, t)
        assert_feat_close(got[tuple(sl)], want, tol=tol, what=f"{what} clip {b} (n={LENGTHS[b]})")
        sl[1 + time_axis] = slice(t, None)
        assert not np.any(got[tuple(sl)]), f"{what}: clip {b} has non-zero rows past its frame count"


@pytest.mark.parametrize("space", ["host", "device"])
def test_whisper_ragged(api, ctx, space):
    x = _batch(LENGTHS)
    xin = x
    if space == "device":
        import torch
        xin = torch.from_numpy(x).cuda()
    for n_mels, padding in ((128, 0), (80, 3000)):
        got, rows = api.whisperLogMelSpectrogramRagged(xin, LENGTHS, nMels=n_mels, padding=padding, ctx=ctx)
        if space == "device":
            ctx.sync()
        _check(got, rows, [R.whisper_log_mel_spectrogram(x[b, :n], n_mels, padding=padding) for b, n in enumerate(LENGTHS)],
               f"whisper ragged {n_mels}", 0)


def test_chatterbox_ragged(api, ctx):
    x = _batch(LENGTHS, seed=9)
    got, rows = api.logMelSpectrogramChatterboxRagged(x, LENGTHS, nMels=128, ctx=ctx)
    _check(got, rows, [R.log_mel_spectrogram_chatterbox(x[b, :n], 128) for b, n in enumerate(LENGTHS)], "chatterbox ragged", 1)


def test_funasr_ragged(api, ctx):
    x = _batch(LENGTHS, seed=11)
    got, rows = api.preprocessAudioRagged(x, LENGTHS, ctx=ctx)
    # CMVN divides by per-column statistics: 2e-4; a clip with a single LFR row has zero variance
    want = [R.preprocess_audio(x[b, :n]) for b, n in enumerate(LENGTHS)]
    keep = [b for b, w in enumerate(want) if w.shape[0] > 2]
    g = np.asarray(got)
    for b in keep:
        t = want[b].shape[0]
        assert rows[b] == t
        assert_feat_close(g[b, :t], want[b], tol=2e-4, what=f"funasr ragged clip {b}")
        assert not np.any(g[b, t:])
    got2, rows2 = api.preprocessAudioRagged(x, LENGTHS, applyNormalization=False, ctx=ctx)
    _check(got2, rows2, [R.preprocess_audio(x[b, :n], apply_normalization=False) for b, n in enumerate(LENGTHS)], "funasr LFR ragged", 0)


def test_funasr_log_mel_and_voice_encoder_ragged(api, ctx):
    x = _batch(LENGTHS, seed=29)
    got, rows = api.funASRLogMelSpectrogramRagged(x, LENGTHS, ctx=ctx)
    _check(got, rows, [R.funasr_log_mel_spectrogram(x[b, :n]) for b, n in enumerate(LENGTHS)], "funasr log-mel ragged", 0)
    got, rows = api.voiceEncoderMelspectrogramRagged(x, LENGTHS, ctx=ctx)
    want = [R.voice_encoder_melspectrogram(x[b, :n]) for b, n in enumerate(LENGTHS)]
    g = np.asarray(got)
    for b, w in enumerate(want):   # linear-power mel (no log): relative to the clip's peak like the non-ragged test
        t = w.shape[1]
        assert rows[b] == t
        assert np.abs(g[b, :, :t] - w).max() <= 1e-4 * max(np.abs(w).max(), 1e-30), f"voice encoder ragged clip {b}"
        assert not np.any(g[b, :, t:])


def test_kaldi_ragged(api, ctx):
    lengths = [n for n in LENGTHS if n >= 400]
    x = _batch(lengths, seed=13)
    got, rows = api.kaldiFbankCAMPPlusRagged(x, lengths, ctx=ctx)
    g = np.asarray(got)
    for b, n in enumerate(lengths):
        want = R.kaldi_fbank_camp_plus(x[b, :n])
        assert rows[b] == want.shape[0]
        assert_feat_close(g[b, :rows[b]], want, what=f"kaldi ragged clip {b}")
        assert not np.any(g[b, rows[b]:])
    got, rows = api.kaldiFbankCAMPPlusRagged(x, lengths, meanNorm=True, ctx=ctx)
    g = np.asarray(got)
    for b, n in enumerate(lengths):
        want = R.kaldi_fbank_mean_norm(R.kaldi_fbank_camp_plus(x[b, :n]))
        assert_feat_close(g[b, :rows[b]], want, tol=2e-4, what=f"kaldi mean-norm ragged clip {b}")


@pytest.mark.parametrize("longest", [100000, 500000])
def test_kaldi_mean_norm_ragged_long_clips(api, ctx, longest):
    # the statistics kernel is chosen by the LONGEST clip's row count: 623 rows -> the shared-memory-resident kernel, 3123 rows -> the
    # streaming kernels; both take the per-clip row counts from the clip table
    lengths = [5120, longest, 48011, 401, 16000]
    x = _batch(lengths, seed=17)
    got, rows = api.kaldiFbankCAMPPlusRagged(x, lengths, meanNorm=True, ctx=ctx)
    g = np.asarray(got)
    for b, n in enumerate(lengths):
        want = R.kaldi_fbank_mean_norm(R.kaldi_fbank_camp_plus(x[b, :n]))
        assert rows[b] == want.shape[0]
        assert_feat_close(g[b, :rows[b]], want, tol=2e-4, what=f"kaldi mean-norm ragged clip {b} (longest {longest})")
        assert not np.any(g[b, rows[b]:])


def test_s3gen_ragged(api, ctx):
    lengths = [24000, 480 * 17, 24000 * 2 + 5, 1921, 9600]
    x = _batch(lengths, sr=24000, seed=17)
    got, rows = api.s3genMelSpectrogramRagged(x, lengths, ctx=ctx)
    g = np.asarray(got)
    for b, n in enumerate(lengths):
        want = R.s3gen_mel_spectrogram(x[b:b + 1, :n])[0]
        assert rows[b] == want.shape[1]
        assert_feat_close(g[b, :, :rows[b]], want, what=f"s3gen ragged clip {b}")
        assert not np.any(g[b, :, rows[b]:])


def test_ragged_host_pipeline_chunks(api, ctx, monkeypatch):
    # 1 MB chunks: the batch crosses several chunks of the host pipeline, each with its own clip / tile tables
    monkeypatch.setenv("B2A_HOST_CHUNK_MB", "1")
    rng = np.random.default_rng(3)
    lengths = [int(v) for v in rng.integers(300, 40000, 24)]
    x = _batch(lengths, seed=23)
    got, rows = api.whisperLogMelSpectrogramRagged(x, lengths, nMels=80, ctx=ctx)
    g = np.asarray(got)
    for b, n in enumerate(lengths):
        want = R.whisper_log_mel_spectrogram(x[b, :n], 80)
        assert rows[b] == want.shape[0]
        assert_feat_close(g[b, :rows[b]], want, what=f"chunked ragged clip {b}")
        assert not np.any(g[b, rows[b]:])


def test_ragged_rejects_bad_lengths(api, ctx):
    x = _batch([4000, 4000])
    with pytest.raises(api.B2AError):
        api.whisperLogMelSpectrogramRagged(x, [4000, 4001], nMels=80, ctx=ctx)
    with pytest.raises(api.B2AError):
        api.whisperLogMelSpectrogramRagged(x, [4000, 0], nMels=80, ctx=ctx)
    with pytest.raises(api.B2ATooShort):
        api.kaldiFbankCAMPPlusRagged(x, [4000, 100], ctx=ctx)
