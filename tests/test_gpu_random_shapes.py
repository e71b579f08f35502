"""Seeded random shapes against the oracle for the round-2 paths whose index arithmetic is new: the warp-per-frame n_fft 1920 front end
(equal-length, ragged), and the Whisper / Chatterbox front ends with a zero tail (skipped tiles, clamp fill) -- clip lengths and paddings
that are not aligned to anything."""
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


def test_s3gen_random_lengths(api, ctx):
    rng = np.random.default_rng(4101)
    for trial in range(10):
        n = int(rng.integers(601, 40000))
        b = int(rng.integers(1, 5))
        x = synth.pcm(b, n, sample_rate=24000, seed=4200 + trial, zero_tail_frac=0.0)
        if n <= 600:
            continue
        want = R.s3gen_mel_spectrogram(x)
        got = api.s3genMelSpectrogram(x, ctx=ctx)
        assert got.shape == want.shape, (n, b)
        assert_feat_close(got, want, what=f"s3gen mel, n = {n}, batch = {b}")


def test_s3gen_random_ragged_batches(api, ctx):
    rng = np.random.default_rng(4102)
    for trial in range(4):
        b = int(rng.integers(2, 7))
        lengths = [int(v) for v in rng.integers(721, 30000, b)]
        n = max(lengths)
        x = synth.pcm(b, n, sample_rate=24000, seed=4300 + trial, zero_tail_frac=0.0)
        got, rows = api.s3genMelSpectrogramRagged(x, lengths, ctx=ctx)
        g = np.asarray(got)
        for i, ln in enumerate(lengths):
            want = R.s3gen_mel_spectrogram(x[i:i + 1, :ln])[0]
            assert rows[i] == want.shape[1]
            assert_feat_close(g[i, :, :rows[i]], want, what=f"ragged s3gen trial {trial} clip {i} ({ln} samples)")
            assert not np.any(g[i, :, rows[i]:])


def test_whisper_random_lengths_and_paddings(api, ctx):
    rng = np.random.default_rng(4103)
    for trial in range(12):
        n = int(rng.integers(200, 60000))
        padding = int(rng.choice([0, int(rng.integers(1, 700)), int(rng.integers(700, 20000)), 480000]))
        n_mels = int(rng.choice([80, 128]))
        x = synth.pcm(2, n, seed=4400 + trial, zero_tail_frac=float(rng.choice([0.0, 0.3])))
        want = np.stack([R.whisper_log_mel_spectrogram(c, n_mels, padding=padding) for c in x])
        got = api.whisperLogMelSpectrogram(x, nMels=n_mels, padding=padding, ctx=ctx)
        assert got.shape == want.shape, (n, padding)
        assert_feat_close(got, want, what=f"whisper, n = {n}, padding = {padding}, {n_mels} mels")
        if trial % 3 == 0:
            want_c = np.stack([R.log_mel_spectrogram_chatterbox(c, 128, padding=padding) for c in x])
            assert_feat_close(api.logMelSpectrogramChatterbox(x, nMels=128, padding=padding, ctx=ctx), want_c,
                              what=f"chatterbox log-mel, n = {n}, padding = {padding}")
