"""GPU tests of the dynamic tile walk of the tiled front-end kernels (frontend.cu: a persistent CTA takes its next tile from a launch-wide
atomic counter; ragged batches ask two tiles ahead; Whisper's zero tail shortens the walk) against the static walk it replaces
(b2a_debug_dyn_tiles switches): every tile is computed the same way whoever computes it, so the results must be bit-identical -- on batches of
more than two (ragged: three) rounds of the 444 / 296 persistent CTAs, where the counter is actually used -- and a few clips are also
checked against the oracle (WhisperAudio.swift:78-137, S3TokenizerUtils.swift:160-208, FunASRAudio.swift:197-216, CAMPPlus.swift:32-106,
VoiceEncoderMelspec.swift:17-68)."""
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


@pytest.fixture()
def both(ctx):
    def run(fn):
        try:
            ctx.lib.b2a_debug_dyn_tiles(1)
            dyn = fn()
            ctx.lib.b2a_debug_dyn_tiles(0)
            sta = fn()
        finally:
            ctx.lib.b2a_debug_dyn_tiles(1)
        return dyn, sta
    return run


def _tiled(b, n, seed, distinct=6):
    """b clips of n samples: `distinct` different ones, repeated (keeps the host-side synthesis short)"""
    x = synth.pcm(distinct, n, seed=seed)
    return np.ascontiguousarray(x[np.arange(b) % distinct])


def test_whisper_equal_lengths(api, ctx, both):
    x = _tiled(24, 16000 * 30, 5001)            # 24 x 94 = 2256 tiles
    for n_mels in (128, 80):
        dyn, sta = both(lambda: api.whisperLogMelSpectrogram(x, nMels=n_mels, ctx=ctx))
        assert np.array_equal(dyn, sta)
        assert_feat_close(dyn[7], R.whisper_log_mel_spectrogram(x[7], n_mels), what=f"whisper {n_mels} under the dynamic walk")
    d16, s16 = both(lambda: api.whisperLogMelSpectrogramF16(x, nMels=128, ctx=ctx))
    assert np.array_equal(d16.view(np.uint16), s16.view(np.uint16))
    assert np.array_equal(d16.view(np.uint16), api.whisperLogMelSpectrogram(x, nMels=128, ctx=ctx).astype(np.float16).view(np.uint16))


def test_whisper_zero_tail_shortened_walk(api, ctx, both):
    # 30 s + 30 s of padding as WhisperSTT calls it (WhisperSTT.swift:139-144): 6000 frames, the walk covers the first 94 of 188 tiles
    x = _tiled(12, 16000 * 30, 5002)
    dyn, sta = both(lambda: api.whisperLogMelSpectrogram(x, nMels=128, padding=480000, ctx=ctx))   # static switch = the ZS instantiation
    assert dyn.shape == (12, 6000, 128) and np.array_equal(dyn, sta)
    explicit = api.whisperLogMelSpectrogram(np.concatenate([x[:2], np.zeros((2, 480000), np.float32)], axis=1), nMels=128, ctx=ctx)
    assert np.array_equal(dyn[:2], explicit)
    # content that ends inside a tile, a short tail, and a tail shorter than the reflection (falls back to the per-tile test)
    y = _tiled(40, 16000 * 7 + 123, 5003)
    for pad in (16000 * 9, 300, 150):
        dyn, sta = both(lambda: api.whisperLogMelSpectrogram(y, nMels=128, padding=pad, ctx=ctx))
        assert np.array_equal(dyn, sta)
        assert_feat_close(dyn[3], R.whisper_log_mel_spectrogram(y[3], 128, padding=pad), what=f"whisper padding {pad}")


def test_mt_layout_lfr_kaldi_voice_encoder(api, ctx, both):
    x = _tiled(40, 16000 * 10, 5004)             # 40 x 32 tiles
    dyn, sta = both(lambda: api.logMelSpectrogramChatterbox(x, ctx=ctx))
    assert np.array_equal(dyn, sta)
    assert_feat_close(dyn[11], R.log_mel_spectrogram_chatterbox(x[11]), what="chatterbox under the dynamic walk")
    dyn, sta = both(lambda: api.voiceEncoderMelspectrogram(x, ctx=ctx))
    assert np.array_equal(dyn, sta)
    dyn, sta = both(lambda: api.preprocessAudio(x, ctx=ctx))
    assert np.array_equal(dyn, sta)
    assert_feat_close(dyn[39], R.preprocess_audio(x[39]), tol=2e-4, what="preprocessAudio under the dynamic walk")
    dyn, sta = both(lambda: api.kaldiFbankCAMPPlus(x, meanNorm=True, ctx=ctx))
    assert np.array_equal(dyn, sta)
    assert_feat_close(dyn[0], R.kaldi_fbank_mean_norm(R.kaldi_fbank_camp_plus(x[0])), tol=2e-4, what="fbank under the dynamic walk")
    # a bank that is not baked (64 mels): the run-time-configured kernel walks the same way
    dyn, sta = both(lambda: api.whisperLogMelSpectrogram(x, nMels=64, ctx=ctx))
    assert np.array_equal(dyn, sta)


def test_ragged_two_tiles_ahead(api, ctx, both):
    rng = np.random.default_rng(5005)
    b, n = 96, 16000 * 20
    x = _tiled(b, n, 5006, distinct=8)
    lengths = [int(v) for v in rng.integers(16000 * 3, n + 1, size=b)]
    lengths[0], lengths[1], lengths[-1] = n, 400, 16000 * 3 + 1     # longest first, a one-frame clip, an odd one last
    for b_, ln in enumerate(lengths):
        x[b_, ln:] = 5.0
    (dyn, rows_d), (sta, rows_s) = both(lambda: api.whisperLogMelSpectrogramRagged(x, lengths, nMels=128, ctx=ctx))
    assert list(rows_d) == list(rows_s) and np.array_equal(np.asarray(dyn), np.asarray(sta))
    for i in (0, 1, 17, b - 1):
        want = R.whisper_log_mel_spectrogram(x[i, :lengths[i]], 128)
        assert_feat_close(np.asarray(dyn)[i, :rows_d[i]], want, what=f"ragged whisper clip {i} under the dynamic walk")
    for fn in (lambda: api.logMelSpectrogramChatterboxRagged(x, lengths, ctx=ctx), lambda: api.preprocessAudioRagged(x, lengths, ctx=ctx),
               lambda: api.kaldiFbankCAMPPlusRagged(x, lengths, meanNorm=True, ctx=ctx), lambda: api.funASRLogMelSpectrogramRagged(x, lengths, ctx=ctx)):
        (dyn, rows_d), (sta, rows_s) = both(fn)
        assert list(rows_d) == list(rows_s) and np.array_equal(np.asarray(dyn), np.asarray(sta))


def test_tiles_below_the_clamp_threshold_are_filled(api, ctx):
    # whisper_clamp_kernel writes the threshold without reading a tile whose largest value does not exceed it (per-tile maxima from the
    # main kernel): a loud stretch, a stretch 120 dB quieter (non-zero, every value below Lmax - 8), a stretch of exact zeros, and a loud
    # burst at the very end so that tiles of all three kinds sit inside the clip.  max(v, thr) == thr exactly, so the features must agree
    # with the oracle like any other clip, in every layout / type, alone and ragged.
    rng = np.random.default_rng(5100)
    n = 16000 * 20
    x = np.zeros((5, n), np.float32)
    for b in range(5):
        loud = (0.3 * rng.standard_normal(n)).astype(np.float32)
        x[b, :16000 * 6] = loud[:16000 * 6]
        x[b, 16000 * 6:16000 * 12] = 1e-6 * loud[16000 * 6:16000 * 12]
        x[b, 16000 * 19:] = loud[16000 * 19:]
    for n_mels in (128, 80):
        got = api.whisperLogMelSpectrogram(x, nMels=n_mels, ctx=ctx)
        for b in (0, 4):
            want = R.whisper_log_mel_spectrogram(x[b], n_mels)
            assert_feat_close(got[b], want, what=f"whisper {n_mels}: quiet / silent tiles")
            thr = want.max() - 2.0
            assert np.all(got[b, 700:1100] == got[b, 700, 0]) and abs(float(got[b, 700, 0]) - thr) <= 1e-4   # the quiet stretch: one constant
            assert np.all(got[b, 1300:1800] == got[b, 700, 0])                                               # the zeros: the same constant
    f16 = api.whisperLogMelSpectrogramF16(x, nMels=128, ctx=ctx)
    assert np.array_equal(f16.view(np.uint16), api.whisperLogMelSpectrogram(x, nMels=128, ctx=ctx).astype(np.float16).view(np.uint16))
    mt = api.logMelSpectrogramChatterbox(x, ctx=ctx)
    assert_feat_close(mt[2], R.log_mel_spectrogram_chatterbox(x[2]), what="chatterbox: quiet / silent tiles")
    lengths = [n, 16000 * 13, 16000 * 7 + 5, n - 1, 16000 * 12]
    rg, rows = api.whisperLogMelSpectrogramRagged(x, lengths, nMels=128, ctx=ctx)
    for b, ln in enumerate(lengths):
        assert_feat_close(np.asarray(rg)[b, :rows[b]], R.whisper_log_mel_spectrogram(x[b, :ln], 128), what=f"ragged clip {b}: quiet / silent tiles")
