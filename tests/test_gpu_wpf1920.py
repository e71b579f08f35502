"""GPU parity tests of the warp-per-frame n_fft 1920 front end (csrc/wpf1920.cu; S3GenMel.swift:43-102) against the oracle, next to the
tiled lane == frame kernel it replaces for equal-length batches (b2a_debug_wpf1920 switches between the two): frame counts around the
8-frame strips, clips shorter than the reflect pad, device buffers that are not 128-byte aligned, the power / dB / normalised branches
through a 1920-point voice-encoder configuration (rotation 0), and the launch counter (one kernel per call)."""
import ctypes as C

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
def both_kernels(ctx):
    """Runs fn() under the warp-per-frame kernel and under the tiled kernel; leaves the default (warp-per-frame) switched on."""
    def run(fn):
        try:
            ctx.lib.b2a_debug_wpf1920(1)
            new = fn()
            ctx.lib.b2a_debug_wpf1920(0)
            old = fn()
        finally:
            ctx.lib.b2a_debug_wpf1920(1)
        return new, old
    return run


# frames = 1 + n // 480 for n >= 721 (pad 720 each side): 2, 7, 8, 9, 16, 17, 25 frames and two long clips
@pytest.mark.parametrize("n", [721, 480 * 6 + 5, 480 * 7, 480 * 8 + 479, 480 * 15 + 1, 480 * 16, 480 * 24 + 100, 24000 * 2 + 17, 24000 * 3])
def test_s3gen_mel_strip_boundaries(api, ctx, both_kernels, n):
    x = synth.pcm(3, n, sample_rate=24000, seed=3001)
    want = R.s3gen_mel_spectrogram(x)
    new, old = both_kernels(lambda: api.s3genMelSpectrogram(x, ctx=ctx))
    assert new.shape == want.shape == old.shape
    assert_feat_close(new, want, what=f"s3gen mel (warp per frame), n = {n}")
    assert_feat_close(old, want, what=f"s3gen mel (tiled), n = {n}")


@pytest.mark.parametrize("n", [600 + 61, 700, 900, 1441, 1920])
def test_s3gen_mel_short_clips_truncated_reflection(api, ctx, n):
    # reflectPad2D truncates the reflection for clips of <= pad samples (S3GenMel.swift:17-25): every frame is an edge frame
    x = synth.pcm(2, n, sample_rate=24000, seed=3002, zero_tail_frac=0.0)
    want = R.s3gen_mel_spectrogram(x)
    got = api.s3genMelSpectrogram(x, ctx=ctx)
    assert_feat_close(got, want, what=f"s3gen mel short clip {n}")


def test_one_kernel_per_call_and_batch_invariance(api, ctx):
    x = synth.pcm(37, 24000 + 1000, sample_rate=24000, seed=3003)
    before = ctx.launch_count
    got = api.s3genMelSpectrogram(x, ctx=ctx)
    assert ctx.launch_count - before == 1, "the S3Gen mel is one kernel launch"
    for b in (0, 17, 36):
        alone = api.s3genMelSpectrogram(x[b], ctx=ctx)
        assert np.array_equal(alone, got[b]), "a clip's features do not depend on the batch around it"
    assert np.array_equal(api.s3genMelSpectrogram(x, ctx=ctx), got), "run-to-run bit identical"


def test_unaligned_device_buffers(api, ctx):
    torch = pytest.importorskip("torch")
    n = 24000 + 333
    x = synth.pcm(4, n, sample_rate=24000, seed=3004)
    want = R.s3gen_mel_spectrogram(x)
    for shift in (0, 1, 3, 16):
        buf = torch.zeros(4 * n + 64, dtype=torch.float32, device="cuda")
        view = buf[shift:shift + 4 * n].view(4, n)
        view.copy_(torch.from_numpy(x))
        got = api.s3genMelSpectrogram(view, ctx=ctx)
        assert_feat_close(got.cpu().numpy(), want, what=f"s3gen mel, device buffer shifted by {shift} floats")


@pytest.mark.parametrize("branch", ["amp", "db", "db_normalized"])
def test_voice_encoder_config_with_1920_point_frames(api, ctx, both_kernels, branch):
    # VoiceEncoderMelspec.swift:17-68 with nFft 1920 / hop 480: centre pad 960 (rotation 0), power spectrum, 40 mels, no log / dB / normalised
    from mlx_swift_audio_b200 import _lib as L
    cfg = L.VoiceEncConfig()
    ctx.lib.b2a_voice_enc_config_default(C.byref(cfg))
    cfg.n_fft, cfg.hop_size, cfg.win_size, cfg.sample_rate = 1920, 480, 1920, 24000
    kw = dict(n_fft=1920, hop_size=480, win_size=1920, sample_rate=24000)
    if branch != "amp":
        cfg.mel_type_db = 1
        kw["mel_type"] = "db"
    if branch == "db_normalized":
        cfg.normalized_mels = 1
        kw["normalized_mels"] = True
    x = synth.pcm(3, 24000 + 77, sample_rate=24000, seed=3005)
    want = np.stack([R.voice_encoder_melspectrogram(c, **kw) for c in x])
    new, old = both_kernels(lambda: api.voiceEncoderMelspectrogram(x, config=cfg, ctx=ctx))
    assert_feat_close(new, want, what=f"voice encoder 1920 ({branch}), warp per frame")
    assert_feat_close(old, want, what=f"voice encoder 1920 ({branch}), tiled")


def test_s3gen_ragged_batch_on_both_kernels(api, ctx, both_kernels):
    # per-clip lengths: the warp-per-frame kernel walks the tile table's 16-frame tiles with each clip's own length and frame count
    lengths = [24000, 480 * 17, 24000 * 2 + 5, 1921, 9600, 480 * 16, 480 * 15 + 479, 24000 * 2 + 5]
    n = max(lengths)
    x = synth.pcm(len(lengths), n, sample_rate=24000, seed=3006, zero_tail_frac=0.0)
    for b, ln in enumerate(lengths):
        x[b, ln:] = 7.0   # whatever lies past a clip's length must not be read
    (new, rows_new), (old, rows_old) = both_kernels(lambda: api.s3genMelSpectrogramRagged(x, lengths, ctx=ctx))
    assert list(rows_new) == list(rows_old)
    for b, ln in enumerate(lengths):
        want = R.s3gen_mel_spectrogram(x[b:b + 1, :ln])[0]
        assert rows_new[b] == want.shape[1]
        for got, name in ((np.asarray(new), "warp per frame"), (np.asarray(old), "tiled")):
            assert_feat_close(got[b, :, :rows_new[b]], want, what=f"s3gen ragged clip {b} ({name})")
            assert not np.any(got[b, :, rows_new[b]:]), "rows past a clip's own frame count are zero"


def test_s3gen_pure_tone_against_fp64_truth(api, ctx, both_kernels):
    # The reference's own test input (1 s, 440 Hz).  Un-clamped ln features of a pure tone reach bins ~100 dB below the peak, where fp32
    # holds rounding noise only: as for the Fun-ASR pure-tone case (exception 1) the yardstick is fp64 truth and the
    # error the fp32 oracle itself has there -- for BOTH kernels, and the warp-per-frame kernel must not be the noisier one by more than 2x.
    t = np.arange(24000, dtype=np.float64) / 24000.0
    x = np.sin(2 * np.pi * 440.0 * t).astype(np.float32)
    t64 = R.s3gen_mel_spectrogram(x.astype(np.float64), dt=np.float64)
    o32 = R.s3gen_mel_spectrogram(x)
    err = lambda a: float(np.max(np.abs(np.asarray(a, np.float64) - t64) / np.maximum(1.0, np.abs(t64))))
    e32 = err(o32)
    new, old = both_kernels(lambda: api.s3genMelSpectrogram(x, ctx=ctx))
    bar = max(1e-4, 4.0 * e32)
    assert err(new) <= bar and err(old) <= bar, (err(new), err(old), e32)
    assert err(new) <= max(1e-4, 2.0 * err(old)), (err(new), err(old))


def test_pruned_stage_b_is_bit_identical(api, ctx):
    # S3Gen's bank ends at bin 640 of 961 (fmax 8000 Hz at 24 kHz), so the kernel skips the ten outputs of every 32-point stage-B transform
    # that only feed bins 641..960 (wpf1920.cu, PRUNE); b2a_debug_wpf1920(2) runs the same kernel with every bin formed.  Interior frames,
    # edge frames (staged through the exchange buffer, whose pad words lie in the pruned band) and ragged tiles must agree bit for bit,
    # and a NaN sample in an edge frame must not leak into later frames of the same warp.
    x = synth.pcm(5, 24000 * 2 + 123, sample_rate=24000, seed=3007)
    short = synth.pcm(3, 1441, sample_rate=24000, seed=3008, zero_tail_frac=0.0)
    lengths = [24000, 480 * 17, 24000 * 2 + 5, 1921, 9600]
    bad = x.copy()
    bad[:, 3] = np.nan   # frames 0 and 1 of every clip (reflect pad 720: sample 3 also appears mirrored in frame 0)
    try:
        ctx.lib.b2a_debug_wpf1920(1)
        p = [api.s3genMelSpectrogram(x, ctx=ctx), api.s3genMelSpectrogram(short, ctx=ctx),
             np.asarray(api.s3genMelSpectrogramRagged(x[:, :24000 * 2 + 5], lengths, ctx=ctx)[0]), api.s3genMelSpectrogram(bad, ctx=ctx)]
        ctx.lib.b2a_debug_wpf1920(2)
        u = [api.s3genMelSpectrogram(x, ctx=ctx), api.s3genMelSpectrogram(short, ctx=ctx),
             np.asarray(api.s3genMelSpectrogramRagged(x[:, :24000 * 2 + 5], lengths, ctx=ctx)[0]), api.s3genMelSpectrogram(bad, ctx=ctx)]
    finally:
        ctx.lib.b2a_debug_wpf1920(1)
    for a, b, name in zip(p, u, ("interior", "short clips", "ragged", "NaN sample")):
        assert np.array_equal(a, b, equal_nan=True), f"pruned and unpruned stage B differ: {name}"
    # frames that do not contain sample 3 (frame t covers samples 480 t - 720 .. 480 t + 1199): untouched by the NaN
    assert np.array_equal(p[3][:, :, 2:], p[0][:, :, 2:])
