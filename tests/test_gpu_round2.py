"""GPU parity tests of the round-2 entry points and of the branches round 1 left untested (VERDICT r01 "What's weak" 4,
"Next round" 3 and 6): fp16 / 16-bit-PCM Whisper entries (bit-exact), voice-encoder dB / normalised branches, b2a_stft n_fft 512,
stand-alone mlxIstft, resample of a one-sample clip, own-stream contexts with torch tensors, tiny clips through the host pipeline."""
import ctypes as C

import numpy as np
import pytest

from oracle import reference_dsp as R
from tests import synth
from tests.test_gpu_parity import assert_feat_close, ISTFT_ATOL

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def api(ctx):
    from mlx_swift_audio_b200 import api as A
    return A


def _bits(a):
    return np.asarray(a).view(np.uint16)


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("n", [16000 * 3 + 37, 5000, 480000 // 4])
def test_whisper_f16_is_the_cast_of_the_fp32_entry(api, ctx, n_mels, n):
    # asType(.float16) of the fp32 feature (WhisperSTT.swift:156-157): round-to-nearest-even is what NumPy's astype does too
    x = synth.pcm(3, n, seed=2001)
    f32 = api.whisperLogMelSpectrogram(x, nMels=n_mels, ctx=ctx)
    f16 = api.whisperLogMelSpectrogramF16(x, nMels=n_mels, ctx=ctx)
    assert f16.dtype == np.float16 and f16.shape == f32.shape
    assert np.array_equal(_bits(f16), _bits(f32.astype(np.float16)))
    # and against the oracle: fp16 spacing at |v| <= 2 is <= 9.8e-4, on top of the 1e-4 fp32 tolerance
    want = np.stack([R.whisper_log_mel_spectrogram(c, n_mels) for c in x]).astype(np.float16)
    assert np.abs(f16.astype(np.float32) - want.astype(np.float32)).max() <= 2e-3


def test_whisper_f16_device_space_and_padding(api, ctx):
    import torch
    x = synth.pcm(2, 20000, seed=2002)
    f32 = api.whisperLogMelSpectrogram(x, nMels=128, padding=4000, ctx=ctx)
    xd = torch.from_numpy(x).cuda()
    f16 = api.whisperLogMelSpectrogramF16(xd, nMels=128, padding=4000)
    assert f16.dtype == torch.float16
    assert np.array_equal(_bits(f16.cpu().numpy()), _bits(f32.astype(np.float16)))


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("n", [16001, 48000])     # (odd length: the 16-bit rows are not 4-byte multiples)
def test_whisper_pcm16_equals_float_entry_on_scaled_samples(api, ctx, f16, n):
    rng = np.random.default_rng(2003)
    xi = rng.integers(-32768, 32768, (3, n)).astype(np.int16)
    xi[:, n - n // 10:] = 0
    xi[0, :5] = [-32768, 32767, 0, 1, -1]
    xf = xi.astype(np.float32) / np.float32(32768.0)       # exact
    if f16:
        got = api.whisperLogMelSpectrogramF16(xi, nMels=128, ctx=ctx)
        want = api.whisperLogMelSpectrogramF16(xf, nMels=128, ctx=ctx)
        assert np.array_equal(_bits(got), _bits(want))
    else:
        got = api.whisperLogMelSpectrogramPCM16(xi, nMels=128, ctx=ctx)
        want = api.whisperLogMelSpectrogram(xf, nMels=128, ctx=ctx)
        assert np.array_equal(got, want)
        assert_feat_close(got, np.stack([R.whisper_log_mel_spectrogram(c, 128) for c in xf]), what="pcm16 vs oracle")


def test_whisper_pcm16_device_space(api, ctx):
    import torch
    rng = np.random.default_rng(2004)
    xi = rng.integers(-20000, 20000, (2, 32000)).astype(np.int16)
    want = api.whisperLogMelSpectrogramF16(xi, nMels=80, ctx=ctx)
    got = api.whisperLogMelSpectrogramF16(torch.from_numpy(xi).cuda(), nMels=80)
    assert np.array_equal(_bits(got.cpu().numpy()), _bits(want))


def test_whisper_f16_ragged(api, ctx):
    from mlx_swift_audio_b200 import _lib as L
    n = 30000
    x = synth.pcm(4, n, seed=2005)
    lens = np.array([n, 5120, 12345, 401], np.int64)
    frames = int(ctx.lib.b2a_whisper_num_frames(n, 0))
    out = np.full((4, frames, 128), 7.0, np.float16)
    rows = np.zeros(4, np.int64)
    I64 = C.POINTER(C.c_int64)
    ctx.check(ctx.lib.b2a_whisper_log_mel_spectrogram_f16_ragged(ctx.h, C.c_void_p(x.ctypes.data), 4, n, lens.ctypes.data_as(I64), 128, 0,
                                                               C.c_void_p(out.ctypes.data), rows.ctypes.data_as(I64), L.B2A_HOST))
    for b in range(4):
        want = api.whisperLogMelSpectrogram(x[b, :lens[b]], nMels=128, ctx=ctx).astype(np.float16)
        assert rows[b] == want.shape[0]
        assert np.array_equal(_bits(out[b, :rows[b]]), _bits(want))
        assert not out[b, rows[b]:].any()


def test_voice_encoder_db_and_normalized_branches(api, ctx):
    # VoiceEncoderMelspec.swift:52-65: melType "db" (20 log10 max(., stft_magnitude_min)) and normalized_mels
    from mlx_swift_audio_b200 import _lib as L
    x = synth.pcm(2, 16000, seed=2006)
    for kw in (dict(mel_type="db"), dict(mel_type="db", normalized_mels=True), dict(mel_power=1.0, mel_type="db"),
               dict(normalized_mels=True)):
        cfg = L.VoiceEncConfig()
        ctx.lib.b2a_voice_enc_config_default(C.byref(cfg))
        cfg.mel_type_db = int(kw.get("mel_type") == "db")
        cfg.normalized_mels = int(kw.get("normalized_mels", False))
        cfg.mel_power = kw.get("mel_power", 2.0)
        got = api.voiceEncoderMelspectrogram(x, config=cfg, ctx=ctx)
        want = np.stack([R.voice_encoder_melspectrogram(c, **kw) for c in x])
        # dB values reach -80: the scaled tolerance 1e-4 * max(1, |want|) is the north-star's relative one
        assert_feat_close(got, want, what=f"voice encoder {kw}")


def test_stft_n_fft_512(api, ctx):
    # plain stft() on the 512-point plan (400-tap window zero-extended, S3TokenizerUtils.swift:235-239)
    x = synth.pcm(2, 8000, seed=2007)
    w = api.hanningWindow(401)[:400]
    for center in (True, False):
        got = api.stft(x, window=w, nFft=512, hopLength=160, winLength=400, center=center, ctx=ctx)
        want = np.stack([R.stft(c, w, 512, 160, center=center) for c in x])
        scale = np.abs(want).max()
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-5 * scale


def test_mlx_istft_standalone(api, ctx):
    # mlxIstft (MLXSTFT.swift:115-163): complex (F, T') spectrum in, no unwrap; Kokoro's window-sum normalisation
    mag, ph = synth.mag_phase(2, 11, 501, seed=2008)
    re, im = (mag * np.cos(ph)).astype(np.float32), (mag * np.sin(ph)).astype(np.float32)
    got = api.mlxIstft(re, im, hopLength=5, winLength=20, ctx=ctx)
    want = np.stack([R.mlx_istft((re[b] + 1j * im[b]).astype(np.complex64), 5, 20) for b in range(2)])
    one = api.mlxIstft((re[0] + 1j * im[0]).astype(np.complex64), hopLength=5, winLength=20, ctx=ctx)
    assert np.array_equal(one, got[0])
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= ISTFT_ATOL


def test_resample_single_sample_clip(api, ctx):
    # T == 1: the clip bound Float(T) - 1.001 is negative, floor gives -1 and the reference's gather wraps to x[0]
    x = np.array([[0.75], [-0.3]], np.float32)
    got = api.resampleAudio(x, 16000, 24000, ctx=ctx)
    want = np.stack([R.resample_audio(c, 16000, 24000) for c in x])
    assert np.array_equal(got, want)


def test_tiny_clips_through_the_host_pipeline(api, ctx):
    # 70 000 rows of 8 frames: the byte-sized chunk of the host pipeline alone would exceed gridDim.y
    ph = np.random.default_rng(2009).uniform(-3, 3, (70000, 8)).astype(np.float32)
    got = api.unwrap(ph, ctx=ctx)
    assert np.abs(got - R.unwrap(ph)).max() <= 1e-5


def test_own_stream_context_is_ordered_against_torch(api):
    import torch
    own = api.Context(0)           # private non-blocking stream
    try:
        x = torch.from_numpy(synth.pcm(8, 160000, seed=2010)).cuda()
        ref = api.whisperLogMelSpectrogram(x, nMels=128)       # torch's stream
        for _ in range(3):
            y = torch.zeros_like(x)
            y.copy_(x)                                         # producer on torch's stream
            out = api.whisperLogMelSpectrogram(y, nMels=128, ctx=own)
            z = out.clone()                                    # consumer on torch's stream
            assert torch.equal(z, ref)
    finally:
        own.close()


def test_tensor_core_whisper_prototype(api, ctx):
    """The opt-in tcgen05 front end (csrc/tc_frontend.cu): same features as the FFT kernel.  Broadband input
    meets the 1e-4 bar; the reference's pure-tone test input does NOT (1.1e-4 .. 1.3e-4 measured: the tensor core's truncating
    fp32 accumulation over 21 MMA steps) -- which is why the path is not the default.  Checked here at 2e-4 so that it stays alive."""
    from tests.test_gpu_parity import RTOL
    x = synth.pcm(3, 16000 * 5 + 123, seed=2011)
    t = np.arange(16000, dtype=np.float32) / np.float32(16000)
    tone = np.sin(np.float32(2 * np.pi * 440.0) * t).astype(np.float32)
    ctx.lib.b2a_debug_whisper_tc(1)
    try:
        got = api.whisperLogMelSpectrogram(x, nMels=128, ctx=ctx)
        got80 = api.whisperLogMelSpectrogram(x, nMels=80, ctx=ctx)
        got_tone = api.whisperLogMelSpectrogram(tone, nMels=80, ctx=ctx)
        launches0 = ctx.launch_count
        api.whisperLogMelSpectrogram(x, nMels=128, ctx=ctx)
        assert ctx.launch_count - launches0 == 2          # tc_whisper_kernel + whisper_clamp_kernel
    finally:
        ctx.lib.b2a_debug_whisper_tc(0)
    assert_feat_close(got, np.stack([R.whisper_log_mel_spectrogram(c, 128) for c in x]), tol=RTOL, what="tensor-core whisper 128")
    assert_feat_close(got80, np.stack([R.whisper_log_mel_spectrogram(c, 80) for c in x]), tol=RTOL, what="tensor-core whisper 80")
    assert_feat_close(got_tone, R.whisper_log_mel_spectrogram(tone, 80), tol=2e-4, what="tensor-core whisper, pure tone")


@pytest.mark.parametrize("n_fft,hop,win", [(256, 64, 256), (1024, 256, 1024), (600, 150, 480), (50, 7, 50)])
def test_generic_stft_any_size(api, ctx, n_fft, hop, win):
    # stft() for sizes without a tuned plan (S3TokenizerUtils.swift:224-263 takes any nFft / hopLength): direct-DFT kernel
    x = synth.pcm(2, 6000, seed=2012)
    w = api.hanningWindow(win + 1)[:win]
    for center in (True, False):
        got = api.stft(x, window=w, nFft=n_fft, hopLength=hop, center=center, ctx=ctx)
        want = np.stack([R.stft(c, w, n_fft, hop, center=center) for c in x])
        assert got.shape == want.shape
        assert np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


def test_generic_stft_short_clip_reflect_loops(api, ctx):
    # a clip shorter than the centre pad: the reference's repeated reflection (S3TokenizerUtils.swift:287-295) through the generic kernel
    x = synth.pcm(1, 300, seed=2013, zero_tail_frac=0.0)[0]
    w = api.hanningWindow(1025)[:1024]
    got = api.stft(x, window=w, nFft=1024, hopLength=256, ctx=ctx)
    want = R.stft(x, w, 1024, 256)
    assert got.shape == want.shape and np.abs(got - want).max() <= 2e-5 * np.abs(want).max()


def test_voice_encoder_and_s3gen_other_fft_sizes(api, ctx):
    from mlx_swift_audio_b200 import _lib as L
    x = synth.pcm(2, 16000, seed=2014)
    cfg = L.VoiceEncConfig()
    ctx.lib.b2a_voice_enc_config_default(C.byref(cfg))
    cfg.n_fft, cfg.hop_size, cfg.win_size = 512, 128, 512
    got = api.voiceEncoderMelspectrogram(x, config=cfg, ctx=ctx)
    want = np.stack([R.voice_encoder_melspectrogram(c, n_fft=512, hop_size=128, win_size=512) for c in x])
    assert_feat_close(got, want, what="voice encoder n_fft 512")
    y = synth.pcm(2, 24000, sample_rate=24000, seed=2015)
    got = api.s3genMelSpectrogram(y, nFft=1024, numMels=80, samplingRate=24000, hopSize=256, winSize=1024, ctx=ctx)
    want = R.s3gen_mel_spectrogram(y, n_fft=1024, hop_size=256, win_size=1024)
    assert_feat_close(got, want, what="s3gen n_fft 1024")


@pytest.mark.parametrize("fr,to", [(24000, 16000), (16000, 24000), (44100, 16000), (48000, 16000), (22050, 24000)])
def test_resample_poly_matches_scipy(api, ctx, fr, to):
    # non-parity extension (stand-in for AVAudioConverter, Audio/AudioResampler.swift:15-88): scipy.signal.resample_poly is the oracle
    from math import gcd
    from scipy import signal
    x = synth.pcm(2, fr // 2 + 13, sample_rate=fr, seed=2016, zero_tail_frac=0.0)
    got = api.resamplePoly(x, fr, to, ctx=ctx)
    g = gcd(fr, to)
    want = np.stack([signal.resample_poly(c.astype(np.float64), to // g, fr // g) for c in x])
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-5
    assert np.array_equal(api.resamplePoly(x[0], fr, fr, ctx=ctx), x[0])


def test_compiled_c_program_runs_the_front_end_on_a_golden_clip(ctx, tmp_path):
    """A plain C99 program built against include/b200audio.h (no Python in the call path): b2a_whisper_log_mel_spectrogram on the
    golden fixture's clips with host buffers, compared with the frozen oracle output."""
    import os
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    gold = np.load(os.path.join(root, "tests", "golden", "oracle_fp32_v1.npz"))
    x = np.ascontiguousarray(gold["in_x16"], np.float32)
    want = gold["whisper80"]
    src = tmp_path / "main.c"
    src.write_text(r'''
#include <stdio.h>
#include <stdlib.h>
#include "b200audio.h"
int main(int argc, char** argv) {
  long batch = atol(argv[3]), n = atol(argv[4]);
  int n_mels = atoi(argv[5]);
  long frames = (long)b2a_whisper_num_frames(n, 0);
  float* x = (float*)malloc(sizeof(float) * batch * n);
  float* y = (float*)malloc(sizeof(float) * batch * frames * n_mels);
  FILE* f = fopen(argv[1], "rb");
  if (!f || fread(x, sizeof(float), batch * n, f) != (size_t)(batch * n)) return 2;
  fclose(f);
  b2a_ctx* ctx = NULL;
  if (b2a_ctx_create(&ctx, 0) != B2A_OK) { fprintf(stderr, "no device\n"); return 3; }
  int rc = b2a_whisper_log_mel_spectrogram(ctx, x, batch, n, n_mels, 0, y, B2A_HOST);
  if (rc != B2A_OK) { fprintf(stderr, "%s\n", b2a_last_error(ctx)); return 4; }
  /* the reference's fatalError path: a clip too short for one frame */
  if (b2a_whisper_log_mel_spectrogram(ctx, x, 1, 100, n_mels, 0, y, B2A_HOST) != B2A_E_TOO_SHORT) return 5;
  f = fopen(argv[2], "wb");
  fwrite(y, sizeof(float), batch * frames * n_mels, f);
  fclose(f);
  printf("%ld launches\n", (long)b2a_ctx_launch_count(ctx));
  b2a_ctx_destroy(ctx);
  return 0;
}
''')
    exe = tmp_path / "main"
    libdir = os.path.join(root, "mlx_swift_audio_b200")
    subprocess.run(["gcc", "-std=c99", "-O1", "-I", os.path.join(root, "include"), str(src), "-o", str(exe), "-L", libdir, "-l:libb200audio.so",
                    "-Wl,-rpath," + libdir], check=True)
    fin, fout = tmp_path / "in.f32", tmp_path / "out.f32"
    x.tofile(fin)
    r = subprocess.run([str(exe), str(fin), str(fout), str(x.shape[0]), str(x.shape[1]), "80"], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.strip().endswith("launches") and int(r.stdout.split()[0]) >= 2
    got = np.fromfile(fout, np.float32).reshape(want.shape)
    assert_feat_close(got, want, what="compiled C program, golden whisper80")


@pytest.mark.parametrize("n_mels", [80, 128])
@pytest.mark.parametrize("n,padding", [(16000 * 3 + 37, 480000), (5000, 480000), (16000, 201), (16000, 202), (16000 * 2, 160 * 32 * 3),
                                       (16000 * 2, 160 * 32 + 399), (480000 // 8, 480000 // 8), (300, 480000)])
def test_whisper_zero_tail_tiles_are_filled_not_transformed(api, ctx, n_mels, n, padding):
    # WhisperSTT.swift:139-144: every clip gets 30 s of zeros appended before whisperLogMelSpectrogram.  Tiles that lie entirely in that
    # zero tail (incl. its reflection at the right edge) are skipped by the main kernel and filled by the clamp kernel with
    # max(floor, Lmax - 8): the same values the oracle computes from explicit zeros -- fp32, fp16 and 16-bit PCM entries alike.
    x = synth.pcm(3, n, seed=2101, zero_tail_frac=0.0)
    want = np.stack([R.whisper_log_mel_spectrogram(c, n_mels, padding=padding) for c in x])
    got = api.whisperLogMelSpectrogram(x, nMels=n_mels, padding=padding, ctx=ctx)
    assert got.shape == want.shape
    assert_feat_close(got, want, what=f"whisper, n = {n}, padding = {padding}")
    f16 = api.whisperLogMelSpectrogramF16(x, nMels=n_mels, padding=padding, ctx=ctx)
    assert np.array_equal(_bits(f16), _bits(got.astype(np.float16))), "fp16 entry = cast of the fp32 entry, filled tiles included"
    # the switch-free check: padding given as explicit zeros (no zero tail for the library to know about) gives the same features
    xz = np.concatenate([x, np.zeros((3, padding), np.float32)], axis=1)
    explicit = api.whisperLogMelSpectrogram(xz, nMels=n_mels, ctx=ctx)
    assert np.array_equal(explicit, got), "virtual zero tail == explicit zeros, bit for bit"


def test_whisper_zero_tail_silent_clip_and_ragged(api, ctx):
    # a clip of digital silence: every value is the log floor ((-10 + 4) / 4 = -1.5), with and without a zero tail
    z = np.zeros((2, 16000), np.float32)
    for padding in (0, 480000):
        got = api.whisperLogMelSpectrogram(z, nMels=80, padding=padding, ctx=ctx)
        assert np.abs(got + 1.5).max() <= 1e-6
    # ragged batch with a zero tail: per-clip lengths, tiles of silence per clip
    lengths = [16000 * 2, 5000, 16000 * 3 + 11, 700]
    n = max(lengths)
    x = synth.pcm(len(lengths), n, seed=2102, zero_tail_frac=0.0)
    got, rows = api.whisperLogMelSpectrogramRagged(x, lengths, nMels=128, padding=48000, ctx=ctx)
    for b, ln in enumerate(lengths):
        want = R.whisper_log_mel_spectrogram(x[b, :ln], 128, padding=48000)
        assert rows[b] == want.shape[0]
        assert_feat_close(np.asarray(got)[b, :rows[b]], want, what=f"ragged whisper with padding, clip {b}")


@pytest.mark.parametrize("n,padding", [(16000 * 2 + 5, 48000), (4000, 160 * 32 * 2 + 201), (16000, 202)])
def test_chatterbox_log_mel_zero_tail_in_the_mel_major_layout(api, ctx, n, padding):
    # S3TokenizerUtils.swift:160-208 takes the same `padding`; (M, T') layout: the clamp kernel fills the skipped tiles column-wise
    x = synth.pcm(2, n, seed=2103, zero_tail_frac=0.0)
    want = np.stack([R.log_mel_spectrogram_chatterbox(c, 128, padding=padding) for c in x])
    got = api.logMelSpectrogramChatterbox(x, nMels=128, padding=padding, ctx=ctx)
    assert_feat_close(got, want, what=f"chatterbox log-mel, n = {n}, padding = {padding}")
    xz = np.concatenate([x, np.zeros((2, padding), np.float32)], axis=1)
    assert np.array_equal(api.logMelSpectrogramChatterbox(xz, nMels=128, ctx=ctx), got), "virtual zero tail == explicit zeros, bit for bit"


def test_pcm16_entries_of_the_other_front_ends_are_bit_identical(api, ctx):
    # 16-bit PCM in (sample = int16 / 32768, exact in fp32): Fun-ASR preprocessAudio, CAM++ Kaldi fbank (+ mean-norm), S3Gen mel -- the same
    # kernels behind a conversion pass, so the features equal those of the fp32 entry on the converted samples bit for bit (host and device space)
    torch = pytest.importorskip("torch")
    rng = np.random.default_rng(2201)
    for sr, n, fn in ((16000, 16000 * 2 + 7, lambda a: api.preprocessAudio(a, ctx=ctx)),
                      (16000, 16000 * 2 + 7, lambda a: api.kaldiFbankCAMPPlus(a, meanNorm=True, ctx=ctx)),
                      (24000, 24000 + 333, lambda a: api.s3genMelSpectrogram(a, ctx=ctx))):
        i16 = rng.integers(-20000, 20000, (3, n), dtype=np.int16)
        f32 = i16.astype(np.float32) / np.float32(32768.0)
        want = fn(f32)
        got = fn(i16)
        assert got.dtype == np.float32 and np.array_equal(got, want)
        got_dev = fn(torch.from_numpy(i16).cuda())
        assert np.array_equal(got_dev.cpu().numpy(), want)
    # and against the oracle on the converted samples
    f32 = (rng.integers(-20000, 20000, (2, 24000 + 333), dtype=np.int16).astype(np.float32) / np.float32(32768.0))
    i16 = np.rint(f32 * 32768.0).astype(np.int16)
    assert_feat_close(api.s3genMelSpectrogram(i16, ctx=ctx), R.s3gen_mel_spectrogram(f32), what="s3gen mel from 16-bit PCM")
