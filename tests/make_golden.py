#!/usr/bin/env python3
"""Generates the golden fixtures under tests/golden/ from the fp32 CPU oracle with fixed seeds.

The reference (smdesai/mlx-swift-audio) ships no golden vectors for this path and cannot run
here (Swift + MLX + Apple frameworks), so these vectors pin the ORACLE, not the reference:
they catch accidental changes to the restatement and give the GPU tests a second, frozen target.
The inputs reproduce the reference's own test inputs where it has any: a 1 s 440 Hz unit sine at
16 kHz (Tests/FunASRTests.swift:145-156) and constant-0.5 arrays (Tests/WhisperTests.swift:87-93).

    python tests/make_golden.py          # rewrites tests/golden/*.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import reference_dsp as R  # noqa: E402
from tests import synth  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def sine_1s():
    t = np.arange(16000, dtype=np.float32) / np.float32(16000)
    return np.sin(np.float32(2 * np.pi * 440.0) * t).astype(np.float32)


def main():
    os.makedirs(OUT, exist_ok=True)
    sine = sine_1s()
    x16 = synth.pcm(2, 8000, 16000, seed=2001)          # 0.5 s clips keep the fixtures small
    x24 = synth.pcm(2, 12000, 24000, seed=2002)
    mag16, ph16 = synth.mag_phase(2, 9, 301, seed=2003)
    mag20, ph20 = synth.mag_phase(2, 11, 241, seed=2004)
    w16 = R.hann_window_periodic(16)
    g = {}
    g["in_sine"] = sine
    g["in_x16"] = x16
    g["in_x24"] = x24
    g["in_mag16"], g["in_ph16"], g["in_mag20"], g["in_ph20"] = mag16, ph16, mag20, ph20
    # the reference's own unit-test inputs
    g["sine_funasr_logmel"] = R.funasr_log_mel_spectrogram(sine)
    g["sine_funasr_lfr"] = R.apply_lfr(g["sine_funasr_logmel"])
    g["sine_funasr_preprocess"] = R.preprocess_audio(sine)
    g["sine_whisper80"] = R.whisper_log_mel_spectrogram(sine, 80)
    # fp64 truth for the pure tone: its un-clamped log-mel spans 29 nepers, so bins ~100 dB below the peak carry
    # nothing but fp32 rounding noise and two fp32 implementations differ there by ~1e-3 ("tolerances")
    g["sine_funasr_logmel_f64"] = R.funasr_log_mel_spectrogram(sine, dt=np.float64).astype(np.float64)
    # seeded synthetic clips
    g["whisper80"] = np.stack([R.whisper_log_mel_spectrogram(c, 80) for c in x16])
    g["whisper128"] = np.stack([R.whisper_log_mel_spectrogram(c, 128) for c in x16])
    g["chatterbox128"] = np.stack([R.log_mel_spectrogram_chatterbox(c, 128) for c in x16])
    g["funasr_preprocess"] = np.stack([R.preprocess_audio(c) for c in x16])
    g["kaldi_fbank"] = np.stack([R.kaldi_fbank_camp_plus(c) for c in x16])
    g["kaldi_fbank_meannorm"] = np.stack([R.kaldi_fbank_mean_norm(f) for f in g["kaldi_fbank"]])
    g["s3gen_mel"] = R.s3gen_mel_spectrogram(x24)
    g["voice_encoder_mel"] = np.stack([R.voice_encoder_melspectrogram(c) for c in x16])
    g["stft400"] = np.stack([R.stft(c, R.whisper_hann_window(400), 400, 160) for c in x16]).view(np.float32)
    re, im = R.stft_hifigan(x24[:, :2000], 16, 4, w16)
    g["stft_hifigan_re"], g["stft_hifigan_im"] = re, im
    re, im = R.cosyvoice3_stft(x24[:, :2000], 16, 4, w16)
    g["cv3_stft_re"], g["cv3_stft_im"] = re, im
    m, p = R.kokoro_transform(x24[:, :2000])
    g["kokoro_mag"], g["kokoro_phase"] = m, p
    g["istft_hifigan"] = R.istft_hifigan(mag16, ph16, 16, 4, w16)
    g["cv3_istft"] = R.cosyvoice3_istft(mag16, ph16, 16, 4, w16)
    g["kokoro_inverse"] = R.kokoro_inverse(mag20, ph20)
    # tables
    g["win_whisper_hann_400"] = R.whisper_hann_window(400)
    g["win_hann_periodic_400"] = R.hann_periodic_via_hanning(400)
    g["win_hamming_400"] = R.hamming_window(400)
    g["win_povey_400"] = R.povey_window(400)
    g["win_hann_periodic_16"] = w16
    g["fb_slaney_80"] = R.mel_filters(16000, 400, 80, 0.0, 8000.0)
    g["fb_slaney_128"] = R.mel_filters(16000, 400, 128, 0.0, 8000.0)
    g["fb_funasr_80"] = R.funasr_mel_filters()
    g["fb_htk_80"] = R.mel_filters_htk(16000, 512, 80, 20.0, 8000.0)
    np.savez_compressed(os.path.join(OUT, "oracle_fp32_v1.npz"), **{k: np.asarray(v) for k, v in g.items()})
    size = os.path.getsize(os.path.join(OUT, "oracle_fp32_v1.npz"))
    print("wrote", len(g), "arrays,", size // 1024, "KiB")

    # adjacent rows: separate file, the v1 vectors stay byte-identical
    a = {}
    rng = np.random.default_rng(2010)
    h16 = np.concatenate([rng.normal(-2.0, 1.0, (2, 9, 301)), rng.normal(0.0, 2.0, (2, 9, 301))], axis=1).astype(np.float32)
    h16[0, 2, 100] = 5.2                                  # exp -> 181 -> magnitude clip at 100 -> output limiter
    h20 = np.concatenate([rng.normal(-2.0, 1.0, (2, 11, 241)), rng.normal(0.0, 2.0, (2, 11, 241))], axis=1).astype(np.float32)
    a["in_h16"], a["in_h20"] = h16, h20
    a["hift_head_istft"] = R.hift_head_istft(h16, 16, 4, w16, 0.99)
    a["kokoro_head_istft"] = R.kokoro_head_istft(h20)
    mel = R.whisper_log_mel_spectrogram(np.concatenate([x16[0], np.zeros(4000, np.float32)]), 80)   # 75 frames, 50 of content
    a["in_mel"] = mel
    a["mel_segment_seek0"] = R.whisper_mel_segment(mel, 0, 50, length=64)
    a["mel_segment_seek37"] = R.whisper_mel_segment(mel, 37, 50, length=64)
    a["resample_24k_16k"] = R.resample_audio(x24, 24000, 16000)
    a["resample_16k_24k"] = R.resample_audio(x16, 16000, 24000)
    np.savez_compressed(os.path.join(OUT, "oracle_fp32_v2_adjacent.npz"), **{k: np.asarray(v) for k, v in a.items()})
    print("wrote", len(a), "arrays,", os.path.getsize(os.path.join(OUT, "oracle_fp32_v2_adjacent.npz")) // 1024, "KiB")


if __name__ == "__main__":
    main()
