"""Pure-tone accuracy probe: GPU vs fp32 oracle vs fp64 oracle on the reference's own test input."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import reference_dsp as R
from mlx_swift_audio_b200 import api as A
t = np.arange(16000, dtype=np.float32) / np.float32(16000)
sine = np.sin(np.float32(2 * np.pi * 440.0) * t).astype(np.float32)
for name, gpu, o32, o64 in [
    ("funasr logmel", A.funASRLogMelSpectrogram(sine), R.funasr_log_mel_spectrogram(sine), R.funasr_log_mel_spectrogram(sine, dt=np.float64)),
    ("whisper80", A.whisperLogMelSpectrogram(sine, nMels=80), R.whisper_log_mel_spectrogram(sine, 80), R.whisper_log_mel_spectrogram(sine, 80, dt=np.float64)),
    ("kaldi", A.kaldiFbankCAMPPlus(sine), R.kaldi_fbank_camp_plus(sine), R.kaldi_fbank_camp_plus(sine, dt=np.float64)),
]:
    e = lambda a, b: float(np.max(np.abs(a - b) / np.maximum(1, np.abs(b))))
    print(f"{name:14s} gpu-vs-o32 {e(gpu, o32):.2e}  gpu-vs-o64 {e(gpu, o64):.2e}  o32-vs-o64 {e(o32, o64):.2e}   min value {o64.min():.2f} max {o64.max():.2f}")
