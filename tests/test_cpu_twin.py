"""The compiled twin of the oracle (oracle/cpu_twin.cpp, the second CPU baseline of BASELINE.md section 4) against the NumPy oracle."""
import subprocess
import os

import numpy as np
import pytest

from oracle import cpu_twin as T
from oracle import reference_dsp as R
from tests import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle")], check=True)
    assert T.available()


@pytest.mark.parametrize("n_mels,n", [(128, 16000 * 3), (80, 16000 * 2 + 77), (128, 201)])
def test_whisper_twin_matches_numpy_oracle(n_mels, n):
    x = synth.pcm(3, n, seed=7)
    got = T.whisper_log_mel_spectrogram(x, n_mels, n_threads=2)
    want = np.stack([R.whisper_log_mel_spectrogram(c, n_mels) for c in x])
    assert got.shape == want.shape
    assert np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))) <= 1e-4


def test_whisper_twin_too_short():
    with pytest.raises(ValueError):
        T.whisper_log_mel_spectrogram(np.zeros((1, 200), np.float32), 80)


def test_istft_twin_matches_numpy_oracle():
    mag, ph = synth.mag_phase(2, 9, 1501, seed=8)
    got = T.istft_hifigan(mag, ph, n_threads=2)
    want = R.istft_hifigan(mag, ph, 16, 4, R.hann_window_periodic(16))
    assert got.shape == want.shape
    assert np.abs(got - want).max() <= 1e-5
