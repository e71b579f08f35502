"""ctypes loader for oracle/libcputwin.so -- the compiled, multi-threaded twin of the NumPy oracle (TEST INFRASTRUCTURE
ONLY, see oracle/__init__.py).  Built by `make -C oracle` (and by __graft_entry__.build()).  Used by tests/test_cpu_twin.py and
by bench.py's cpu_baseline / --impl reference legs; never by the product path."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import reference_dsp as R

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libcputwin.so")
_lib = None


def available() -> bool:
    return os.path.exists(LIB_PATH)


def _load():
    global _lib
    if _lib is None:
        lib = C.CDLL(LIB_PATH)
        fp = C.POINTER(C.c_float)
        lib.twin_whisper_log_mel.argtypes = [fp, C.c_int64, C.c_int64, C.c_int, fp, fp, fp, C.c_int]
        lib.twin_whisper_log_mel.restype = C.c_int
        lib.twin_istft_hifigan.argtypes = [fp, fp, C.c_int64, C.c_int64, fp, fp, C.c_int]
        lib.twin_istft_hifigan.restype = C.c_int
        _lib = lib
    return _lib


def _p(a):
    return a.ctypes.data_as(C.POINTER(C.c_float))


def whisper_log_mel_spectrogram(audio: np.ndarray, n_mels: int, n_threads: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    """(batch, n) fp32 -> (batch, n // 160, n_mels); window and filterbank come from the NumPy oracle.
    ``out``: optional preallocated result (the timing arm reuses one buffer so that page faults are not part of the measurement)."""
    x = np.ascontiguousarray(np.atleast_2d(audio), dtype=np.float32)
    b, n = x.shape
    w = np.ascontiguousarray(R.whisper_hann_window(400), dtype=np.float32)
    fb = np.ascontiguousarray(R.mel_filters(16000, 400, n_mels), dtype=np.float32)
    if out is None:
        out = np.empty((b, n // 160, n_mels), np.float32)
    assert out.shape == (b, n // 160, n_mels) and out.dtype == np.float32 and out.flags.c_contiguous
    rc = _load().twin_whisper_log_mel(_p(x), b, n, n_mels, _p(w), _p(fb), _p(out), n_threads or (os.cpu_count() or 1))
    if rc != 0:
        raise ValueError("Input is too short for STFT")
    return out


def istft_hifigan(magnitude: np.ndarray, phase: np.ndarray, n_threads: int = 0, out: np.ndarray | None = None) -> np.ndarray:
    """(batch, 9, frames) x 2 -> (batch, (frames - 1) * 4), periodic Hann 16 / hop 4."""
    m = np.ascontiguousarray(magnitude, dtype=np.float32)
    p = np.ascontiguousarray(phase, dtype=np.float32)
    b, f, frames = m.shape
    assert f == 9 and p.shape == m.shape
    w = np.ascontiguousarray(R.hann_window_periodic(16), dtype=np.float32)
    if out is None:
        out = np.empty((b, (frames - 1) * 4), np.float32)
    assert out.shape == (b, (frames - 1) * 4) and out.dtype == np.float32 and out.flags.c_contiguous
    rc = _load().twin_istft_hifigan(_p(m), _p(p), b, frames, _p(w), _p(out), n_threads or (os.cpu_count() or 1))
    if rc != 0:
        raise ValueError("too few frames")
    return out
