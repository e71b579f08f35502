"""Synthetic inputs of SURVEY.md section 8(d): seeded, reproducible, shared by tests and bench."""
import numpy as np


def pcm(batch: int, n_samples: int, sample_rate: int = 16000, seed: int = 1000, zero_tail_frac: float = 0.1) -> np.ndarray:
    """0.1*N(0,1) + three sines (220 Hz, 1 kHz, 3.3 kHz; amplitude 0.2; random phase per clip),
    clipped to [-1, 1], last 10 % of every clip zeroed (exercises the 1e-10 floor and the max-8 clamp)."""
    out = np.empty((batch, n_samples), np.float32)
    t = np.arange(n_samples, dtype=np.float64) / sample_rate
    for b in range(batch):
        rng = np.random.default_rng([seed, b])
        x = 0.1 * rng.standard_normal(n_samples)
        for f in (220.0, 1000.0, 3300.0):
            x += 0.2 * np.sin(2 * np.pi * f * t + rng.uniform(0, 2 * np.pi))
        x = np.clip(x, -1.0, 1.0)
        nz = int(n_samples * zero_tail_frac)
        if nz > 0:
            x[n_samples - nz:] = 0.0
        out[b] = x.astype(np.float32)
    return out


def mag_phase(batch: int, n_bins: int, n_frames: int, seed: int = 1005, phase_mode: str = "sin"):
    """mag = exp(N(-2,1)) with 0.1 % of entries forced to 150 (exercises the clip at 100);
    phase = sin(N(0, 2^2)) in [-1, 1] as the vocoders produce, or U(-pi, pi) ("uniform", exercises unwrap)."""
    rng = np.random.default_rng(seed)
    mag = np.exp(rng.normal(-2.0, 1.0, (batch, n_bins, n_frames))).astype(np.float32)
    mask = rng.random((batch, n_bins, n_frames)) < 1e-3
    mag[mask] = 150.0
    if phase_mode == "sin":
        ph = np.sin(rng.normal(0.0, 2.0, (batch, n_bins, n_frames))).astype(np.float32)
    else:
        ph = rng.uniform(-np.pi, np.pi, (batch, n_bins, n_frames)).astype(np.float32)
    return mag, ph
