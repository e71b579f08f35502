"""NumPy emulation of the tcgen05 split-precision DFT of tools/tcgen05_dft/tc_dft.cu (same folding, same operand split, fp32
accumulation) -> Whisper log-mel, compared with the fp32 oracle and fp64 truth.  Also the builder of the DFT operand matrices.

Math (N = 400, windowed frame xw[n], X[k] = sum xw[n] e^{-2 pi i k n / N}, k = 0..200), two symmetric folds:
  s1 = xw[n], s2 = xw[400-n] (xw[400] := 0), s3 = xw[200-n], s4 = xw[200+n] (:= 0 for n = 0), n = 0..100
  a1 = s1 + s2, b1 = s1 - s2, a2 = s3 + s4, b2 = s3 - s4
  ee = a1 + a2 -> Re X[2j]   = sum_n ee[n] cos(2 pi 2j n / N)      (row n = 100 halved: that term is counted twice)
  eo = a1 - a2 -> Re X[2j+1] = sum_n eo[n] cos(2 pi (2j+1) n / N)
  oe = b1 - b2 -> Im X[2j]   = -sum_n oe[n] sin(2 pi 2j n / N)
  oo = b1 + b2 -> Im X[2j+1] = -sum_n oo[n] sin(2 pi (2j+1) n / N) (row n = 100 halved)
Operands: v' = v * 2^12 and F' = F * 2^4 are split as hi = fp16(v'), lo = fp16(v' - hi) (no per-term scaling: the global
pre-scale keeps the lo terms in fp16's normal range); D = Ah Bh + Al Bh + Ah Bl in one fp32 accumulator; power * 2^-32.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
F32 = np.float32
N, KP, NP = 400, 112, 112          # K and N padded to multiples of 16
SX, SF = 2.0 ** 12, 2.0 ** 4
QUADS = ("ee", "eo", "oe", "oo")


def dft_matrices():
    """-> dict q -> (KP, NP) float64 matrix M_q with D_q = A_q @ M_q (already scaled by SF, zero padded)."""
    n = np.arange(101, dtype=np.float64)[:, None]
    j = np.arange(NP, dtype=np.float64)[None, :]
    m = {}
    th_e = 2 * np.pi * (2 * j) * n / N
    th_o = 2 * np.pi * (2 * j + 1) * n / N
    m["ee"] = np.cos(th_e) * (j <= 100)
    m["eo"] = np.cos(th_o) * (j <= 99)
    m["oe"] = -np.sin(th_e) * (j <= 100)
    m["oo"] = -np.sin(th_o) * (j <= 99)
    m["ee"][100, :] *= 0.5
    m["oo"][100, :] *= 0.5
    m["eo"][100, :] = 0.0
    m["oe"][100, :] = 0.0
    m["oe"][0, :] = 0.0
    m["oo"][0, :] = 0.0
    out = {}
    for q in QUADS:
        z = np.zeros((KP, NP))
        z[:101] = m[q]
        z[np.abs(z) < 1e-15] = 0.0
        out[q] = z * SF
    return out


def window_tables(window):
    """-> (4, KP) float32: the window taps of s1..s4 for n = 0..KP-1, pre-scaled by SX, zero where the sample does not exist."""
    w = np.asarray(window, np.float64)
    t = np.zeros((4, KP))
    for n in range(101):
        t[0, n] = w[n]
        t[1, n] = w[400 - n] if n >= 1 else 0.0
        t[2, n] = w[200 - n]
        t[3, n] = w[200 + n] if n >= 1 else 0.0
    return (t * SX).astype(F32)


def split16(a):
    hi = a.astype(np.float16)
    lo = (a.astype(F32) - hi.astype(F32)).astype(np.float16)
    return hi, lo


def folded_operands(frames, wt):
    """frames (T, 401...) raw samples (frame start .. +400 inclusive: index 400 is multiplied by 0) -> dict q -> (T, KP) fp32"""
    n = np.arange(KP)
    nn = np.minimum(n, 100)
    valid = (n <= 100)
    s1 = frames[:, nn] * wt[0]
    s2 = frames[:, 400 - nn] * wt[1]
    s3 = frames[:, 200 - nn] * wt[2]
    s4 = frames[:, 200 + nn] * wt[3]
    a1 = (s1 + s2).astype(F32); b1 = (s1 - s2).astype(F32)
    a2 = (s3 + s4).astype(F32); b2 = (s3 - s4).astype(F32)
    ops = {"ee": a1 + a2, "eo": a1 - a2, "oe": b1 - b2, "oo": b1 + b2}
    return {q: (v * valid).astype(F32) for q, v in ops.items()}


def tc_power(frames, window, products=3):
    """-> (T, 201) fp32 power spectrum through the emulated tensor-core path."""
    wt = window_tables(window)
    mats = dft_matrices()
    ops = folded_operands(frames.astype(F32), wt)
    d = {}
    for q in QUADS:
        ah, al = split16(ops[q])
        bh, bl = split16(mats[q].astype(F32))
        acc = ah.astype(F32) @ bh.astype(F32)
        if products >= 3:
            acc = acc + al.astype(F32) @ bh.astype(F32) + ah.astype(F32) @ bl.astype(F32)
        if products >= 4:
            acc = acc + al.astype(F32) @ bl.astype(F32)
        d[q] = acc
    p = np.zeros((frames.shape[0], 201), F32)
    sc = F32(1.0 / (SX * SF) ** 2)
    p[:, 0::2] = (d["ee"][:, :101] ** 2 + d["oe"][:, :101] ** 2) * sc
    p[:, 1::2] = (d["eo"][:, :100] ** 2 + d["oo"][:, :100] ** 2) * sc
    return p


def whisper_from_power(p, n_mels):
    from oracle import reference_dsp as R
    fb = R.mel_filters(16000, 400, n_mels, 0.0, 8000.0)
    mel = (p @ fb.T).astype(F32)
    lg = np.log10(np.maximum(mel, F32(1e-10))).astype(F32)
    lg = np.maximum(lg, lg.max() - F32(8.0))
    return ((lg + F32(4.0)) / F32(4.0)).astype(F32)


def frames_of(x):
    from oracle import reference_dsp as R
    xp = R.reflect_pad(np.asarray(x, F32), 200)
    nf = 1 + (xp.shape[0] - 400) // 160 - 1          # Whisper drops the last frame
    xp = np.concatenate([xp, np.zeros(1, F32)])
    idx = (np.arange(nf) * 160)[:, None] + np.arange(401)[None, :]
    return xp[idx]


if __name__ == "__main__":
    from oracle import reference_dsp as R
    from tests import synth
    t = np.arange(16000, dtype=F32) / F32(16000)
    cases = {
        "broadband (bench clip)": synth.pcm(1, 48000, seed=1001)[0],
        "pure tone 440 Hz (reference test input)": np.sin(F32(2 * np.pi * 440.0) * t).astype(F32),
        "quiet broadband x 1e-4": synth.pcm(1, 48000, seed=7)[0] * F32(1e-4),
        "tone + silence": np.concatenate([np.sin(F32(2 * np.pi * 1000.0) * t[:8000]), np.zeros(8000, F32)]).astype(F32),
        "constant 0.5 (reference test input)": np.full(16000, 0.5, F32),
    }
    w = R.whisper_hann_window(400)
    for name, x in cases.items():
        want32 = R.whisper_log_mel_spectrogram(x, 128)
        want64 = R.whisper_log_mel_spectrogram(x, 128, dt=np.float64) if "dt" in R.whisper_log_mel_spectrogram.__code__.co_varnames else want32
        fr = frames_of(x)
        line = [name]
        for prod in (1, 3, 4):
            got = whisper_from_power(tc_power(fr, w, prod), 128)
            e32 = np.max(np.abs(got - want32) / np.maximum(1.0, np.abs(want32)))
            e64 = np.max(np.abs(got - want64) / np.maximum(1.0, np.abs(want64)))
            line.append(f"{prod} products: vs fp32 oracle {e32:.2e}, vs fp64 {e64:.2e}")
        e = np.max(np.abs(want32 - want64) / np.maximum(1.0, np.abs(want64)))
        line.append(f"(fp32 oracle vs fp64 {e:.2e})")
        print(" | ".join(line))
