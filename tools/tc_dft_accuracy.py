"""Accuracy of tensor-core formulations of the Whisper front end's 400-point DFT, emulated in NumPy (groundwork for DESIGN.md
section 6 / 11: the one design that could beat the CUDA-core FFT is a split-precision DFT-as-GEMM on tcgen05).

The DFT of a windowed real frame is folded once (even / odd parts: two real GEMMs with K = 201 / 199), the operands are split
into low-precision terms the tensor core multiplies exactly and accumulates in fp32 (emulated: fp32 matmul of the rounded terms),
and the result goes through the reference's |X|^2 -> mel -> log10 -> clamp -> (x + 4) / 4 chain.  Reported: the test metric
max |a - b| / max(1, |b|) against the fp32 oracle and against fp64 truth, on the benchmark's synthetic clip (noise + three
sines + 10 % silent tail) and on the reference's pure-tone test input.

    python tools/tc_dft_accuracy.py

Formulations: fp16 x2 (hi + scaled lo, three or four products), bf16 x2 (three products), bf16 x3 (six products), tf32 x1, tf32 x2
(three products).  Tolerance to meet: 1e-4.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import reference_dsp as R  # noqa: E402
from tests import synth  # noqa: E402

F32 = np.float32


def rnd_fp16(a):
    return a.astype(np.float16).astype(F32)


def rnd_bits(a, keep):
    """round-to-nearest-even of fp32 to `keep` explicit mantissa bits (bf16: 7, tf32: 10)"""
    u = a.astype(F32).view(np.uint32).astype(np.uint64)
    drop = 23 - keep
    u = u + ((1 << (drop - 1)) - 1) + ((u >> drop) & 1)
    return ((u >> drop) << drop).astype(np.uint32).view(F32)


def split(a, rnd, terms, scale):
    """a ~= sum_i parts[i] / scale**i with parts[i] representable in the low-precision format"""
    parts, rest, s = [], a.astype(np.float64), 1.0
    for _ in range(terms):
        p = rnd((rest * s).astype(F32))
        parts.append(p)
        rest = rest - p.astype(np.float64) / s
        s *= scale
    return parts


def gemm_split(a, w, rnd, terms, scale, max_order):
    """sum over term pairs (i, j) with i + j <= max_order of a_i @ w_j / scale**(i + j), fp32 accumulation"""
    ap, wp = split(a, rnd, terms, scale), split(w, rnd, terms, scale)
    acc = np.zeros((a.shape[0], w.shape[1]), F32)
    for i in range(terms):
        for j in range(terms):
            if i + j <= max_order:
                acc = acc + (ap[i] @ wp[j]) * F32(1.0 / scale ** (i + j))
    return acc


def frames_of(x):
    w = R.whisper_hann_window(400)
    xp = R.reflect_pad(np.asarray(x, F32), 200)
    n = 1 + (xp.size - 400) // 160
    idx = np.arange(400)[None, :] + 160 * np.arange(n)[:, None]
    return (xp[idx] * w[None, :]).astype(F32)[:-1]     # the last frame is dropped (WhisperAudio.swift:105)


def log_mel_from_power(p, n_mels=128):
    mel = np.matmul(p.astype(F32), R.mel_filters(16000, 400, n_mels, 0.0, 8000.0).T).astype(F32)
    ls = np.log10(np.maximum(mel, F32(1e-10))).astype(F32)
    ls = np.maximum(ls, ls.max() - F32(8.0))
    return ((ls + F32(4.0)) / F32(4.0)).astype(F32)


def folded_operands(fr):
    n = np.arange(201)
    k = np.arange(201)
    e = fr[:, :201].copy()
    e[:, 1:200] += fr[:, :200:-1]            # x[n] + x[400 - n], n = 1..199
    o = fr[:, 1:200] - fr[:, :200:-1]
    ang = 2.0 * np.pi * np.outer(n, k) / 400.0
    return e, np.cos(ang), o, -np.sin(ang[1:200])


def power_tc(fr, rnd, terms, scale, max_order):
    e, wc, o, ws = folded_operands(fr)
    re = gemm_split(e, wc.astype(F32) if terms == 0 else wc, rnd, terms, scale, max_order)
    im = gemm_split(o, ws, rnd, terms, scale, max_order)
    return re * re + im * im


def err(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(1.0, np.abs(b.astype(np.float64)))))


def main():
    t = np.arange(16000 * 3, dtype=F32) / F32(16000)
    clips = {
        "bench clip (noise + sines + silent tail), 6 s": synth.pcm(1, 96000, seed=1001)[0],
        "pure 440 Hz tone, 3 s": np.sin(F32(2 * np.pi * 440.0) * t).astype(F32),
    }
    forms = [
        ("fp16 x2, 3 products (lo scaled by 2^11)", rnd_fp16, 2, 2048.0, 1),
        ("fp16 x2, 4 products", rnd_fp16, 2, 2048.0, 2),
        ("bf16 x2, 3 products", lambda a: rnd_bits(a, 7), 2, 256.0, 1),
        ("bf16 x3, 6 products", lambda a: rnd_bits(a, 7), 3, 256.0, 2),
        ("tf32 x1", lambda a: rnd_bits(a, 10), 1, 1.0, 0),
        ("tf32 x2, 3 products", lambda a: rnd_bits(a, 10), 2, 2048.0, 1),
    ]
    for name, x in clips.items():
        o32 = R.whisper_log_mel_spectrogram(x, 128)
        o64 = R.whisper_log_mel_spectrogram(x, 128, dt=np.float64)
        fr = frames_of(x)
        print(f"== {name}: {fr.shape[0]} frames; fp32 oracle vs fp64 truth {err(o32, o64):.2e}")
        e, wc, o, ws = folded_operands(fr)
        exact = log_mel_from_power(((e.astype(np.float64) @ wc) ** 2 + (o.astype(np.float64) @ ws) ** 2).astype(F32))
        print(f"   folded DFT in fp64 + fp32 tail           vs fp32 oracle {err(exact, o32):.2e}   vs fp64 truth {err(exact, o64):.2e}")
        for fname, rnd, terms, scale, order in forms:
            got = log_mel_from_power(power_tc(fr, rnd, terms, scale, order))
            print(f"   {fname:40s} vs fp32 oracle {err(got, o32):.2e}   vs fp64 truth {err(got, o64):.2e}")


if __name__ == "__main__":
    main()
