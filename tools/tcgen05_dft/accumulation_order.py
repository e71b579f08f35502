"""NumPy emulation of the tcgen05 split-precision DFT's ACCUMULATION (tools/tcgen05_dft/emulate.py supplies the folding and the operand split):
every MMA forms its K = 16 partial sum exactly and adds it to the fp32 TMEM accumulator with truncation (round toward zero) -- the model
profiles/r02_tcgen05_dft.txt arrives at for the measured 1.14e-4 on the reference's pure-tone input.  Compared, per Whisper log-mel case:

  slice-major   the prototype's order: per K slice the three products Ah Bh, Al Bh, Ah Bl into ONE accumulator (21 truncating adds per output)
  split         the four quadrant GEMMs are independent, so run ONE QUADRANT AT A TIME: a quadrant needs 112 TMEM columns per accumulator, which
                leaves room for the large products (Ah Bh: 7 truncating adds) and the small ones (Al Bh, Ah Bl: 14 adds of terms 2^-11 as large)
                in SEPARATE accumulators, summed once by the epilogue in fp32 round-to-nearest
  split, 2 x 4  ... and the large products alternating between two accumulators (4 + 3 truncating adds)
  rn            round to nearest after every MMA (what a correctly rounding accumulator would give), for scale

    python tools/tcgen05_dft/accumulation_order.py

CPU only (a design aid for DESIGN.md section 11, "Next"); nothing here is on the product path.
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from tools.tcgen05_dft import emulate as E  # noqa: E402

F32, F64 = np.float32, np.float64


def rz32(v64):
    """fp64 -> fp32 with round toward zero"""
    f = v64.astype(F32)
    over = np.abs(f.astype(F64)) > np.abs(v64)
    return np.where(over, np.nextafter(f, F32(0.0)), f).astype(F32)


def accumulate(terms, mode):
    """terms: list of (T, N) fp64 exact per-MMA partial sums, in issue order -> fp32 accumulator after all of them"""
    acc = np.zeros(terms[0].shape, F32)
    for t in terms:
        s = acc.astype(F64) + t
        acc = rz32(s) if mode == "rz" else s.astype(F32)
    return acc


def quadrant(a32, b32, order):
    ah, al = E.split16(a32)
    bh, bl = E.split16(b32)
    ah, al, bh, bl = (v.astype(F64) for v in (ah, al, bh, bl))
    sl = [slice(16 * i, 16 * i + 16) for i in range(E.KP // 16)]
    hh = [ah[:, s] @ bh[s] for s in sl]
    lh = [al[:, s] @ bh[s] for s in sl]
    hl = [ah[:, s] @ bl[s] for s in sl]
    if order == "slice-major":
        return accumulate([t for i in range(len(sl)) for t in (hh[i], lh[i], hl[i])], "rz")
    if order == "rn":
        return accumulate([t for i in range(len(sl)) for t in (hh[i], lh[i], hl[i])], "rn")
    small = accumulate([t for i in range(len(sl)) for t in (lh[i], hl[i])], "rz")
    if order == "split":
        return (accumulate(hh, "rz") + small).astype(F32)
    if order == "split, 2 x 4":
        return ((accumulate(hh[0::2], "rz") + accumulate(hh[1::2], "rz")).astype(F32) + small).astype(F32)
    raise ValueError(order)


def tc_power(frames, window, order):
    wt = E.window_tables(window)
    mats = E.dft_matrices()
    ops = E.folded_operands(frames.astype(F32), wt)
    d = {q: quadrant(ops[q], mats[q].astype(F32), order) for q in E.QUADS}
    p = np.zeros((frames.shape[0], 201), F32)
    sc = F32(1.0 / (E.SX * E.SF) ** 2)
    p[:, 0::2] = (d["ee"][:, :101] ** 2 + d["oe"][:, :101] ** 2) * sc
    p[:, 1::2] = (d["eo"][:, :100] ** 2 + d["oo"][:, :100] ** 2) * sc
    return p


if __name__ == "__main__":
    from oracle import reference_dsp as R
    from tests import synth
    t = np.arange(16000, dtype=F32) / F32(16000)
    cases = {
        "pure tone 440 Hz, 1 s (reference test input)": np.sin(F32(2 * np.pi * 440.0) * t).astype(F32),
        "broadband (bench clip, 3 s)": synth.pcm(1, 48000, seed=1001)[0],
        "tone + silence": np.concatenate([np.sin(F32(2 * np.pi * 1000.0) * t[:8000]), np.zeros(8000, F32)]).astype(F32),
        "constant 0.5 (reference test input)": np.full(16000, 0.5, F32),
    }
    w = R.whisper_hann_window(400)
    for name, x in cases.items():
        fr = E.frames_of(x)
        for n_mels in (80, 128):
            want = R.whisper_log_mel_spectrogram(x, n_mels)
            line = [f"{name}, {n_mels} mel"]
            for order in ("slice-major", "split", "split, 2 x 4", "rn"):
                got = E.whisper_from_power(tc_power(fr, w, order), n_mels)
                line.append(f"{order}: {np.max(np.abs(got - want) / np.maximum(1.0, np.abs(want))):.2e}")
            print(" | ".join(line), flush=True)
