"""Bring-up / accuracy / timing of the tensor-core Whisper front end (B2A_WHISPER_TC=1) on a B200.
    B2A_WHISPER_TC=1 python tools/tcgen05_dft/bringup.py [--time]
Compares |X|^2 with an fp64 DFT, the log-mel with the fp32 oracle (test metric) on the broadband clip, the reference's pure-tone
test input and the golden fixtures, and times the full 1024 x 30 s step against the FFT kernel (separate process: B2A_WHISPER_TC=0)."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

from mlx_swift_audio_b200 import api  # noqa: E402
from oracle import reference_dsp as R  # noqa: E402
from tests import synth  # noqa: E402


def metric(a, b):
    return float(np.max(np.abs(a.astype(np.float64) - b.astype(np.float64)) / np.maximum(1.0, np.abs(b.astype(np.float64)))))


def main():
    tc = os.environ.get("B2A_WHISPER_TC", "0")
    print("B2A_WHISPER_TC =", tc)
    ctx = api.Context(0)
    lib = ctx.lib
    # ---- power spectrum against an fp64 DFT ----
    x = synth.pcm(2, 16000 * 3 + 37, seed=1001)
    frames = x.shape[1] // 160
    if tc == "1":
        pw = torch.zeros((2, frames, 201), device="cuda")
        lib.b2a_debug_tc_power_buffer(pw.data_ptr())
    got = api.whisperLogMelSpectrogram(x, nMels=128, ctx=ctx)
    if tc == "1":
        torch.cuda.synchronize()
        lib.b2a_debug_tc_power_buffer(None)
        w = R.whisper_hann_window(400).astype(np.float64)
        worst = 0.0
        for b in range(2):
            xp = R.reflect_pad(x[b], 200).astype(np.float64)
            idx = (np.arange(frames) * 160)[:, None] + np.arange(400)[None, :]
            ref = np.abs(np.fft.rfft(xp[idx] * w, axis=1)) ** 2
            g = pw[b].cpu().numpy().astype(np.float64)
            rel = np.abs(g - ref) / ref.max(axis=1, keepdims=True)
            worst = max(worst, rel.max())
            if rel.max() > 1e-3:
                i = np.unravel_index(np.argmax(rel), rel.shape)
                print("  power mismatch clip", b, "frame/bin", i, "got", g[i], "want", ref[i])
                print("  frame 5 bins 0..7 got ", g[5, :8])
                print("  frame 5 bins 0..7 want", ref[5, :8])
        print("power spectrum: max |err| / frame max = %.3e" % worst)
    want = np.stack([R.whisper_log_mel_spectrogram(c, 128) for c in x])
    print("broadband 2 x 3 s, 128 mel: %.3e" % metric(got, want))
    t = np.arange(16000, dtype=np.float32) / np.float32(16000)
    cases = {
        "pure tone 440 Hz 1 s (reference test input)": np.sin(np.float32(2 * np.pi * 440.0) * t).astype(np.float32),
        "constant 0.5 (reference test input)": np.full(16000, 0.5, np.float32),
        "tone + silence": np.concatenate([np.sin(np.float32(2 * np.pi * 1000.0) * t[:8000]), np.zeros(8000, np.float32)]).astype(np.float32),
        "quiet broadband x 1e-4": synth.pcm(1, 48000, seed=7)[0] * np.float32(1e-4),
        "short clip 5000": synth.pcm(1, 5000, seed=3)[0],
        "clip of 401 samples": synth.pcm(1, 401, seed=4, zero_tail_frac=0.0)[0],
    }
    for name, c in cases.items():
        for nm in (80, 128):
            g = api.whisperLogMelSpectrogram(c, nMels=nm, ctx=ctx)
            wnt = R.whisper_log_mel_spectrogram(c, nm)
            print(f"{name}, {nm} mel: {metric(g, wnt):.3e}")
    gold = np.load(os.path.join(ROOT, "tests", "golden", "oracle_fp32_v1.npz"))
    print("golden sine_whisper80: %.3e" % metric(api.whisperLogMelSpectrogram(gold["in_sine"], nMels=80, ctx=ctx), gold["sine_whisper80"]))
    print("golden whisper80: %.3e" % metric(api.whisperLogMelSpectrogram(gold["in_x16"], nMels=80, ctx=ctx), gold["whisper80"]))
    print("golden whisper128: %.3e" % metric(api.whisperLogMelSpectrogram(gold["in_x16"], nMels=128, ctx=ctx), gold["whisper128"]))
    if "--time" in sys.argv:
        B, n = 1024, 480000
        xu = torch.from_numpy(synth.pcm(16, n, seed=1001)).cuda()
        xd = xu.repeat(B // 16, 1).contiguous()
        out = torch.empty((B, n // 160, 128), device="cuda")
        dctx = api.Context(0, torch.cuda.current_stream().cuda_stream)
        for _ in range(3):
            api.whisperLogMelSpectrogram(xd, 128, ctx=dctx, out=out)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(10):
            api.whisperLogMelSpectrogram(xd, 128, ctx=dctx, out=out)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 10
        print("1024 x 30 s, 128 mel: %.4f ms / step, %.3f of the HBM roofline (3.539 GB / 6453.7 GB/s)" % (ms, 3.538944 / ms / 6453.7 * 1e3))
        ref = np.stack([R.whisper_log_mel_spectrogram(c, 128) for c in xu[:2].cpu().numpy()])
        print("  full-size clips 0..1 vs oracle: %.3e" % metric(out[:2].cpu().numpy(), ref))


if __name__ == "__main__":
    main()
