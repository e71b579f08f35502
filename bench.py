#!/usr/bin/env python3
"""Benchmark of the STFT-family DSP hot path on B200 (contract: see the task statement / DESIGN.md).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload NAME] [--impl ours|reference]

One "step" = one pass of the hot path over one batch of synthetic input.  The default workload is
BASELINE.json configs[1]: Whisper large-v3-turbo 128-mel log-mel of 1024 x 30 s clips per GPU (weak
scaling: every rank owns its own 1024 clips, no data-path collective).  Prints ONE JSON line.

  value     whole-job audio-seconds per second with inputs resident in HBM (CUDA events, max over ranks)
  e2e       same metric through the C ABI with pinned HOST buffers (H2D + kernels + D2H inside the timed region)
  roofline  algorithmic bytes / device time of the call vs the measured HBM copy bandwidth
  cpu_baseline  the CPU oracle (port of the reference's algorithm) on a bounded sample, all host cores

--impl reference times the CPU restatement of the reference (oracle/: the NumPy port and, for the headline
workloads, its compiled multi-threaded twin; the reference itself cannot be built outside macOS) on the host cores for the same workload; under torchrun only rank 0 works.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s"

# name -> dict(batch, seconds per clip, sample rate, algorithmic bytes per clip, description)
WORKLOADS = {
    # SURVEY.md 8(d) / BASELINE.md section 3
    "whisper128": dict(batch=1024, clip_s=30.0, sr=16000, desc="Whisper large-v3-turbo 128-mel log-mel, 1024 x 30 s @16 kHz (configs[1])"),
    "whisper80_1clip": dict(batch=1, clip_s=30.0, sr=16000, desc="Whisper 80-mel log-mel, 1 x 30 s (configs[0])"),
    "funasr": dict(batch=512, clip_s=20.0, sr=16000, desc="Fun-ASR preprocessAudio (log-mel + LFR 7/6 + CMVN), 512 x 20 s (configs[2], 3a)"),
    "kaldi": dict(batch=512, clip_s=20.0, sr=16000, desc="Kaldi-style 80-dim fbank (CAM++) + mean-norm, 512 x 20 s (configs[2], 3b)"),
    "s3gen": dict(batch=256, clip_s=10.0, sr=24000, desc="CosyVoice2/Chatterbox 24 kHz 80-mel (n_fft 1920, hop 480), 256 x 10 s (configs[3])"),
    "istft_hift": dict(batch=512, clip_s=30.0, sr=24000, desc="CosyVoice HiFT iSTFT (n_fft 16, hop 4), 512 x 30 s of mag/phase (configs[4], 5a)"),
    "whisper_segment": dict(batch=1024, clip_s=30.0, sr=16000, desc="Whisper seek window: slice + zero-pad + fp16 cast of the 128-mel log-mel, 1024 clips (SURVEY 8f rank 1)"),
    "hift_head": dict(batch=512, clip_s=30.0, sr=24000, desc="HiFT vocoder head: exp/sin split + iSTFT (16/4) + limiter fused, 512 x 30 s of conv output (SURVEY 8f rank 2)"),
    "whisper128_ragged": dict(batch=1024, clip_s=30.0, sr=16000, desc="Whisper 128-mel log-mel of a RAGGED batch: 1024 clips of 5..30 s (uniform) in one launch (b2a_whisper_log_mel_spectrogram_ragged)"),
    "istft_kokoro": dict(batch=512, clip_s=30.0, sr=24000, desc="Kokoro iSTFTNet iSTFT (n_fft 20, hop 5), 512 x 30 s of mag/phase (configs[4], 5b)"),
    "chatterbox128": dict(batch=1024, clip_s=10.0, sr=16000, desc="S3Tokenizer / Chatterbox 128-mel log-mel (periodic Hann, (M, T') layout), 1024 x 10 s @16 kHz (SURVEY 8a row a14)"),
    "voice_encoder": dict(batch=1024, clip_s=10.0, sr=16000, desc="Chatterbox voice-encoder 40-mel power mel ((M, T') layout, interpreted bank), 1024 x 10 s @16 kHz (SURVEY 8a row a21)"),
    "stft_kokoro": dict(batch=512, clip_s=30.0, sr=24000, desc="Kokoro MLXSTFT.transform (n_fft 20, hop 5): magnitude and atan2 phase, 512 x 30 s (SURVEY 8a row a26)"),
    "stft_hift": dict(batch=512, clip_s=30.0, sr=24000, desc="HiFT forward STFT of the source signal (stftHiFiGAN, n_fft 16, hop 4, reflect pad), 512 x 30 s -> real / imag (SURVEY 8a row a22)"),
}


def _peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            f = [x.strip() for x in r.split(",")]
            if len(f) < 8:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for nm, v in zip(names, f[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


# ------------------------------------------------------------------------------------------------------
# workloads on the GPU (direct C-ABI calls on preallocated buffers)
# ------------------------------------------------------------------------------------------------------

class GpuWorkload:
    def __init__(self, name, batch_override=None):
        import torch
        from mlx_swift_audio_b200 import api, _lib
        self.torch = torch
        self.name = name
        w = WORKLOADS[name]
        self.batch = batch_override or w["batch"]
        self.sr = w["sr"]
        self.n = int(round(w["clip_s"] * w["sr"]))
        self.audio_s = self.batch * w["clip_s"]
        dev = torch.device("cuda", torch.cuda.current_device())
        self.dev = dev
        self.ctx = api.Context(dev.index, torch.cuda.current_stream(dev).cuda_stream)
        self.hctx = api.Context(dev.index)  # own stream for the host-buffer (e2e) path
        lib = self.ctx.lib
        self.lib = lib
        g = torch.Generator(device=dev)
        g.manual_seed(1000 + (int(os.environ.get("RANK", "0"))))
        B, n = self.batch, self.n
        DEV = _lib.B2A_DEVICE
        if name == "whisper_segment":
            frames = 6000                      # mel of 30 s of audio + the 30 s of padding transcribe() appends
            mel = torch.randn((B, frames, 128), generator=g, device=dev)
            self.inputs = [mel]
            self.out = torch.empty((B, 3000, 128), dtype=torch.float16, device=dev)
            rs = np.random.default_rng(3)
            seek = np.ascontiguousarray(rs.integers(0, 3000, B), np.int64)     # arbitrary frame offsets, as the decoder produces
            content = np.full(B, 3000, np.int64)                               # 30 s of content: windows near the end are zero-padded
            self._keep = (seek, content)
            I64 = C.POINTER(C.c_int64)
            self.call = lambda c, i, o, sp: lib.b2a_whisper_mel_segment_f16(c.h, i[0], B, frames, 128, seek.ctypes.data_as(I64),
                                                                            content.ctypes.data_as(I64), 3000, o, sp)
        elif name == "hift_head":
            frames = n // 4 + 1
            h = torch.randn((B, 18, frames), generator=g, device=dev)
            h[:, :9] -= 2.0
            h[:, 9:] *= 2.0
            self.inputs = [h]
            self.out = torch.empty((B, (frames - 1) * 4), device=dev)
            win = np.ascontiguousarray(api.hannWindowPeriodic(16), np.float32)
            self._keep = win
            wp = win.ctypes.data_as(C.POINTER(C.c_float))
            self.call = lambda c, i, o, sp: lib.b2a_hift_head_istft(c.h, i[0], B, frames, 16, 4, wp, C.c_float(0.99), o, sp)
        elif name == "stft_kokoro":
            frames = int(lib.b2a_vocoder_stft_num_frames(n, 20, 5))
            self.inputs = [0.1 * torch.randn((B, n), generator=g, device=dev)]
            self.out = torch.empty((2, B, 11, frames), device=dev)   # magnitude and phase, one buffer
            half = B * 11 * frames * 4
            self.call = lambda c, i, o, sp: lib.b2a_kokoro_stft_transform(c.h, i[0], B, n, 20, 5, 20, o, C.c_void_p(o.value + half), sp)
        elif name == "stft_hift":
            frames = int(lib.b2a_vocoder_stft_num_frames(n, 16, 4))
            self.inputs = [0.1 * torch.randn((B, n), generator=g, device=dev)]
            self.out = torch.empty((2, B, 9, frames), device=dev)   # real and imaginary parts, one buffer
            win = np.ascontiguousarray(api.hannWindowPeriodic(16), np.float32)
            self._keep = win
            wp = win.ctypes.data_as(C.POINTER(C.c_float))
            half = B * 9 * frames * 4
            self.call = lambda c, i, o, sp: lib.b2a_stft_hifigan(c.h, i[0], B, n, 16, 4, wp, o, C.c_void_p(o.value + half), sp)
        elif name.startswith("istft"):
            nfft, hop = (16, 4) if name == "istft_hift" else (20, 5)
            F = nfft // 2 + 1
            frames = n // hop + 1
            self.inputs = [torch.exp(torch.randn((B, F, frames), generator=g, device=dev) - 2.0),
                           torch.sin(2.0 * torch.randn((B, F, frames), generator=g, device=dev))]
            self.inputs[0][torch.rand((B, F, frames), generator=g, device=dev) < 1e-3] = 150.0
            self.out = torch.empty((B, (frames - 1) * hop), device=dev)
            win = np.ascontiguousarray(api.hannWindowPeriodic(16), np.float32)
            self._keep = win
            wp = win.ctypes.data_as(C.POINTER(C.c_float))
            if name == "istft_hift":
                self.call = lambda c, i, o, sp: lib.b2a_istft_hifigan(c.h, i[0], i[1], B, frames, 16, 4, wp, o, sp)
            else:
                self.call = lambda c, i, o, sp: lib.b2a_kokoro_stft_inverse(c.h, i[0], i[1], B, frames, 20, 5, 20, o, sp)
        else:
            t = torch.arange(n, device=dev, dtype=torch.float32) / self.sr
            x = 0.1 * torch.randn((B, n), generator=g, device=dev)
            for f in (220.0, 1000.0, 3300.0):
                ph = 2 * np.pi * torch.rand((B, 1), generator=g, device=dev)
                x += 0.2 * torch.sin(2 * np.pi * f * t[None, :] + ph)
            x.clamp_(-1.0, 1.0)
            x[:, n - n // 10:] = 0.0
            self.inputs = [x]
            if name == "whisper128_ragged":
                rs = np.random.default_rng(4)
                lens = np.ascontiguousarray(rs.integers(5 * self.sr, n + 1, B), np.int64)
                lens[0] = n
                rows = np.zeros(B, np.int64)
                self._keep = (lens, rows)
                I64 = C.POINTER(C.c_int64)
                frames = int(lib.b2a_whisper_num_frames(n, 0))
                self.out = torch.empty((B, frames, 128), device=dev)
                self.audio_s = float(lens.sum()) / self.sr
                self.call = lambda c, i, o, sp: lib.b2a_whisper_log_mel_spectrogram_ragged(c.h, i[0], B, n, lens.ctypes.data_as(I64), 128, 0, o,
                                                                                           rows.ctypes.data_as(I64), sp)
            elif name in ("whisper128", "whisper80_1clip"):
                nm = 128 if name == "whisper128" else 80
                frames = int(lib.b2a_whisper_num_frames(n, 0))
                self.out = torch.empty((B, frames, nm), device=dev)
                self.call = lambda c, i, o, sp: lib.b2a_whisper_log_mel_spectrogram(c.h, i[0], B, n, nm, 0, o, sp)
            elif name == "chatterbox128":
                frames = int(lib.b2a_whisper_num_frames(n, 0))   # the last STFT frame is dropped, as in the Whisper front end
                self.out = torch.empty((B, 128, frames), device=dev)
                self.call = lambda c, i, o, sp: lib.b2a_log_mel_spectrogram_chatterbox(c.h, i[0], B, n, 128, 0, o, sp)
            elif name == "voice_encoder":
                cfg = _lib.VoiceEncConfig()
                lib.b2a_voice_enc_config_default(C.byref(cfg))
                frames = int(lib.b2a_stft_num_frames(n, cfg.n_fft, cfg.hop_size, 1))
                self.out = torch.empty((B, cfg.num_mels, frames), device=dev)
                self._keep = cfg
                self.call = lambda c, i, o, sp: lib.b2a_voice_encoder_melspectrogram(c.h, i[0], B, n, C.byref(cfg), o, sp)
            elif name == "funasr":
                frames = int(lib.b2a_funasr_num_frames(n))
                rows = int(lib.b2a_lfr_num_rows(frames, 6))
                self.out = torch.empty((B, rows, 560), device=dev)
                self.call = lambda c, i, o, sp: lib.b2a_funasr_preprocess_audio(c.h, i[0], B, n, 80, 7, 6, 1, o, sp)
            elif name == "kaldi":
                frames = int(lib.b2a_kaldi_num_frames(n, 400, 160))
                self.out = torch.empty((B, frames, 80), device=dev)
                self.call = lambda c, i, o, sp: lib.b2a_kaldi_fbank_campplus(c.h, i[0], B, n, 16000, 80, 25.0, 10.0, 1, o, sp)
            elif name == "s3gen":
                frames = int(lib.b2a_s3gen_num_frames(n, 1920, 480))
                self.out = torch.empty((B, 80, frames), device=dev)
                self.call = lambda c, i, o, sp: lib.b2a_s3gen_mel_spectrogram(c.h, i[0], B, n, 1920, 80, 24000, 480, 1920, 0, 8000, o, sp)
            else:
                raise SystemExit(f"unknown workload {name}")
        self.in_bytes = sum(t.numel() * t.element_size() for t in self.inputs)
        self.out_bytes = self.out.numel() * self.out.element_size()
        self.algo_bytes = self.in_bytes + self.out_bytes
        if name == "whisper128_ragged":   # valid samples in, the whole (zero-filled) output out
            self.algo_bytes = int(self._keep[0].sum()) * 4 + self.out_bytes
        if name == "whisper_segment":   # only the windows' rows are read (here: the rows up to frame 3000 from each seek)
            self.algo_bytes = int(sum(3000 - int(s0) for s0 in self._keep[0])) * 128 * 4 + self.out_bytes
        self.DEV, self.HOST = _lib.B2A_DEVICE, _lib.B2A_HOST
        self.h_in = self.h_out = None

    def step_device(self):
        rc = self.call(self.ctx, [C.c_void_p(t.data_ptr()) for t in self.inputs], C.c_void_p(self.out.data_ptr()), self.DEV)
        if rc != 0:
            raise RuntimeError("b200audio call failed: " + self.lib.b2a_last_error(self.ctx.h).decode())

    def prepare_host(self):
        torch = self.torch
        self.h_in = [torch.empty(t.shape, dtype=torch.float32, pin_memory=True) for t in self.inputs]
        for h, d in zip(self.h_in, self.inputs):
            h.copy_(d)
        self.h_out = torch.empty(self.out.shape, dtype=self.out.dtype, pin_memory=True)
        torch.cuda.synchronize()

    def step_host(self):
        rc = self.call(self.hctx, [C.c_void_p(t.data_ptr()) for t in self.h_in], C.c_void_p(self.h_out.data_ptr()), self.HOST)
        if rc != 0:
            raise RuntimeError("b200audio host call failed: " + self.lib.b2a_last_error(self.hctx.h).decode())


# ------------------------------------------------------------------------------------------------------
# CPU arm: the oracle (port of the reference's algorithm) over all host cores on a bounded sample
# ------------------------------------------------------------------------------------------------------

def _cpu_clip_job(args):
    name, n, sr, seed, reps = args
    from oracle import reference_dsp as R
    from tests import synth
    try:  # one BLAS / FFT thread per worker process: the processes already cover every core
        import threadpoolctl
        threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    t0 = time.perf_counter()
    if name == "whisper_segment":
        rng = np.random.default_rng(seed)
        mel = rng.standard_normal((6000, 128)).astype(np.float32)
        gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(reps):
            R.whisper_mel_segment(mel, int(seed) % 3000, 3000)
    elif name == "hift_head":
        frames = n // 4 + 1
        rng = np.random.default_rng(seed)
        h = rng.standard_normal((1, 18, frames)).astype(np.float32)
        h[:, :9] -= 2.0
        h[:, 9:] *= 2.0
        gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(reps):
            R.hift_head_istft(h, 16, 4, R.hann_window_periodic(16))
    elif name == "stft_kokoro":
        x = synth.pcm(1, n, sample_rate=sr, seed=seed)
        gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(reps):
            R.kokoro_transform(x)
    elif name == "stft_hift":
        x = synth.pcm(1, n, sample_rate=sr, seed=seed)
        gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(reps):
            R.stft_hifigan(x, 16, 4, R.hann_window_periodic(16))
    elif name.startswith("istft"):
        nfft, hop = (16, 4) if name == "istft_hift" else (20, 5)
        frames = n // hop + 1
        mag, ph = synth.mag_phase(1, nfft // 2 + 1, frames, seed=seed)
        gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(reps):
            if name == "istft_hift":
                R.istft_hifigan(mag, ph, 16, 4, R.hann_window_periodic(16))
            else:
                R.kokoro_inverse(mag, ph)
    else:
        x = synth.pcm(1, n, sample_rate=sr, seed=seed)[0]
        gen = time.perf_counter() - t0
        t0 = time.perf_counter()
        for _ in range(reps):
            if name == "whisper128":
                R.whisper_log_mel_spectrogram(x, 128)
            elif name == "whisper80_1clip":
                R.whisper_log_mel_spectrogram(x, 80)
            elif name == "funasr":
                R.preprocess_audio(x)
            elif name == "kaldi":
                R.kaldi_fbank_mean_norm(R.kaldi_fbank_camp_plus(x))
            elif name == "s3gen":
                R.s3gen_mel_spectrogram(x)
            elif name == "chatterbox128":
                R.log_mel_spectrogram_chatterbox(x, 128)
            elif name == "voice_encoder":
                R.voice_encoder_melspectrogram(x)
    return time.perf_counter() - t0, gen


TWIN_WORKLOADS = ("whisper128", "whisper80_1clip", "istft_hift")


def cpu_arm_twin(name: str, seconds: float = 2.0):
    """The compiled multi-threaded twin of the oracle (oracle/cpu_twin.cpp; BASELINE.md section 4, baseline B) on a bounded
    sample.  -> (audio-s/s, threads, sample description) or None when the twin does not cover the workload / is not built."""
    from oracle import cpu_twin as T
    from tests import synth
    if name not in TWIN_WORKLOADS or not T.available():
        return None
    w = WORKLOADS[name]
    n = int(round(w["clip_s"] * w["sr"]))
    cores = os.cpu_count() or 1

    def make(clips):
        if name == "istft_hift":
            mag, ph = synth.mag_phase(clips, 9, n // 4 + 1, seed=5000)
            return lambda: T.istft_hifigan(mag, ph, n_threads=cores)
        x = synth.pcm(clips, n, sample_rate=w["sr"], seed=5000)
        return lambda: T.whisper_log_mel_spectrogram(x, 128 if name == "whisper128" else 80, n_threads=cores)

    probe = make(cores)
    probe()
    t0 = time.perf_counter()
    probe()
    t1 = time.perf_counter() - t0
    clips = int(min(16 * cores, max(cores, cores * round(seconds / max(t1, 1e-4)))))
    run = make(clips) if clips != cores else probe
    t0 = time.perf_counter()
    run()
    dt = time.perf_counter() - t0
    return clips * w["clip_s"] / dt, cores, (f"{clips} clips x {w['clip_s']:.0f} s, C++ twin of the oracle (oracle/cpu_twin.cpp: same op order, generated "
                                             f"FFT codelets) on {cores} threads (all host cores); input synthesis excluded")


def cpu_arm(name: str, core_seconds: float = 2.0, reps: int = 1):
    """-> (audio-s/s, cores, sample description).  Warm-up (filterbank caches) excluded.
    The sample is sized so that every core is busy for about `core_seconds` (10-30 s of CPU work in total).
    Where the compiled twin covers the workload, the faster of the two CPU restatements is the reported value and the
    other one is quoted in the sample text."""
    if name == "whisper128_ragged":   # the CPU restatements process one clip at a time anyway: audio-s/s of the equal-length workload
        name = "whisper128"
    twin = cpu_arm_twin(name, core_seconds)
    if twin is not None:
        nv, ncores, nsample = cpu_arm_numpy(name, core_seconds / 2, reps)
        best = twin if twin[0] >= nv else (nv, ncores, nsample)
        other = f"NumPy/SciPy oracle on {ncores} processes: {nv:.3g} audio-s/s" if best is twin else f"C++ twin: {twin[0]:.3g} audio-s/s"
        return best[0], best[1], best[2] + "; " + other
    return cpu_arm_numpy(name, core_seconds, reps)


def cpu_arm_numpy(name: str, core_seconds: float = 2.0, reps: int = 1):
    import multiprocessing as mp
    w = WORKLOADS[name]
    n = int(round(w["clip_s"] * w["sr"]))
    cores = os.cpu_count() or 1
    with mp.get_context("fork").Pool(cores) as pool:
        pool.map(_cpu_clip_job, [(name, min(n, 16000), w["sr"], 1, 1)] * cores)  # warm caches in every worker
        t1 = max(r[0] for r in pool.map(_cpu_clip_job, [(name, n, w["sr"], 4999, 1)] * cores, chunksize=1))
        clips_per_core = int(min(64, max(1, round(core_seconds / max(t1, 1e-4)))))
        jobs = [(name, n, w["sr"], 5000 + i, reps) for i in range(cores * clips_per_core)]
        res = pool.map(_cpu_clip_job, jobs, chunksize=clips_per_core)
    # all workers run concurrently; input synthesis (second field) is excluded from the timed work
    busy = sum(r[0] for r in res) / cores
    audio = len(jobs) * reps * w["clip_s"]
    return audio / busy, cores, (f"{len(jobs) * reps} clips x {w['clip_s']:.0f} s on {cores} processes (all host cores), "
                                 "NumPy/SciPy(pocketfft) fp32 port of the reference's algorithm; input synthesis excluded")


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name = args.workload
    vals = []
    for i in range(args.warmup + args.steps):
        v, cores, sample = cpu_arm(name, core_seconds=1.0)
        if i >= args.warmup:
            vals.append(v)
    value = float(np.mean(vals))
    w = WORKLOADS[name]
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": 1e3 * (cores * w["clip_s"]) / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": w["desc"], "note": "CPU restatement of the reference (oracle/): the reference's Swift+MLX path cannot be built on Linux"},
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--workload", default="whisper128", choices=sorted(WORKLOADS))
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=None, help="override clips per GPU (debug)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--gather", default="none", choices=["none", "fused", "nccl"],
                    help="N > 1 only: also time the step WITH the features delivered to rank 0 -- 'fused' = the kernels store "
                         "straight into rank 0's HBM through peer-mapped memory (shard.FusedGather), 'nccl' = kernel then dist.gather")
    ap.add_argument("--no-e2e", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3:
        args.warmup = 3
    if args.impl == "reference":
        run_reference(args)
        return

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(v: float) -> float:
        if world == 1:
            return v
        t = torch.tensor([v], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    wl = GpuWorkload(args.workload, args.batch)
    W = WORKLOADS[args.workload]
    peak_gbs, peak_src = _peaks()

    # ---- device-resident timing --------------------------------------------------------------------
    for _ in range(args.warmup):
        wl.step_device()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    launches0 = wl.ctx.launch_count
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        wl.step_device()
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    launches = wl.ctx.launch_count - launches0
    if ms < 400.0:  # keep the sampler alive long enough to catch the clocks under this load
        t_end = time.time() + 0.5
        while time.time() < t_end:
            wl.step_device()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_step = max_over_ranks(ms / args.steps)
    value = wl.audio_s * world / (ms_step * 1e-3)

    achieved = wl.algo_bytes / (ms / args.steps * 1e-3) / 1e9
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak_gbs, "unit": "GB/s", "frac": achieved / peak_gbs, "traffic": None,
                "peak_source": peak_src, "algorithmic_bytes_per_launch": wl.algo_bytes,
                "kernel": "whole C-ABI call on the device (main fused kernel + its small fix-up / statistics kernels)"}
    prof = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(prof):
        try:
            roofline["traffic"] = json.load(open(prof)).get(args.workload)
        except Exception:
            pass

    # ---- optional: the same step with the features delivered to a consumer rank ("with and without the gather") ----
    gather = None
    if world > 1 and args.gather != "none":
        from mlx_swift_audio_b200.shard import FusedGather, gather_features
        per_clip = tuple(wl.out.shape[1:])
        if args.gather == "fused":
            if wl.out.dtype != torch.float32:
                raise SystemExit("--gather fused: float32 outputs only")
            fg = FusedGather(wl.ctx, wl.batch * world, per_clip, dst=0)
            peer_out = C.c_void_p(fg.local_out().data_ptr())

            def gstep():
                rc = wl.call(wl.ctx, [C.c_void_p(t.data_ptr()) for t in wl.inputs], peer_out, wl.DEV)
                if rc != 0:
                    raise RuntimeError("b200audio call failed: " + wl.lib.b2a_last_error(wl.ctx.h).decode())
        else:
            def gstep():
                wl.step_device()
                gather_features(wl.out, wl.batch * world, dst=0)
        for _ in range(2):
            gstep()
        barrier()
        g0 = torch.cuda.Event(enable_timing=True)
        g1 = torch.cuda.Event(enable_timing=True)
        g0.record()
        for _ in range(args.steps):
            gstep()
        g1.record()
        barrier()
        gms = max_over_ranks(g0.elapsed_time(g1) / args.steps)
        gather = {"mode": args.gather, "value": wl.audio_s * world / (gms * 1e-3), "unit": UNIT, "ms_per_step": gms,
                  "bytes_to_consumer_per_step": wl.out_bytes * (world - 1),
                  "consumer_ingress_GBps": wl.out_bytes * (world - 1) / (gms * 1e-3) / 1e9,
                  "note": "all ranks' features land in rank 0's HBM; bounded by rank 0's NVLink ingress, not by the kernels"}
        if args.gather == "fused":
            fg.close()

    # ---- end-to-end through the C ABI with pinned host buffers ---------------------------------------
    e2e = None
    if not args.no_e2e:
        wl.prepare_host()
        wl.step_host()  # warm the staging buffers
        barrier()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            wl.step_host()
        barrier()
        dt = max_over_ranks((time.perf_counter() - t0) / args.e2e_steps)
        e2e = {"value": wl.audio_s * world / dt, "unit": UNIT, "h2d_bytes_per_step": wl.in_bytes, "d2h_bytes_per_step": wl.out_bytes,
               "ms_per_step": dt * 1e3, "steps": args.e2e_steps}

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu:
        v, cores, sample = cpu_arm(args.workload)
        cpu = {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample}

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": W["desc"], "clips_per_gpu": wl.batch, "samples_per_clip": wl.n,
                           "l2_policy": "inputs+outputs per step (%.2f GB) exceed the 126 MB L2; no flush needed" % (wl.algo_bytes / 1e9)
                           if wl.algo_bytes > 4 * 126e6 else "working set fits L2: numbers are L2-warm",
                           "parallelism": f"dp{world} (clips sharded by rank, no collective)"},
                "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clocks}
        if gather is not None:
            line["gather"] = gather
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
