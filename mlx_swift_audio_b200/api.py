"""Host-side mirror of the reference's Swift DSP helpers over the C ABI (ctypes).

Function names and argument meaning follow the Swift free functions they replace (SURVEY.md
section 8b); the Swift labels become keyword arguments.  Inputs may be

  * NumPy float32 arrays  -> B2A_HOST  (the library stages them through the GPU), or
  * torch CUDA tensors    -> B2A_DEVICE (kernels run on the current torch stream, async).

and the result comes back as the same kind.  Where the reference takes one clip ``(T,)``, a
leading batch axis ``(B, T)`` is also accepted (independent clips; per-clip statistics).

Error behaviour: the reference ``fatalError``s on bad input; here ``B2AError`` is raised
(``B2ATooShort`` for "Input is too short for STFT").  There is no CPU fallback: every call
reaches a CUDA kernel or raises.
"""
from __future__ import annotations

import ctypes as C
import threading

import numpy as np

from . import _lib as L


class B2AError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"b200audio {L.STATUS_NAMES.get(status, status)}: {message}")
        self.status = status


class B2ATooShort(B2AError):
    pass


def _raise(status: int, msg: str):
    raise (B2ATooShort if status == L.B2A_E_TOO_SHORT else B2AError)(status, msg)


class Context:
    """Owns a b2a_ctx (stream, device tables, scratch).  One per host thread / stream."""

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = L.load()
        h = C.c_void_p()
        if stream is None:
            rc = self.lib.b2a_ctx_create(C.byref(h), device)
        else:
            rc = self.lib.b2a_ctx_create_on_stream(C.byref(h), device, C.c_void_p(stream))
        if rc != L.B2A_OK:
            _raise(rc, "cannot create a context (no sm_100 CUDA device?)")
        self.h = h
        self.device = device
        self.stream = int(self.lib.b2a_ctx_stream(h) or 0)   # the cudaStream_t the context enqueues on
        self._torch_order = None                            # torch stream that must wait for the call just issued (see _ctx_for)

    def close(self):
        if getattr(self, "h", None):
            self.lib.b2a_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def sync(self):
        self.check(self.lib.b2a_ctx_sync(self.h))

    def check(self, rc: int):
        pending, self._torch_order = self._torch_order, None
        if rc != L.B2A_OK:
            _raise(rc, self.lib.b2a_last_error(self.h).decode())
        if pending is not None:
            # the context enqueued on its OWN stream while the caller's tensors live on torch's stream: the result must not be
            # consumed (nor the inputs reused) on torch's stream before the kernels are done
            import torch
            ext = torch.cuda.ExternalStream(self.stream, device=pending.device)
            pending.wait_event(ext.record_event())

    @property
    def launch_count(self) -> int:
        return int(self.lib.b2a_ctx_launch_count(self.h))

    def enable_timing(self, on: bool = True):
        self.check(self.lib.b2a_ctx_enable_timing(self.h, int(on)))

    def last_kernel_ms(self) -> float:
        ms = C.c_float()
        self.check(self.lib.b2a_ctx_last_kernel_ms(self.h, C.byref(ms)))
        return float(ms.value)


_tls = threading.local()


def default_context(device: int = 0, stream: int | None = None) -> Context:
    key = (device, stream)
    cache = getattr(_tls, "ctx", None)
    if cache is None:
        cache = _tls.ctx = {}
    if key not in cache:
        cache[key] = Context(device, stream)
    return cache[key]


# ---- array plumbing -------------------------------------------------------------------------------

def _is_torch(x) -> bool:
    return type(x).__module__.startswith("torch")


class _Arr:
    """Uniform view of a NumPy array (host) or a torch CUDA tensor (device)."""

    def __init__(self, x):
        if _is_torch(x):
            import torch
            if not x.is_cuda:
                raise B2AError(L.B2A_E_BAD_ARG, "torch tensors must live on a CUDA device (use NumPy arrays for host data)")
            self.t = x.detach().to(torch.float32).contiguous()
            self.space = L.B2A_DEVICE
            self.shape = tuple(self.t.shape)
            self.ptr = C.c_void_p(self.t.data_ptr())
            self.device = self.t.device.index or 0
        else:
            self.t = np.ascontiguousarray(x, dtype=np.float32)
            self.space = L.B2A_HOST
            self.shape = self.t.shape
            self.ptr = C.c_void_p(self.t.ctypes.data)
            self.device = None

    def empty(self, shape):
        if self.space == L.B2A_DEVICE:
            import torch
            return torch.empty(shape, dtype=torch.float32, device=self.t.device)
        return np.empty(shape, np.float32)


class DevicePtr:
    """A raw device address used as ``out=`` of a front end: e.g. this rank's slice of a peer-mapped gather buffer
    (shard.FusedGather), so that the kernel's own stores deliver the features to the consumer GPU."""

    def __init__(self, address: int, shape=None):
        self.address = int(address)
        self.shape = None if shape is None else tuple(shape)

    def data_ptr(self) -> int:
        return self.address


def _arr_pcm(audio):
    """(_Arr, is_int16): 16-bit PCM (NumPy int16 array or torch int16 CUDA tensor) keeps its bytes -- the library converts
    (sample = int16 / 32768, as AVAudioFile decodes a 16-bit file) on the device; anything else goes through _Arr as float32."""
    is_t = _is_torch(audio)
    if is_t:
        import torch
        i16 = audio.dtype == torch.int16
    else:
        audio = np.asarray(audio)
        i16 = audio.dtype == np.int16
    if not i16:
        return _Arr(audio), False
    a = _Arr.__new__(_Arr)
    if is_t:
        if not audio.is_cuda:
            raise B2AError(L.B2A_E_BAD_ARG, "torch tensors must live on a CUDA device (use NumPy arrays for host data)")
        a.t, a.space, a.device = audio.contiguous(), L.B2A_DEVICE, audio.device.index or 0
        a.ptr = C.c_void_p(a.t.data_ptr())
    else:
        a.t, a.space, a.device = np.ascontiguousarray(audio), L.B2A_HOST, None
        a.ptr = C.c_void_p(a.t.ctypes.data)
    a.shape = tuple(a.t.shape)
    return a, True


def _ptr(a):
    if _is_torch(a) or isinstance(a, DevicePtr):
        return C.c_void_p(a.data_ptr())
    return C.c_void_p(a.ctypes.data)


def _out_for(a: "_Arr", shape, out):
    """Result buffer: a fresh array of the input's kind, or the caller's ``out`` (device inputs only: a contiguous float32
    torch CUDA tensor or a DevicePtr of the batched result shape)."""
    if out is None:
        return a.empty(shape)
    if a.space != L.B2A_DEVICE:
        raise B2AError(L.B2A_E_BAD_ARG, "out= needs device-resident input (torch CUDA tensor)")
    if isinstance(out, DevicePtr):
        if out.shape is not None and tuple(out.shape) != tuple(shape):
            raise B2AError(L.B2A_E_BAD_ARG, f"out has shape {out.shape}, the result has shape {tuple(shape)}")
        return out
    import torch
    if not (_is_torch(out) and out.is_cuda and out.dtype == torch.float32 and out.is_contiguous() and tuple(out.shape) == tuple(shape)):
        raise B2AError(L.B2A_E_BAD_ARG, f"out must be a contiguous float32 CUDA tensor of shape {tuple(shape)}")
    return out


def _ctx_for(a: _Arr, ctx: Context | None) -> Context:
    """The context of a call.  An explicit context that enqueues on a stream other than torch's current one (``Context(dev)``
    owns a non-blocking stream) is ordered against it for device-resident tensors: the context's stream waits for the work that
    produced the inputs before the launch, torch's stream waits for the kernels after it (``Context.check``)."""
    if ctx is not None:
        if a.space == L.B2A_DEVICE:
            import torch
            cur = torch.cuda.current_stream(a.device)
            if int(cur.cuda_stream) != ctx.stream:
                torch.cuda.ExternalStream(ctx.stream, device=cur.device).wait_event(cur.record_event())
                ctx._torch_order = cur
        return ctx
    if a.space == L.B2A_DEVICE:
        import torch
        return default_context(a.device, torch.cuda.current_stream(a.device).cuda_stream)
    return default_context(0, None)


def _batched(a: _Arr, base_ndim: int):
    """-> (batch, n, had_batch_axis)"""
    if len(a.shape) == base_ndim:
        return 1, a.shape[-1], False
    if len(a.shape) == base_ndim + 1:
        return a.shape[0], a.shape[-1], True
    raise B2AError(L.B2A_E_BAD_ARG, f"expected {base_ndim}-D input (or one leading batch axis), got shape {a.shape}")


def _fptr(w: np.ndarray):
    return w.ctypes.data_as(C.POINTER(C.c_float))


# ---- host-only helpers (windows, filterbanks, shape rules) --------------------------------------------

def _window(kind: int, length: int) -> np.ndarray:
    out = np.empty(length, np.float32)
    rc = L.load().b2a_window(kind, length, _fptr(out))
    if rc != L.B2A_OK:
        _raise(rc, "bad window parameters")
    return out


def whisperHannWindow(length: int) -> np.ndarray:
    """STT/Whisper/WhisperAudio.swift:32-44"""
    return _window(L.WIN_WHISPER_HANN, length)


def hanningWindow(length: int) -> np.ndarray:
    """Codec/S3Tokenizer/S3TokenizerUtils.swift:213-221 (== Kokoro ``hanning``, MLXSTFT.swift:12-20)"""
    return _window(L.WIN_HANNING, length)


hanning = hanningWindow


def hammingWindow(length: int) -> np.ndarray:
    """STT/FunASR/FunASRAudio.swift:35-45"""
    return _window(L.WIN_HAMMING, length)


def poveyWindow(size: int) -> np.ndarray:
    """Codec/S3Gen/CAMPPlus.swift:15-19"""
    return _window(L.WIN_POVEY, size)


def hannWindowPeriodic(size: int) -> np.ndarray:
    """Codec/S3Gen/HiFiGAN.swift:15-20"""
    return _window(L.WIN_HANN_PERIODIC, size)


cosyVoice3HannWindowPeriodic = hannWindowPeriodic  # CausalHiFTGenerator.swift:429-432


def melFilters(sampleRate: int, nFft: int, nMels: int, fMin: float = 0.0, fMax: float | None = None) -> np.ndarray:
    """Codec/S3Tokenizer/S3TokenizerUtils.swift:301-375 -> (nMels, nFft/2+1)"""
    out = np.empty((nMels, nFft // 2 + 1), np.float32)
    rc = L.load().b2a_mel_filters(sampleRate, nFft, nMels, fMin, -1.0 if fMax is None else fMax, _fptr(out))
    if rc != L.B2A_OK:
        _raise(rc, "bad filterbank parameters")
    return out


def funASRMelFilters(sampleRate: int = 16000, nFft: int = 400, nMels: int = 80) -> np.ndarray:
    """STT/FunASR/FunASRAudio.swift:322-396 -> (nMels, nFft/2)"""
    out = np.empty((nMels, nFft // 2), np.float32)
    rc = L.load().b2a_funasr_mel_filters(sampleRate, nFft, nMels, _fptr(out))
    if rc != L.B2A_OK:
        _raise(rc, "bad filterbank parameters")
    return out


def melFiltersHTK(sampleRate: int, nFft: int, nMels: int, fMin: float, fMax: float) -> np.ndarray:
    """Codec/S3Gen/CAMPPlus.swift:134-175 -> (nFft/2+1, nMels)"""
    out = np.empty((nFft // 2 + 1, nMels), np.float32)
    rc = L.load().b2a_mel_filters_htk(sampleRate, nFft, nMels, fMin, fMax, _fptr(out))
    if rc != L.B2A_OK:
        _raise(rc, "bad filterbank parameters")
    return out


def nextPowerOf2(n: int) -> int:
    """Codec/S3Gen/CAMPPlus.swift:22-29"""
    return int(L.load().b2a_next_power_of_2(n))


def computeFeatureLength(audioLength: int, hopLength: int = 160, lfrN: int = 6) -> int:
    """STT/FunASR/FunASRAudio.swift:225-235"""
    return int(L.load().b2a_funasr_compute_feature_length(audioLength, hopLength, lfrN))


def reflectPadIndex(i: int, n: int, padding: int) -> int:
    """Source index of padded position i under reflectPad (S3TokenizerUtils.swift:266-298)."""
    return int(L.load().b2a_reflect_pad_index(i, n, padding))


# ---- front ends ------------------------------------------------------------------------------------

def padOrTrim(array, length: int = 480000, ctx: Context | None = None):
    """STT/Whisper/WhisperAudio.swift:54-67"""
    a = _Arr(array)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    out = a.empty((b, length))
    c.check(c.lib.b2a_pad_or_trim(c.h, a.ptr, b, n, length, _ptr(out), a.space))
    return out if had else out[0]


def reflectPad(x, padding: int, ctx: Context | None = None):
    """Codec/S3Tokenizer/S3TokenizerUtils.swift:266-298 (== reflectPad1D, STT/FunASR/FunASRAudio.swift:280-310): (T,) -> (T + 2*padding,)"""
    a = _Arr(x)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    out = a.empty((b, n + 2 * padding))
    c.check(c.lib.b2a_reflect_pad(c.h, a.ptr, b, n, padding, _ptr(out), a.space))
    return out if had else out[0]


reflectPad1D = reflectPad


def whisperLogMelSpectrogram(audio, nMels: int, padding: int = 0, ctx: Context | None = None, out=None):
    """STT/Whisper/WhisperAudio.swift:78-137 -> (T', nMels)"""
    a = _Arr(audio)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    frames = int(c.lib.b2a_whisper_num_frames(n, padding))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = _out_for(a, (b, frames, nMels), out)
    c.check(c.lib.b2a_whisper_log_mel_spectrogram(c.h, a.ptr, b, n, nMels, padding, _ptr(out), a.space))
    return out if had or isinstance(out, DevicePtr) else out[0]


def whisperLogMelSpectrogramF16(audio, nMels: int, padding: int = 0, ctx: Context | None = None, out=None):
    """``whisperLogMelSpectrogram(...).asType(.float16)`` in one kernel (STT/Whisper/WhisperAudio.swift:78-137 followed by the cast
    of STT/Whisper/WhisperSTT.swift:156-157,181-182) -> (T', nMels) float16.  ``audio`` may also be 16-bit PCM (NumPy int16 array
    or torch int16 CUDA tensor): sample = int16 / 32768, as AVAudioFile decodes a 16-bit file (WhisperEngine.swift:327-369)."""
    return _whisper_typed(audio, nMels, padding, ctx, out, True)


def whisperLogMelSpectrogramPCM16(audio_i16, nMels: int, padding: int = 0, ctx: Context | None = None, out=None):
    """whisperLogMelSpectrogram of 16-bit PCM (sample = int16 / 32768) -> (T', nMels) float32."""
    return _whisper_typed(audio_i16, nMels, padding, ctx, out, False)


def _whisper_typed(audio, nMels, padding, ctx, out, f16: bool):
    a, i16 = _arr_pcm(audio)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    frames = int(c.lib.b2a_whisper_num_frames(n, padding))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    shape = (b, frames, nMels)
    if out is None:
        if a.space == L.B2A_DEVICE:
            import torch
            out = torch.empty(shape, dtype=torch.float16 if f16 else torch.float32, device=a.t.device)
        else:
            out = np.empty(shape, np.float16 if f16 else np.float32)
    elif not isinstance(out, DevicePtr):
        import torch
        want = torch.float16 if f16 else torch.float32
        if not (_is_torch(out) and out.is_cuda and out.dtype == want and out.is_contiguous() and tuple(out.shape) == shape and a.space == L.B2A_DEVICE):
            raise B2AError(L.B2A_E_BAD_ARG, f"out must be a contiguous {want} CUDA tensor of shape {shape} (device-resident input)")
    if i16:
        rc = c.lib.b2a_whisper_log_mel_spectrogram_pcm16(c.h, a.ptr, b, n, nMels, padding, int(f16), _ptr(out), a.space)
    elif f16:
        rc = c.lib.b2a_whisper_log_mel_spectrogram_f16(c.h, a.ptr, b, n, nMels, padding, _ptr(out), a.space)
    else:
        rc = c.lib.b2a_whisper_log_mel_spectrogram(c.h, a.ptr, b, n, nMels, padding, _ptr(out), a.space)
    c.check(rc)
    return out if had or isinstance(out, DevicePtr) else out[0]


def logMelSpectrogramChatterbox(audio, nMels: int = 128, padding: int = 0, ctx: Context | None = None, out=None):
    """Codec/S3Tokenizer/S3TokenizerUtils.swift:160-208 -> (nMels, T')"""
    a = _Arr(audio)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    frames = int(c.lib.b2a_whisper_num_frames(n, padding))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = _out_for(a, (b, nMels, frames), out)
    c.check(c.lib.b2a_log_mel_spectrogram_chatterbox(c.h, a.ptr, b, n, nMels, padding, _ptr(out), a.space))
    return out if had or isinstance(out, DevicePtr) else out[0]


def logMelSpectrogramCAMPPlus(audio, sampleRate: int = 16000, numMelBins: int = 128, ctx: Context | None = None):
    """TTS/CosyVoice2/CosyVoice2TTS.swift:787-795 (thin wrapper over logMelSpectrogramChatterbox)"""
    return logMelSpectrogramChatterbox(audio, nMels=numMelBins, padding=0, ctx=ctx)


def funASRLogMelSpectrogram(audio, nMels: int = 80, nFft: int = 400, hopLength: int = 160, ctx: Context | None = None, out=None):
    """STT/FunASR/FunASRAudio.swift:57-94 -> (T', nMels)"""
    if (nFft, hopLength) != (400, 160):
        _raise(L.B2A_E_UNSUPPORTED, "Fun-ASR log-mel is built for nFft 400 / hopLength 160")
    a = _Arr(audio)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    frames = int(c.lib.b2a_funasr_num_frames(n))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = _out_for(a, (b, frames, nMels), out)
    c.check(c.lib.b2a_funasr_log_mel_spectrogram(c.h, a.ptr, b, n, nMels, _ptr(out), a.space))
    return out if had or isinstance(out, DevicePtr) else out[0]


def applyLFR(features, lfrM: int = 7, lfrN: int = 6, ctx: Context | None = None):
    """STT/FunASR/FunASRAudio.swift:108-154 : (T, M) -> (ceil(T/lfrN), lfrM*M)"""
    a = _Arr(features)
    if len(a.shape) == 2:
        b, had = 1, False
    elif len(a.shape) == 3:
        b, had = a.shape[0], True
    else:
        raise B2AError(L.B2A_E_BAD_ARG, "applyLFR expects (T, M) or (B, T, M)")
    t, m = a.shape[-2], a.shape[-1]
    c = _ctx_for(a, ctx)
    rows = int(c.lib.b2a_lfr_num_rows(t, lfrN))
    out = a.empty((b, rows, lfrM * m))
    c.check(c.lib.b2a_apply_lfr(c.h, a.ptr, b, t, m, lfrM, lfrN, _ptr(out), a.space))
    return out if had else out[0]


def applyCMVN(features, cmvnMean=None, cmvnIstd=None, ctx: Context | None = None):
    """STT/FunASR/FunASRAudio.swift:165-180"""
    a = _Arr(features)
    if len(a.shape) == 2:
        b, had = 1, False
    elif len(a.shape) == 3:
        b, had = a.shape[0], True
    else:
        raise B2AError(L.B2A_E_BAD_ARG, "applyCMVN expects (T, D) or (B, T, D)")
    t, d = a.shape[-2], a.shape[-1]
    c = _ctx_for(a, ctx)
    out = a.empty((b, t, d))
    pm = pi = None
    keep = []
    if cmvnMean is not None or cmvnIstd is not None:
        if cmvnMean is None or cmvnIstd is None:
            raise B2AError(L.B2A_E_BAD_ARG, "cmvnMean and cmvnIstd must both be given")
        if a.space == L.B2A_DEVICE:
            import torch
            m_ = torch.as_tensor(cmvnMean, dtype=torch.float32, device=a.t.device).contiguous()
            i_ = torch.as_tensor(cmvnIstd, dtype=torch.float32, device=a.t.device).contiguous()
        else:
            m_ = np.ascontiguousarray(cmvnMean, np.float32)
            i_ = np.ascontiguousarray(cmvnIstd, np.float32)
        keep = [m_, i_]
        pm, pi = _ptr(m_), _ptr(i_)
    c.check(c.lib.b2a_apply_cmvn(c.h, a.ptr, b, t, d, pm, pi, _ptr(out), a.space))
    del keep
    return out if had else out[0]


def preprocessAudio(audio, nMels: int = 80, lfrM: int = 7, lfrN: int = 6, applyNormalization: bool = True,
                    ctx: Context | None = None, out=None):
    """STT/FunASR/FunASRAudio.swift:197-216 -> (ceil(T'/lfrN), nMels*lfrM).  ``audio`` may be 16-bit PCM (int16 / 32768)."""
    a, i16 = _arr_pcm(audio)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    frames = int(c.lib.b2a_funasr_num_frames(n))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    rows = int(c.lib.b2a_lfr_num_rows(frames, lfrN))
    out = _out_for(a, (b, rows, nMels * lfrM), out)
    fn = c.lib.b2a_funasr_preprocess_audio_pcm16 if i16 else c.lib.b2a_funasr_preprocess_audio
    c.check(fn(c.h, a.ptr, b, n, nMels, lfrM, lfrN, int(applyNormalization), _ptr(out), a.space))
    return out if had or isinstance(out, DevicePtr) else out[0]


def kaldiFbankCAMPPlus(audio, sampleRate: int = 16000, numMelBins: int = 80, frameLength: float = 25.0,
                       frameShift: float = 10.0, meanNorm: bool = False, ctx: Context | None = None, out=None):
    """Codec/S3Gen/CAMPPlus.swift:32-106 -> (T', numMelBins).  ``meanNorm`` adds the caller-side
    ``fbank - mean(fbank, axis: 0)`` of CAMPPlus.inference (:797-802).  ``audio`` may be 16-bit PCM (int16 / 32768)."""
    a, i16 = _arr_pcm(audio)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    win = int(np.float32(sampleRate) * np.float32(frameLength) / np.float32(1000))
    hop = int(np.float32(sampleRate) * np.float32(frameShift) / np.float32(1000))
    frames = int(c.lib.b2a_kaldi_num_frames(n, win, hop))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "signal shorter than one analysis window")
    out = _out_for(a, (b, frames, numMelBins), out)
    fn = c.lib.b2a_kaldi_fbank_campplus_pcm16 if i16 else c.lib.b2a_kaldi_fbank_campplus
    c.check(fn(c.h, a.ptr, b, n, sampleRate, numMelBins, frameLength, frameShift, int(meanNorm), _ptr(out), a.space))
    return out if had or isinstance(out, DevicePtr) else out[0]


def s3genMelSpectrogram(y, nFft: int = 1920, numMels: int = 80, samplingRate: int = 24000, hopSize: int = 480,
                        winSize: int = 1920, fmin: int = 0, fmax: int = 8000, center: bool = False,
                        ctx: Context | None = None, out=None):
    """Codec/S3Gen/Mel/S3GenMel.swift:43-102 : (B, T) or (T,) -> (B, numMels, T') or (numMels, T').  ``y`` may be 16-bit PCM (int16 / 32768)."""
    a, i16 = _arr_pcm(y)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    frames = int(c.lib.b2a_s3gen_num_frames(n, nFft, hopSize))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = _out_for(a, (b, numMels, frames), out)
    fn = c.lib.b2a_s3gen_mel_spectrogram_pcm16 if i16 else c.lib.b2a_s3gen_mel_spectrogram
    c.check(fn(c.h, a.ptr, b, n, nFft, numMels, samplingRate, hopSize, winSize, fmin, fmax, _ptr(out), a.space))
    return out if had or isinstance(out, DevicePtr) else out[0]


# ---- ragged batches: one launch for clips of different lengths (include/b200audio.h, "ragged batches") ------------

def _ragged(audio, lengths, ctx):
    a = _Arr(audio)
    if len(a.shape) != 2:
        raise B2AError(L.B2A_E_BAD_ARG, f"ragged batches take a (B, T_max) array, got shape {a.shape}")
    ln = np.ascontiguousarray(lengths, dtype=np.int64)
    if ln.shape != (a.shape[0],):
        raise B2AError(L.B2A_E_BAD_ARG, "lengths must hold one entry per clip")
    rows = np.zeros(a.shape[0], np.int64)
    I64 = C.POINTER(C.c_int64)
    return a, _ctx_for(a, ctx), ln, rows, ln.ctypes.data_as(I64), rows.ctypes.data_as(I64)


def whisperLogMelSpectrogramRagged(audio, lengths, nMels: int, padding: int = 0, ctx: Context | None = None):
    """whisperLogMelSpectrogram (WhisperAudio.swift:78-137) of every audio[b, :lengths[b]] in one launch
    -> ((B, T'max, nMels) with rows past a clip's frame count zero, frames per clip)"""
    a, c, ln, rows, lp, rp = _ragged(audio, lengths, ctx)
    b, n = a.shape
    frames = int(c.lib.b2a_whisper_num_frames(n, padding))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = a.empty((b, frames, nMels))
    c.check(c.lib.b2a_whisper_log_mel_spectrogram_ragged(c.h, a.ptr, b, n, lp, nMels, padding, _ptr(out), rp, a.space))
    return out, rows


def logMelSpectrogramChatterboxRagged(audio, lengths, nMels: int = 128, padding: int = 0, ctx: Context | None = None):
    """logMelSpectrogramChatterbox (S3TokenizerUtils.swift:160-208) per clip -> ((B, nMels, T'max), frames per clip)"""
    a, c, ln, rows, lp, rp = _ragged(audio, lengths, ctx)
    b, n = a.shape
    frames = int(c.lib.b2a_whisper_num_frames(n, padding))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = a.empty((b, nMels, frames))
    c.check(c.lib.b2a_log_mel_spectrogram_chatterbox_ragged(c.h, a.ptr, b, n, lp, nMels, padding, _ptr(out), rp, a.space))
    return out, rows


def preprocessAudioRagged(audio, lengths, nMels: int = 80, lfrM: int = 7, lfrN: int = 6, applyNormalization: bool = True,
                          ctx: Context | None = None):
    """preprocessAudio (FunASRAudio.swift:197-216) per clip -> ((B, rows_max, nMels*lfrM), LFR rows per clip)"""
    a, c, ln, rows, lp, rp = _ragged(audio, lengths, ctx)
    b, n = a.shape
    frames = int(c.lib.b2a_funasr_num_frames(n))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = a.empty((b, int(c.lib.b2a_lfr_num_rows(frames, lfrN)), nMels * lfrM))
    c.check(c.lib.b2a_funasr_preprocess_audio_ragged(c.h, a.ptr, b, n, lp, nMels, lfrM, lfrN, int(applyNormalization), _ptr(out), rp, a.space))
    return out, rows


def kaldiFbankCAMPPlusRagged(audio, lengths, sampleRate: int = 16000, numMelBins: int = 80, frameLength: float = 25.0,
                             frameShift: float = 10.0, meanNorm: bool = False, ctx: Context | None = None):
    """kaldiFbankCAMPPlus (CAMPPlus.swift:32-106) per clip -> ((B, T'max, numMelBins), frames per clip)"""
    a, c, ln, rows, lp, rp = _ragged(audio, lengths, ctx)
    b, n = a.shape
    win = int(np.float32(sampleRate) * np.float32(frameLength) / np.float32(1000))
    hop = int(np.float32(sampleRate) * np.float32(frameShift) / np.float32(1000))
    frames = int(c.lib.b2a_kaldi_num_frames(n, win, hop))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "signal shorter than one analysis window")
    out = a.empty((b, frames, numMelBins))
    c.check(c.lib.b2a_kaldi_fbank_campplus_ragged(c.h, a.ptr, b, n, lp, sampleRate, numMelBins, frameLength, frameShift, int(meanNorm),
                                                  _ptr(out), rp, a.space))
    return out, rows


def funASRLogMelSpectrogramRagged(audio, lengths, nMels: int = 80, ctx: Context | None = None):
    """funASRLogMelSpectrogram (FunASRAudio.swift:57-94) per clip -> ((B, T'max, nMels), frames per clip)"""
    a, c, ln, rows, lp, rp = _ragged(audio, lengths, ctx)
    b, n = a.shape
    frames = int(c.lib.b2a_funasr_num_frames(n))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = a.empty((b, frames, nMels))
    c.check(c.lib.b2a_funasr_log_mel_spectrogram_ragged(c.h, a.ptr, b, n, lp, nMels, _ptr(out), rp, a.space))
    return out, rows


def voiceEncoderMelspectrogramRagged(wav, lengths, config: L.VoiceEncConfig | None = None, ctx: Context | None = None):
    """voiceEncoderMelspectrogram (VoiceEncoderMelspec.swift:17-68) per clip -> ((B, numMels, T'max), frames per clip)"""
    cfg = config
    if cfg is None:
        cfg = L.VoiceEncConfig()
        L.load().b2a_voice_enc_config_default(C.byref(cfg))
    a, c, ln, rows, lp, rp = _ragged(wav, lengths, ctx)
    b, n = a.shape
    frames = int(c.lib.b2a_stft_num_frames(n, cfg.n_fft, cfg.hop_size, 1))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = a.empty((b, cfg.num_mels, frames))
    c.check(c.lib.b2a_voice_encoder_melspectrogram_ragged(c.h, a.ptr, b, n, lp, C.byref(cfg), _ptr(out), rp, a.space))
    return out, rows


def s3genMelSpectrogramRagged(y, lengths, nFft: int = 1920, numMels: int = 80, samplingRate: int = 24000, hopSize: int = 480,
                              winSize: int = 1920, fmin: int = 0, fmax: int = 8000, ctx: Context | None = None):
    """s3genMelSpectrogram (S3GenMel.swift:43-102) per clip -> ((B, numMels, T'max), frames per clip)"""
    a, c, ln, rows, lp, rp = _ragged(y, lengths, ctx)
    b, n = a.shape
    frames = int(c.lib.b2a_s3gen_num_frames(n, nFft, hopSize))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = a.empty((b, numMels, frames))
    c.check(c.lib.b2a_s3gen_mel_spectrogram_ragged(c.h, a.ptr, b, n, lp, nFft, numMels, samplingRate, hopSize, winSize, fmin, fmax,
                                                   _ptr(out), rp, a.space))
    return out, rows


def voiceEncoderMelspectrogram(wav, config: L.VoiceEncConfig | None = None, pad: bool = True, ctx: Context | None = None):
    """TTS/Chatterbox/VoiceEncoder/VoiceEncoderMelspec.swift:17-68 -> (numMels, T')"""
    cfg = config
    if cfg is None:
        cfg = L.VoiceEncConfig()
        L.load().b2a_voice_enc_config_default(C.byref(cfg))
    a = _Arr(wav)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    frames = int(c.lib.b2a_stft_num_frames(n, cfg.n_fft, cfg.hop_size, 1))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = a.empty((b, cfg.num_mels, frames))
    c.check(c.lib.b2a_voice_encoder_melspectrogram(c.h, a.ptr, b, n, C.byref(cfg), _ptr(out), a.space))
    return out if had else out[0]


def stft(x, window, nFft: int, hopLength: int, winLength: int | None = None, center: bool = True,
         padMode: str = "reflect", ctx: Context | None = None):
    """Codec/S3Tokenizer/S3TokenizerUtils.swift:224-263 -> complex64 (T', nFft/2+1).
    ``winLength`` and ``padMode`` are ignored, as in the reference (:229,231)."""
    a = _Arr(x)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    w = np.ascontiguousarray(window, np.float32)
    frames = int(c.lib.b2a_stft_num_frames(n, nFft, hopLength, int(center)))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short for STFT")
    out = a.empty((b, frames, nFft // 2 + 1, 2))
    c.check(c.lib.b2a_stft(c.h, a.ptr, b, n, _fptr(w), w.shape[0], nFft, hopLength, int(center), _ptr(out), a.space))
    if a.space == L.B2A_DEVICE:
        import torch
        z = torch.view_as_complex(out)
    else:
        z = out.view(np.complex64)[..., 0]
    return z if had else z[0]


# ---- vocoder STFT / iSTFT ----------------------------------------------------------------------------

def _small_stft(fn_name, x, nFft, hopLength, window, ctx):
    a = _Arr(x)
    if len(a.shape) != 2:
        raise B2AError(L.B2A_E_BAD_ARG, "expected (B, T)")
    b, n = a.shape
    c = _ctx_for(a, ctx)
    frames = int(c.lib.b2a_vocoder_stft_num_frames(n, nFft, hopLength))
    if frames <= 0:
        _raise(L.B2A_E_TOO_SHORT, "Input is too short")
    o0 = a.empty((b, nFft // 2 + 1, frames))
    o1 = a.empty((b, nFft // 2 + 1, frames))
    w = np.ascontiguousarray(window, np.float32)
    c.check(getattr(c.lib, fn_name)(c.h, a.ptr, b, n, nFft, hopLength, _fptr(w), _ptr(o0), _ptr(o1), a.space))
    return o0, o1


def stftHiFiGAN(x, nFft: int, hopLength: int, window, ctx: Context | None = None):
    """Codec/S3Gen/HiFiGAN.swift:257-295 : (B, T) -> (real, imag) each (B, nFft/2+1, frames)"""
    return _small_stft("b2a_stft_hifigan", x, nFft, hopLength, window, ctx)


def cosyVoice3Stft(x, nFft: int, hopLength: int, window, ctx: Context | None = None):
    """TTS/CosyVoice3/HiFiGAN/CausalHiFTGenerator.swift:435-460"""
    return _small_stft("b2a_cosyvoice3_stft", x, nFft, hopLength, window, ctx)


def _istft(fn_name, magnitude, phase, nFft, hopLength, window, ctx):
    m = _Arr(magnitude)
    p = _Arr(phase)
    if len(m.shape) != 3 or m.shape != p.shape or m.space != p.space:
        raise B2AError(L.B2A_E_BAD_ARG, "magnitude and phase must both be (B, F, frames) in the same memory space")
    b, f, frames = m.shape
    if f != nFft // 2 + 1:
        raise B2AError(L.B2A_E_BAD_ARG, f"expected {nFft // 2 + 1} frequency bins, got {f}")
    c = _ctx_for(m, ctx)
    if frames < 2:
        _raise(L.B2A_E_TOO_SHORT, "iSTFT needs at least 2 frames")
    out = m.empty((b, (frames - 1) * hopLength))
    w = np.ascontiguousarray(window, np.float32)
    c.check(getattr(c.lib, fn_name)(c.h, m.ptr, p.ptr, b, frames, nFft, hopLength, _fptr(w), _ptr(out), m.space))
    return out


def istftHiFiGAN(magnitude, phase, nFft: int, hopLength: int, window, ctx: Context | None = None):
    """Codec/S3Gen/HiFiGAN.swift:298-367 -> (B, (frames-1)*hop)"""
    return _istft("b2a_istft_hifigan", magnitude, phase, nFft, hopLength, window, ctx)


def cosyVoice3Istft(magnitude, phase, nFft: int, hopLength: int, window, ctx: Context | None = None):
    """TTS/CosyVoice3/HiFiGAN/CausalHiFTGenerator.swift:463-514"""
    return _istft("b2a_cosyvoice3_istft", magnitude, phase, nFft, hopLength, window, ctx)


def s3TokenizerSegments(mel, melLen, window: int = 3000, stride: int = 2600, ctx: Context | None = None):
    """Unified segment batch of S3Tokenizer.quantize / quantizeMixedBatch (Codec/S3Tokenizer/S3Tokenizer.swift:474-571):
    mel (B, M, Tmax), melLen (B,) -> (segments (S, M, window), lengths (S,) int32, [(batch index, segment index), ...])."""
    a = _Arr(mel)
    if len(a.shape) != 3:
        raise B2AError(L.B2A_E_BAD_ARG, "mel must be (B, M, Tmax)")
    b, m, t = a.shape
    ml = np.ascontiguousarray(melLen, np.int64)
    if ml.shape != (b,) or (ml > t).any():
        raise B2AError(L.B2A_E_BAD_ARG, "melLen must hold one length <= Tmax per clip")
    c = _ctx_for(a, ctx)
    I64, I32 = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    n = int(c.lib.b2a_s3tokenizer_plan_segments(ml.ctypes.data_as(I64), b, window, stride, None, None, None, 0))
    if n < 0:
        raise B2AError(L.B2A_E_BAD_ARG, "bad segment plan arguments")
    bi, st, ln = (np.empty(n, np.int32) for _ in range(3))
    c.lib.b2a_s3tokenizer_plan_segments(ml.ctypes.data_as(I64), b, window, stride, bi.ctypes.data_as(I32), st.ctypes.data_as(I32),
                                        ln.ctypes.data_as(I32), n)
    out = a.empty((n, m, window))
    c.check(c.lib.b2a_s3tokenizer_gather_segments(c.h, a.ptr, b, m, t, n, bi.ctypes.data_as(I32), st.ctypes.data_as(I32),
                                                  ln.ctypes.data_as(I32), window, _ptr(out), a.space))
    info, k, prev = [], 0, -1
    for i in range(n):
        k = k + 1 if bi[i] == prev else 0
        prev = bi[i]
        info.append((int(bi[i]), k))
    return out, ln, info


def resampleAudio(audio, fromRate: int, toRate: int, ctx: Context | None = None):
    """resampleAudio (TTS/CosyVoice2/CosyVoice2TTS.swift:733-744) = linearInterpolate1d
    (TTS/CosyVoice2/HiFiGAN/CosyHiFTGenerator.swift:17-58) on (T,) or (B, T); bit-exact fp32."""
    a = _Arr(audio)
    b, n, had_batch = _batched(a, 1)
    c = _ctx_for(a, ctx)
    new_t = int(c.lib.b2a_resample_linear_length(n, fromRate, toRate))
    out = a.empty((b, new_t))
    c.check(c.lib.b2a_resample_linear(c.h, a.ptr, b, n, fromRate, toRate, _ptr(out), a.space))
    return out if had_batch else out[0]


def resamplePoly(audio, fromRate: int, toRate: int, ctx: Context | None = None):
    """NON-PARITY EXTENSION: anti-aliased polyphase resampling, defined as ``scipy.signal.resample_poly(audio, up, down)`` with its
    default Kaiser design -- the stand-in for ``AudioResampler.resample`` (Audio/AudioResampler.swift:15-88: AVAudioConverter, no
    reproducible definition).  (T,) or (B, T) -> (ceil(T up / down),)"""
    a = _Arr(audio)
    b, n, had = _batched(a, 1)
    c = _ctx_for(a, ctx)
    new_t = int(c.lib.b2a_resample_poly_length(n, fromRate, toRate))
    out = a.empty((b, new_t))
    c.check(c.lib.b2a_resample_poly(c.h, a.ptr, b, n, fromRate, toRate, _ptr(out), a.space))
    return out if had else out[0]


def whisperMelSegment(mel, seek, contentFrames, length: int = 3000, ctx: Context | None = None):
    """Seek window of the Whisper decode loop (STT/Whisper/WhisperSTT.swift:171-182,624-635): rows
    [seek, seek + min(length, contentFrames - seek)) of the fp32 log-mel (T', M) or (B, T', M), zero-padded to ``length`` rows,
    as float16.  ``seek`` / ``contentFrames``: ints, or one per clip."""
    a = _Arr(mel)
    had_batch = len(a.shape) == 3
    shape = a.shape if had_batch else (1,) + tuple(a.shape)
    b, t, m = shape
    sk = np.ascontiguousarray(np.broadcast_to(np.asarray(seek, np.int64), (b,)))
    cf = np.ascontiguousarray(np.broadcast_to(np.asarray(contentFrames, np.int64), (b,)))
    c = _ctx_for(a, ctx)
    if a.space == L.B2A_DEVICE:
        import torch
        out = torch.empty((b, length, m), dtype=torch.float16, device=a.t.device)
    else:
        out = np.empty((b, length, m), np.float16)
    I64 = C.POINTER(C.c_int64)
    c.check(c.lib.b2a_whisper_mel_segment_f16(c.h, a.ptr, b, t, m, sk.ctypes.data_as(I64), cf.ctypes.data_as(I64), length, _ptr(out),
                                              a.space))
    return out if had_batch else out[0]


def _head_istft(conv_out, n_fft, hop, ctx, call, out_shape):
    h = _Arr(conv_out)
    if len(h.shape) != 3 or h.shape[1] != n_fft + 2:
        raise B2AError(L.B2A_E_BAD_ARG, f"expected the convolution output (B, {n_fft + 2}, frames)")
    b, _, frames = h.shape
    c = _ctx_for(h, ctx)
    if frames < 2:
        _raise(L.B2A_E_TOO_SHORT, "iSTFT needs at least 2 frames")
    out = h.empty(out_shape(b, (frames - 1) * hop))
    c.check(call(c, h, b, frames, out))
    return out


def hiftHeadIstft(convOut, nFft: int, hopLength: int, window, audioLimit: float = 0.99, ctx: Context | None = None):
    """Tail of HiFTGenerator.decode (Codec/S3Gen/HiFiGAN.swift:577-589) as one kernel: exp / sin split of the convPost
    output (B, nFft+2, frames), istftHiFiGAN, clip to +-audioLimit.  -> (B, (frames-1)*hop)"""
    w = np.ascontiguousarray(window, np.float32)
    return _head_istft(convOut, nFft, hopLength, ctx,
                       lambda c, h, b, frames, out: c.lib.b2a_hift_head_istft(c.h, h.ptr, b, frames, nFft, hopLength, _fptr(w),
                                                                              float(audioLimit), _ptr(out), h.space),
                       lambda b, n: (b, n))


def unwrap(p, ctx: Context | None = None):
    """TTS/Kokoro/Decoder/MLXSTFT.swift:23-46: numpy-style phase unwrap along the last axis"""
    a = _Arr(p)
    if len(a.shape) < 1 or a.shape[-1] < 1:
        raise B2AError(L.B2A_E_BAD_ARG, "unwrap needs at least one sample along the last axis")
    n = a.shape[-1]
    rows = int(np.prod(a.shape[:-1])) if len(a.shape) > 1 else 1
    c = _ctx_for(a, ctx)
    out = a.empty(a.shape)
    c.check(c.lib.b2a_unwrap(c.h, a.ptr, rows, n, _ptr(out), a.space))
    return out


def mlxStft(x, nFft: int = 20, hopLength: int = 5, ctx: Context | None = None):
    """TTS/Kokoro/Decoder/MLXSTFT.swift:69-113 for the configuration the package uses (periodic Hann of nFft taps, centre reflect
    padding): (T,) or (B, T) -> complex64 (F, frames) or (B, F, frames), i.e. the reference's transposed rfft of the frames."""
    w = np.ascontiguousarray(hanningWindow(nFft + 1)[:nFft], np.float32)
    a = _Arr(x)
    single = len(a.shape) == 1
    re, im = stftHiFiGAN(a.t[None] if single else a.t, nFft, hopLength, w, ctx=ctx)
    if _is_torch(re):
        import torch
        z = torch.complex(re, im)
    else:
        z = (re + 1j * im).astype(np.complex64)
    return z[0] if single else z


def s3genTrimFade(samplingRate: int = 24000) -> np.ndarray:
    """Fade-in window of S3Token2Wav (Codec/S3Gen/S3Gen.swift:259-262): zeros(sr/50) ++ (cos(linspace(pi, 0, sr/50)) + 1) / 2"""
    out = np.empty(2 * (samplingRate // 50), np.float32)
    rc = L.load().b2a_s3gen_trim_fade(samplingRate, _fptr(out))
    if rc != L.B2A_OK:
        _raise(rc, "bad sampling rate")
    return out


def hiftHeadIstftFade(convOut, nFft: int, hopLength: int, window, trimFade, audioLimit: float = 0.99, ctx: Context | None = None):
    """hiftHeadIstft followed by ``result[..., :fadeLen] *= trimFade`` of S3Token2Wav.callAsFunction (S3Gen.swift:284-289), one kernel."""
    w = np.ascontiguousarray(window, np.float32)
    fd = np.ascontiguousarray(trimFade, np.float32)
    return _head_istft(convOut, nFft, hopLength, ctx,
                       lambda c, h, b, frames, out: c.lib.b2a_hift_head_istft_fade(c.h, h.ptr, b, frames, nFft, hopLength, _fptr(w),
                                                                                   float(audioLimit), _fptr(fd), len(fd), _ptr(out), h.space),
                       lambda b, n: (b, n))


def kokoroHeadIstft(convOut, filterLength: int = 20, hopLength: int = 5, winLength: int = 20, ctx: Context | None = None):
    """Tail of the Kokoro generator (TTS/Kokoro/Decoder/Generator.swift:182-190) as one kernel: exp / sin split of the
    conv_post output (B, filterLength+2, frames) and MLXSTFT.inverse.  -> (B, 1, (frames-1)*hop)"""
    return _head_istft(convOut, filterLength, hopLength, ctx,
                       lambda c, h, b, frames, out: c.lib.b2a_kokoro_head_istft(c.h, h.ptr, b, frames, filterLength, hopLength,
                                                                                winLength, _ptr(out), h.space),
                       lambda b, n: (b, 1, n))


def mlxIstft(real, imag=None, hopLength: int | None = None, winLength: int | None = None, window: str = "hann", center: bool = True,
             ctx: Context | None = None):
    """TTS/Kokoro/Decoder/MLXSTFT.swift:115-163: complex spectrum (F, frames) [or (B, F, frames)] -> (L,).  The spectrum is
    given as a complex64 array, or as separate real / imaginary float32 arrays."""
    if window.lower() != "hann":
        raise B2AError(L.B2A_E_BAD_ARG, f"Only hanning is supported for window, not {window}")
    if not center:
        raise B2AError(L.B2A_E_UNSUPPORTED, "mlxIstft is built for center = true (the only form the reference calls)")
    if imag is None:
        z = np.asarray(real)
        if not np.iscomplexobj(z):
            raise B2AError(L.B2A_E_BAD_ARG, "mlxIstft needs a complex spectrum")
        inter = np.ascontiguousarray(np.stack([z.real, z.imag], axis=-1), np.float32)
    elif _is_torch(real):
        import torch
        inter = torch.stack([real, imag], dim=-1).contiguous()
    else:
        inter = np.ascontiguousarray(np.stack([np.asarray(real, np.float32), np.asarray(imag, np.float32)], axis=-1))
    a = _Arr(inter)
    had = len(a.shape) == 4
    if len(a.shape) not in (3, 4):
        raise B2AError(L.B2A_E_BAD_ARG, "spectrum must be (F, frames) or (B, F, frames)")
    b = a.shape[0] if had else 1
    f, frames = a.shape[-3], a.shape[-2]
    win = winLength if winLength is not None else (f - 1) * 2
    hop = hopLength if hopLength is not None else win // 4
    if f != win // 2 + 1:
        raise B2AError(L.B2A_E_BAD_ARG, "spectrum rows must equal winLength / 2 + 1")
    if frames < 2:
        _raise(L.B2A_E_TOO_SHORT, "iSTFT needs at least 2 frames")
    c = _ctx_for(a, ctx)
    out = a.empty((b, (frames - 1) * hop))
    c.check(c.lib.b2a_mlx_istft(c.h, a.ptr, b, frames, win, hop, _ptr(out), a.space))
    return out if had else out[0]


def mergeTokenizedSegments(tokenizedSegments, overlap: int, tokenRate: int):
    """Codec/S3Tokenizer/S3TokenizerUtils.swift:71-88 (host-side integer rule) -> list[int]"""
    segs = [np.asarray(t, np.int32).reshape(-1) for t in tokenizedSegments]
    lens = np.asarray([len(t) for t in segs], np.int64)
    flat = np.ascontiguousarray(np.concatenate(segs) if segs else np.zeros(0, np.int32), np.int32)
    if flat.size == 0:
        flat = np.zeros(1, np.int32)
    out = np.empty(max(1, int(lens.sum())), np.int32)
    I32, I64 = C.POINTER(C.c_int32), C.POINTER(C.c_int64)
    n = int(L.load().b2a_merge_tokenized_segments(flat.ctypes.data_as(I32), lens.ctypes.data_as(I64), len(segs), overlap, tokenRate,
                                                  out.ctypes.data_as(I32), out.shape[0]))
    if n < 0:
        _raise(L.B2A_E_BAD_ARG, "mergeTokenizedSegments: bad segment lengths")
    return [int(v) for v in out[:n]]


class MLXSTFT:
    """TTS/Kokoro/Decoder/MLXSTFT.swift:165-235 (``transform`` / ``inverse`` / call)."""

    def __init__(self, filterLength: int = 800, hopLength: int = 200, winLength: int = 800, window: str = "hann",
                 ctx: Context | None = None):
        if window.lower() != "hann":
            raise B2AError(L.B2A_E_BAD_ARG, f"Only hanning is supported for window, not {window}")  # MLXSTFT.swift:54
        self.filterLength, self.hopLength, self.winLength, self.window = filterLength, hopLength, winLength, window
        self.ctx = ctx
        self.magnitude = None
        self.phase = None

    def transform(self, inputData):
        a = _Arr(inputData)
        if len(a.shape) == 1:
            if a.space == L.B2A_DEVICE:
                a = _Arr(a.t[None])
            else:
                a = _Arr(a.t[None])
        b, n = a.shape
        c = _ctx_for(a, self.ctx)
        frames = int(c.lib.b2a_vocoder_stft_num_frames(n, self.filterLength, self.hopLength))
        if frames <= 0:
            _raise(L.B2A_E_TOO_SHORT, "Input is too short")
        f = self.filterLength // 2 + 1
        mag, ph = a.empty((b, f, frames)), a.empty((b, f, frames))
        c.check(c.lib.b2a_kokoro_stft_transform(c.h, a.ptr, b, n, self.filterLength, self.hopLength, self.winLength,
                                                _ptr(mag), _ptr(ph), a.space))
        return mag, ph

    def inverse(self, magnitude, phase):
        m, p = _Arr(magnitude), _Arr(phase)
        if len(m.shape) != 3 or m.shape != p.shape or m.space != p.space:
            raise B2AError(L.B2A_E_BAD_ARG, "magnitude and phase must both be (B, F, frames)")
        b, f, frames = m.shape
        c = _ctx_for(m, self.ctx)
        if frames < 2:
            _raise(L.B2A_E_TOO_SHORT, "iSTFT needs at least 2 frames")
        out = m.empty((b, 1, (frames - 1) * self.hopLength))
        c.check(c.lib.b2a_kokoro_stft_inverse(c.h, m.ptr, p.ptr, b, frames, self.filterLength, self.hopLength, self.winLength,
                                              _ptr(out), m.space))
        return out

    def __call__(self, inputData):
        mag, ph = self.transform(inputData)
        self.magnitude, self.phase = mag, ph
        rec = self.inverse(mag, ph)
        return rec[..., None, :] if not _is_torch(rec) else rec.unsqueeze(-2)
