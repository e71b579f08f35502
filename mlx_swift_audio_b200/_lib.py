"""ctypes loader for libb200audio.so (the C ABI in include/b200audio.h).

Fails loudly when the CUDA library is missing or cannot be loaded: there is no CPU path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libb200audio.so")

B2A_OK, B2A_E_BAD_ARG, B2A_E_TOO_SHORT, B2A_E_CUDA, B2A_E_UNSUPPORTED, B2A_E_NOMEM = range(6)
B2A_HOST, B2A_DEVICE = 0, 1
WIN_WHISPER_HANN, WIN_HANNING, WIN_HAMMING, WIN_POVEY, WIN_HANN_PERIODIC = range(5)

IPC_HANDLE_BYTES = 64   # B2A_IPC_HANDLE_BYTES
STATUS_NAMES = {0: "OK", 1: "BAD_ARG", 2: "TOO_SHORT", 3: "CUDA", 4: "UNSUPPORTED", 5: "NOMEM"}

_f = C.POINTER(C.c_float)
_i64 = C.c_int64
_ctx = C.c_void_p


class VoiceEncConfig(C.Structure):
    _fields_ = [("num_mels", C.c_int), ("sample_rate", C.c_int), ("n_fft", C.c_int), ("hop_size", C.c_int),
                ("win_size", C.c_int), ("fmin", C.c_int), ("fmax", C.c_int), ("mel_power", C.c_float),
                ("mel_type_db", C.c_int), ("normalized_mels", C.c_int), ("stft_magnitude_min", C.c_float)]


# name -> (restype, argtypes); exactly the symbols include/b200audio.h declares
SIGNATURES = {
    "b2a_version": (C.c_char_p, []),
    "b2a_ctx_create": (C.c_int, [C.POINTER(_ctx), C.c_int]),
    "b2a_ctx_create_on_stream": (C.c_int, [C.POINTER(_ctx), C.c_int, C.c_void_p]),
    "b2a_ctx_destroy": (C.c_int, [_ctx]),
    "b2a_ctx_sync": (C.c_int, [_ctx]),
    "b2a_ctx_stream": (C.c_void_p, [_ctx]),
    "b2a_last_error": (C.c_char_p, [_ctx]),
    "b2a_ctx_launch_count": (_i64, [_ctx]),
    "b2a_host_alloc": (C.c_int, [C.POINTER(C.c_void_p), C.c_uint64]),
    "b2a_host_free": (C.c_int, [C.c_void_p]),
    "b2a_device_alloc": (C.c_int, [_ctx, C.POINTER(C.c_void_p), C.c_uint64]),
    "b2a_device_free": (C.c_int, [_ctx, C.c_void_p]),
    "b2a_ipc_export": (C.c_int, [_ctx, C.c_void_p, C.c_char_p]),
    "b2a_ipc_open": (C.c_int, [_ctx, C.c_char_p, C.POINTER(C.c_void_p)]),
    "b2a_ipc_close": (C.c_int, [_ctx, C.c_void_p]),
    "b2a_memcpy_d2h": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, C.c_uint64]),
    "b2a_window": (C.c_int, [C.c_int, C.c_int, _f]),
    "b2a_mel_filters": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _f]),
    "b2a_funasr_mel_filters": (C.c_int, [C.c_int, C.c_int, C.c_int, _f]),
    "b2a_mel_filters_htk": (C.c_int, [C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, _f]),
    "b2a_reflect_pad_index": (_i64, [_i64, _i64, _i64]),
    "b2a_next_power_of_2": (C.c_int, [C.c_int]),
    "b2a_funasr_compute_feature_length": (_i64, [_i64, C.c_int, C.c_int]),
    "b2a_stft_num_frames": (_i64, [_i64, C.c_int, C.c_int, C.c_int]),
    "b2a_whisper_num_frames": (_i64, [_i64, _i64]),
    "b2a_funasr_num_frames": (_i64, [_i64]),
    "b2a_lfr_num_rows": (_i64, [_i64, C.c_int]),
    "b2a_kaldi_num_frames": (_i64, [_i64, C.c_int, C.c_int]),
    "b2a_s3gen_num_frames": (_i64, [_i64, C.c_int, C.c_int]),
    "b2a_vocoder_stft_num_frames": (_i64, [_i64, C.c_int, C.c_int]),
    "b2a_istft_out_length": (_i64, [_i64, C.c_int]),
    "b2a_reflect_pad": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, _i64, C.c_void_p, C.c_int]),
    "b2a_pad_or_trim": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, _i64, C.c_void_p, C.c_int]),
    "b2a_whisper_log_mel_spectrogram": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, _i64, C.c_void_p, C.c_int]),
    "b2a_whisper_log_mel_spectrogram_f16": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, _i64, C.c_void_p, C.c_int]),
    "b2a_whisper_log_mel_spectrogram_pcm16": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, _i64, C.c_int, C.c_void_p, C.c_int]),
    "b2a_whisper_log_mel_spectrogram_f16_ragged": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.POINTER(_i64), C.c_int, _i64, C.c_void_p, C.POINTER(_i64), C.c_int]),
    "b2a_log_mel_spectrogram_chatterbox": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, _i64, C.c_void_p, C.c_int]),
    "b2a_whisper_log_mel_spectrogram_ragged": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.POINTER(_i64), C.c_int, _i64, C.c_void_p, C.POINTER(_i64), C.c_int]),
    "b2a_log_mel_spectrogram_chatterbox_ragged": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.POINTER(_i64), C.c_int, _i64, C.c_void_p, C.POINTER(_i64), C.c_int]),
    "b2a_funasr_log_mel_spectrogram_ragged": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.POINTER(_i64), C.c_int, C.c_void_p, C.POINTER(_i64), C.c_int]),
    "b2a_voice_encoder_melspectrogram_ragged": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.POINTER(_i64), C.POINTER(VoiceEncConfig), C.c_void_p, C.POINTER(_i64), C.c_int]),
    "b2a_funasr_preprocess_audio_ragged": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.POINTER(_i64), C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(_i64), C.c_int]),
    "b2a_kaldi_fbank_campplus_ragged": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.POINTER(_i64), C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.POINTER(_i64), C.c_int]),
    "b2a_s3gen_mel_spectrogram_ragged": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.POINTER(_i64), C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.POINTER(_i64), C.c_int]),
    "b2a_funasr_log_mel_spectrogram": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_void_p, C.c_int]),
    "b2a_apply_lfr": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_apply_cmvn": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]),
    "b2a_funasr_preprocess_audio": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_funasr_preprocess_audio_pcm16": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_kaldi_fbank_campplus_pcm16": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_int]),
    "b2a_s3gen_mel_spectrogram_pcm16": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_kaldi_fbank_campplus": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_float, C.c_float, C.c_int, C.c_void_p, C.c_int]),
    "b2a_s3gen_mel_spectrogram": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_voice_enc_config_default": (None, [C.POINTER(VoiceEncConfig)]),
    "b2a_voice_encoder_melspectrogram": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.POINTER(VoiceEncConfig), C.c_void_p, C.c_int]),
    "b2a_stft": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, _f, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_stft_hifigan": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, _f, C.c_void_p, C.c_void_p, C.c_int]),
    "b2a_istft_hifigan": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, _i64, _i64, C.c_int, C.c_int, _f, C.c_void_p, C.c_int]),
    "b2a_cosyvoice3_stft": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, _f, C.c_void_p, C.c_void_p, C.c_int]),
    "b2a_cosyvoice3_istft": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, _i64, _i64, C.c_int, C.c_int, _f, C.c_void_p, C.c_int]),
    "b2a_kokoro_stft_transform": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int]),
    "b2a_kokoro_stft_inverse": (C.c_int, [_ctx, C.c_void_p, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_s3tokenizer_plan_segments": (_i64, [C.POINTER(_i64), _i64, _i64, _i64, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32), _i64]),
    "b2a_s3tokenizer_gather_segments": (C.c_int, [_ctx, C.c_void_p, _i64, C.c_int, _i64, _i64, C.POINTER(C.c_int32), C.POINTER(C.c_int32),
                                                  C.POINTER(C.c_int32), _i64, C.c_void_p, C.c_int]),
    "b2a_resample_linear_length": (_i64, [_i64, C.c_int, C.c_int]),
    "b2a_resample_poly_length": (_i64, [_i64, C.c_int, C.c_int]),
    "b2a_resample_poly_filter": (_i64, [_i64, C.c_int, C.c_int, _f, _i64, C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(_i64)]),
    "b2a_resample_poly": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_resample_linear": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_whisper_mel_segment_f16": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.POINTER(_i64), C.POINTER(_i64), _i64, C.c_void_p,
                                              C.c_int]),
    "b2a_mlx_istft": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_merge_tokenized_segments": (_i64, [C.POINTER(C.c_int32), C.POINTER(_i64), _i64, C.c_int, C.c_int, C.POINTER(C.c_int32), _i64]),
    "b2a_unwrap": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_void_p, C.c_int]),
    "b2a_hift_head_istft": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, _f, C.c_float, C.c_void_p, C.c_int]),
    "b2a_hift_head_istft_fade": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, _f, C.c_float, _f, _i64, C.c_void_p, C.c_int]),
    "b2a_s3gen_trim_fade": (C.c_int, [C.c_int, _f]),
    "b2a_kokoro_head_istft": (C.c_int, [_ctx, C.c_void_p, _i64, _i64, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]),
    "b2a_debug_mel_program_apply": (C.c_int, [_f, C.c_int, C.c_int, C.c_int, _f, _f]),
    "b2a_debug_wpf_mel_apply": (C.c_int, [_f, C.c_int, C.c_int, C.c_int, _f, _f]),
    "b2a_debug_plan_layout": (C.c_int, [C.c_int, C.c_int, C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b2a_debug_mel_program_dump": (C.c_int, [_f, C.c_int, C.c_int, C.c_int, C.c_int, C.POINTER(C.c_uint), C.c_int,
                                             C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]),
    "b2a_debug_whisper_tc": (C.c_int, [C.c_int]),
    "b2a_debug_wpf1920": (C.c_int, [C.c_int]),
    "b2a_debug_dyn_tiles": (C.c_int, [C.c_int]),
    "b2a_debug_tc_power_buffer": (C.c_int, [C.c_void_p]),
    "b2a_ctx_enable_timing": (C.c_int, [_ctx, C.c_int]),
    "b2a_ctx_last_kernel_ms": (C.c_int, [_ctx, _f]),
}

_lib = None


def load() -> C.CDLL:
    """Loads the CUDA extension; raises (never falls back) when it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m mlx_swift_audio_b200.build` "
            "(needs nvcc).  There is no CPU fallback for this path.")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
