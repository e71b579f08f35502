#!/usr/bin/env python3
"""Static check of mlx_swift_audio_b200/swift/B200AudioShim.swift against include/b200audio.h (no Swift toolchain in this image).

  * every `b2a_*(...)` call in the shim names a function the header declares and passes exactly as many arguments as its C
    prototype has parameters;
  * every helper of SURVEY.md section 8b has a Swift binding of the same name;
  * every compute entry point of the header is either bound by the shim or listed in NOT_BOUND with the reason.

    python tools/check_swift_shim.py     # exit code 0 when consistent
"""
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "b200audio.h")
SHIM = os.path.join(ROOT, "mlx_swift_audio_b200", "swift", "B200AudioShim.swift")

# the reference's helper names that must exist in the shim under the same name
REQUIRED_SWIFT = [
    "stft", "melFilters", "hanningWindow", "hammingWindow", "poveyWindow", "hannWindowPeriodic", "whisperHannWindow", "padOrTrim",
    "whisperLogMelSpectrogram", "logMelSpectrogramChatterbox", "funASRLogMelSpectrogram", "applyLFR", "applyCMVN", "preprocessAudio",
    "kaldiFbankCAMPPlus", "s3genMelSpectrogram", "voiceEncoderMelspectrogram", "stftHiFiGAN", "istftHiFiGAN", "cosyVoice3Stft",
    "cosyVoice3Istft", "mlxStft", "mlxIstft", "transform", "inverse", "unwrap", "reflectPad", "reflectPad1D", "funASRMelFilters",
    "computeMelFiltersHTK", "nextPowerOf2", "computeFeatureLength", "logMelSpectrogramCAMPPlus", "cosyVoice3HannWindowPeriodic",
    "mergeTokenizedSegments", "resampleAudio",
]
# header entry points the shim deliberately leaves to the caller (with the reason)
NOT_BOUND = {
    "b2a_ctx_create_on_stream": "stream sharing is a host-runtime decision", "b2a_ctx_sync": "host buffers: calls are synchronous",
    "b2a_ctx_stream": "host buffers only", "b2a_ctx_launch_count": "diagnostics", "b2a_host_alloc": "Swift arrays are used directly",
    "b2a_host_free": "see b2a_host_alloc", "b2a_device_alloc": "multi-GPU: INTEGRATION.md section 6", "b2a_device_free": "multi-GPU",
    "b2a_ipc_export": "multi-GPU", "b2a_ipc_open": "multi-GPU", "b2a_ipc_close": "multi-GPU", "b2a_memcpy_d2h": "multi-GPU",
    "b2a_version": "diagnostics", "b2a_reflect_pad_index": "index rule used by tests", "b2a_istft_out_length": "shape rule, inlined",
    "b2a_s3tokenizer_plan_segments": "S3Tokenizer windows: INTEGRATION.md section 4", "b2a_s3tokenizer_gather_segments": "see plan_segments",
    "b2a_debug_mel_program_apply": "test hook", "b2a_debug_plan_layout": "test hook", "b2a_debug_mel_program_dump": "build-time generator hook",
    "b2a_debug_whisper_tc": "experimental switch", "b2a_debug_wpf1920": "A/B switch", "b2a_debug_dyn_tiles": "A/B switch", "b2a_debug_wpf_mel_apply": "host test hook", "b2a_debug_tc_power_buffer": "bring-up hook", "b2a_ctx_enable_timing": "diagnostics",
    "b2a_ctx_last_kernel_ms": "diagnostics", "b2a_resample_poly_filter": "host table used by tests",
    "b2a_log_mel_spectrogram_chatterbox_ragged": "ragged: bound like whisperLogMelSpectrogramRagged", "b2a_funasr_log_mel_spectrogram_ragged": "ragged",
    "b2a_voice_encoder_melspectrogram_ragged": "ragged", "b2a_funasr_preprocess_audio_ragged": "ragged", "b2a_kaldi_fbank_campplus_ragged": "ragged",
    "b2a_s3gen_mel_spectrogram_ragged": "ragged", "b2a_whisper_log_mel_spectrogram_f16_ragged": "ragged",
}


C_TYPES = {"b2a_voice_enc_config", "b2a_ctx"}


def c_prototypes(text):
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    protos = {}
    for m in re.finditer(r"B2A_API\s+[\w\s\*]+?\b(b2a_\w+)\s*\(([^;]*?)\)\s*;", text, flags=re.S):
        name, params = m.group(1), m.group(2).strip()
        n = 0 if params in ("", "void") else len(split_top(params))
        protos[name] = n
    return protos


def split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "([{":
            depth += 1
        elif ch in ")]}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur)
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur)
    return out


def swift_calls(text):
    text = re.sub(r"//[^\n]*", "", text)
    calls = []
    for m in re.finditer(r"\b(b2a_\w+)\s*\(", text):
        if m.group(1) in C_TYPES:      # `b2a_voice_enc_config()` is the struct's initialiser, not a call into the library
            continue
        i, depth = m.end(), 1
        while depth and i < len(text):
            depth += text[i] in "([{"
            depth -= text[i] in ")]}"
            i += 1
        args = text[m.end(): i - 1]
        calls.append((m.group(1), 0 if not args.strip() else len(split_top(args)), text.count("\n", 0, m.start()) + 1))
    return calls


def main():
    protos = c_prototypes(open(HEADER).read())
    shim = open(SHIM).read()
    errors = []
    bound = set()
    for name, n, line in swift_calls(shim):
        if name not in protos:
            errors.append(f"{SHIM}:{line}: {name} is not declared in b200audio.h")
            continue
        bound.add(name)
        if n != protos[name]:
            errors.append(f"{SHIM}:{line}: {name} called with {n} arguments, the C prototype has {protos[name]}")
    funcs = set(re.findall(r"\bfunc\s+(\w+)", shim))
    for name in REQUIRED_SWIFT:
        if name not in funcs:
            errors.append(f"shim has no binding named {name} (SURVEY.md section 8b)")
    for name in sorted(protos):
        if name not in bound and name not in NOT_BOUND:
            errors.append(f"{name} is declared in b200audio.h but neither bound by the shim nor listed in NOT_BOUND")
    for name in NOT_BOUND:
        if name not in protos:
            errors.append(f"NOT_BOUND lists {name}, which b200audio.h does not declare")
    print(f"{len(protos)} C entry points, {len(bound)} bound by the shim, {len(NOT_BOUND)} deliberately unbound, {len(funcs)} Swift functions")
    for e in errors:
        print("ERROR:", e)
    return 1 if errors else 0


if __name__ == "__main__":
    sys.exit(main())
