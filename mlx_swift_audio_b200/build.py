"""Builds libb200audio.so (CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

    python -m mlx_swift_audio_b200.build [--force]

The shared library is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libb200audio.so")
OBJ_DIR = os.path.join(HERE, "_build")

SOURCES = ["host_tables.cpp", "frontend.cu", "wpf1920.cu", "vocoder.cu", "tc_frontend.cu", "generic_stft.cu", "capi.cu"]
HEADERS = ["codelets.h", "mel_baked.h", "internal.h", "pad_index.cuh", os.path.join("..", "..", "include", "b200audio.h")]

NVCC_FLAGS = [
    "-O3", "-std=c++17", "-lineinfo",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA extension cannot be built (there is no CPU fallback)")


def _digest() -> str:
    h = hashlib.sha256()
    for f in SOURCES + HEADERS:
        with open(os.path.join(CSRC, f), "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def regenerate_codelets() -> None:
    gen = os.path.join(ROOT, "tools", "gen_codelets.py")
    out = subprocess.run([sys.executable, gen], check=True, capture_output=True, text=True).stdout
    path = os.path.join(CSRC, "codelets.h")
    if not os.path.exists(path) or open(path).read() != out:
        with open(path, "w") as fh:
            fh.write(out)


def regenerate_mel_baked() -> None:
    """csrc/mel_baked.h: straight-line mel projections for the standard banks, generated from the library's own
    host-side filterbank code (host_tables.cpp compiled stand-alone with g++; see tools/gen_mel_baked.py)."""
    os.makedirs(OBJ_DIR, exist_ok=True)
    gxx = shutil.which("g++") or shutil.which("c++")
    if gxx is None:
        raise RuntimeError("g++ not found: cannot generate csrc/mel_baked.h")
    host_lib = os.path.join(OBJ_DIR, "libhosttables.so")
    subprocess.run([gxx, "-O2", "-std=c++17", "-shared", "-fPIC", "-o", host_lib, os.path.join(CSRC, "host_tables.cpp")], check=True)
    gen = os.path.join(ROOT, "tools", "gen_mel_baked.py")
    out = subprocess.run([sys.executable, gen, host_lib], check=True, capture_output=True, text=True).stdout
    path = os.path.join(CSRC, "mel_baked.h")
    if not os.path.exists(path) or open(path).read() != out:
        with open(path, "w") as fh:
            fh.write(out)


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ_DIR, exist_ok=True)
    if not os.path.exists(os.path.join(CSRC, "mel_baked.h")):
        regenerate_mel_baked()
    stamp = os.path.join(OBJ_DIR, "stamp")
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read() == dig:
        return LIB
    nvcc = _nvcc()

    def compile_one(src: str) -> str:
        obj = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-x", "cu", "-c", os.path.join(CSRC, src), "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        log = os.path.join(OBJ_DIR, os.path.splitext(src)[0] + ".log")
        with open(log, "w") as fh:
            fh.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed for %s:\n%s" % (src, r.stderr[-6000:]))
        if verbose:
            sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=len(SOURCES)) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed:\n" + r.stderr[-4000:])
    with open(stamp, "w") as fh:
        fh.write(dig)
    return LIB


if __name__ == "__main__":
    regenerate_codelets()
    regenerate_mel_baked()
    path = build(force="--force" in sys.argv, verbose="-v" in sys.argv)
    print(path)
