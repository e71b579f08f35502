"""CPU: every generated FFT codelet (the same straight-line code nvcc compiles) against numpy.fft under g++."""
import ctypes
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HDR = os.path.join(ROOT, "mlx_swift_audio_b200", "csrc", "codelets.h")


@pytest.fixture(scope="module")
def lib():
    names = re.findall(r"B2A_CODELET void (b2a_(rdftodd|rdft|cdft|c2r)(\d+))\(", open(HDR).read())
    src = ['#include "codelets.h"', 'extern "C" {']
    for full, kind, n in names:
        n = int(n)
        if kind == "rdft":
            h = n // 2 + 1
            src.append(f"void t_{full}(const float* x, float* yr, float* yi){{ float a[{n}], br[{h}], bi[{h}]; for(int i=0;i<{n};i++)a[i]=x[i]; {full}(a,br,bi); for(int i=0;i<{h};i++){{yr[i]=br[i];yi[i]=bi[i];}} }}")
        elif kind == "rdftodd":
            h = (n - 1) // 2 + 1
            src.append(f"void t_{full}(const float* x, float* yr, float* yi){{ float a[{n}], br[{h}], bi[{h}]; for(int i=0;i<{n};i++)a[i]=x[i]; {full}(a,br,bi); for(int i=0;i<{h};i++){{yr[i]=br[i];yi[i]=bi[i];}} }}")
        elif kind == "cdft":
            src.append(f"void t_{full}(const float* xr, const float* xi, float* yr, float* yi){{ float a[{n}],b[{n}],c[{n}],d[{n}]; for(int i=0;i<{n};i++){{a[i]=xr[i];b[i]=xi[i];}} {full}(a,b,c,d); for(int i=0;i<{n};i++){{yr[i]=c[i];yi[i]=d[i];}} }}")
        else:
            h = n // 2 + 1
            src.append(f"void t_{full}(const float* xr, const float* xi, float* y){{ float a[{h}],b[{h}],c[{n}]; for(int i=0;i<{h};i++){{a[i]=xr[i];b[i]=xi[i];}} {full}(a,b,c); for(int i=0;i<{n};i++)y[i]=c[i]; }}")
    src.append("}")
    d = tempfile.mkdtemp()
    cpp, so = os.path.join(d, "t.cpp"), os.path.join(d, "t.so")
    open(cpp, "w").write("\n".join(src))
    subprocess.run(["g++", "-O1", "-shared", "-fPIC", "-ffp-contract=off", "-I", os.path.dirname(HDR), cpp, "-o", so], check=True)
    return ctypes.CDLL(so), names


def test_generated_header_is_up_to_date():
    out = subprocess.run([sys.executable, os.path.join(ROOT, "tools", "gen_codelets.py")], check=True, capture_output=True, text=True).stdout
    assert out == open(HDR).read(), "run: python tools/gen_codelets.py > mlx_swift_audio_b200/csrc/codelets.h"


def test_codelets_match_numpy_fft(lib):
    L, names = lib
    P = ctypes.POINTER(ctypes.c_float)
    p = lambda a: a.ctypes.data_as(P)  # noqa: E731
    rng = np.random.default_rng(0)
    assert len(names) >= 16
    for full, kind, n in names:
        n = int(n)
        fn = getattr(L, "t_" + full)
        for _ in range(3):
            if kind == "rdft":
                x = rng.standard_normal(n).astype(np.float32)
                yr = np.zeros(n // 2 + 1, np.float32); yi = np.zeros_like(yr)
                fn(p(x), p(yr), p(yi))
                got, ref = yr + 1j * yi, np.fft.rfft(x.astype(np.float64))
            elif kind == "rdftodd":
                h = (n - 1) // 2 + 1
                x = rng.standard_normal(n).astype(np.float32)
                yr = np.zeros(h, np.float32); yi = np.zeros_like(yr)
                fn(p(x), p(yr), p(yi))
                j = np.arange(n)
                got, ref = yr + 1j * yi, np.array([np.sum(x * np.exp(-2j * np.pi * j * (k + 0.5) / n)) for k in range(h)])
            elif kind == "cdft":
                xr = rng.standard_normal(n).astype(np.float32); xi = rng.standard_normal(n).astype(np.float32)
                yr = np.zeros(n, np.float32); yi = np.zeros_like(yr)
                fn(p(xr), p(xi), p(yr), p(yi))
                got, ref = yr + 1j * yi, np.fft.fft(xr.astype(np.float64) + 1j * xi)
            else:
                h = n // 2 + 1
                xr = rng.standard_normal(h).astype(np.float32); xi = rng.standard_normal(h).astype(np.float32)
                y = np.zeros(n, np.float32)
                fn(p(xr), p(xi), p(y))
                got, ref = y, np.fft.irfft(xr.astype(np.float64) + 1j * xi, n) * n
            assert np.abs(got - ref).max() <= 4e-6 * max(1.0, np.abs(ref).max()), (full, np.abs(got - ref).max())
