"""CPU: the C++ mirror header compiles as C++17 and links against the shared library (no compute call)."""
import os
import subprocess
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_cpp_mirror_compiles_and_links(built_lib):
    src = r'''
#include "b200audio.hpp"
#include <cstdio>
int main() {
  auto w = b2a::hannWindowPeriodic(16);
  auto f = b2a::melFilters(16000, 400, 80, 0.0f, 8000.0f);
  std::printf("%s %.3f %zu\n", b2a_version(), w[4], f.size());
  // (taking the addresses instantiates the ragged-batch and peer-memory wrappers without needing a GPU)
  auto pr = &b2a::Context::whisperLogMelSpectrogramRagged;
  auto pf = &b2a::Context::preprocessAudioRagged;
  auto pi = &b2a::Context::ipcOpen;
  if (!pr || !pf || !pi) return 2;
  try { b2a::Context c(0); } catch (const b2a::Error& e) { std::printf("no gpu: %s\n", e.what()); }
  return (w[4] > 0.49f && w[4] < 0.51f && f.size() == 80u * 201u) ? 0 : 1;
}
'''
    d = tempfile.mkdtemp()
    cpp, exe = os.path.join(d, "t.cpp"), os.path.join(d, "t")
    open(cpp, "w").write(src)
    libdir = os.path.join(ROOT, "mlx_swift_audio_b200")
    subprocess.run(["g++", "-std=c++17", "-I", os.path.join(ROOT, "include"), cpp, "-o", exe, "-L", libdir, "-l:libb200audio.so",
                    "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
    assert "b200audio" in r.stdout


def test_c_header_is_plain_c99(built_lib):
    """include/b200audio.h is the drop-in boundary: it has to compile as C (no C++-isms), and a C program has to link
    against the shared library and call the host-side functions."""
    src = r'''
#include "b200audio.h"
#include <stdio.h>
int main(void) {
  float w[16];
  unsigned char handle[B2A_IPC_HANDLE_BYTES];
  b2a_ctx* ctx = 0;
  int64_t lengths[2] = {4000, 3000}, rows[2];
  if (b2a_window(B2A_WIN_HANN_PERIODIC, 16, w) != B2A_OK) return 1;
  if (b2a_whisper_num_frames(480000, 0) != 3000) return 2;
  if (b2a_reflect_pad_index(0, 10, 3) != 3) return 3;
  /* no context: every compute / peer-memory entry point must refuse, not crash */
  if (b2a_whisper_log_mel_spectrogram_ragged(ctx, w, 2, 4000, lengths, 80, 0, w, rows, B2A_HOST) == B2A_OK) return 4;
  if (b2a_ipc_export(ctx, w, handle) == B2A_OK) return 5;
  printf("%s %.3f\n", b2a_version(), w[4]);
  return 0;
}
'''
    d = tempfile.mkdtemp()
    c, exe = os.path.join(d, "t.c"), os.path.join(d, "t")
    open(c, "w").write(src)
    libdir = os.path.join(ROOT, "mlx_swift_audio_b200")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-pedantic", "-I", os.path.join(ROOT, "include"), c, "-o", exe,
                    "-L", libdir, "-l:libb200audio.so", "-Wl,-rpath," + libdir], check=True)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode == 0, (r.returncode, r.stdout + r.stderr)
    assert "b200audio" in r.stdout


def test_swift_shim_matches_the_c_header():
    """No Swift toolchain here: tools/check_swift_shim.py checks every b2a_* call of the shim against the C prototypes (name, argument
    count), that every helper of SURVEY.md section 8b is bound under the reference's name, and that every entry point of the header
    is either bound or listed as deliberately unbound."""
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, os.path.join(root, "tools", "check_swift_shim.py")], capture_output=True, text=True)
    assert r.returncode == 0, r.stdout + r.stderr
