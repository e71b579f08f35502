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
