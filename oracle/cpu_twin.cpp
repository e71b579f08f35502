// C++ (std::thread) twin of the oracle for the two headline workloads (TEST INFRASTRUCTURE ONLY -- see oracle/__init__.py).
//
// BASELINE.md section 4 plans two CPU baselines: (A) the NumPy/SciPy restatement (oracle/reference_dsp.py, which is also the
// parity oracle) and (B) a compiled twin of the same restatement, parallel over clips on all host cores.  This file is (B) for
//   * whisperLogMelSpectrogram (STT/Whisper/WhisperAudio.swift:78-137): reflect pad 200, frames of 400 at hop 160, symmetric
//     Hann, rfft, drop the last frame, |X|^2, Slaney filterbank, log10(max(., 1e-10)), max(., global max - 8), (. + 4) / 4;
//   * istftHiFiGAN (Codec/S3Gen/HiFiGAN.swift:298-367): min(mag, 100), polar -> rectangular, irfft(16), window, overlap-add,
//     division by max(sum w^2, 1e-8), trim 8 samples on each side.
// The window and the filterbank are passed in by the caller (tests and bench.py take them from the NumPy oracle), so the twin
// restates the per-clip arithmetic only.  The 400-point real FFT is the same two-stage 20 x 20 decomposition as the CUDA
// kernel, built from the generated codelets (csrc/codelets.h compiles as plain C++); tests/test_cpu_twin.py checks the twin
// against the NumPy oracle.  Only bench.py's cpu_baseline / --impl reference legs and tests/ load the library built from this.
#include <algorithm>
#include <atomic>
#include <thread>
#include <cmath>
#include <cstdint>
#include <vector>

#include "../mlx_swift_audio_b200/csrc/codelets.h"

namespace {

constexpr int N = 400, N1 = 20, N2 = 20, H1 = N1 / 2, HOP = 160, NB = N / 2 + 1;

struct Twiddles {
  float c[N2][H1 + 1], s[N2][H1 + 1];
  Twiddles() {
    for (int n2 = 0; n2 < N2; ++n2)
      for (int k1 = 0; k1 <= H1; ++k1) {
        const double a = -2.0 * M_PI * double(n2 * k1) / double(N);
        c[n2][k1] = float(std::cos(a));
        s[n2][k1] = float(std::sin(a));
      }
  }
};
const Twiddles kTw;

// clips are independent: a shared counter hands them out to n_threads workers (BASELINE.md's "OpenMP twin"; plain std::thread
// so that the build needs no libgomp)
template <class F>
void parallel_clips(int64_t batch, int n_threads, F&& body) {
  std::atomic<int64_t> next{0};
  auto worker = [&] { for (int64_t b; (b = next.fetch_add(1)) < batch;) body(b); };
  std::vector<std::thread> pool;
  for (int t = 1; t < std::max(1, n_threads); ++t) pool.emplace_back(worker);
  worker();
  for (auto& th : pool) th.join();
}

// power spectrum of one windowed frame (400 samples) -> 201 bins
void power_spectrum_400(const float* fr, float* pw) {
  float yr[H1 + 1][N2], yi[H1 + 1][N2];   // [k1][n2]
  for (int n2 = 0; n2 < N2; ++n2) {
    float in[N1], r[H1 + 1], i[H1 + 1];
    for (int n1 = 0; n1 < N1; ++n1) in[n1] = fr[N2 * n1 + n2];
    b2a_rdft20(in, r, i);
    for (int k1 = 0; k1 <= H1; ++k1) {   // inter-stage twiddle W_N^(n2 k1)
      const float c = kTw.c[n2][k1], s = kTw.s[n2][k1];
      yr[k1][n2] = r[k1] * c - i[k1] * s;
      yi[k1][n2] = r[k1] * s + i[k1] * c;
    }
  }
  for (int k1 = 0; k1 <= H1; ++k1) {
    float ur[N2], ui[N2];
    b2a_cdft20(yr[k1], yi[k1], ur, ui);
    for (int k2 = 0; k2 < N2; ++k2) {
      const int k = k1 + N1 * k2;
      const float p = ur[k2] * ur[k2] + ui[k2] * ui[k2];
      if (k <= N / 2) pw[k] = p;
      else if (N - k <= N / 2 && k1 != 0 && k1 != H1) pw[N - k] = p;   // conjugate mirror: same power
    }
  }
}

}  // namespace

extern "C" {

// audio (batch, n) -> out (batch, n / 160, n_mels); window (400), filters (n_mels, 201)
int twin_whisper_log_mel(const float* audio, int64_t batch, int64_t n, int n_mels, const float* window, const float* filters, float* out,
                         int n_threads) {
  if (n < 201) return -1;
  const int64_t frames = n / HOP;   // 1 + n / 160 frames of the centred STFT, the last one dropped
  // sparse form of the triangular bank: first / last non-zero bin per filter
  std::vector<int> lo(n_mels, 0), hi(n_mels, -1);
  for (int m = 0; m < n_mels; ++m)
    for (int k = 0; k < NB; ++k)
      if (filters[size_t(m) * NB + k] != 0.0f) {
        if (hi[m] < 0) lo[m] = k;
        hi[m] = k;
      }
  parallel_clips(batch, n_threads, [&](int64_t b) {
    const float* x = audio + b * n;
    float* o = out + b * frames * n_mels;
    float gmax = -3.0e38f;
    std::vector<float> fr(N), pw(NB);
    for (int64_t f = 0; f < frames; ++f) {
      for (int t = 0; t < N; ++t) {
        int64_t j = f * HOP + t - 200;   // reflect pad by 200 (numpy "reflect")
        if (j < 0) j = -j;
        else if (j >= n) j = 2 * (n - 1) - j;
        fr[t] = x[j] * window[t];
      }
      power_spectrum_400(fr.data(), pw.data());
      for (int m = 0; m < n_mels; ++m) {
        float acc = 0.0f;
        const float* w = filters + size_t(m) * NB;
        for (int k = lo[m]; k <= hi[m]; ++k) acc += w[k] * pw[k];
        const float l = std::log10(std::max(acc, 1e-10f));
        o[f * n_mels + m] = l;
        gmax = std::max(gmax, l);
      }
    }
    const float floor_ = gmax - 8.0f;
    for (int64_t i = 0; i < frames * n_mels; ++i) o[i] = (std::max(o[i], floor_) + 4.0f) / 4.0f;
  });
  return 0;
}

// mag, phase (batch, 9, frames) -> out (batch, (frames - 1) * 4); window (16)
int twin_istft_hifigan(const float* mag, const float* phase, int64_t batch, int64_t frames, const float* window, float* out, int n_threads) {
  constexpr int NF = 16, F = 9, H = 4, P = NF / 2;
  if (frames < 2) return -1;
  const int64_t full = (frames - 1) * H + NF, out_len = (frames - 1) * H;
  parallel_clips(batch, n_threads, [&](int64_t b) {
    const float* mp = mag + b * F * frames;
    const float* pp = phase + b * F * frames;
    std::vector<float> y(full, 0.0f), env(full, 0.0f);
    for (int64_t f = 0; f < frames; ++f) {
      float xr[F], xi[F], t[NF];
      for (int k = 0; k < F; ++k) {
        const float m = std::min(mp[k * frames + f], 100.0f), p = pp[k * frames + f];
        xr[k] = m * std::cos(p);
        xi[k] = m * std::sin(p);
      }
      b2a_c2r16(xr, xi, t);   // unnormalised inverse real DFT
      for (int i = 0; i < NF; ++i) {
        y[f * H + i] += t[i] * (1.0f / NF) * window[i];
        env[f * H + i] += window[i] * window[i];
      }
    }
    float* o = out + b * out_len;
    for (int64_t i = 0; i < out_len; ++i) o[i] = y[i + P] / std::max(env[i + P], 1e-8f);
  });
  return 0;
}

}  // extern "C"
