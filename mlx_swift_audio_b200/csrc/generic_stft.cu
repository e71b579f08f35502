// Generic STFT for any (n_fft, hop, window) -- the sizes without a tuned FFT plan (frontend.cu builds 400 / 512 / 1920; the
// vocoder kernels 16 / 20) -- and the matching spectrum -> (|X| or |X|^2) -> sparse mel -> log kernel, so that stft()
// (Codec/S3Tokenizer/S3TokenizerUtils.swift:224-263), voiceEncoderMelspectrogram (VoiceEncoderMelspec.swift:17-68) and
// s3genMelSpectrogram (S3GenMel.swift:43-102) accept the configurations the reference accepts (any nFft / hopLength), not only
// the ones its shipped models use.  Correctness path, not a tuned one: a direct DFT from a shared-memory twiddle table,
// four frames per CTA, thread == bin.
//
//   * frames: reflect (center) padding through the reference's index map (incl. its repeated reflection for short clips),
//     window folded into the staging copy, [n][4 frames] float4 layout so that one broadcast LDS.128 feeds four frames;
//   * X[k] = sum_n xw[n] e^{-2 pi i k n / N}: twiddle index (k n) mod N advanced incrementally, table entries computed on the host
//     in double precision; fp32 accumulation (error ~ sqrt(N) eps, the FFT's order of magnitude for the sizes in question).
#include <cuda_runtime.h>

#include <cmath>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200audio.h"
#include "internal.h"

namespace b2a {

namespace {

constexpr int kFB = 4;         // frames per CTA
constexpr int kThreads = 256;

__device__ __forceinline__ long long reflect_or_zero(long long j, long long n, int pad_mode) {
  if (j < 0 || j >= n) {
    if (pad_mode != PAD_REFLECT) return -1;
    if (n == 1) return 0;
    long long t = j < 0 ? -j - 1 : j - n;
    if (t >= n - 1) t %= n - 1;
    j = j < 0 ? t + 1 : n - 2 - t;
  }
  return j;
}

__global__ void __launch_bounds__(kThreads) generic_stft_kernel(const float* __restrict__ x, float2* __restrict__ out, const float* __restrict__ window,
                                                                const float2* __restrict__ tw, long long n_samples, long long n_frames, int n_fft,
                                                                int hop, long long pad_left, int pad_mode) {
  extern __shared__ __align__(16) float smem[];
  float4* s_x = reinterpret_cast<float4*>(smem);                  // [n_fft] x 4 frames, windowed
  float2* s_tw = reinterpret_cast<float2*>(smem + 4 * n_fft);     // [n_fft] (cos, -sin)(2 pi m / N)
  const long long clip = blockIdx.y;
  const long long f0 = (long long)blockIdx.x * kFB;
  const float* __restrict__ xc = x + clip * n_samples;
  for (int i = threadIdx.x; i < n_fft; i += kThreads) {
    s_tw[i] = __ldg(tw + i);
    const float w = __ldg(window + i);
    float v[kFB];
#pragma unroll
    for (int j = 0; j < kFB; ++j) {
      const long long src = reflect_or_zero((f0 + j) * hop + i - pad_left, n_samples, pad_mode);
      v[j] = (f0 + j < n_frames && src >= 0) ? __ldg(xc + src) * w : 0.0f;
    }
    s_x[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
  __syncthreads();
  const int n_bins = n_fft / 2 + 1;
  for (int k = threadIdx.x; k < n_bins; k += kThreads) {
    float re[kFB] = {0.0f, 0.0f, 0.0f, 0.0f}, im[kFB] = {0.0f, 0.0f, 0.0f, 0.0f};
    int idx = 0;
    for (int n = 0; n < n_fft; ++n) {
      const float2 t = s_tw[idx];
      const float4 v = s_x[n];
      re[0] = fmaf(v.x, t.x, re[0]); im[0] = fmaf(v.x, t.y, im[0]);
      re[1] = fmaf(v.y, t.x, re[1]); im[1] = fmaf(v.y, t.y, im[1]);
      re[2] = fmaf(v.z, t.x, re[2]); im[2] = fmaf(v.z, t.y, im[2]);
      re[3] = fmaf(v.w, t.x, re[3]); im[3] = fmaf(v.w, t.y, im[3]);
      idx += k;
      if (idx >= n_fft) idx -= n_fft;
    }
#pragma unroll
    for (int j = 0; j < kFB; ++j)
      if (f0 + j < n_frames) out[(clip * n_frames + f0 + j) * n_bins + k] = make_float2(re[j], im[j]);
  }
}

// spectrum (batch, T', F) complex -> sparse mel of |X| / |X|^2 -> optional log -> optional affine -> (T', M) or (M, T').
// One warp per (frame, 32 filters); lanes over filters.
__global__ void __launch_bounds__(256) generic_mel_kernel(const float2* __restrict__ spec, float* __restrict__ out, long long n_frames, int n_bins,
                                                          const int4* __restrict__ fb_desc, const float* __restrict__ fb_w, int n_mels, int spec_mode,
                                                          int log_mode, float log_floor, int post_affine, float post_sub, float post_div, int out_mode) {
  const long long clip = blockIdx.y;
  const long long f = (long long)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (f >= n_frames) return;
  const float2* __restrict__ sp = spec + (clip * n_frames + f) * n_bins;
  for (int m = threadIdx.x & 31; m < n_mels; m += 32) {
    const int4 d = __ldg(fb_desc + m);
    float v = 0.0f;
    for (int i = 0; i < d.y; ++i) {
      const float2 z = __ldg(sp + d.x + i);
      const float pw = z.x * z.x + z.y * z.y;
      v = fmaf(__ldg(fb_w + d.z + i), spec_mode == SPEC_POWER ? pw : sqrtf(pw), v);
    }
    if (log_mode == LOG_LN) v = logf(fmaxf(v, log_floor));
    else if (log_mode == LOG_LOG10) v = log10f(fmaxf(v, log_floor));
    else if (log_mode == LOG_DB20) v = 20.0f * log10f(fmaxf(v, log_floor));
    if (post_affine) v = (v - post_sub) / post_div;
    if (out_mode == OUT_MT) out[(clip * n_mels + m) * n_frames + f] = v;
    else out[(clip * n_frames + f) * n_mels + m] = v;
  }
}

std::mutex g_mu;
std::map<std::pair<int, int>, float2*> g_tw;   // (device, n_fft) -> twiddle table

int twiddles(int dev, int n_fft, const float2** out, std::string* err) {
  std::lock_guard<std::mutex> lk(g_mu);
  auto it = g_tw.find({dev, n_fft});
  if (it == g_tw.end()) {
    std::vector<float2> h(static_cast<size_t>(n_fft));
    for (int m = 0; m < n_fft; ++m) {
      const double a = -2.0 * M_PI * double(m) / double(n_fft);
      h[size_t(m)] = make_float2(float(cos(a)), float(sin(a)));
    }
    float2* d = nullptr;
    cudaError_t e = cudaMalloc(&d, sizeof(float2) * h.size());
    if (e == cudaSuccess) e = cudaMemcpy(d, h.data(), sizeof(float2) * h.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) {
      if (d) cudaFree(d);
      if (err) *err = std::string("twiddle table: ") + cudaGetErrorString(e);
      return B2A_E_CUDA;
    }
    it = g_tw.emplace(std::make_pair(dev, n_fft), d).first;
  }
  *out = it->second;
  return B2A_OK;
}

}  // namespace

bool generic_stft_supported(int n_fft, int hop) { return n_fft >= 2 && n_fft <= 8192 && hop >= 1; }

// window_dev: device, n_fft floats (zero-extended).  out: (batch, n_frames, n_fft / 2 + 1) complex64.
int launch_generic_stft(const float* x, int64_t batch, int64_t n_samples, int64_t n_frames, int n_fft, int hop, int64_t pad_left, int pad_mode,
                        const float* window_dev, float* out_complex, void* stream, int* launches, std::string* err) {
  if (!generic_stft_supported(n_fft, hop) || batch <= 0 || batch > 65535 || n_frames <= 0) {
    if (err) *err = "generic stft: n_fft must be in [2, 8192], hop >= 1, batch <= 65535 per launch";
    return B2A_E_UNSUPPORTED;
  }
  int dev = 0;
  cudaGetDevice(&dev);
  const float2* tw = nullptr;
  int rc = twiddles(dev, n_fft, &tw, err);
  if (rc != B2A_OK) return rc;
  const size_t smem = size_t(n_fft) * (sizeof(float4) + sizeof(float2));
  cudaError_t e;
  if (smem > 48 * 1024 && (e = cudaFuncSetAttribute(generic_stft_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, int(smem))) != cudaSuccess) {
    if (err) *err = std::string("cudaFuncSetAttribute: ") + cudaGetErrorString(e);
    return B2A_E_CUDA;
  }
  dim3 grid(unsigned((n_frames + kFB - 1) / kFB), unsigned(batch));
  generic_stft_kernel<<<grid, kThreads, smem, static_cast<cudaStream_t>(stream)>>>(x, reinterpret_cast<float2*>(out_complex), window_dev, tw, n_samples,
                                                                                   n_frames, n_fft, hop, pad_left, pad_mode);
  if ((e = cudaGetLastError()) != cudaSuccess) {
    if (err) *err = std::string("generic_stft_kernel launch: ") + cudaGetErrorString(e);
    return B2A_E_CUDA;
  }
  *launches += 1;
  return B2A_OK;
}

int launch_generic_mel(const float* spec_complex, float* out, int64_t batch, int64_t n_frames, int n_bins, const DeviceBank& bank, int spec_mode,
                       int log_mode, float log_floor, int post_affine, float post_sub, float post_div, int out_mode, void* stream, int* launches,
                       std::string* err) {
  if (batch <= 0 || batch > 65535 || (out_mode != OUT_TM && out_mode != OUT_MT)) {
    if (err) *err = "generic mel: bad launch";
    return B2A_E_BAD_ARG;
  }
  dim3 grid(unsigned((n_frames + 7) / 8), unsigned(batch));
  generic_mel_kernel<<<grid, 256, 0, static_cast<cudaStream_t>(stream)>>>(reinterpret_cast<const float2*>(spec_complex), out, n_frames, n_bins,
                                                                          reinterpret_cast<const int4*>(bank.desc), bank.weights, bank.n_mels, spec_mode,
                                                                          log_mode, log_floor, post_affine, post_sub, post_div, out_mode);
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    if (err) *err = std::string("generic_mel_kernel launch: ") + cudaGetErrorString(e);
    return B2A_E_CUDA;
  }
  *launches += 1;
  return B2A_OK;
}

}  // namespace b2a
