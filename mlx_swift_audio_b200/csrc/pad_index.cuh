// Padding index map shared by the front-end kernels (frontend.cu, wpf1920.cu): the one statement of the padding rules.
#pragma once
#include "internal.h"

#ifndef B2A_DEV
#define B2A_DEV __device__ __forceinline__
#endif

namespace b2a {

// Source index of padded coordinate p of one clip, or -1 where the padding is zero (the one statement of the padding rules:
// reflectPad with the reference's repeated same-direction reflection for clips shorter than the pad, S3TokenizerUtils.swift:266-298,
// and MLX.padded's zeros).  The modulo only runs when the overshoot exceeds one reflection.
B2A_DEV long long padded_index(long long p, long long pad_left, long long n_eff, int pad_mode) {
  long long j = p - pad_left;
  if (j < 0 || j >= n_eff) {
    if (pad_mode != PAD_REFLECT) return -1;
    if (n_eff == 1) return 0;
    long long t = j < 0 ? -j - 1 : j - n_eff;
    if (t >= n_eff - 1) t %= n_eff - 1;
    j = j < 0 ? t + 1 : n_eff - 2 - t;
  }
  return j;
}

// value of padded coordinate p of one clip (samples in [n_samples, n_eff) are the caller's zero tail)
B2A_DEV float fetch_padded(const float* __restrict__ xc, long long p, long long pad_left, long long n_samples,
                           long long n_eff, int pad_mode) {
  const long long j = padded_index(p, pad_left, n_eff, pad_mode);
  return j >= 0 && j < n_samples ? __ldg(xc + j) : 0.0f;
}

}  // namespace b2a
