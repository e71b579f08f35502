// Host-side tables and integer rules of the STFT-family DSP path: window generators,
// mel filterbanks, reflect-pad index map, frame-count rules.  Pure C++ (no CUDA), so the
// CPU test-suite can check every one of them against the oracle through the C ABI.
//
// Each function restates (does not copy) the fp32 scalar arithmetic of the reference
// helper it names; paths are relative to the reference's package/ directory.
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <functional>
#include <vector>

#include "../../include/b200audio.h"
#include "internal.h"

namespace b2a {

static const float kPi = 3.14159265358979323846f;  // Swift Float.pi rounds to the same fp32 value

// ---- windows -----------------------------------------------------------------------------
int make_window(int kind, int length, float* out) {
  if (length <= 0 || out == nullptr) return B2A_E_BAD_ARG;
  if (length == 1 && kind != B2A_WIN_POVEY && kind != B2A_WIN_HANN_PERIODIC) {
    out[0] = 1.0f;  // the `length == 1` early return of WhisperAudio.swift:33, S3TokenizerUtils.swift:214, FunASRAudio.swift:36
    return B2A_OK;
  }
  switch (kind) {
    case B2A_WIN_WHISPER_HANN: {  // STT/Whisper/WhisperAudio.swift:32-44
      const float factor = 2.0f * kPi / float(length - 1);
      for (int n = 0; n < length; ++n) out[n] = 0.5f * (1.0f - cosf(float(n) * factor));
      return B2A_OK;
    }
    case B2A_WIN_HANNING: {  // Codec/S3Tokenizer/S3TokenizerUtils.swift:213-221, TTS/Kokoro/Decoder/MLXSTFT.swift:12-20
      const float factor = kPi / float(length - 1);
      for (int i = 0; i < length; ++i) {
        const float n = float(1 - length) + 2.0f * float(i);
        out[i] = 0.5f + 0.5f * cosf(n * factor);
      }
      return B2A_OK;
    }
    case B2A_WIN_HAMMING: {  // STT/FunASR/FunASRAudio.swift:35-45
      const float factor = 2.0f * kPi / float(length - 1);
      for (int n = 0; n < length; ++n) out[n] = 0.54f - 0.46f * cosf(float(n) * factor);
      return B2A_OK;
    }
    case B2A_WIN_POVEY: {  // Codec/S3Gen/CAMPPlus.swift:15-19
      for (int n = 0; n < length; ++n) {
        const float hann = 0.5f - 0.5f * cosf(2.0f * kPi * float(n) / float(length - 1));
        out[n] = powf(hann, 0.85f);
      }
      return B2A_OK;
    }
    case B2A_WIN_HANN_PERIODIC: {  // Codec/S3Gen/HiFiGAN.swift:15-20, CosyVoice3 CausalHiFTGenerator.swift:429-432
      for (int n = 0; n < length; ++n) out[n] = 0.5f * (1.0f - cosf(2.0f * kPi * float(n) / float(length)));
      return B2A_OK;
    }
    default:
      return B2A_E_BAD_ARG;
  }
}

// hanningWindow(length: n + 1)[0 ..< n]  (S3TokenizerUtils.swift:172, S3GenMel.swift:66, MLXSTFT.swift:52)
void hann_periodic_via_hanning(int n, std::vector<float>& w) {
  std::vector<float> full(n + 1);
  make_window(B2A_WIN_HANNING, n + 1, full.data());
  w.assign(full.begin(), full.begin() + n);
}

// ---- Slaney filterbank: S3TokenizerUtils.swift:301-375 -------------------------------------
int mel_filters_slaney(int sample_rate, int n_fft, int n_mels, float f_min, float f_max, float* out) {
  if (sample_rate <= 0 || n_fft <= 0 || n_mels <= 0 || out == nullptr) return B2A_E_BAD_ARG;
  const float actual_fmax = f_max >= 0.0f ? f_max : float(sample_rate) / 2.0f;
  const float f_sp = 200.0f / 3.0f;
  const float min_log_hz = 1000.0f;
  const float min_log_mel = min_log_hz / f_sp;
  const float logstep = logf(6.4f) / 27.0f;
  auto hz_to_mel = [&](float hz) { return hz >= min_log_hz ? min_log_mel + logf(hz / min_log_hz) / logstep : hz / f_sp; };
  auto mel_to_hz = [&](float mel) { return mel >= min_log_mel ? min_log_hz * expf(logstep * (mel - min_log_mel)) : f_sp * mel; };
  const float mel_min = hz_to_mel(f_min), mel_max = hz_to_mel(actual_fmax);
  std::vector<float> pts(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i) pts[i] = mel_to_hz(mel_min + float(i) * (mel_max - mel_min) / float(n_mels + 1));
  const int nb = n_fft / 2 + 1;
  std::vector<float> freqs(nb);
  for (int i = 0; i < nb; ++i) freqs[i] = float(i) * float(sample_rate) / float(n_fft);
  for (int m = 0; m < n_mels; ++m) {
    const float fl = pts[m], fc = pts[m + 1], fr = pts[m + 2];
    const float enorm = 2.0f / (pts[m + 2] - pts[m]);
    for (int k = 0; k < nb; ++k) {
      const float f = freqs[k];
      float v = 0.0f;
      if (f >= fl && f <= fc) v = (f - fl) / (fc - fl);
      else if (f > fc && f <= fr) v = (fr - f) / (fr - fc);
      out[size_t(m) * nb + k] = v * enorm;
    }
  }
  return B2A_OK;
}

// ---- Fun-ASR HTK filterbank on the 200-point linspace grid: FunASRAudio.swift:322-396 -------
static void linspace_f32(float a, float b, int num, std::vector<float>& v) {
  // MLX linspace evaluates (1 - t) * start + t * stop with t = arange(num) / (num - 1) in fp32
  v.resize(num);
  for (int i = 0; i < num; ++i) {
    const float t = float(i) / float(num - 1);
    v[i] = (1.0f - t) * a + t * b;
  }
}

int mel_filters_funasr(int sample_rate, int n_fft, int n_mels, float* out) {
  if (sample_rate <= 0 || n_fft < 4 || n_mels <= 0 || out == nullptr) return B2A_E_BAD_ARG;
  const int n_freqs = n_fft / 2;
  std::vector<float> all_freqs, m_pts;
  linspace_f32(0.0f, float(sample_rate) / 2.0f, n_freqs, all_freqs);
  const float m_min = 2595.0f * log10f(1.0f + 0.0f / 700.0f);
  const float m_max = 2595.0f * log10f(1.0f + (float(sample_rate) / 2.0f) / 700.0f);
  linspace_f32(m_min, m_max, n_mels + 2, m_pts);
  std::vector<float> f_pts(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i) f_pts[i] = 700.0f * (powf(10.0f, m_pts[i] / 2595.0f) - 1.0f);
  for (int m = 0; m < n_mels; ++m) {
    const float enorm = 2.0f / (f_pts[m + 2] - f_pts[m]);
    const float d0 = f_pts[m + 1] - f_pts[m], d1 = f_pts[m + 2] - f_pts[m + 1];
    for (int k = 0; k < n_freqs; ++k) {
      const float down = -(f_pts[m] - all_freqs[k]) / d0;   // rising edge
      const float up = (f_pts[m + 2] - all_freqs[k]) / d1;  // falling edge
      out[size_t(m) * n_freqs + k] = fmaxf(0.0f, fminf(down, up)) * enorm;
    }
  }
  return B2A_OK;
}

// ---- integer-bin HTK triangles: CAMPPlus.swift:134-175 ---------------------------------------
int mel_filters_htk_int(int sample_rate, int n_fft, int n_mels, float f_min, float f_max, float* out) {
  if (sample_rate <= 0 || n_fft <= 0 || n_mels <= 0 || out == nullptr) return B2A_E_BAD_ARG;
  auto hz_to_mel = [](float hz) { return 2595.0f * log10f(1.0f + hz / 700.0f); };
  auto mel_to_hz = [](float mel) { return 700.0f * (powf(10.0f, mel / 2595.0f) - 1.0f); };
  const float mel_min = hz_to_mel(f_min), mel_max = hz_to_mel(f_max);
  std::vector<long> bins(n_mels + 2);
  for (int i = 0; i < n_mels + 2; ++i) {
    const float mel = mel_min + float(i) * (mel_max - mel_min) / float(n_mels + 1);
    bins[i] = lroundf(mel_to_hz(mel) * float(n_fft) / float(sample_rate));  // Swift round(): half away from zero
  }
  const int nb = n_fft / 2 + 1;
  for (size_t i = 0; i < size_t(nb) * n_mels; ++i) out[i] = 0.0f;
  for (int m = 1; m <= n_mels; ++m) {
    const long lo = bins[m - 1], ce = bins[m], hi = bins[m + 1];
    for (long k = lo; k < ce; ++k)
      if (k >= 0 && k < nb && ce != lo) out[size_t(k) * n_mels + (m - 1)] = float(k - lo) / float(ce - lo);
    for (long k = ce; k < hi; ++k)
      if (k >= 0 && k < nb && hi != ce) out[size_t(k) * n_mels + (m - 1)] = float(hi - k) / float(hi - ce);
  }
  return B2A_OK;
}

// ---- reflect index map: S3TokenizerUtils.swift:266-298 == FunASRAudio.swift:280-310 -----------
// The while-loops there prepend / append chunks x[1..a] reversed, so walking outwards from the
// signal edge the source index cycles 1, 2, ..., n-1, 1, 2, ... (left) and n-2, ..., 0, n-2, ... (right).
int64_t reflect_pad_index(int64_t i, int64_t n, int64_t pad) {
  const int64_t j = i - pad;  // position relative to the signal
  if (j >= 0 && j < n) return j;
  if (n == 1) return 0;
  if (j < 0) return ((-j - 1) % (n - 1)) + 1;
  return n - 2 - ((j - n) % (n - 1));
}

// ---- sparse filterbank: each filter = contiguous run of bins with non-zero weight ----------
// bank is (n_mels, n_bins) row-major when !bin_major, else (n_bins, n_mels).
// Mel schedule of the warp-per-frame front end (layout: internal.h)
bool build_wpf_mel(const SparseBank& sb, int n_bins_spectrum, std::vector<uint32_t>& blob) {
  blob.clear();
  const int M = sb.n_mels;
  if (M <= 0 || M > kWpfRounds * 32 || sb.max_bin >= n_bins_spectrum) return false;
  const int span = (n_bins_spectrum + 3) & ~3;   // the kernel may read the spectrum row up to its next 16-byte boundary
  struct Seg { int m, a, b; };   // filter m, bins [a, b)
  std::vector<Seg> segs;
  std::vector<int> per_filter(static_cast<size_t>(M), 1);
  for (int m = 0; m < M; ++m) segs.push_back({m, sb.start[m], sb.start[m] + sb.count[m]});
  const int rounds = (M + 31) / 32, want = rounds * 32;
  while (int(segs.size()) < want) {   // halve the longest segment of a filter that still has fewer than four
    int best = -1;
    for (int i = 0; i < int(segs.size()); ++i)
      if (per_filter[size_t(segs[i].m)] < 4 && segs[i].b - segs[i].a >= 2 && (best < 0 || segs[i].b - segs[i].a > segs[best].b - segs[best].a)) best = i;
    if (best < 0) break;
    const Seg s = segs[size_t(best)];
    const int mid = s.a + (s.b - s.a + 1) / 2;
    segs[size_t(best)] = {s.m, s.a, mid};
    segs.push_back({s.m, mid, s.b});
    per_filter[size_t(s.m)] += 1;
  }
  while (int(segs.size()) < want) segs.push_back({-1, 0, 0});   // idle lanes
  std::stable_sort(segs.begin(), segs.end(), [](const Seg& x, const Seg& y) { return x.b - x.a > y.b - y.a; });
  // Round r = segments [32 r, 32 r + 32).  A lane reads len4[r] groups of four bins from a 16-byte aligned start a0 <= a with
  // a0 + 4 len4[r] >= b (weights zero outside [a, b)).  Within a quarter-warp the eight 16-byte reads of one instruction are
  // conflict-free when (a0 / 4) mod 8 differs from lane to lane: the slack of the shorter segments is used for that, and the
  // segments of a round may sit on any lane.
  int len4[4] = {0, 0, 0, 0}, woff[4] = {0, 0, 0, 0};
  std::vector<int> lane_of(size_t(want), -1), a0_of(size_t(want), 0);
  int conflicts = 0;
  int words = kWpfMelHeader;
  for (int r = 0; r < rounds; ++r) {
    int need = 0;
    for (int i = r * 32; i < r * 32 + 32; ++i) {
      const Seg& s = segs[size_t(i)];
      if (s.b > s.a) need = std::max(need, (s.b - (s.a & ~3) + 3) / 4);
    }
    if (4 * need > span) return false;
    // bipartite matching (augmenting paths) of the round's segments to the 32 (quarter-warp, residue) slots; one more group of slack is
    // tried when the minimal length leaves a segment without a free residue
    int best_unmatched = 33;
    std::vector<int> best_slot;
    int best_need = need;
    for (int extra = 0; extra <= 1 && best_unmatched > 0; ++extra) {
      const int nd = need + extra;
      if (4 * nd > span) break;
      bool ok[32][8];
      for (int j = 0; j < 32; ++j) {
        const Seg& s = segs[size_t(r * 32 + j)];
        int lo = std::max(0, s.b - 4 * nd), hi = std::min(s.a, span - 4 * nd);
        if (s.b <= s.a) { lo = 0; hi = span - 4 * nd; }
        lo = (lo + 3) & ~3;
        hi &= ~3;
        for (int rho = 0; rho < 8; ++rho) ok[j][rho] = false;
        for (int a0 = lo; a0 <= hi; a0 += 4) ok[j][(a0 / 4) & 7] = true;
      }
      std::vector<int> slot_owner(32, -1), slot_of(32, -1);   // slot = quarter * 8 + residue
      std::function<bool(int, std::vector<char>&)> augment = [&](int j, std::vector<char>& seen) -> bool {
        for (int sl = 0; sl < 32; ++sl) {
          if (!ok[j][sl & 7] || seen[size_t(sl)]) continue;
          seen[size_t(sl)] = 1;
          if (slot_owner[size_t(sl)] < 0 || augment(slot_owner[size_t(sl)], seen)) {
            slot_owner[size_t(sl)] = j;
            slot_of[size_t(j)] = sl;
            return true;
          }
        }
        return false;
      };
      int unmatched = 0;
      for (int j = 0; j < 32; ++j) {
        std::vector<char> seen(32, 0);
        if (!augment(j, seen)) unmatched += 1;
      }
      if (unmatched < best_unmatched) {
        best_unmatched = unmatched;
        best_slot = slot_of;
        best_need = nd;
      }
    }
    need = best_need;
    len4[r] = need;
    {
      std::vector<char> taken(32, 0);
      for (int j = 0; j < 32; ++j)
        if (best_slot[size_t(j)] >= 0) taken[size_t(best_slot[size_t(j)])] = 1;
      for (int j = 0; j < 32; ++j) {
        const Seg& s = segs[size_t(r * 32 + j)];
        int lo = std::max(0, s.b - 4 * need), hi = std::min(s.a, span - 4 * need);
        if (s.b <= s.a) { lo = 0; hi = span - 4 * need; }
        lo = (lo + 3) & ~3;
        hi &= ~3;
        int sl = best_slot[size_t(j)], a0 = hi;
        if (sl < 0) {   // no free residue: any free slot, a two-way bank conflict on that quarter-warp's loads
          for (sl = 0; taken[size_t(sl)]; ++sl) {}
          taken[size_t(sl)] = 1;
          if (s.b > s.a) conflicts += 1;
        } else {
          for (a0 = hi; a0 >= lo && ((a0 / 4) & 7) != (sl & 7); a0 -= 4) {}
        }
        lane_of[size_t(r * 32 + j)] = sl;   // lane = quarter * 8 + position; positions within a quarter are interchangeable
        a0_of[size_t(r * 32 + j)] = a0;
      }
    }
    woff[r] = words;
    words += need * 32 * 4;
  }
  const int start_off = words;
  words += rounds * 32;
  words = (words + 3) & ~3;
  const int fin_off = words;
  words += M * 4;
  if (words > kWpfMelMaxWords) return false;
  blob.assign(size_t(words), 0u);
  blob[0] = uint32_t(M);
  blob[1] = uint32_t(rounds);
  for (int r = 0; r < rounds; ++r) {
    blob[size_t(2 + r)] = uint32_t(len4[r]);
    blob[size_t(6 + r)] = uint32_t(woff[r]);
  }
  blob[10] = uint32_t(start_off);
  blob[11] = uint32_t(fin_off);
  blob[12] = uint32_t(words);
  blob[13] = uint32_t(conflicts);
  for (int m = 0; m < M; ++m)
    for (int j = 0; j < 4; ++j) blob[size_t(fin_off + 4 * m + j)] = uint32_t(rounds * 32);   // the zero slot
  std::vector<std::vector<std::pair<int, int>>> of_filter{size_t(M)};   // (first bin, slot)
  for (int i = 0; i < want; ++i) {
    const Seg& s = segs[size_t(i)];
    const int r = i / 32, lane = lane_of[size_t(i)], a0 = a0_of[size_t(i)];
    blob[size_t(start_off + r * 32 + lane)] = uint32_t(a0);
    if (s.m < 0 || s.b <= s.a) continue;
    for (int k = s.a; k < s.b; ++k) {
      const float w = sb.weights[size_t(sb.offset[size_t(s.m)] + (k - sb.start[size_t(s.m)]))];
      const int g = (k - a0) / 4, e = (k - a0) % 4;
      memcpy(&blob[size_t(woff[r] + (g * 32 + lane) * 4 + e)], &w, 4);   // float4 [group][lane]
    }
    of_filter[size_t(s.m)].push_back({s.a, r * 32 + lane});
  }
  for (int m = 0; m < M; ++m) {
    std::sort(of_filter[size_t(m)].begin(), of_filter[size_t(m)].end());
    for (size_t j = 0; j < of_filter[size_t(m)].size(); ++j) blob[size_t(fin_off + 4 * m) + j] = uint32_t(of_filter[size_t(m)][j].second);
  }
  return true;
}

void build_sparse_bank(const float* bank, int n_mels, int n_bins, bool bin_major, SparseBank& sb) {
  sb.n_mels = n_mels;
  sb.n_bins = n_bins;
  sb.start.assign(n_mels, 0);
  sb.count.assign(n_mels, 0);
  sb.offset.assign(n_mels, 0);
  sb.weights.clear();
  sb.max_bin = -1;
  for (int m = 0; m < n_mels; ++m) {
    int first = -1, last = -1;
    for (int k = 0; k < n_bins; ++k) {
      const float v = bin_major ? bank[size_t(k) * n_mels + m] : bank[size_t(m) * n_bins + k];
      if (v != 0.0f) {
        if (first < 0) first = k;
        last = k;
      }
    }
    sb.offset[m] = int(sb.weights.size());
    if (first >= 0) {
      sb.start[m] = first;
      sb.count[m] = last - first + 1;
      for (int k = first; k <= last; ++k)
        sb.weights.push_back(bin_major ? bank[size_t(k) * n_mels + m] : bank[size_t(m) * n_bins + k]);
      if (last > sb.max_bin) sb.max_bin = last;
    }
  }
  // per-bin form: bin k feeds filters m_lo(k) and m_lo(k)+1 only, with m_lo non-decreasing in k.
  // (True for every triangular bank of the reference; checked by reconstructing the dense bank.)
  auto at = [&](int m, int k) -> float {
    if (m < 0 || m >= n_mels) return 0.0f;
    return bin_major ? bank[size_t(k) * n_mels + m] : bank[size_t(m) * n_bins + k];
  };
  sb.two_adjacent = true;
  sb.bin_mlo.assign(n_bins, 0);
  int prev_m = 0;
  for (int k = 0; k < n_bins && sb.two_adjacent; ++k) {
    int first = -1, last = -1;
    for (int m = 0; m < n_mels; ++m)
      if (at(m, k) != 0.0f) {
        if (first < 0) first = m;
        last = m;
      }
    int m_lo = prev_m;
    if (first >= 0) {
      if (last - first > 1) { sb.two_adjacent = false; break; }
      if (last == first && first - 1 >= prev_m) m_lo = first - 1;  // single filter: keep the lower one open as long as possible
      else m_lo = first;
      if (m_lo < prev_m) { sb.two_adjacent = false; break; }
    }
    for (int m = 0; m < n_mels; ++m) {  // reconstruction check
      const float want = at(m, k);
      const float got = m == m_lo ? at(m_lo, k) : (m == m_lo + 1 ? at(m_lo + 1, k) : 0.0f);
      if (want != got) { sb.two_adjacent = false; break; }
    }
    sb.bin_mlo[k] = m_lo;
    prev_m = m_lo;
  }
  if (!sb.two_adjacent) sb.bin_mlo.clear();
}

// Compiles the bank into the frontend kernel's mel "step program" (see mel_steps in frontend.cu) for
// n_chunks warps.  Filter m owns the bins whose lower filter is m (m_lo == m); each of its steps adds
// W[m][k]*P[k] to the open accumulator and W[m+1][k]*P[k] to the next one, and its last step emits.  A chunk
// (the filters of one warp) starts with zero-emit steps for the bins its first filter shares with the
// previous chunk's last filter.  Chunks are cut at filter boundaries, balanced by step count.
void spectrum_slots(const PlanShape& ps, std::vector<int>& slot_of_bin) {
  const int N = ps.n_fft, N1 = ps.n1, N2 = ps.n2, H1 = N1 / 2;
  slot_of_bin.assign(N / 2 + 1, -1);
  for (int k = 0; k <= N / 2; ++k) {
    const int r = k % N1;
    int slot;
    if (r == 0) slot = k / N1;                                   // item 0, real part: bins N1*k2, k2 = 0..N2/2
    else if (r == H1) slot = N2 / 2 + 1 + (k - H1) / N1;         // item 0, odd part: bins H1 + N1*k2
    else if (r < H1) slot = r * 2 * N2 + k / N1;                 // item r, direct output k2 = k / N1
    else slot = (N1 - r) * 2 * N2 + (N - (N1 - r) - k) / N1;      // item N1-r, mirrored output k2: k = N - N1*k2 - (N1-r)
    slot_of_bin[k] = slot;
  }
}

// Staging word of filter m: the finished mel values wait for the store phase in the rows of the exchange buffer that stage B
// leaves unused (rows N2+1 .. 2*N2-1 of every item's block).  Inside a block the filters sit at pitch frame_tile+1 and block q
// is skewed by (C*q mod frame_tile) words, C = filters per block, so that word(m, f) = base(m) + f has bank (m + f) mod 32:
// the frame-major writes of the mel stage and the filter-major reads of the store stage are both conflict free.
// emit_word[m] = byte offset of base(m) in the exchange buffer (never 0).
int output_block_capacity(const PlanShape& ps) {
  const int free_words = (ps.n2 - 1) * ps.frame_tile;
  return (free_words - (ps.frame_tile - 1)) / (ps.frame_tile + 1);
}
bool output_words(const PlanShape& ps, int n_mels, std::vector<int>& emit_word) {
  const int cap = output_block_capacity(ps);
  emit_word.assign(n_mels, 0);
  for (int m = 0; m < n_mels; ++m) {
    const int q = m / cap, j = m % cap;
    if (q >= ps.n1 / 2) return false;
    const int word = (q * 2 * ps.n2 + ps.n2 + 1) * ps.frame_tile + (cap * q) % ps.frame_tile + j * (ps.frame_tile + 1);
    emit_word[m] = word * int(sizeof(float));
  }
  return true;
}

void build_mel_program(const float* bank, int n_mels, int n_bins, bool bin_major, int frame_tile, const int* emit_word, int n_chunks,
                       const int* bin_slot, SparseBank& sb) {
  sb.steps.clear();
  sb.chunk_m.assign(n_chunks + 1, 0);
  sb.chunk_s.assign(n_chunks + 1, 0);
  if (!sb.two_adjacent) return;
  auto at = [&](int m, int k) -> float {
    if (m < 0 || m >= n_mels) return 0.0f;
    return bin_major ? bank[size_t(k) * n_mels + m] : bank[size_t(m) * n_bins + k];
  };
  auto bits = [](int v) {
    float f;
    memcpy(&f, &v, sizeof(float));
    return f;
  };
  std::vector<std::vector<int>> own(n_mels);  // bins whose lower filter is m
  for (int k = 0; k < n_bins; ++k) {
    const int m = sb.bin_mlo[k];
    if (m >= 0 && m < n_mels && (at(m, k) != 0.0f || at(m + 1, k) != 0.0f)) own[m].push_back(k);
  }
  // balance: cost of a filter = 3 instructions per step (spectrum load, two multiply-adds) + 1 for the store of the sum
  auto cost = [&](int m) { return (long long)(3 * own[m].size() + 1); };
  long long total = 0;
  for (int m = 0; m < n_mels; ++m) total += cost(m);
  {
    long long run = 0;
    int w = 1;
    for (int m = 0; m < n_mels && w < n_chunks; ++m) {
      run += cost(m);
      while (w < n_chunks && run * n_chunks >= total * w) sb.chunk_m[w++] = m + 1;
    }
    for (; w <= n_chunks; ++w) sb.chunk_m[w] = n_mels;
    sb.chunk_m[n_chunks] = n_mels;
  }
  auto push = [&](float wlo, float whi, int k, int emit_m) {   // emit_m: filter finished by this step, or -1
    sb.steps.push_back(wlo);
    sb.steps.push_back(whi);
    sb.steps.push_back(bits((bin_slot ? bin_slot[k] : k) * frame_tile * int(sizeof(float))));
    sb.steps.push_back(bits(emit_m >= 0 ? emit_word[emit_m] : 0));
  };
  for (int c = 0; c < n_chunks; ++c) {
    sb.chunk_s[c] = int(sb.steps.size() / 4);
    const int ma = sb.chunk_m[c], mb = sb.chunk_m[c + 1];
    if (ma < mb && ma > 0)
      for (int k : own[ma - 1])
        if (at(ma, k) != 0.0f) push(at(ma, k), 0.0f, k, -1);
    for (int m = ma; m < mb; ++m) {
      if (own[m].empty()) {
        push(0.0f, 0.0f, 0, m);
        continue;
      }
      for (size_t i = 0; i < own[m].size(); ++i) {
        const int k = own[m][i];
        push(at(m, k), at(m + 1, k), k, i + 1 == own[m].size() ? m : -1);
      }
    }
  }
  sb.chunk_s[n_chunks] = int(sb.steps.size() / 4);
}

}  // namespace b2a

// ---- C ABI (host-only part) ------------------------------------------------------------------
extern "C" {

int b2a_window(int kind, int length, float* out) { return b2a::make_window(kind, length, out); }

int b2a_mel_filters(int sample_rate, int n_fft, int n_mels, float f_min, float f_max, float* out) {
  return b2a::mel_filters_slaney(sample_rate, n_fft, n_mels, f_min, f_max, out);
}

int b2a_funasr_mel_filters(int sample_rate, int n_fft, int n_mels, float* out) {
  return b2a::mel_filters_funasr(sample_rate, n_fft, n_mels, out);
}

int b2a_mel_filters_htk(int sample_rate, int n_fft, int n_mels, float f_min, float f_max, float* out) {
  return b2a::mel_filters_htk_int(sample_rate, n_fft, n_mels, f_min, f_max, out);
}

int64_t b2a_reflect_pad_index(int64_t i, int64_t n, int64_t pad) {
  if (n <= 0 || pad < 0 || i < 0 || i >= n + 2 * pad) return -1;
  return b2a::reflect_pad_index(i, n, pad);
}

int b2a_next_power_of_2(int n) {  // CAMPPlus.swift:22-29
  if (n <= 1) return 1;
  int p = 1;
  while (p < n) p *= 2;
  return p;
}

int64_t b2a_funasr_compute_feature_length(int64_t audio_length, int hop_length, int lfr_n) {  // FunASRAudio.swift:225-235
  if (hop_length <= 0 || lfr_n <= 0) return -1;
  return (audio_length / hop_length + lfr_n - 1) / lfr_n;
}

// stft: numFrames = 1 + (T + 2*(nFft/2)*center - nFft) / hop   (S3TokenizerUtils.swift:245-251)
int64_t b2a_stft_num_frames(int64_t n_samples, int n_fft, int hop, int center) {
  if (n_samples <= 0 || n_fft <= 0 || hop <= 0) return -1;
  const int64_t len = n_samples + (center ? 2 * int64_t(n_fft / 2) : 0);
  if (len < n_fft) return -1;
  return 1 + (len - n_fft) / hop;
}

int64_t b2a_whisper_num_frames(int64_t n_samples, int64_t padding) {  // WhisperAudio.swift:91-105
  if (padding < 0) return -1;
  const int64_t f = b2a_stft_num_frames(n_samples + padding, 400, 160, 1);
  return f < 0 ? -1 : f - 1;
}

int64_t b2a_funasr_num_frames(int64_t n_samples) { return b2a_stft_num_frames(n_samples, 400, 160, 1); }

int64_t b2a_lfr_num_rows(int64_t n_frames, int lfr_n) {  // FunASRAudio.swift:117
  if (n_frames <= 0 || lfr_n <= 0) return -1;
  return (n_frames + lfr_n - 1) / lfr_n;
}

int64_t b2a_kaldi_num_frames(int64_t n_samples, int win_length, int hop) {  // CAMPPlus.swift:51-54
  if (n_samples < win_length || win_length <= 0 || hop <= 0) return -1;  // shorter than one window: MLX.take would read out of bounds
  return (n_samples - win_length) / hop + 1;
}

int64_t b2a_s3gen_num_frames(int64_t n_samples, int n_fft, int hop) {  // S3GenMel.swift:10-28,61-77
  if (n_samples <= 0 || n_fft <= 0 || hop <= 0) return -1;
  const int64_t pad = (n_fft - hop) / 2;
  const int64_t eff = pad < n_samples - 1 ? pad : n_samples - 1;  // reflectPad2D truncates the reflection, no loop
  const int64_t len = n_samples + 2 * eff;
  if (len < n_fft) return -1;
  return 1 + (len - n_fft) / hop;
}

int64_t b2a_vocoder_stft_num_frames(int64_t n_samples, int n_fft, int hop) {  // HiFiGAN.swift:262-270
  if (n_samples <= n_fft / 2 || n_fft <= 0 || hop <= 0) return -1;  // needs x[1 ..< pad+1]
  return (n_samples + 2 * int64_t(n_fft / 2) - n_fft) / hop + 1;
}

int64_t b2a_istft_out_length(int64_t n_frames, int hop) {  // HiFiGAN.swift:331,362-364
  if (n_frames <= 0 || hop <= 0) return -1;
  return (n_frames - 1) * hop;
}

void b2a_voice_enc_config_default(b2a_voice_enc_config* c) {  // Config/ChatterboxConfig.swift:139-156
  if (!c) return;
  c->num_mels = 40;
  c->sample_rate = 16000;
  c->n_fft = 400;
  c->hop_size = 160;
  c->win_size = 400;
  c->fmin = 0;
  c->fmax = 8000;
  c->mel_power = 2.0f;
  c->mel_type_db = 0;
  c->normalized_mels = 0;
  c->stft_magnitude_min = 1e-4f;
}

// S3Tokenizer.quantize / quantizeMixedBatch segment plan (Codec/S3Tokenizer/S3Tokenizer.swift:474-571): a clip of at most `window`
// frames is one segment of its own length; a longer clip is cut into windows of `window` frames at stride `stride` (start = 0,
// stride, ... while start < length; the last ones are shorter).  Writes (batch index, start, length) per segment when the output
// pointers are non-null (capacity `cap` segments) and returns the number of segments (or -1 on bad arguments / overflow).
int64_t b2a_s3tokenizer_plan_segments(const int64_t* mel_len, int64_t batch, int64_t window, int64_t stride, int32_t* batch_idx,
                                      int32_t* start, int32_t* length, int64_t cap) {
  if (!mel_len || batch < 0 || window <= 0 || stride <= 0) return -1;
  int64_t n = 0;
  for (int64_t b = 0; b < batch; ++b) {
    const int64_t len = mel_len[b];
    if (len < 0 || len >= (int64_t(1) << 31)) return -1;
    if (len <= window) {   // "short audio": one segment (also for length 0: an all-zero window)
      if (batch_idx) {
        if (n >= cap) return -1;
        batch_idx[n] = int32_t(b); start[n] = 0; length[n] = int32_t(len);
      }
      ++n;
      continue;
    }
    for (int64_t st = 0; st < len; st += stride) {
      if (batch_idx) {
        if (n >= cap) return -1;
        batch_idx[n] = int32_t(b); start[n] = int32_t(st); length[n] = int32_t(std::min(window, len - st));
      }
      ++n;
    }
  }
  return n;
}

// mergeTokenizedSegments (Codec/S3Tokenizer/S3TokenizerUtils.swift:71-88; call sites S3Tokenizer.swift:622,837 with overlap = 4 s,
// tokenRate = 25): the token sequences of consecutive long-audio windows are joined by dropping (overlap / 2) * tokenRate tokens at
// every inner edge -- the left edge of every segment but the first, the right edge of every segment but the last; a segment whose
// kept range is empty contributes nothing.  tokens: the segments back to back; seg_len[n_segments].  Writes at most `cap` tokens
// to `out` (may be NULL to query) and returns the merged length, or -1 on bad arguments / overflow of `cap`.
int64_t b2a_merge_tokenized_segments(const int32_t* tokens, const int64_t* seg_len, int64_t n_segments, int overlap, int token_rate,
                                     int32_t* out, int64_t cap) {
  if (!seg_len || n_segments < 0 || (n_segments > 0 && !tokens)) return -1;
  const int64_t overlap_tokens = int64_t(overlap / 2) * token_rate;   // integer division first, as in the reference
  int64_t n = 0, base = 0;
  for (int64_t i = 0; i < n_segments; ++i) {
    const int64_t len = seg_len[i];
    if (len < 0) return -1;
    const int64_t left = i == 0 ? 0 : overlap_tokens;
    const int64_t right = i != n_segments - 1 ? len - overlap_tokens : len;
    if (left < right) {
      if (left < 0 || right > len) return -1;   // (the reference would trap on the out-of-range slice)
      if (out) {
        if (n + (right - left) > cap) return -1;
        for (int64_t t = left; t < right; ++t) out[n + (t - left)] = tokens[base + t];
      }
      n += right - left;
    }
    base += len;
  }
  return n;
}

// ---- polyphase resampler (non-parity extension: the reference's AudioResampler.resample, Audio/AudioResampler.swift:15-88, is
// Apple's AVAudioConverter and has no reproducible definition).  The design is scipy.signal.resample_poly's, which serves as its
// documented oracle: up / down reduced by their gcd, linear-phase low-pass firwin(2 * 10 * max(up, down) + 1, 1 / max(up, down),
// window = kaiser(beta 5)) scaled by `up`, zero-phase alignment, output length ceil(n * up / down), zeros outside the signal.
namespace {
double bessel_i0(double x) {   // power series, converges fast for the |x| <= 5 used here
  double sum = 1.0, term = 1.0;
  const double q = 0.25 * x * x;
  for (int k = 1; k < 200; ++k) {
    term *= q / (double(k) * double(k));
    sum += term;
    if (term < 1e-18 * sum) break;
  }
  return sum;
}
int64_t gcd64(int64_t a, int64_t b) {
  while (b) {
    const int64_t t = a % b;
    a = b;
    b = t;
  }
  return a;
}
int64_t upfirdn_len(int64_t len_h, int64_t n_in, int64_t up, int64_t down) {   // scipy.signal._upfirdn._output_len
  const int64_t nt = len_h + (n_in - 1) * up;
  return (nt + down - 1) / down;
}
}  // namespace

int64_t b2a_resample_poly_length(int64_t n_samples, int from_rate, int to_rate) {
  if (n_samples <= 0 || from_rate <= 0 || to_rate <= 0) return 0;
  const int64_t g = gcd64(to_rate, from_rate);
  const int64_t up = to_rate / g, down = from_rate / g;
  return (n_samples * up + down - 1) / down;
}

// Writes the zero-padded polyphase filter h (scaled by `up`) and the alignment: y[j] = sum_i x[i] * h[(j + pre_remove) * down - i * up].
// Returns the number of taps written (<= cap), or -1.
int64_t b2a_resample_poly_filter(int64_t n_samples, int from_rate, int to_rate, float* h_out, int64_t cap, int* up_out, int* down_out,
                                 int64_t* pre_remove_out) {
  if (n_samples <= 0 || from_rate <= 0 || to_rate <= 0 || from_rate == to_rate) return -1;
  const int64_t g = gcd64(to_rate, from_rate);
  const int64_t up = to_rate / g, down = from_rate / g;
  if (up > 4096 || down > 4096) return -1;
  const int64_t max_rate = std::max(up, down), half_len = 10 * max_rate, taps = 2 * half_len + 1;
  const double fc = 1.0 / double(max_rate), alpha = 0.5 * double(taps - 1), pi = 3.14159265358979323846;
  std::vector<double> h(static_cast<size_t>(taps));
  double sum = 0.0;
  const double i0b = bessel_i0(5.0);
  for (int64_t k = 0; k < taps; ++k) {
    const double m = double(k) - alpha, a = pi * fc * m;
    const double sinc = m == 0.0 ? 1.0 : sin(a) / a;
    const double r = 2.0 * double(k) / double(taps - 1) - 1.0;
    const double win = bessel_i0(5.0 * sqrt(std::max(0.0, 1.0 - r * r))) / i0b;
    h[size_t(k)] = fc * sinc * win;
    sum += h[size_t(k)];
  }
  const int64_t n_out = b2a_resample_poly_length(n_samples, from_rate, to_rate);
  const int64_t n_pre_pad = down - half_len % down;
  int64_t n_post_pad = 0;
  const int64_t n_pre_remove = (half_len + n_pre_pad) / down;
  while (upfirdn_len(taps + n_pre_pad + n_post_pad, n_samples, up, down) < n_out + n_pre_remove) ++n_post_pad;
  const int64_t total = n_pre_pad + taps + n_post_pad;
  if (up_out) *up_out = int(up);
  if (down_out) *down_out = int(down);
  if (pre_remove_out) *pre_remove_out = n_pre_remove;
  if (h_out) {
    if (total > cap) return -1;
    for (int64_t k = 0; k < total; ++k) h_out[k] = 0.0f;
    for (int64_t k = 0; k < taps; ++k) h_out[n_pre_pad + k] = float(h[size_t(k)] / sum * double(up));
  }
  return total;
}

int64_t b2a_resample_linear_length(int64_t n_samples, int from_rate, int to_rate) {  // CosyVoice2TTS.swift:733-739, CosyHiFTGenerator.swift:26-31
  if (n_samples <= 0 || from_rate <= 0 || to_rate <= 0) return 0;
  if (from_rate == to_rate) return n_samples;
  const float ratio = float(to_rate) / float(from_rate);
  int64_t n = int64_t(float(n_samples) * ratio);   // Int(Float(T) * scaleFactor): truncation, fp32 product
  return n == 0 ? 1 : n;
}

const char* b2a_version(void) { return "b200audio 0.1 (sm_100a)"; }

// Test hook (host only): compiles `bank` into the frontend kernel's mel step program and interprets it on the
// host exactly as the kernel does (per-chunk accumulators, emit-on-flag), for one spectrum `p` (n_bins).
// Returns the number of steps, or -1 when the bank is not of the two-adjacent-filters form.
int b2a_debug_mel_program_apply(const float* bank, int n_mels, int n_bins, int bin_major, const float* p, float* out) {
  b2a::SparseBank sb;
  b2a::build_sparse_bank(bank, n_mels, n_bins, bin_major != 0, sb);
  const int kChunks = 9;
  std::vector<int> words(n_mels);
  for (int m = 0; m < n_mels; ++m) words[m] = (m + 1) * 128;   // emit word = "row" m+1, no swizzle
  b2a::build_mel_program(bank, n_mels, n_bins, bin_major != 0, 1, words.data(), kChunks, nullptr, sb);
  if (sb.steps.empty()) return -1;
  std::vector<int> seen(n_mels, 0);
  for (int c = 0; c < kChunks; ++c) {
    float acc0 = 0.0f, acc1 = 0.0f;
    for (int s = sb.chunk_s[c]; s < sb.chunk_s[c + 1]; ++s) {
      const float* t = &sb.steps[size_t(s) * 4];
      int off, w;
      memcpy(&off, t + 2, 4);
      memcpy(&w, t + 3, 4);
      const float pk = p[off / 4];
      acc0 = fmaf(t[0], pk, acc0);
      acc1 = fmaf(t[1], pk, acc1);
      if (w != 0) {
        const int m = w / 128 - 1;
        if (m < sb.chunk_m[c] || m >= sb.chunk_m[c + 1]) return -2;  // a chunk emits its own filters only
        out[m] = acc0;
        seen[m] += 1;
        acc0 = acc1;
        acc1 = 0.0f;
      }
    }
  }
  for (int m = 0; m < n_mels; ++m)
    if (seen[m] != 1) return -2;  // every filter must have been emitted exactly once
  return int(sb.steps.size() / 4);
}

// Test hook (host only): builds the warp-per-frame mel schedule (build_wpf_mel) of `bank` and interprets it on the host as wpf1920.cu does
// (per-lane segment sums over len[r] bins from the lane's start bin, then the <= 4 segment sums of a filter in slot order), for one
// spectrum `p` (n_bins).  Returns the schedule's size in words, or -1 when the bank does not fit.
int b2a_debug_wpf_mel_apply(const float* bank, int n_mels, int n_bins, int bin_major, const float* p, float* out) {
  b2a::SparseBank sb;
  b2a::build_sparse_bank(bank, n_mels, n_bins, bin_major != 0, sb);
  std::vector<uint32_t> blob;
  if (!b2a::build_wpf_mel(sb, n_bins, blob)) return -1;
  const int rounds = int(blob[1]), span = (n_bins + 3) & ~3;
  std::vector<float> spec(size_t(span), 12345.0f);   // the words between the last bin and the 16-byte boundary hold stale finite data
  for (int k = 0; k < n_bins; ++k) spec[size_t(k)] = p[k];
  std::vector<float> part(size_t(rounds) * 32 + 1, 0.0f);
  for (int r = 0; r < rounds; ++r)
    for (int lane = 0; lane < 32; ++lane) {
      const int a0 = int(blob[blob[10] + size_t(r) * 32 + lane]);
      if (a0 % 4 != 0) return -3;   // 16-byte loads
      float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
      for (int g = 0; g < int(blob[2 + size_t(r)]); ++g)
        for (int e = 0; e < 4; ++e) {
          float w;
          memcpy(&w, &blob[blob[6 + size_t(r)] + (size_t(g) * 32 + lane) * 4 + e], 4);
          if (a0 + 4 * g + e >= span) return -2;   // the kernel would read past the spectrum row
          acc[e] = fmaf(w, spec[size_t(a0 + 4 * g + e)], acc[e]);
        }
      part[size_t(r) * 32 + lane] = (acc[0] + acc[1]) + (acc[2] + acc[3]);
    }
  for (int m = 0; m < n_mels; ++m) {
    const uint32_t* f = &blob[blob[11] + 4 * size_t(m)];
    out[m] = ((part[f[0]] + part[f[1]]) + part[f[2]]) + part[f[3]];
  }
  return int(blob.size()) + (int(blob[13]) << 16);   // schedule words, and the number of unavoidable bank-conflict pairs above bit 16
}

// Test hook (host only): the shared-memory layout of the FFT plan of `n_fft` -- row of the exchange buffer that holds each spectrum
// bin after stage B (slots_out, n_fft/2+1 entries, in rows of frame_tile floats) and the word offset at which each of `n_mels`
// mel sums is staged (words_out).  Returns frame_tile * 65536 + n1 * 256 + n2 / 1 packed as (frame_tile << 16) | (n1 << 8) | n2,
// or -1 if there is no plan / n_mels does not fit.
int b2a_debug_plan_layout(int n_fft, int n_mels, int* slots_out, int* words_out) {
  const b2a::PlanShape* ps = b2a::plan_shape(n_fft);
  if (!ps) return -1;
  std::vector<int> slots, words;
  b2a::spectrum_slots(*ps, slots);
  if (!b2a::output_words(*ps, n_mels, words)) return -1;
  for (size_t k = 0; k < slots.size(); ++k) slots_out[k] = slots[k];
  for (int m = 0; m < n_mels; ++m) words_out[m] = words[m] / int(sizeof(float));
  return (ps->frame_tile << 16) | (ps->n1 << 8) | ps->n2;
}

// Build hook (host only): the mel step program of `bank` for the FFT plan of `n_fft`, as raw 32-bit words (4 per step) plus the
// per-chunk filter / step boundaries.  tools/gen_mel_baked.py turns the programs of the reference's standard banks into
// straight-line device code at build time; at run time the program built for the caller's bank is compared with the baked
// one word for word before the baked kernel is chosen.  Returns the number of steps, -1 if the bank has no step program,
// -2 if `cap_steps` is too small.
int b2a_debug_mel_program_dump(const float* bank, int n_mels, int n_bins, int bin_major, int n_fft,
                               unsigned* steps_out, int cap_steps, int* chunk_m, int* chunk_s, int* n_chunks_out, int* frame_tile_out) {
  const b2a::PlanShape* ps = b2a::plan_shape(n_fft);
  if (!ps) return -3;
  const int frame_tile = ps->frame_tile, n_chunks = ps->n_chunks;
  *n_chunks_out = n_chunks;
  *frame_tile_out = frame_tile;
  std::vector<int> slots;
  b2a::spectrum_slots(*ps, slots);
  if (n_bins > int(slots.size())) return -3;
  b2a::SparseBank sb;
  b2a::build_sparse_bank(bank, n_mels, n_bins, bin_major != 0, sb);
  std::vector<int> words;
  if (!b2a::output_words(*ps, n_mels, words)) return -3;
  b2a::build_mel_program(bank, n_mels, n_bins, bin_major != 0, frame_tile, words.data(), n_chunks, slots.data(), sb);
  if (sb.steps.empty()) return -1;
  const int n = int(sb.steps.size() / 4);
  if (n > cap_steps) return -2;
  memcpy(steps_out, sb.steps.data(), sizeof(float) * sb.steps.size());
  for (int c = 0; c <= n_chunks; ++c) {
    chunk_m[c] = sb.chunk_m[c];
    chunk_s[c] = sb.chunk_s[c];
  }
  return n;
}

}  // extern "C"
