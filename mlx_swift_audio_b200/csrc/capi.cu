// C ABI of the B200 STFT-family DSP path (include/b200audio.h): context, device-table caches,
// scratch, the overlapped host<->device chunk pipeline for B2A_HOST buffers, and one entry point
// per reference helper.  There is no CPU fallback anywhere in this file: every compute entry
// point ends in a CUDA kernel launch or fails with B2A_E_CUDA.
#include <cuda_runtime.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <map>
#include <mutex>
#include <string>
#include <vector>

#include "../../include/b200audio.h"
#include "internal.h"

using namespace b2a;

namespace {

struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
};

struct BankStorage {
  SparseBank host;
  int* desc = nullptr;
  float* weights = nullptr;
  float* steps = nullptr;
  int baked_id = 0;
  uint32_t* wpf = nullptr;     // n_fft 1920: mel schedule of the warp-per-frame kernel (build_wpf_mel)
  int wpf_words = 0;
};

constexpr int kSlots = 3;  // chunk ring depth of the host pipeline

}  // namespace

struct b2a_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  cudaStream_t s_h2d = nullptr, s_d2h = nullptr;
  cudaEvent_t ev_h2d[kSlots] = {}, ev_comp[kSlots] = {}, ev_d2h[kSlots] = {};
  cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr;
  bool timing = false, timed = false;
  std::string err;
  int64_t launches = 0;
  std::mutex mu;
  std::map<std::string, BankStorage> banks;      // like the reference's MelFilterCache (CAMPPlus.swift:111-131), per context
  std::map<std::string, std::vector<float>> windows;
  DevBuf in[kSlots][2], out[kSlots][2];          // host-pipeline staging
  DevBuf fade;                                   // fade-in window of the fused vocoder head (device copy of fade_host)
  std::vector<float> fade_host;
  DevBuf scratch[kSlots][5];                     // 0: clip_max / flags, 1: tile_min, 2: temp features / unwrapped phase, 3 / 4: ragged clip / tile tables
  int64_t chunk_clip0 = 0;                       // first clip of the chunk run_batched is handing to the body
  int* h_flag = nullptr;                         // pinned
  int* h_tab[kSlots] = {};                       // pinned staging of the ragged clip tables (one per ring slot) ...
  size_t h_tab_ints[kSlots] = {};
  cudaEvent_t ev_tab[kSlots] = {};               // ... and the event behind each slot's last upload
};

namespace {

int fail(b2a_ctx* c, int code, const std::string& msg) {
  if (c) c->err = msg;
  return code;
}

int cu(b2a_ctx* c, cudaError_t e, const char* what) {
  if (e == cudaSuccess) return B2A_OK;
  return fail(c, B2A_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

int ensure(b2a_ctx* c, DevBuf& b, size_t bytes) {
  if (b.bytes >= bytes && b.p) return B2A_OK;
  if (b.p) {
    cudaError_t e = cudaFree(b.p);  // synchronises the device: no kernel still reads the old buffer
    b.p = nullptr;
    b.bytes = 0;
    if (e != cudaSuccess) return cu(c, e, "cudaFree");
  }
  size_t want = std::max<size_t>(bytes, 256);
  cudaError_t e = cudaMalloc(&b.p, want);
  if (e != cudaSuccess) {
    b.p = nullptr;
    return fail(c, e == cudaErrorMemoryAllocation ? B2A_E_NOMEM : B2A_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  b.bytes = want;
  return B2A_OK;
}

const std::vector<float>& cached_window(b2a_ctx* c, const std::string& key, const std::function<void(std::vector<float>&)>& make) {
  auto it = c->windows.find(key);
  if (it != c->windows.end()) return it->second;
  std::vector<float> w;
  make(w);
  return c->windows.emplace(key, std::move(w)).first->second;
}

int cached_bank(b2a_ctx* c, const std::string& key0, int n_fft, int n_mels, int n_bins, bool bin_major,
                const std::function<int(float*)>& make_dense, DeviceBank* out) {
  const PlanShape* ps = plan_shape(n_fft);   // null: no tuned plan -- the bank is only used by the generic kernels (desc + weights)
  const int frame_tile = ps ? ps->frame_tile : 0, n_chunks = ps ? ps->n_chunks : 0;
  const std::string key = key0 + "_ft" + std::to_string(frame_tile) + "_c" + std::to_string(n_chunks);
  auto it = c->banks.find(key);
  if (it == c->banks.end()) {
    std::vector<float> dense(size_t(n_mels) * n_bins);
    int rc = make_dense(dense.data());
    if (rc != B2A_OK) return fail(c, rc, "bad filterbank parameters");
    // validation and the (host-only) mel program first: nothing is allocated on the device for a bank that cannot run
    std::vector<int> slots, words;
    if (ps) {
      spectrum_slots(*ps, slots);
      if (n_bins > int(slots.size())) return fail(c, B2A_E_BAD_ARG, "filterbank has more bins than the spectrum");
      if (!output_words(*ps, n_mels, words)) return fail(c, B2A_E_BAD_ARG, "n_mels out of range for this FFT plan");
    }
    BankStorage bs;
    build_sparse_bank(dense.data(), n_mels, n_bins, bin_major, bs.host);
    if (ps) build_mel_program(dense.data(), n_mels, n_bins, bin_major, frame_tile, words.data(), n_chunks, slots.data(), bs.host);
    std::vector<uint32_t> wpf_blob;   // n_fft 1920: the warp-per-frame kernel's schedule (none when the bank does not fit: the tiled kernel runs)
    if (n_fft == 1920) build_wpf_mel(bs.host, n_fft / 2 + 1, wpf_blob);
    const size_t nw = std::max<size_t>(bs.host.weights.size(), 1);
    std::vector<int> desc(size_t(n_mels) * 4, 0);
    for (int m = 0; m < n_mels; ++m) {
      desc[4 * m + 0] = bs.host.start[m];
      desc[4 * m + 1] = bs.host.count[m];
      desc[4 * m + 2] = bs.host.offset[m];
    }
    // synchronous uploads (once per configuration), ordered before any later launch; a failure frees what was allocated
    auto upload = [&]() -> cudaError_t {
      cudaError_t e;
      if ((e = cudaMalloc(&bs.desc, sizeof(int) * desc.size())) != cudaSuccess) return e;
      if ((e = cudaMalloc(&bs.weights, sizeof(float) * nw)) != cudaSuccess) return e;
      if ((e = cudaMemcpy(bs.desc, desc.data(), sizeof(int) * desc.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
      if (!bs.host.weights.empty() &&
          (e = cudaMemcpy(bs.weights, bs.host.weights.data(), sizeof(float) * bs.host.weights.size(), cudaMemcpyHostToDevice)) != cudaSuccess)
        return e;
      if (!bs.host.steps.empty()) {
        if ((e = cudaMalloc(&bs.steps, sizeof(float) * bs.host.steps.size())) != cudaSuccess) return e;
        if ((e = cudaMemcpy(bs.steps, bs.host.steps.data(), sizeof(float) * bs.host.steps.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
      }
      if (!wpf_blob.empty()) {
        if ((e = cudaMalloc(&bs.wpf, sizeof(uint32_t) * wpf_blob.size())) != cudaSuccess) return e;
        if ((e = cudaMemcpy(bs.wpf, wpf_blob.data(), sizeof(uint32_t) * wpf_blob.size(), cudaMemcpyHostToDevice)) != cudaSuccess) return e;
        bs.wpf_words = int(wpf_blob.size());
      }
      return cudaSuccess;
    };
    const cudaError_t ue = upload();
    if (ue != cudaSuccess) {
      if (bs.desc) cudaFree(bs.desc);
      if (bs.weights) cudaFree(bs.weights);
      if (bs.steps) cudaFree(bs.steps);
      if (bs.wpf) cudaFree(bs.wpf);
      return cu(c, ue, "filterbank upload");
    }
    if (!bs.host.steps.empty())
      bs.baked_id = frontend_match_baked(bs.host.steps.data(), int(bs.host.steps.size() / 4), bs.host.chunk_m.data(), bs.host.chunk_s.data(),
                                         n_chunks, frame_tile, n_mels);
    it = c->banks.emplace(key, std::move(bs)).first;
  }
  const BankStorage& bs = it->second;
  out->desc = bs.desc;
  out->weights = bs.weights;
  out->steps = bs.steps;
  out->n_steps = int(bs.host.steps.size() / 4);
  out->n_chunks = n_chunks;
  out->frame_tile = frame_tile;
  out->host_chunk_m = bs.host.chunk_m.empty() ? nullptr : bs.host.chunk_m.data();
  out->host_chunk_s = bs.host.chunk_s.empty() ? nullptr : bs.host.chunk_s.data();
  out->n_mels = n_mels;
  out->n_bins_used = bs.host.max_bin + 1;
  out->baked_id = bs.baked_id;
  out->wpf_mel = bs.wpf;
  out->wpf_words = bs.wpf_words;
  return B2A_OK;
}

struct Guard {
  b2a_ctx* c;
  std::unique_lock<std::mutex> lk;
  int prev_dev = -1;
  bool ok = true;
  explicit Guard(b2a_ctx* ctx) : c(ctx), lk(ctx->mu) {
    cudaGetDevice(&prev_dev);
    if (prev_dev != c->device) ok = cudaSetDevice(c->device) == cudaSuccess;
    c->err.clear();
  }
  ~Guard() {
    if (prev_dev >= 0 && prev_dev != c->device) cudaSetDevice(prev_dev);
  }
};

// Runs `body(d_in0, d_in1, d_out0, d_out1, n_clips, slot)` over the batch.
//  B2A_DEVICE: a single call on the caller's pointers.
//  B2A_HOST  : clips are streamed through device staging buffers in chunks; H2D of chunk i+1, the
//              kernels of chunk i and D2H of chunk i-1 overlap on three streams.
using Body = std::function<int(const float*, const float*, float*, float*, int64_t, int)>;
constexpr int64_t kMaxClipsPerLaunch = 32768;   // several kernels index clips with gridDim.y (limit 65535): both memory spaces split here

// Sizes per clip in BYTES (the buffers may hold fp32, fp16 or 16-bit PCM); the body sees them through float-typed pointers.
int run_batched_bytes(b2a_ctx* c, int space, int64_t batch, const void* in0_v, size_t in0_per_clip, const void* in1_v,
                      size_t in1_per_clip, void* out0_v, size_t out0_per_clip, void* out1_v, size_t out1_per_clip, const Body& body) {
  const char* in0 = static_cast<const char*>(in0_v);
  const char* in1 = static_cast<const char*>(in1_v);
  char* out0 = static_cast<char*>(out0_v);
  char* out1 = static_cast<char*>(out1_v);
  auto F = [](const char* p) { return reinterpret_cast<const float*>(p); };
  auto FM = [](char* p) { return reinterpret_cast<float*>(p); };
  if (space == B2A_DEVICE) {
    if (c->timing) cudaEventRecord(c->ev_t0, c->stream);
    int rc = B2A_OK;
    for (int64_t c0 = 0; c0 < batch && rc == B2A_OK; c0 += kMaxClipsPerLaunch) {
      const int64_t n = std::min(kMaxClipsPerLaunch, batch - c0);
      c->chunk_clip0 = c0;
      rc = body(F(in0 + c0 * in0_per_clip), in1 ? F(in1 + c0 * in1_per_clip) : nullptr, FM(out0 + c0 * out0_per_clip),
                out1 ? FM(out1 + c0 * out1_per_clip) : nullptr, n, 0);
    }
    if (c->timing) {
      cudaEventRecord(c->ev_t1, c->stream);
      c->timed = true;
    }
    return rc;
  }
  if (space != B2A_HOST) return fail(c, B2A_E_BAD_ARG, "space must be B2A_HOST or B2A_DEVICE");
  const size_t per_clip = in0_per_clip + in1_per_clip + out0_per_clip + out1_per_clip;
  // bytes per chunk (inputs + outputs).  Measured on the B200 box (tools/gpu/e2e_sweep.sh): the host path is bound by the H2D copies
  // (~54 GB/s) for any chunk of 32..192 MB; 64 MB keeps fill / drain small (measured on the pcm16 -> fp16 Whisper call: 16 MB 22.97 ms, 32 MB 20.62, 64 MB 19.88, 128 MB 20.30, 256 MB 21.28; ramping the first and last
  // chunks down to 1/8 of the size changed nothing: the step is bound by the duplex PCIe traffic itself).  B2A_HOST_CHUNK_MB overrides.
  size_t target = size_t(64) << 20;
  if (const char* mb = getenv("B2A_HOST_CHUNK_MB")) target = size_t(std::max(1, atoi(mb))) << 20;
  int64_t chunk = std::max<int64_t>(1, int64_t(target / std::max<size_t>(per_clip, 1)));
  chunk = std::min(std::min(chunk, batch), kMaxClipsPerLaunch);   // (tiny clips: the byte target alone would exceed gridDim.y)
  const int64_t n_chunks = (batch + chunk - 1) / chunk;
  const int slots = int(std::min<int64_t>(kSlots, n_chunks));
  for (int s = 0; s < slots; ++s) {
    int rc;
    if ((rc = ensure(c, c->in[s][0], in0_per_clip * chunk)) != B2A_OK) return rc;
    if (in1 && (rc = ensure(c, c->in[s][1], in1_per_clip * chunk)) != B2A_OK) return rc;
    if ((rc = ensure(c, c->out[s][0], out0_per_clip * chunk)) != B2A_OK) return rc;
    if (out1 && (rc = ensure(c, c->out[s][1], out1_per_clip * chunk)) != B2A_OK) return rc;
  }
  cudaError_t e;
  if (n_chunks == 1) {
    // A call that fits one chunk (a single clip, a short batch): nothing to overlap, so the copies ride the compute stream -- no
    // cross-stream events in the latency path of the reference's one-clip calls
    if ((e = cudaMemcpyAsync(c->in[0][0].p, in0, in0_per_clip * batch, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return cu(c, e, "H2D copy");
    if (in1 && (e = cudaMemcpyAsync(c->in[0][1].p, in1, in1_per_clip * batch, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess)
      return cu(c, e, "H2D copy");
    c->chunk_clip0 = 0;
    int rc = body(static_cast<const float*>(c->in[0][0].p), static_cast<const float*>(c->in[0][1].p), static_cast<float*>(c->out[0][0].p),
                  static_cast<float*>(c->out[0][1].p), batch, 0);
    if (rc != B2A_OK) {
      cudaStreamSynchronize(c->stream);
      return rc;
    }
    if ((e = cudaMemcpyAsync(out0, c->out[0][0].p, out0_per_clip * batch, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return cu(c, e, "D2H copy");
    if (out1 && (e = cudaMemcpyAsync(out1, c->out[0][1].p, out1_per_clip * batch, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess)
      return cu(c, e, "D2H copy");
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return cu(c, e, "sync");
    return B2A_OK;
  }
  // order the copy streams after whatever is already queued on the compute stream
  if ((e = cudaEventRecord(c->ev_comp[0], c->stream)) != cudaSuccess) return cu(c, e, "event record");
  if ((e = cudaStreamWaitEvent(c->s_h2d, c->ev_comp[0], 0)) != cudaSuccess) return cu(c, e, "stream wait");
  for (int64_t i = 0; i < n_chunks; ++i) {
    const int s = int(i % kSlots);
    const int64_t c0 = i * chunk, n = std::min(chunk, batch - c0);
    if (i >= kSlots) {  // slot reuse: its previous D2H must have drained
      if ((e = cudaStreamWaitEvent(c->s_h2d, c->ev_d2h[s], 0)) != cudaSuccess) return cu(c, e, "stream wait");
    }
    if ((e = cudaMemcpyAsync(c->in[s][0].p, in0 + c0 * in0_per_clip, in0_per_clip * n, cudaMemcpyHostToDevice, c->s_h2d)) != cudaSuccess)
      return cu(c, e, "H2D copy");
    if (in1 && (e = cudaMemcpyAsync(c->in[s][1].p, in1 + c0 * in1_per_clip, in1_per_clip * n, cudaMemcpyHostToDevice, c->s_h2d)) != cudaSuccess)
      return cu(c, e, "H2D copy");
    if ((e = cudaEventRecord(c->ev_h2d[s], c->s_h2d)) != cudaSuccess) return cu(c, e, "event record");
    if ((e = cudaStreamWaitEvent(c->stream, c->ev_h2d[s], 0)) != cudaSuccess) return cu(c, e, "stream wait");
    c->chunk_clip0 = c0;
    int rc = body(static_cast<const float*>(c->in[s][0].p), static_cast<const float*>(c->in[s][1].p),
                  static_cast<float*>(c->out[s][0].p), static_cast<float*>(c->out[s][1].p), n, s);
    if (rc != B2A_OK) {
      cudaStreamSynchronize(c->s_h2d);
      cudaStreamSynchronize(c->stream);
      cudaStreamSynchronize(c->s_d2h);
      return rc;
    }
    if ((e = cudaEventRecord(c->ev_comp[s], c->stream)) != cudaSuccess) return cu(c, e, "event record");
    if ((e = cudaStreamWaitEvent(c->s_d2h, c->ev_comp[s], 0)) != cudaSuccess) return cu(c, e, "stream wait");
    if ((e = cudaMemcpyAsync(out0 + c0 * out0_per_clip, c->out[s][0].p, out0_per_clip * n, cudaMemcpyDeviceToHost, c->s_d2h)) != cudaSuccess)
      return cu(c, e, "D2H copy");
    if (out1 && (e = cudaMemcpyAsync(out1 + c0 * out1_per_clip, c->out[s][1].p, out1_per_clip * n, cudaMemcpyDeviceToHost, c->s_d2h)) != cudaSuccess)
      return cu(c, e, "D2H copy");
    if ((e = cudaEventRecord(c->ev_d2h[s], c->s_d2h)) != cudaSuccess) return cu(c, e, "event record");
  }
  if ((e = cudaStreamSynchronize(c->s_d2h)) != cudaSuccess) return cu(c, e, "sync");
  if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return cu(c, e, "sync");
  return B2A_OK;
}

// fp32 buffers: sizes per clip in floats
int run_batched(b2a_ctx* c, int space, int64_t batch, const float* in0, size_t in0_per_clip, const float* in1,
                size_t in1_per_clip, float* out0, size_t out0_per_clip, float* out1, size_t out1_per_clip, const Body& body) {
  return run_batched_bytes(c, space, batch, in0, sizeof(float) * in0_per_clip, in1, sizeof(float) * in1_per_clip, out0,
                           sizeof(float) * out0_per_clip, out1, sizeof(float) * out1_per_clip, body);
}

int check_common(b2a_ctx* c, const void* in, const void* out, int64_t batch, int64_t n) {
  if (!c) return B2A_E_BAD_ARG;
  if (!in || !out) return fail(c, B2A_E_BAD_ARG, "null buffer");
  if (batch <= 0 || n <= 0) return fail(c, B2A_E_BAD_ARG, "batch and sizes must be positive");
  return B2A_OK;
}

// One preset of the fused front-end.
struct Preset {
  int n_fft = 400, hop = 160, win_len = 400;
  const std::vector<float>* window = nullptr;  // zero-extended to n_fft
  int pad_mode = PAD_REFLECT;
  int64_t pad_left = 200;
  int64_t zero_tail = 0;
  int pre_mode = PRE_NONE;
  int spec_mode = SPEC_POWER;
  DeviceBank bank;
  int log_mode = LOG_NONE;
  float log_floor = 0.0f;
  int whisper_norm = 0;
  int post_affine = 0;
  float post_sub = 0.0f, post_div = 1.0f;
  int out_mode = OUT_TM;
  int64_t n_frames = 0;
  int lfr_m = 7, lfr_n = 6;
  int64_t lfr_rows = 0;
  int in_i16 = 0;          // the clips are 16-bit PCM (converted to fp32 in [-1, 1) on the device, x / 32768)
  int out_f16 = 0;         // the features are written as fp16 by the store loop (Whisper (T', M) front end)
  int post_cmvn = 0;       // Fun-ASR per-utterance CMVN on the LFR features
  int post_mean_norm = 0;  // CAM++ time-mean removal
};

// Per-clip lengths of a ragged batch (the *_ragged entry points): clip b holds lengths[b] <= n_samples valid samples at the
// start of its row; `frames_of` is the front end's own frame-count rule.  Outputs keep the strides of an n_samples-long clip;
// rows past a clip's own count are zero.
struct Ragged {
  const int64_t* lengths = nullptr;             // host
  std::function<int64_t(int64_t)> frames_of;
  int64_t* out_rows = nullptr;                  // host, optional: rows (frames / LFR rows) written per clip
};

int run_preset(b2a_ctx* c, const Preset& p, const void* audio, int64_t batch, int64_t n_samples, void* out, int space,
               const Ragged* rg = nullptr) {
  if (p.out_f16 && (p.bank.n_mels & 1)) return fail(c, B2A_E_BAD_ARG, "fp16 output needs an even n_mels");
  std::vector<int64_t> clip_frames;
  if (rg) {
    if (p.out_mode == OUT_COMPLEX) return fail(c, B2A_E_UNSUPPORTED, "ragged batches are built for the mel front ends only");
    if (n_samples > 0x7fffffffLL) return fail(c, B2A_E_BAD_ARG, "ragged batches hold clips of fewer than 2^31 samples");   // (32-bit clip table)
    clip_frames.resize(size_t(batch));
    for (int64_t b = 0; b < batch; ++b) {
      const int64_t len = rg->lengths[b];
      if (len <= 0 || len > n_samples) return fail(c, B2A_E_BAD_ARG, "lengths[b] must lie in [1, n_samples]");
      const int64_t fr = rg->frames_of(len);
      if (fr <= 0) return fail(c, B2A_E_TOO_SHORT, "Input is too short for STFT");
      clip_frames[size_t(b)] = fr;
      if (rg->out_rows) rg->out_rows[b] = p.out_mode == OUT_LFR ? b2a_lfr_num_rows(fr, p.lfr_n) : fr;
    }
  }
  size_t out_per_clip;
  switch (p.out_mode) {
    case OUT_LFR: out_per_clip = size_t(p.lfr_rows) * p.lfr_m * p.bank.n_mels; break;
    case OUT_COMPLEX: out_per_clip = size_t(p.n_frames) * (p.n_fft / 2 + 1) * 2; break;
    default: out_per_clip = size_t(p.n_frames) * p.bank.n_mels; break;
  }
  const int tiles = frontend_tiles_per_clip(p.n_fft, p.n_frames);
  Body body = [&](const float* d_in, const float*, float* d_out, float*, int64_t n, int slot) -> int {
    FrontendArgs a;
    a.n_fft = p.n_fft; a.hop = p.hop; a.win_len = p.win_len;
    a.x = d_in; a.batch = n; a.n_samples = n_samples; a.zero_tail = p.zero_tail; a.out_f16 = p.out_f16;
    a.pad_mode = p.pad_mode; a.pad_left = p.pad_left; a.pre_mode = p.pre_mode;
    a.window = p.window->data();
    a.spec_mode = p.spec_mode; a.bank = p.bank; a.log_mode = p.log_mode; a.log_floor = p.log_floor;
    a.whisper_norm = p.whisper_norm; a.post_affine = p.post_affine; a.post_sub = p.post_sub; a.post_div = p.post_div;
    a.out_mode = p.out_mode; a.n_frames = p.n_frames; a.out = d_out; a.lfr_m = p.lfr_m; a.lfr_n = p.lfr_n; a.lfr_rows = p.lfr_rows;
    int64_t max_tail = 0;        // ragged: rows the shortest clip of this chunk falls short of the batch stride
    int launches = 0;
    std::string err;
    if (p.in_i16) {   // 16-bit PCM -> fp32 scratch (x / 32768, exact), then the fp32 front end
      int rc;
      if ((rc = ensure(c, c->scratch[slot][2], sizeof(float) * size_t(n) * size_t(n_samples))) != B2A_OK) return rc;
      if ((rc = launch_pcm16_to_f32(d_in, static_cast<float*>(c->scratch[slot][2].p), n * n_samples, c->stream, &launches, &err)) != B2A_OK) {
        c->err = err;
        return rc;
      }
      a.x = static_cast<const float*>(c->scratch[slot][2].p);
    }
    if (rg) {
      const PlanShape* ps = plan_shape(p.n_fft);
      const int ft = ps ? ps->frame_tile : 32;
      const int64_t c0 = c->chunk_clip0;
      // the table is written into pinned staging (a pageable source would make the "async" upload a synchronous staged copy);
      // the slot's previous upload has long finished, the event only makes that a guarantee
      const size_t n_ints = size_t(n) * 4;
      cudaError_t e;
      if (c->h_tab_ints[slot] < n_ints) {
        if (c->h_tab[slot]) {
          if ((e = cudaEventSynchronize(c->ev_tab[slot])) != cudaSuccess) return cu(c, e, "event sync");
          cudaFreeHost(c->h_tab[slot]);
          c->h_tab[slot] = nullptr;
          c->h_tab_ints[slot] = 0;
        }
        if ((e = cudaHostAlloc(reinterpret_cast<void**>(&c->h_tab[slot]), sizeof(int) * n_ints, cudaHostAllocDefault)) != cudaSuccess)
          return fail(c, B2A_E_NOMEM, std::string("cudaHostAlloc: ") + cudaGetErrorString(e));
        c->h_tab_ints[slot] = n_ints;
      } else if ((e = cudaEventSynchronize(c->ev_tab[slot])) != cudaSuccess) {
        return cu(c, e, "event sync");
      }
      int* clip_tab = c->h_tab[slot];
      int64_t total = 0, min_rows = 0x7fffffffffffffffLL;
      for (int64_t b = 0; b < n; ++b) {
        const int64_t fr = clip_frames[size_t(c0 + b)];
        const int64_t lfr = p.out_mode == OUT_LFR ? b2a_lfr_num_rows(fr, p.lfr_n) : 0;
        clip_tab[4 * b + 0] = int(rg->lengths[c0 + b]);
        clip_tab[4 * b + 1] = int(fr);
        clip_tab[4 * b + 2] = int(lfr);
        clip_tab[4 * b + 3] = int(total);
        total += (fr + ft - 1) / ft;
        min_rows = std::min(min_rows, p.out_mode == OUT_LFR ? lfr : fr);
      }
      max_tail = (p.out_mode == OUT_LFR ? p.lfr_rows : p.n_frames) - min_rows;
      if (total > 0x7fffffffLL) return fail(c, B2A_E_BAD_ARG, "too many tiles in one launch");
      int rc;
      if ((rc = ensure(c, c->scratch[slot][3], sizeof(int) * n_ints)) != B2A_OK) return rc;
      if ((e = cudaMemcpyAsync(c->scratch[slot][3].p, clip_tab, sizeof(int) * n_ints, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess)
        return cu(c, e, "table upload");
      if ((e = cudaEventRecord(c->ev_tab[slot], c->stream)) != cudaSuccess) return cu(c, e, "event record");
      if ((rc = ensure(c, c->scratch[slot][4], sizeof(int) * 4 * size_t(total))) != B2A_OK) return rc;
      a.clip_tab = c->scratch[slot][3].p;
      a.tile_tab = c->scratch[slot][4].p;
      a.total_tiles = total;
      if ((rc = launch_tile_table(a.clip_tab, n, total, c->scratch[slot][4].p, c->stream, &launches, &err)) != B2A_OK) {
        c->err = err;
        return rc;
      }
      // rows past a clip's own count are zero (only those: the kernels write the rest)
      if (p.out_mode == OUT_MT) rc = launch_zero_tails(d_out, a.clip_tab, 1, n, p.bank.n_mels, p.n_frames, 1, max_tail, c->stream, &launches, &err);
      else if (p.out_mode == OUT_LFR) rc = launch_zero_tails(d_out, a.clip_tab, 2, n, p.lfr_rows, int64_t(p.lfr_m) * p.bank.n_mels, 0, max_tail, c->stream, &launches, &err);
      else rc = launch_zero_tails(d_out, a.clip_tab, 1, n, p.n_frames, p.out_f16 ? p.bank.n_mels / 2 : p.bank.n_mels, 0, max_tail, c->stream, &launches, &err);   // (fp16 rows: n_mels / 2 words of zero bits)
      if (rc != B2A_OK) {
        c->err = err;
        return rc;
      }
    }
    if (p.whisper_norm) {
      int rc;
      // clip maxima and tile minima back to back: one memset initialises both (launch_plan)
      const size_t n_tiles = rg ? size_t(a.total_tiles) : size_t(n) * tiles;
      if ((rc = ensure(c, c->scratch[slot][0], sizeof(int) * (size_t(n) + 2 * n_tiles + 1))) != B2A_OK) return rc;
      a.clip_max = static_cast<int*>(c->scratch[slot][0].p);
      a.tile_min = reinterpret_cast<float*>(a.clip_max + n);
      a.tile_max = a.clip_max + n + n_tiles;       // per-tile maxima (used by the kernels that track them)
      a.tile_ctr = a.clip_max + n + 2 * n_tiles;   // the dynamic tile walk's counter, initialised by the same memset
    } else {
      int rc;
      if ((rc = ensure(c, c->scratch[slot][0], sizeof(int))) != B2A_OK) return rc;
      a.tile_ctr = static_cast<int*>(c->scratch[slot][0].p);
    }
    int rc;
    if (!frontend_plan_exists(p.n_fft, p.hop, p.win_len)) {
      // no tuned plan for this (n_fft, hop): generic direct-DFT kernel (+ spectrum -> mel kernel for the mel front ends)
      if (rg || p.pre_mode != PRE_NONE || p.whisper_norm || p.out_mode == OUT_LFR || p.in_i16 || p.out_f16 || p.zero_tail != 0) {
        c->err = "this front end is built for its standard (n_fft, hop) only";
        return B2A_E_UNSUPPORTED;
      }
      const int n_bins = p.n_fft / 2 + 1;
      if ((rc = ensure(c, c->scratch[slot][4], sizeof(float) * size_t(p.n_fft))) != B2A_OK) return rc;
      float* d_win = static_cast<float*>(c->scratch[slot][4].p);
      cudaError_t e = cudaMemcpyAsync(d_win, p.window->data(), sizeof(float) * size_t(p.n_fft), cudaMemcpyHostToDevice, c->stream);   // (pageable: staged at once)
      if (e != cudaSuccess) return cu(c, e, "window upload");
      float* d_spec = d_out;
      if (p.out_mode != OUT_COMPLEX) {
        if ((rc = ensure(c, c->scratch[slot][2], sizeof(float) * 2 * size_t(n) * size_t(p.n_frames) * n_bins)) != B2A_OK) return rc;
        d_spec = static_cast<float*>(c->scratch[slot][2].p);
      }
      rc = launch_generic_stft(a.x, n, n_samples, p.n_frames, p.n_fft, p.hop, p.pad_left, p.pad_mode, d_win, d_spec, c->stream, &launches, &err);
      if (rc == B2A_OK && p.out_mode != OUT_COMPLEX)
        rc = launch_generic_mel(d_spec, d_out, n, p.n_frames, n_bins, p.bank, p.spec_mode, p.log_mode, p.log_floor, p.post_affine, p.post_sub, p.post_div,
                                p.out_mode, c->stream, &launches, &err);
    } else {
      rc = launch_frontend(a, c->stream, &launches, &err);
    }
    if (rc == B2A_OK && p.post_cmvn)
      rc = launch_cmvn(d_out, d_out, n, p.lfr_rows, p.lfr_m * p.bank.n_mels, nullptr, nullptr, c->stream, &launches, &err, a.clip_tab);
    if (rc == B2A_OK && p.post_mean_norm) rc = launch_mean_norm(d_out, n, p.n_frames, p.bank.n_mels, c->stream, &launches, &err, a.clip_tab);
    c->launches += launches;
    if (rc != B2A_OK) c->err = err;
    return rc;
  };
  return run_batched_bytes(c, space, batch, audio, size_t(n_samples) * (p.in_i16 ? sizeof(int16_t) : sizeof(float)), nullptr, 0, out,
                           out_per_clip * (p.out_f16 ? sizeof(uint16_t) : sizeof(float)), nullptr, 0, body);
}

std::string key_of(const char* tag, std::initializer_list<double> v) {
  std::string k = tag;
  for (double d : v) {
    k += '_';
    k += std::to_string(d);
  }
  return k;
}

int slaney_bank(b2a_ctx* c, int sr, int n_fft, int n_mels, float fmin, float fmax, DeviceBank* out) {
  if (n_mels <= 0) return fail(c, B2A_E_BAD_ARG, "n_mels must be positive");
  return cached_bank(c, key_of("slaney", {double(sr), double(n_fft), double(n_mels), fmin, fmax}), n_fft, n_mels, n_fft / 2 + 1, false,
                     [&](float* d) { return mel_filters_slaney(sr, n_fft, n_mels, fmin, fmax, d); }, out);
}

const std::vector<float>& window_of(b2a_ctx* c, int kind, int length, int n_fft, bool via_plus_one) {
  return cached_window(c, key_of("win", {double(kind), double(length), double(n_fft), double(via_plus_one)}), [&](std::vector<float>& w) {
    if (via_plus_one) hann_periodic_via_hanning(length, w);
    else {
      w.resize(length);
      make_window(kind, length, w.data());
    }
    w.resize(n_fft, 0.0f);  // zero-extend (S3TokenizerUtils.swift:235-239)
  });
}

}  // namespace

extern "C" {

// ---- context ------------------------------------------------------------------------------------
static int ctx_create(b2a_ctx** out, int device, void* stream, bool own) {
  if (!out) return B2A_E_BAD_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || count <= 0 || device < 0 || device >= count) return B2A_E_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return B2A_E_CUDA;
  if (prop.major != 10) return B2A_E_CUDA;  // kernels are built for sm_100a only
  int prev = 0;
  cudaGetDevice(&prev);
  if (cudaSetDevice(device) != cudaSuccess) return B2A_E_CUDA;
  b2a_ctx* c = new b2a_ctx();
  c->device = device;
  int rc = B2A_OK;
  do {
    if (own) {
      if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { rc = B2A_E_CUDA; break; }
      c->own_stream = true;
    } else {
      c->stream = static_cast<cudaStream_t>(stream);
    }
    if (cudaStreamCreateWithFlags(&c->s_h2d, cudaStreamNonBlocking) != cudaSuccess) { rc = B2A_E_CUDA; break; }
    if (cudaStreamCreateWithFlags(&c->s_d2h, cudaStreamNonBlocking) != cudaSuccess) { rc = B2A_E_CUDA; break; }
    for (int s = 0; s < kSlots && rc == B2A_OK; ++s) {
      if (cudaEventCreateWithFlags(&c->ev_h2d[s], cudaEventDisableTiming) != cudaSuccess) rc = B2A_E_CUDA;
      if (cudaEventCreateWithFlags(&c->ev_comp[s], cudaEventDisableTiming) != cudaSuccess) rc = B2A_E_CUDA;
      if (cudaEventCreateWithFlags(&c->ev_d2h[s], cudaEventDisableTiming) != cudaSuccess) rc = B2A_E_CUDA;
    }
    if (rc != B2A_OK) break;
    if (cudaEventCreate(&c->ev_t0) != cudaSuccess || cudaEventCreate(&c->ev_t1) != cudaSuccess) { rc = B2A_E_CUDA; break; }
    if (cudaHostAlloc(reinterpret_cast<void**>(&c->h_flag), sizeof(int), cudaHostAllocDefault) != cudaSuccess) { rc = B2A_E_CUDA; break; }
    for (int s = 0; s < kSlots && rc == B2A_OK; ++s)
      if (cudaEventCreateWithFlags(&c->ev_tab[s], cudaEventDisableTiming) != cudaSuccess) rc = B2A_E_CUDA;
    if (rc != B2A_OK) break;
    std::string err;
    rc = init_frontend_tables(&err);
  } while (false);
  cudaSetDevice(prev);
  if (rc != B2A_OK) {
    b2a_ctx_destroy(c);
    return rc;
  }
  *out = c;
  return B2A_OK;
}

int b2a_ctx_create(b2a_ctx** ctx, int device) { return ctx_create(ctx, device, nullptr, true); }
int b2a_ctx_create_on_stream(b2a_ctx** ctx, int device, void* cuda_stream) { return ctx_create(ctx, device, cuda_stream, false); }

int b2a_ctx_destroy(b2a_ctx* c) {
  if (!c) return B2A_OK;
  int prev = 0;
  cudaGetDevice(&prev);
  cudaSetDevice(c->device);
  cudaDeviceSynchronize();
  for (int s = 0; s < kSlots; ++s) {
    for (int j = 0; j < 2; ++j) {
      if (c->in[s][j].p) cudaFree(c->in[s][j].p);
      if (c->out[s][j].p) cudaFree(c->out[s][j].p);
    }
    for (int j = 0; j < 5; ++j)
      if (c->scratch[s][j].p) cudaFree(c->scratch[s][j].p);
    if (c->ev_h2d[s]) cudaEventDestroy(c->ev_h2d[s]);
    if (c->ev_comp[s]) cudaEventDestroy(c->ev_comp[s]);
    if (c->ev_d2h[s]) cudaEventDestroy(c->ev_d2h[s]);
  }
  for (auto& kv : c->banks) {
    cudaFree(kv.second.desc);
    cudaFree(kv.second.weights);
    if (kv.second.steps) cudaFree(kv.second.steps);
    if (kv.second.wpf) cudaFree(kv.second.wpf);
  }
  if (c->ev_t0) cudaEventDestroy(c->ev_t0);
  if (c->ev_t1) cudaEventDestroy(c->ev_t1);
  if (c->fade.p) cudaFree(c->fade.p);
  if (c->h_flag) cudaFreeHost(c->h_flag);
  for (int s = 0; s < kSlots; ++s) {
    if (c->h_tab[s]) cudaFreeHost(c->h_tab[s]);
    if (c->ev_tab[s]) cudaEventDestroy(c->ev_tab[s]);
  }
  if (c->s_h2d) cudaStreamDestroy(c->s_h2d);
  if (c->s_d2h) cudaStreamDestroy(c->s_d2h);
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  cudaSetDevice(prev);
  delete c;
  return B2A_OK;
}

int b2a_ctx_sync(b2a_ctx* c) {
  if (!c) return B2A_E_BAD_ARG;
  return cu(c, cudaStreamSynchronize(c->stream), "sync");
}

void* b2a_ctx_stream(const b2a_ctx* c) { return c ? static_cast<void*>(c->stream) : nullptr; }

const char* b2a_last_error(const b2a_ctx* c) { return c ? c->err.c_str() : "null context"; }
int64_t b2a_ctx_launch_count(const b2a_ctx* c) { return c ? c->launches : 0; }

int b2a_host_alloc(void** ptr, uint64_t bytes) {
  if (!ptr) return B2A_E_BAD_ARG;
  return cudaHostAlloc(ptr, bytes, cudaHostAllocDefault) == cudaSuccess ? B2A_OK : B2A_E_NOMEM;
}
int b2a_host_free(void* ptr) { return cudaFreeHost(ptr) == cudaSuccess ? B2A_OK : B2A_E_CUDA; }

// ---- multi-GPU: peer-mapped output buffers (the kernels' epilogue stores are the gather) ----------
namespace {
struct DeviceGuard {   // the calling thread's current device is restored on scope exit
  int prev = 0;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    cudaSetDevice(dev);
  }
  ~DeviceGuard() { cudaSetDevice(prev); }
};
}  // namespace

int b2a_device_alloc(b2a_ctx* c, void** ptr, uint64_t bytes) {
  if (!c || !ptr || bytes == 0) return fail(c, B2A_E_BAD_ARG, "b2a_device_alloc: bad argument");
  DeviceGuard g(c->device);
  cudaError_t e = cudaMalloc(ptr, bytes);
  if (e != cudaSuccess) {
    *ptr = nullptr;
    return fail(c, e == cudaErrorMemoryAllocation ? B2A_E_NOMEM : B2A_E_CUDA, std::string("cudaMalloc: ") + cudaGetErrorString(e));
  }
  return B2A_OK;
}

int b2a_device_free(b2a_ctx* c, void* ptr) {
  if (!c) return B2A_E_BAD_ARG;
  DeviceGuard g(c->device);
  return cu(c, cudaFree(ptr), "cudaFree");
}

int b2a_ipc_export(b2a_ctx* c, void* device_ptr, unsigned char handle[B2A_IPC_HANDLE_BYTES]) {
  static_assert(sizeof(cudaIpcMemHandle_t) == B2A_IPC_HANDLE_BYTES, "CUDA IPC handle size");
  if (!c || !device_ptr || !handle) return fail(c, B2A_E_BAD_ARG, "b2a_ipc_export: bad argument");
  DeviceGuard g(c->device);
  cudaIpcMemHandle_t h;
  int rc = cu(c, cudaIpcGetMemHandle(&h, device_ptr), "cudaIpcGetMemHandle");
  if (rc == B2A_OK) std::memcpy(handle, &h, sizeof(h));
  return rc;
}

int b2a_ipc_open(b2a_ctx* c, const unsigned char handle[B2A_IPC_HANDLE_BYTES], void** peer_ptr) {
  if (!c || !handle || !peer_ptr) return fail(c, B2A_E_BAD_ARG, "b2a_ipc_open: bad argument");
  DeviceGuard g(c->device);
  cudaIpcMemHandle_t h;
  std::memcpy(&h, handle, sizeof(h));
  // (peer access between the two devices is enabled by the runtime as part of the open)
  return cu(c, cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess), "cudaIpcOpenMemHandle");
}

int b2a_ipc_close(b2a_ctx* c, void* peer_ptr) {
  if (!c || !peer_ptr) return fail(c, B2A_E_BAD_ARG, "b2a_ipc_close: bad argument");
  DeviceGuard g(c->device);
  return cu(c, cudaIpcCloseMemHandle(peer_ptr), "cudaIpcCloseMemHandle");
}

int b2a_memcpy_d2h(b2a_ctx* c, void* host_dst, const void* device_src, uint64_t bytes) {
  if (!c || !host_dst || !device_src) return fail(c, B2A_E_BAD_ARG, "b2a_memcpy_d2h: bad argument");
  DeviceGuard g(c->device);
  int rc = cu(c, cudaMemcpyAsync(host_dst, device_src, bytes, cudaMemcpyDeviceToHost, c->stream), "cudaMemcpyAsync");
  return rc != B2A_OK ? rc : cu(c, cudaStreamSynchronize(c->stream), "sync");
}

int b2a_debug_dyn_tiles(int on) {
  frontend_dyn_tiles_enable(on);
  return B2A_OK;
}

int b2a_debug_whisper_tc(int on) {
  tc_whisper_enable(on);
  return B2A_OK;
}

int b2a_debug_wpf1920(int on) {
  wpf1920_enable(on);
  return B2A_OK;
}

int b2a_debug_tc_power_buffer(void* device_ptr) {
  tc_debug_set_power_buffer(static_cast<float*>(device_ptr));
  return B2A_OK;
}

int b2a_ctx_enable_timing(b2a_ctx* c, int on) {
  if (!c) return B2A_E_BAD_ARG;
  c->timing = on != 0;
  c->timed = false;
  return B2A_OK;
}

int b2a_ctx_last_kernel_ms(b2a_ctx* c, float* ms) {
  if (!c || !ms) return B2A_E_BAD_ARG;
  if (!c->timed) return fail(c, B2A_E_BAD_ARG, "no timed B2A_DEVICE call yet");
  cudaError_t e = cudaEventSynchronize(c->ev_t1);
  if (e != cudaSuccess) return cu(c, e, "event sync");
  return cu(c, cudaEventElapsedTime(ms, c->ev_t0, c->ev_t1), "elapsed");
}

// ---- front ends -----------------------------------------------------------------------------------
int b2a_pad_or_trim(b2a_ctx* c, const float* x, int64_t batch, int64_t n_samples, int64_t length, float* out, int space) {
  int rc = check_common(c, x, out, batch, n_samples);
  if (rc != B2A_OK) return rc;
  if (length <= 0) return fail(c, B2A_E_BAD_ARG, "length must be positive");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  Body body = [&](const float* d_in, const float*, float* d_out, float*, int64_t n, int) -> int {
    int launches = 0;
    std::string err;
    int r = launch_pad_or_trim(d_in, d_out, n, n_samples, length, c->stream, &launches, &err);
    c->launches += launches;
    if (r != B2A_OK) c->err = err;
    return r;
  };
  return run_batched(c, space, batch, x, size_t(n_samples), nullptr, 0, out, size_t(length), nullptr, 0, body);
}

int b2a_reflect_pad(b2a_ctx* c, const float* x, int64_t batch, int64_t n_samples, int64_t padding, float* out, int space) {
  int rc = check_common(c, x, out, batch, n_samples);
  if (rc != B2A_OK) return rc;
  if (padding < 0) return fail(c, B2A_E_BAD_ARG, "padding must be >= 0");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  Body body = [&](const float* d_in, const float*, float* d_out, float*, int64_t n, int) -> int {
    int launches = 0;
    std::string err;
    int r = launch_reflect_pad(d_in, d_out, n, n_samples, padding, c->stream, &launches, &err);
    c->launches += launches;
    if (r != B2A_OK) c->err = err;
    return r;
  };
  return run_batched(c, space, batch, x, size_t(n_samples), nullptr, 0, out, size_t(n_samples + 2 * padding), nullptr, 0, body);
}

int b2a_s3tokenizer_gather_segments(b2a_ctx* c, const float* mel, int64_t batch, int n_mels, int64_t t_max, int64_t n_segments,
                                    const int32_t* batch_idx, const int32_t* start, const int32_t* length, int64_t window, float* out,
                                    int space) {
  int rc = check_common(c, mel, out, batch, t_max);
  if (rc != B2A_OK) return rc;
  if (!batch_idx || !start || !length || n_mels <= 0 || n_mels > 65535 || window <= 0 || window > 0x7fffffff || n_segments <= 0 ||
      n_segments > 0x7fffffff)
    return fail(c, B2A_E_BAD_ARG, "bad segment arguments");
  std::vector<int> seg(3 * size_t(n_segments));
  for (int64_t s = 0; s < n_segments; ++s) {
    if (batch_idx[s] < 0 || batch_idx[s] >= batch || start[s] < 0 || length[s] < 0 || length[s] > window ||
        int64_t(start[s]) + length[s] > t_max)
      return fail(c, B2A_E_BAD_ARG, "segment outside the mel");
    seg[3 * s] = batch_idx[s]; seg[3 * s + 1] = start[s]; seg[3 * s + 2] = length[s];
  }
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  // segments index clips freely, so the batch is not streamed clip by clip: host buffers are copied whole
  const size_t in_bytes = sizeof(float) * size_t(batch) * n_mels * size_t(t_max);
  const size_t out_bytes = sizeof(float) * size_t(n_segments) * n_mels * size_t(window);
  const float* d_in = mel;
  float* d_out = out;
  cudaError_t e;
  if (space == B2A_HOST) {
    if ((rc = ensure(c, c->in[0][0], in_bytes)) != B2A_OK) return rc;
    if ((rc = ensure(c, c->out[0][0], out_bytes)) != B2A_OK) return rc;
    if ((e = cudaMemcpyAsync(c->in[0][0].p, mel, in_bytes, cudaMemcpyHostToDevice, c->stream)) != cudaSuccess) return cu(c, e, "H2D copy");
    d_in = static_cast<const float*>(c->in[0][0].p);
    d_out = static_cast<float*>(c->out[0][0].p);
  } else if (space != B2A_DEVICE) {
    return fail(c, B2A_E_BAD_ARG, "space must be B2A_HOST or B2A_DEVICE");
  }
  if ((rc = ensure(c, c->scratch[0][0], sizeof(int) * seg.size())) != B2A_OK) return rc;
  if ((e = cudaMemcpyAsync(c->scratch[0][0].p, seg.data(), sizeof(int) * seg.size(), cudaMemcpyHostToDevice, c->stream)) != cudaSuccess)
    return cu(c, e, "segment upload");
  int launches = 0;
  std::string err;
  rc = launch_mel_windows(d_in, d_out, n_mels, t_max, static_cast<const int*>(c->scratch[0][0].p), int(n_segments), int(window), c->stream,
                          &launches, &err);
  c->launches += launches;
  if (rc != B2A_OK) return fail(c, rc, err);
  if (space == B2A_HOST) {
    if ((e = cudaMemcpyAsync(out, d_out, out_bytes, cudaMemcpyDeviceToHost, c->stream)) != cudaSuccess) return cu(c, e, "D2H copy");
    if ((e = cudaStreamSynchronize(c->stream)) != cudaSuccess) return cu(c, e, "sync");
  }
  return B2A_OK;
}

int b2a_resample_linear(b2a_ctx* c, const float* x, int64_t batch, int64_t n_samples, int from_rate, int to_rate, float* out, int space) {
  int rc = check_common(c, x, out, batch, n_samples);
  if (rc != B2A_OK) return rc;
  if (from_rate <= 0 || to_rate <= 0) return fail(c, B2A_E_BAD_ARG, "sample rates must be positive");
  if (n_samples >= (1LL << 31)) return fail(c, B2A_E_BAD_ARG, "clip too long");   // the reference indexes with int32
  const int64_t new_t = b2a_resample_linear_length(n_samples, from_rate, to_rate);
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  const float step = float(n_samples) / float(new_t);
  const float hi_clip = float(n_samples) - 1.001f;
  Body body = [&](const float* d_in, const float*, float* d_out, float*, int64_t n, int) -> int {
    int launches = 0;
    std::string err;
    int r;
    if (from_rate == to_rate) {   // identity (CosyVoice2TTS.swift:734-736)
      r = cu(c, cudaMemcpyAsync(d_out, d_in, sizeof(float) * size_t(n) * n_samples, cudaMemcpyDeviceToDevice, c->stream), "copy");
    } else {
      r = launch_resample_linear(d_in, d_out, n, n_samples, new_t, step, hi_clip, c->stream, &launches, &err);
      if (r != B2A_OK) c->err = err;
    }
    c->launches += launches;
    return r;
  };
  return run_batched(c, space, batch, x, size_t(n_samples), nullptr, 0, out, size_t(new_t), nullptr, 0, body);
}

int b2a_resample_poly(b2a_ctx* c, const float* x, int64_t batch, int64_t n_samples, int from_rate, int to_rate, float* out, int space) {
  int rc = check_common(c, x, out, batch, n_samples);
  if (rc != B2A_OK) return rc;
  if (from_rate <= 0 || to_rate <= 0) return fail(c, B2A_E_BAD_ARG, "sample rates must be positive");
  const int64_t new_t = b2a_resample_poly_length(n_samples, from_rate, to_rate);
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  int up = 1, down = 1;
  int64_t pre = 0;
  std::vector<float> h;
  if (from_rate != to_rate) {
    const int64_t taps = b2a_resample_poly_filter(n_samples, from_rate, to_rate, nullptr, 0, &up, &down, &pre);
    if (taps <= 0 || taps > 0x7fffffff) return fail(c, B2A_E_UNSUPPORTED, "resample_poly: rate ratio out of range (up, down <= 4096 after reduction)");
    h.resize(size_t(taps));
    b2a_resample_poly_filter(n_samples, from_rate, to_rate, h.data(), taps, &up, &down, &pre);
    if ((rc = ensure(c, c->scratch[0][4], sizeof(float) * h.size())) != B2A_OK) return rc;
    // (synchronous upload: the design depends on the rate pair only in practice, the table is small)
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e == cudaSuccess) e = cudaMemcpy(c->scratch[0][4].p, h.data(), sizeof(float) * h.size(), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) return cu(c, e, "filter upload");
  }
  Body body = [&](const float* d_in, const float*, float* d_out, float*, int64_t n, int) -> int {
    int launches = 0;
    std::string err;
    int r;
    if (from_rate == to_rate) {
      r = cu(c, cudaMemcpyAsync(d_out, d_in, sizeof(float) * size_t(n) * n_samples, cudaMemcpyDeviceToDevice, c->stream), "copy");
    } else {
      r = launch_resample_poly(d_in, d_out, static_cast<const float*>(c->scratch[0][4].p), n, n_samples, new_t, up, down, pre, int(h.size()), c->stream,
                               &launches, &err);
      if (r != B2A_OK) c->err = err;
    }
    c->launches += launches;
    return r;
  };
  return run_batched(c, space, batch, x, size_t(n_samples), nullptr, 0, out, size_t(new_t), nullptr, 0, body);
}

int b2a_whisper_mel_segment_f16(b2a_ctx* c, const float* mel, int64_t batch, int64_t n_frames, int n_mels, const int64_t* seek,
                                const int64_t* content_frames, int64_t length, void* out_f16, int space) {
  int rc = check_common(c, mel, out_f16, batch, n_frames);
  if (rc != B2A_OK) return rc;
  if (!seek || !content_frames) return fail(c, B2A_E_BAD_ARG, "null seek / content_frames");
  if (n_mels <= 0 || length <= 0 || length > 0x7fffffff || ((length * n_mels) & 1)) return fail(c, B2A_E_BAD_ARG, "bad n_mels / length");
  for (int64_t b = 0; b < batch; ++b)
    if (seek[b] < 0 || seek[b] > n_frames) return fail(c, B2A_E_BAD_ARG, "seek outside the mel");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  int64_t done = 0;   // clips already handed to the kernel (run_batched walks the batch in order)
  Body body = [&](const float* d_in, const float*, float* d_out, float*, int64_t n, int slot) -> int {
    int r;
    if ((r = ensure(c, c->scratch[slot][0], sizeof(long long) * 2 * size_t(n))) != B2A_OK) return r;
    std::vector<long long> sc(2 * size_t(n));
    for (int64_t i = 0; i < n; ++i) {
      sc[2 * i] = seek[done + i];
      sc[2 * i + 1] = content_frames[done + i];
    }
    done += n;
    // pageable-source copy: the host vector may be released as soon as the call returns
    cudaError_t e = cudaMemcpyAsync(c->scratch[slot][0].p, sc.data(), sizeof(long long) * sc.size(), cudaMemcpyHostToDevice, c->stream);
    if (e != cudaSuccess) return cu(c, e, "seek upload");
    int launches = 0;
    std::string err;
    r = launch_mel_segment_f16(d_in, d_out, n, n_frames, n_mels, static_cast<const long long*>(c->scratch[slot][0].p), int(length),
                               c->stream, &launches, &err);
    c->launches += launches;
    if (r != B2A_OK) c->err = err;
    return r;
  };
  // the fp16 output travels through the float-typed pipeline as (length * n_mels / 2) 32-bit words per clip
  return run_batched(c, space, batch, mel, size_t(n_frames) * n_mels, nullptr, 0, static_cast<float*>(out_f16),
                     size_t(length) * n_mels / 2, nullptr, 0, body);
}

static int whisper_like(b2a_ctx* c, const void* audio, int64_t batch, int64_t n_samples, int n_mels, int64_t padding, void* out,
                        int space, bool chatterbox, const int64_t* lengths = nullptr, int64_t* out_rows = nullptr, bool in_i16 = false,
                        bool out_f16 = false) {
  int rc = check_common(c, audio, out, batch, n_samples);
  if (rc != B2A_OK) return rc;
  if (padding < 0) return fail(c, B2A_E_BAD_ARG, "padding must be >= 0");
  const int64_t frames = b2a_whisper_num_frames(n_samples, padding);
  if (frames <= 0) return fail(c, B2A_E_TOO_SHORT, "Input is too short for STFT");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  Preset p;
  // Whisper: symmetric Hann (WhisperAudio.swift:89); S3Tokenizer/Chatterbox: hanningWindow(401)[0..<400] (S3TokenizerUtils.swift:172)
  p.window = chatterbox ? &window_of(c, B2A_WIN_HANNING, 400, 400, true) : &window_of(c, B2A_WIN_WHISPER_HANN, 400, 400, false);
  p.zero_tail = padding;
  // Whisper passes fMax 8000 explicitly (WhisperAudio.swift:113-119); Chatterbox leaves it nil = sr/2 (S3TokenizerUtils.swift:190-194)
  if ((rc = slaney_bank(c, 16000, 400, n_mels, 0.0f, chatterbox ? -1.0f : 8000.0f, &p.bank)) != B2A_OK) return rc;
  p.log_mode = LOG_LOG10;
  p.log_floor = 1e-10f;
  p.whisper_norm = 1;
  p.out_mode = chatterbox ? OUT_MT : OUT_TM;
  p.n_frames = frames;  // the last STFT frame is dropped (WhisperAudio.swift:105, S3TokenizerUtils.swift:184)
  p.in_i16 = in_i16;
  p.out_f16 = out_f16;
  Ragged rg{lengths, [&](int64_t len) { return b2a_whisper_num_frames(len, padding); }, out_rows};
  return run_preset(c, p, audio, batch, n_samples, out, space, lengths ? &rg : nullptr);
}

int b2a_whisper_log_mel_spectrogram(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, int n_mels, int64_t padding,
                                    float* out, int space) {
  return whisper_like(c, audio, batch, n_samples, n_mels, padding, out, space, false);
}

int b2a_whisper_log_mel_spectrogram_f16(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, int n_mels, int64_t padding,
                                        void* out_f16, int space) {
  return whisper_like(c, audio, batch, n_samples, n_mels, padding, out_f16, space, false, nullptr, nullptr, false, true);
}

int b2a_whisper_log_mel_spectrogram_pcm16(b2a_ctx* c, const int16_t* audio, int64_t batch, int64_t n_samples, int n_mels, int64_t padding,
                                          int out_is_f16, void* out, int space) {
  return whisper_like(c, audio, batch, n_samples, n_mels, padding, out, space, false, nullptr, nullptr, true, out_is_f16 != 0);
}

int b2a_whisper_log_mel_spectrogram_f16_ragged(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                               int n_mels, int64_t padding, void* out_f16, int64_t* out_frames, int space) {
  if (!lengths) return fail(c, B2A_E_BAD_ARG, "lengths must not be NULL");
  return whisper_like(c, audio, batch, n_samples, n_mels, padding, out_f16, space, false, lengths, out_frames, false, true);
}

int b2a_log_mel_spectrogram_chatterbox(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, int n_mels, int64_t padding,
                                       float* out, int space) {
  return whisper_like(c, audio, batch, n_samples, n_mels, padding, out, space, true);
}

int b2a_whisper_log_mel_spectrogram_ragged(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                           int n_mels, int64_t padding, float* out, int64_t* out_frames, int space) {
  if (!lengths) return fail(c, B2A_E_BAD_ARG, "lengths must not be NULL");
  return whisper_like(c, audio, batch, n_samples, n_mels, padding, out, space, false, lengths, out_frames);
}

int b2a_log_mel_spectrogram_chatterbox_ragged(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                              int n_mels, int64_t padding, float* out, int64_t* out_frames, int space) {
  if (!lengths) return fail(c, B2A_E_BAD_ARG, "lengths must not be NULL");
  return whisper_like(c, audio, batch, n_samples, n_mels, padding, out, space, true, lengths, out_frames);
}

static int funasr_common(b2a_ctx* c, const void* audio, int64_t batch, int64_t n_samples, int n_mels, int lfr_m, int lfr_n,
                         int lfr, int norm, float* out, int space, const int64_t* lengths = nullptr, int64_t* out_rows = nullptr,
                         bool in_i16 = false) {
  int rc = check_common(c, audio, out, batch, n_samples);
  if (rc != B2A_OK) return rc;
  if (n_mels <= 0) return fail(c, B2A_E_BAD_ARG, "n_mels must be positive");
  if (lfr && (lfr_m <= 0 || lfr_n <= 0)) return fail(c, B2A_E_BAD_ARG, "lfr_m and lfr_n must be positive");
  const int64_t frames = b2a_funasr_num_frames(n_samples);
  if (frames <= 0) return fail(c, B2A_E_TOO_SHORT, "Input is too short for STFT");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  Preset p;
  p.window = &window_of(c, B2A_WIN_HAMMING, 400, 400, false);
  if ((rc = cached_bank(c, key_of("funasr", {16000.0, 400.0, double(n_mels)}), 400, n_mels, 200, false,
                        [&](float* d) { return mel_filters_funasr(16000, 400, n_mels, d); }, &p.bank)) != B2A_OK)
    return rc;
  p.log_mode = LOG_LN;
  p.log_floor = 1e-10f;
  p.n_frames = frames;
  if (lfr) {
    p.out_mode = OUT_LFR;
    p.lfr_m = lfr_m;
    p.lfr_n = lfr_n;
    p.lfr_rows = b2a_lfr_num_rows(frames, lfr_n);
    p.post_cmvn = norm;
  }
  p.in_i16 = in_i16;
  Ragged rg{lengths, [](int64_t len) { return b2a_funasr_num_frames(len); }, out_rows};
  return run_preset(c, p, audio, batch, n_samples, out, space, lengths ? &rg : nullptr);
}

int b2a_funasr_log_mel_spectrogram(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, int n_mels, float* out, int space) {
  return funasr_common(c, audio, batch, n_samples, n_mels, 0, 0, 0, 0, out, space);
}

int b2a_funasr_preprocess_audio(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, int n_mels, int lfr_m, int lfr_n,
                                int apply_normalization, float* out, int space) {
  return funasr_common(c, audio, batch, n_samples, n_mels, lfr_m, lfr_n, 1, apply_normalization != 0, out, space);
}

int b2a_funasr_preprocess_audio_pcm16(b2a_ctx* c, const int16_t* audio, int64_t batch, int64_t n_samples, int n_mels, int lfr_m, int lfr_n,
                                      int apply_normalization, float* out, int space) {
  return funasr_common(c, audio, batch, n_samples, n_mels, lfr_m, lfr_n, 1, apply_normalization != 0, out, space, nullptr, nullptr, true);
}

int b2a_funasr_log_mel_spectrogram_ragged(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths, int n_mels,
                                          float* out, int64_t* out_frames, int space) {
  if (!lengths) return fail(c, B2A_E_BAD_ARG, "lengths must not be NULL");
  return funasr_common(c, audio, batch, n_samples, n_mels, 0, 0, 0, 0, out, space, lengths, out_frames);
}

int b2a_funasr_preprocess_audio_ragged(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths, int n_mels,
                                       int lfr_m, int lfr_n, int apply_normalization, float* out, int64_t* out_rows, int space) {
  if (!lengths) return fail(c, B2A_E_BAD_ARG, "lengths must not be NULL");
  return funasr_common(c, audio, batch, n_samples, n_mels, lfr_m, lfr_n, 1, apply_normalization != 0, out, space, lengths, out_rows);
}

int b2a_apply_lfr(b2a_ctx* c, const float* features, int64_t batch, int64_t n_frames, int n_mels, int lfr_m, int lfr_n, float* out,
                  int space) {
  int rc = check_common(c, features, out, batch, n_frames);
  if (rc != B2A_OK) return rc;
  if (n_mels <= 0 || lfr_m <= 0 || lfr_n <= 0) return fail(c, B2A_E_BAD_ARG, "n_mels, lfr_m, lfr_n must be positive");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  const int64_t rows = b2a_lfr_num_rows(n_frames, lfr_n);
  Body body = [&](const float* d_in, const float*, float* d_out, float*, int64_t n, int) -> int {
    int launches = 0;
    std::string err;
    int r = launch_lfr(d_in, d_out, n, n_frames, n_mels, lfr_m, lfr_n, c->stream, &launches, &err);
    c->launches += launches;
    if (r != B2A_OK) c->err = err;
    return r;
  };
  return run_batched(c, space, batch, features, size_t(n_frames) * n_mels, nullptr, 0, out, size_t(rows) * lfr_m * n_mels, nullptr, 0, body);
}

int b2a_apply_cmvn(b2a_ctx* c, const float* features, int64_t batch, int64_t n_rows, int dim, const float* cmvn_mean,
                   const float* cmvn_istd, float* out, int space) {
  int rc = check_common(c, features, out, batch, n_rows);
  if (rc != B2A_OK) return rc;
  if (dim <= 0) return fail(c, B2A_E_BAD_ARG, "dim must be positive");
  if ((cmvn_mean == nullptr) != (cmvn_istd == nullptr)) return fail(c, B2A_E_BAD_ARG, "cmvn_mean and cmvn_istd must both be given or both be NULL");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  const float* d_mean = nullptr;
  const float* d_istd = nullptr;
  if (cmvn_mean) {
    // statistics always come from the host side of the boundary (model config); upload per call
    if ((rc = ensure(c, c->scratch[0][2], sizeof(float) * 2 * size_t(dim))) != B2A_OK) return rc;
    float* d = static_cast<float*>(c->scratch[0][2].p);
    cudaError_t e;
    if (space == B2A_DEVICE) {
      if ((e = cudaMemcpyAsync(d, cmvn_mean, sizeof(float) * dim, cudaMemcpyDeviceToDevice, c->stream)) != cudaSuccess) return cu(c, e, "copy");
      if ((e = cudaMemcpyAsync(d + dim, cmvn_istd, sizeof(float) * dim, cudaMemcpyDeviceToDevice, c->stream)) != cudaSuccess) return cu(c, e, "copy");
    } else {
      if ((e = cudaMemcpy(d, cmvn_mean, sizeof(float) * dim, cudaMemcpyHostToDevice)) != cudaSuccess) return cu(c, e, "copy");
      if ((e = cudaMemcpy(d + dim, cmvn_istd, sizeof(float) * dim, cudaMemcpyHostToDevice)) != cudaSuccess) return cu(c, e, "copy");
    }
    d_mean = d;
    d_istd = d + dim;
  }
  Body body = [&](const float* d_in, const float*, float* d_out, float*, int64_t n, int) -> int {
    int launches = 0;
    std::string err;
    int r = launch_cmvn(d_in, d_out, n, n_rows, dim, d_mean, d_istd, c->stream, &launches, &err);
    c->launches += launches;
    if (r != B2A_OK) c->err = err;
    return r;
  };
  return run_batched(c, space, batch, features, size_t(n_rows) * dim, nullptr, 0, out, size_t(n_rows) * dim, nullptr, 0, body);
}

static int kaldi_common(b2a_ctx* c, const void* audio, int64_t batch, int64_t n_samples, int sample_rate, int num_mel_bins,
                        float frame_length_ms, float frame_shift_ms, int mean_norm, float* out, int space, const int64_t* lengths,
                        int64_t* out_rows, bool in_i16 = false) {
  int rc = check_common(c, audio, out, batch, n_samples);
  if (rc != B2A_OK) return rc;
  if (sample_rate <= 0 || num_mel_bins <= 0) return fail(c, B2A_E_BAD_ARG, "sample_rate and num_mel_bins must be positive");
  const int win_length = int(float(sample_rate) * frame_length_ms / 1000.0f);  // CAMPPlus.swift:40-42
  const int hop = int(float(sample_rate) * frame_shift_ms / 1000.0f);
  const int n_fft = b2a_next_power_of_2(win_length);
  if (!frontend_plan_exists(n_fft, hop, win_length)) return fail(c, B2A_E_UNSUPPORTED, "Kaldi fbank is built for 25 ms / 10 ms at 16 kHz (win 400, hop 160, n_fft 512)");
  const int64_t frames = b2a_kaldi_num_frames(n_samples, win_length, hop);
  if (frames <= 0) return fail(c, B2A_E_TOO_SHORT, "signal shorter than one analysis window");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  Preset p;
  p.n_fft = n_fft; p.hop = hop; p.win_len = win_length;
  p.window = &window_of(c, B2A_WIN_POVEY, win_length, n_fft, false);
  p.pad_mode = PAD_NONE;
  p.pad_left = 0;
  p.pre_mode = PRE_KALDI;
  if ((rc = cached_bank(c, key_of("htkint", {double(sample_rate), double(n_fft), double(num_mel_bins), 20.0, double(sample_rate) / 2}),
                        n_fft, num_mel_bins, n_fft / 2 + 1, true,
                        [&](float* d) { return mel_filters_htk_int(sample_rate, n_fft, num_mel_bins, 20.0f, float(sample_rate) / 2, d); },
                        &p.bank)) != B2A_OK)
    return rc;
  p.log_mode = LOG_LN;
  p.log_floor = 1.1920929e-07f;
  p.n_frames = frames;
  p.post_mean_norm = mean_norm != 0;
  p.in_i16 = in_i16;
  Ragged rg{lengths, [&](int64_t len) { return b2a_kaldi_num_frames(len, win_length, hop); }, out_rows};
  return run_preset(c, p, audio, batch, n_samples, out, space, lengths ? &rg : nullptr);
}

int b2a_kaldi_fbank_campplus(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, int sample_rate, int num_mel_bins,
                             float frame_length_ms, float frame_shift_ms, int mean_norm, float* out, int space) {
  return kaldi_common(c, audio, batch, n_samples, sample_rate, num_mel_bins, frame_length_ms, frame_shift_ms, mean_norm, out, space, nullptr, nullptr);
}

int b2a_kaldi_fbank_campplus_pcm16(b2a_ctx* c, const int16_t* audio, int64_t batch, int64_t n_samples, int sample_rate, int num_mel_bins,
                                   float frame_length_ms, float frame_shift_ms, int mean_norm, float* out, int space) {
  return kaldi_common(c, audio, batch, n_samples, sample_rate, num_mel_bins, frame_length_ms, frame_shift_ms, mean_norm, out, space, nullptr, nullptr, true);
}

int b2a_kaldi_fbank_campplus_ragged(b2a_ctx* c, const float* audio, int64_t batch, int64_t n_samples, const int64_t* lengths, int sample_rate,
                                    int num_mel_bins, float frame_length_ms, float frame_shift_ms, int mean_norm, float* out,
                                    int64_t* out_frames, int space) {
  if (!lengths) return fail(c, B2A_E_BAD_ARG, "lengths must not be NULL");
  return kaldi_common(c, audio, batch, n_samples, sample_rate, num_mel_bins, frame_length_ms, frame_shift_ms, mean_norm, out, space, lengths, out_frames);
}

static int s3gen_common(b2a_ctx* c, const void* y, int64_t batch, int64_t n_samples, int n_fft, int num_mels, int sampling_rate,
                        int hop_size, int win_size, int fmin, int fmax, float* out, int space, const int64_t* lengths, int64_t* out_rows,
                        bool in_i16 = false) {
  int rc = check_common(c, y, out, batch, n_samples);
  if (rc != B2A_OK) return rc;
  if (win_size > n_fft || win_size <= 0) return fail(c, B2A_E_BAD_ARG, "win_size must be in (0, n_fft]");
  if (!frontend_plan_exists(n_fft, hop_size, n_fft) && (!generic_stft_supported(n_fft, hop_size) || lengths))
    return fail(c, B2A_E_UNSUPPORTED, "s3gen mel: n_fft must lie in [2, 8192] (ragged batches: 1920 / 480 and 400 / 160 only)");
  const int64_t frames = b2a_s3gen_num_frames(n_samples, n_fft, hop_size);
  if (frames <= 0) return fail(c, B2A_E_TOO_SHORT, "Input is too short for STFT");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  Preset p;
  p.n_fft = n_fft; p.hop = hop_size; p.win_len = n_fft;
  p.window = &window_of(c, B2A_WIN_HANNING, win_size, n_fft, true);
  p.pad_mode = PAD_REFLECT;
  const int64_t pad = (n_fft - hop_size) / 2;
  p.pad_left = std::min<int64_t>(pad, n_samples - 1);  // reflectPad2D truncates, no loop (S3GenMel.swift:17-25)
  p.spec_mode = SPEC_MAGNITUDE;
  if ((rc = slaney_bank(c, sampling_rate, n_fft, num_mels, float(fmin), float(fmax), &p.bank)) != B2A_OK) return rc;
  p.log_mode = LOG_LN;
  p.log_floor = 1e-5f;
  p.out_mode = OUT_MT;
  p.n_frames = frames;
  if (lengths)   // reflectPad2D truncates the pad for clips of <= pad samples: one pad_left per launch, so those stay out of ragged batches
    for (int64_t b = 0; b < batch; ++b)
      if (lengths[b] <= pad) return fail(c, B2A_E_UNSUPPORTED, "ragged s3gen mel needs every clip longer than (n_fft - hop) / 2 samples");
  p.in_i16 = in_i16;
  Ragged rg{lengths, [&](int64_t len) { return b2a_s3gen_num_frames(len, n_fft, hop_size); }, out_rows};
  return run_preset(c, p, y, batch, n_samples, out, space, lengths ? &rg : nullptr);
}

int b2a_s3gen_mel_spectrogram(b2a_ctx* c, const float* y, int64_t batch, int64_t n_samples, int n_fft, int num_mels, int sampling_rate,
                              int hop_size, int win_size, int fmin, int fmax, float* out, int space) {
  return s3gen_common(c, y, batch, n_samples, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, out, space, nullptr, nullptr);
}

int b2a_s3gen_mel_spectrogram_pcm16(b2a_ctx* c, const int16_t* y, int64_t batch, int64_t n_samples, int n_fft, int num_mels, int sampling_rate,
                                    int hop_size, int win_size, int fmin, int fmax, float* out, int space) {
  if (!frontend_plan_exists(n_fft, hop_size, n_fft)) return fail(c, B2A_E_UNSUPPORTED, "16-bit PCM input is built for the tuned (n_fft, hop) = (1920, 480) / (400, 160)");
  return s3gen_common(c, y, batch, n_samples, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, out, space, nullptr, nullptr, true);
}

int b2a_s3gen_mel_spectrogram_ragged(b2a_ctx* c, const float* y, int64_t batch, int64_t n_samples, const int64_t* lengths, int n_fft,
                                     int num_mels, int sampling_rate, int hop_size, int win_size, int fmin, int fmax, float* out,
                                     int64_t* out_frames, int space) {
  if (!lengths) return fail(c, B2A_E_BAD_ARG, "lengths must not be NULL");
  return s3gen_common(c, y, batch, n_samples, n_fft, num_mels, sampling_rate, hop_size, win_size, fmin, fmax, out, space, lengths, out_frames);
}

static int voice_encoder_common(b2a_ctx* c, const float* wav, int64_t batch, int64_t n_samples, const b2a_voice_enc_config* cfg_in,
                                float* out, int space, const int64_t* lengths, int64_t* out_rows) {
  int rc = check_common(c, wav, out, batch, n_samples);
  if (rc != B2A_OK) return rc;
  b2a_voice_enc_config cfg;
  if (cfg_in) cfg = *cfg_in;
  else b2a_voice_enc_config_default(&cfg);
  if (cfg.win_size > cfg.n_fft || cfg.win_size <= 0) return fail(c, B2A_E_BAD_ARG, "win_size must be in (0, n_fft]");
  if (!frontend_plan_exists(cfg.n_fft, cfg.hop_size, cfg.n_fft) && (!generic_stft_supported(cfg.n_fft, cfg.hop_size) || lengths))
    return fail(c, B2A_E_UNSUPPORTED, "voice-encoder mel: n_fft must lie in [2, 8192] (ragged batches: n_fft 400 / hop 160 only)");
  if (cfg.mel_power != 1.0f && cfg.mel_power != 2.0f) return fail(c, B2A_E_UNSUPPORTED, "mel_power 1.0 and 2.0 are built");
  const int64_t frames = b2a_stft_num_frames(n_samples, cfg.n_fft, cfg.hop_size, 1);
  if (frames <= 0) return fail(c, B2A_E_TOO_SHORT, "Input is too short for STFT");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  Preset p;
  p.n_fft = cfg.n_fft; p.hop = cfg.hop_size; p.win_len = cfg.n_fft;
  p.window = &window_of(c, B2A_WIN_HANNING, cfg.win_size, cfg.n_fft, true);
  p.pad_left = cfg.n_fft / 2;
  p.spec_mode = cfg.mel_power == 2.0f ? SPEC_POWER : SPEC_MAGNITUDE;
  if ((rc = slaney_bank(c, cfg.sample_rate, cfg.n_fft, cfg.num_mels, float(cfg.fmin), float(cfg.fmax), &p.bank)) != B2A_OK) return rc;
  if (cfg.mel_type_db) {
    p.log_mode = LOG_DB20;
    p.log_floor = cfg.stft_magnitude_min;
  }
  if (cfg.normalized_mels) {  // VoiceEncoderMelspec.swift:61-65
    const float min_level_db = 20.0f * log10f(cfg.stft_magnitude_min);
    p.post_affine = 1;
    p.post_sub = min_level_db;
    p.post_div = -min_level_db + 15.0f;
  }
  p.out_mode = OUT_MT;
  p.n_frames = frames;
  Ragged rg{lengths, [&](int64_t len) { return b2a_stft_num_frames(len, cfg.n_fft, cfg.hop_size, 1); }, out_rows};
  return run_preset(c, p, wav, batch, n_samples, out, space, lengths ? &rg : nullptr);
}

int b2a_voice_encoder_melspectrogram(b2a_ctx* c, const float* wav, int64_t batch, int64_t n_samples, const b2a_voice_enc_config* cfg_in,
                                     float* out, int space) {
  return voice_encoder_common(c, wav, batch, n_samples, cfg_in, out, space, nullptr, nullptr);
}

int b2a_voice_encoder_melspectrogram_ragged(b2a_ctx* c, const float* wav, int64_t batch, int64_t n_samples, const int64_t* lengths,
                                            const b2a_voice_enc_config* cfg_in, float* out, int64_t* out_frames, int space) {
  if (!lengths) return fail(c, B2A_E_BAD_ARG, "lengths must not be NULL");
  return voice_encoder_common(c, wav, batch, n_samples, cfg_in, out, space, lengths, out_frames);
}

int b2a_stft(b2a_ctx* c, const float* x, int64_t batch, int64_t n_samples, const float* window, int win_len, int n_fft, int hop,
             int center, float* out_complex, int space) {
  int rc = check_common(c, x, out_complex, batch, n_samples);
  if (rc != B2A_OK) return rc;
  if (!window || win_len <= 0 || win_len > n_fft) return fail(c, B2A_E_BAD_ARG, "window must have 1..n_fft taps");
  // tuned plans: (n_fft, hop) = (400, 160), (1920, 480), (512, 160) with <= 400 taps; every other size runs the generic kernel
  int plan_win = n_fft;
  if (!frontend_plan_exists(n_fft, hop, plan_win)) {
    if (n_fft == 512 && win_len <= 400 && frontend_plan_exists(n_fft, hop, 400)) plan_win = 400;
    else if (!generic_stft_supported(n_fft, hop)) return fail(c, B2A_E_UNSUPPORTED, "stft: n_fft must lie in [2, 8192] and hop must be positive");
  }
  const int64_t frames = b2a_stft_num_frames(n_samples, n_fft, hop, center);
  if (frames <= 0) return fail(c, B2A_E_TOO_SHORT, "Input is too short for STFT");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  std::vector<float> w(window, window + win_len);
  w.resize(n_fft, 0.0f);
  Preset p;
  p.n_fft = n_fft; p.hop = hop; p.win_len = plan_win;
  p.window = &w;
  p.pad_mode = center ? PAD_REFLECT : PAD_NONE;
  p.pad_left = center ? n_fft / 2 : 0;
  p.out_mode = OUT_COMPLEX;
  p.n_frames = frames;
  return run_preset(c, p, x, batch, n_samples, out_complex, space);
}

// ---- vocoder --------------------------------------------------------------------------------------
static int small_stft_common(b2a_ctx* c, const float* x, int64_t batch, int64_t n_samples, int n_fft, int hop, const float* window,
                             int pad_mode, int out_kind, float* o0, float* o1, int space) {
  int rc = check_common(c, x, o0, batch, n_samples);
  if (rc != B2A_OK) return rc;
  if (!o1 || !window) return fail(c, B2A_E_BAD_ARG, "null buffer");
  const int64_t frames = b2a_vocoder_stft_num_frames(n_samples, n_fft, hop);
  if (frames <= 0) return fail(c, B2A_E_TOO_SHORT, "Input is too short");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  const size_t per = size_t(n_fft / 2 + 1) * frames;
  Body body = [&](const float* d_in, const float*, float* d0, float* d1, int64_t n, int) -> int {
    SmallStftArgs a;
    a.n_fft = n_fft; a.hop = hop; a.x = d_in; a.batch = n; a.n_samples = n_samples; a.n_frames = frames;
    a.pad_mode = pad_mode; a.window = window; a.out_kind = out_kind; a.out0 = d0; a.out1 = d1;
    int launches = 0;
    std::string err;
    int r = launch_small_stft(a, c->stream, &launches, &err);
    c->launches += launches;
    if (r != B2A_OK) c->err = err;
    return r;
  };
  return run_batched(c, space, batch, x, size_t(n_samples), nullptr, 0, o0, per, o1, per, body);
}

int b2a_stft_hifigan(b2a_ctx* c, const float* x, int64_t batch, int64_t n_samples, int n_fft, int hop, const float* window,
                     float* real_out, float* imag_out, int space) {
  return small_stft_common(c, x, batch, n_samples, n_fft, hop, window, PAD_REFLECT, SOUT_REAL_IMAG, real_out, imag_out, space);
}

int b2a_cosyvoice3_stft(b2a_ctx* c, const float* x, int64_t batch, int64_t n_samples, int n_fft, int hop, const float* window,
                        float* real_out, float* imag_out, int space) {
  if (c && n_samples > 0 && n_samples <= n_fft / 2) {
    // zero padding has no minimum length beyond one frame (CausalHiFTGenerator.swift:438-440)
    if (n_samples + 2 * (n_fft / 2) < n_fft) return fail(c, B2A_E_TOO_SHORT, "Input is too short");
  }
  return small_stft_common(c, x, batch, n_samples, n_fft, hop, window, PAD_ZERO, SOUT_REAL_IMAG, real_out, imag_out, space);
}

int b2a_kokoro_stft_transform(b2a_ctx* c, const float* x, int64_t batch, int64_t n_samples, int filter_length, int hop_length,
                              int win_length, float* magnitude_out, float* phase_out, int space) {
  if (!c) return B2A_E_BAD_ARG;
  if (win_length <= 0 || win_length > filter_length) return fail(c, B2A_E_BAD_ARG, "win_length must be in (0, filter_length]");
  std::vector<float> w;
  hann_periodic_via_hanning(win_length, w);  // getWindow "hann" (MLXSTFT.swift:48-67)
  w.resize(filter_length, 0.0f);
  return small_stft_common(c, x, batch, n_samples, filter_length, hop_length, w.data(), PAD_REFLECT, SOUT_MAG_PHASE, magnitude_out,
                           phase_out, space);
}

// head != 0: `mag` is the vocoder's convolution output (batch, 2F, frames), `phase` is unused (exp / sin are formed in the kernel)
static int istft_common(b2a_ctx* c, const float* mag, const float* phase, int64_t batch, int64_t n_frames, int n_fft, int hop,
                        const float* window, int use_clip_lo, float clip_hi, int norm, int unwrap, float* out, int space,
                        int head = 0, float out_limit = 0.0f, const float* fade = nullptr, int64_t fade_len = 0) {
  int rc = check_common(c, mag, out, batch, n_frames);
  if (rc != B2A_OK) return rc;
  if ((!head && !phase) || !window) return fail(c, B2A_E_BAD_ARG, "null buffer");
  if (n_frames < 2) return fail(c, B2A_E_TOO_SHORT, "iSTFT needs at least 2 frames");
  if (fade_len < 0 || fade_len > 0x7fffffffLL || (fade_len > 0 && !fade)) return fail(c, B2A_E_BAD_ARG, "bad fade window");
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  // S3Gen applies its fade only to waveforms at least as long as the window (S3Gen.swift:286)
  const bool use_fade = head && fade_len > 0 && (n_frames - 1) * int64_t(hop) >= fade_len;
  if (use_fade && (c->fade_host.size() != size_t(fade_len) || std::memcmp(c->fade_host.data(), fade, sizeof(float) * size_t(fade_len)) != 0)) {
    // a new window (once per model in practice): synchronous upload -- earlier kernels may still read the old table
    cudaError_t e = cudaStreamSynchronize(c->stream);
    if (e != cudaSuccess) return cu(c, e, "sync");
    c->fade_host.clear();
    if ((rc = ensure(c, c->fade, sizeof(float) * size_t(fade_len))) != B2A_OK) return rc;
    if ((e = cudaMemcpy(c->fade.p, fade, sizeof(float) * size_t(fade_len), cudaMemcpyHostToDevice)) != cudaSuccess) return cu(c, e, "fade upload");
    c->fade_host.assign(fade, fade + fade_len);
  }
  const size_t per_in = size_t(n_fft / 2 + 1) * n_frames * (head ? 2 : 1);
  const size_t per_out = size_t(n_frames - 1) * hop;
  Body body = [&](const float* d_mag, const float* d_phase, float* d_out, float*, int64_t n, int slot) -> int {
    IstftArgs a;
    a.n_fft = n_fft; a.hop = hop; a.mag = d_mag; a.phase = d_phase; a.batch = n; a.n_frames = n_frames; a.window = window;
    a.use_clip_lo = use_clip_lo; a.clip_lo = 0.0f; a.clip_hi = clip_hi; a.norm = norm; a.out = d_out;
    a.head = head; a.out_limit = out_limit;
    if (use_fade) {
      a.fade = static_cast<const float*>(c->fade.p);
      a.fade_len = int(fade_len);
    }
    if (unwrap && !head) {   // (head: phase = sin(.) in [-1, 1], unwrap is the identity)
      int r;
      if ((r = ensure(c, c->scratch[slot][0], sizeof(int))) != B2A_OK) return r;
      if ((r = ensure(c, c->scratch[slot][2], sizeof(float) * per_in * size_t(n))) != B2A_OK) return r;
      a.unwrap = 2;
      a.d_flag = static_cast<int*>(c->scratch[slot][0].p);
      a.h_flag = c->h_flag;
      a.scratch_phase = static_cast<float*>(c->scratch[slot][2].p);
    }
    int launches = 0;
    std::string err;
    int r = launch_istft(a, c->stream, &launches, &err);
    c->launches += launches;
    if (r != B2A_OK) c->err = err;
    return r;
  };
  return run_batched(c, space, batch, mag, per_in, head ? nullptr : phase, head ? 0 : per_in, out, per_out, nullptr, 0, body);
}

int b2a_hift_head_istft(b2a_ctx* c, const float* conv_out, int64_t batch, int64_t n_frames, int n_fft, int hop, const float* window,
                        float audio_limit, float* out, int space) {
  // magnitude = exp(h[:, :F]), phase = sin(h[:, F:]), istftHiFiGAN, clip to +-audioLimit (HiFiGAN.swift:577-589)
  return istft_common(c, conv_out, nullptr, batch, n_frames, n_fft, hop, window, 0, 100.0f, NORM_WSQ_FLOOR, 0, out, space, 1, audio_limit);
}

int b2a_hift_head_istft_fade(b2a_ctx* c, const float* conv_out, int64_t batch, int64_t n_frames, int n_fft, int hop, const float* window,
                             float audio_limit, const float* fade, int64_t fade_len, float* out, int space) {
  // ... followed by result[..., 0 ..< fadeLen] *= trimFade of S3Gen.callAsFunction (S3Gen.swift:284-289), in the same kernel
  return istft_common(c, conv_out, nullptr, batch, n_frames, n_fft, hop, window, 0, 100.0f, NORM_WSQ_FLOOR, 0, out, space, 1, audio_limit,
                      fade, fade_len);
}

int b2a_s3gen_trim_fade(int sampling_rate, float* out) {
  // zeros(nTrim) ++ (cos(linspace(pi, 0, nTrim)) + 1) / 2, nTrim = sr / 50 (S3Gen.swift:259-262)
  const int n = sampling_rate / 50;
  if (n < 2 || !out) return B2A_E_BAD_ARG;
  for (int i = 0; i < n; ++i) out[i] = 0.0f;
  const float pi = 3.14159265358979323846f;
  for (int i = 0; i < n; ++i) {
    const float t = pi + float(i) * ((0.0f - pi) / float(n - 1));   // fp32 linspace
    out[n + i] = (cosf(t) + 1.0f) / 2.0f;
  }
  return B2A_OK;
}

int b2a_kokoro_head_istft(b2a_ctx* c, const float* conv_out, int64_t batch, int64_t n_frames, int filter_length, int hop_length,
                          int win_length, float* out, int space) {
  // spec = exp(x[:, :F]), phase = sin(x[:, F:]), MLXSTFT.inverse (Generator.swift:182-190); no limiter
  if (!c) return B2A_E_BAD_ARG;
  if (win_length != filter_length) return fail(c, B2A_E_UNSUPPORTED, "Kokoro inverse is built for win_length == filter_length");
  std::vector<float> w;
  hann_periodic_via_hanning(win_length, w);
  return istft_common(c, conv_out, nullptr, batch, n_frames, filter_length, hop_length, w.data(), 0, 3.402823466e+38f, NORM_WSUM_NONZERO,
                      1, out, space, 1, 0.0f);
}

int b2a_mlx_istft(b2a_ctx* c, const float* spec_complex, int64_t batch, int64_t n_frames, int win_length, int hop_length, float* out,
                  int space) {
  // mlxIstft (MLXSTFT.swift:115-163): x complex64 (F, frames), irfft of every frame, "hann" window, overlap-add, division by the
  // window sum where it is non-zero, trim of win_length / 2 at both ends (center = true).  No unwrap, no magnitude clip.
  if (!c) return B2A_E_BAD_ARG;
  if (win_length <= 0 || hop_length <= 0) return fail(c, B2A_E_BAD_ARG, "win_length and hop_length must be positive");
  std::vector<float> w;
  hann_periodic_via_hanning(win_length, w);   // getWindow "hann" (MLXSTFT.swift:48-67)
  return istft_common(c, spec_complex, nullptr, batch, n_frames, win_length, hop_length, w.data(), 0, 3.402823466e+38f, NORM_WSUM_NONZERO,
                      0, out, space, 2, 0.0f);
}

int b2a_unwrap(b2a_ctx* c, const float* phase, int64_t n_rows, int64_t n_frames, float* out, int space) {
  // numpy-style unwrap along the last axis of (n_rows, n_frames) (MLXSTFT.swift:23-46)
  int rc = check_common(c, phase, out, n_rows, n_frames);
  if (rc != B2A_OK) return rc;
  Guard g(c);
  if (!g.ok) return fail(c, B2A_E_CUDA, "cudaSetDevice failed");
  Body body = [&](const float* d_in, const float*, float* d_out, float*, int64_t n, int) -> int {
    int launches = 0;
    std::string err;
    int r = launch_unwrap(d_in, d_out, n, n_frames, c->stream, &launches, &err);
    c->launches += launches;
    if (r != B2A_OK) c->err = err;
    return r;
  };
  return run_batched(c, space, n_rows, phase, size_t(n_frames), nullptr, 0, out, size_t(n_frames), nullptr, 0, body);
}

int b2a_istft_hifigan(b2a_ctx* c, const float* magnitude, const float* phase, int64_t batch, int64_t n_frames, int n_fft, int hop,
                      const float* window, float* out, int space) {
  // clip(magnitude, max: 1e2) only (HiFiGAN.swift:300)
  return istft_common(c, magnitude, phase, batch, n_frames, n_fft, hop, window, 0, 100.0f, NORM_WSQ_FLOOR, 0, out, space);
}

int b2a_cosyvoice3_istft(b2a_ctx* c, const float* magnitude, const float* phase, int64_t batch, int64_t n_frames, int n_fft, int hop,
                         const float* window, float* out, int space) {
  // clip(magnitude, min: 0, max: 1e2) (CausalHiFTGenerator.swift:464)
  return istft_common(c, magnitude, phase, batch, n_frames, n_fft, hop, window, 1, 100.0f, NORM_WSQ_FLOOR, 0, out, space);
}

int b2a_kokoro_stft_inverse(b2a_ctx* c, const float* magnitude, const float* phase, int64_t batch, int64_t n_frames, int filter_length,
                            int hop_length, int win_length, float* out, int space) {
  if (!c) return B2A_E_BAD_ARG;
  if (win_length != filter_length) return fail(c, B2A_E_UNSUPPORTED, "Kokoro inverse is built for win_length == filter_length");
  std::vector<float> w;
  hann_periodic_via_hanning(win_length, w);
  // no magnitude clip; plain window-sum normalisation with != 0 guard; unwrap first (MLXSTFT.swift:143-155,215)
  return istft_common(c, magnitude, phase, batch, n_frames, filter_length, hop_length, w.data(), 0, 3.402823466e+38f, NORM_WSUM_NONZERO, 1,
                      out, space);
}

}  // extern "C"
