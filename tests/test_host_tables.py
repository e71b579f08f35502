"""CPU: the C-ABI library loads, exports every symbol include/b200audio.h declares, and its host-side
tables / integer rules agree with the oracle.  No compute entry point is called (no GPU here)."""
import ctypes as C
import os
import re

import numpy as np
import pytest

from oracle import reference_dsp as R

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def api(built_lib):
    from mlx_swift_audio_b200 import api as A
    return A


def test_every_declared_symbol_is_exported(built_lib):
    hdr = open(os.path.join(ROOT, "include", "b200audio.h")).read()
    names = set(re.findall(r"B2A_API\s+[\w\s\*]+?\b(b2a_\w+)\s*\(", hdr))
    assert len(names) >= 40
    from mlx_swift_audio_b200 import _lib
    assert names == set(_lib.SIGNATURES), names ^ set(_lib.SIGNATURES)
    for n in names:
        assert hasattr(built_lib, n), n


def test_no_gpu_means_an_error_not_a_fallback(built_lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    assert built_lib.b2a_ctx_create(C.byref(h), 0) != 0
    from mlx_swift_audio_b200 import api as A
    with pytest.raises(A.B2AError):
        A.whisperLogMelSpectrogram(np.zeros(16000, np.float32), nMels=80)


def test_windows(api):
    for got, want in [(api.whisperHannWindow(400), R.whisper_hann_window(400)), (api.hanningWindow(401), R.hanning_window(401)),
                      (api.hanningWindow(1921), R.hanning_window(1921)), (api.hammingWindow(400), R.hamming_window(400)),
                      (api.poveyWindow(400), R.povey_window(400)), (api.hannWindowPeriodic(16), R.hann_window_periodic(16)),
                      (api.hannWindowPeriodic(20), R.hann_window_periodic(20)), (api.hanningWindow(21), R.hanning_window(21))]:
        assert np.abs(got - want).max() <= 2.5e-7
    assert api.whisperHannWindow(1).tolist() == [1.0] and api.hammingWindow(1).tolist() == [1.0]


@pytest.mark.parametrize("args", [(16000, 400, 80, 0.0, 8000.0), (16000, 400, 128, 0.0, 8000.0), (16000, 400, 128, 0.0, None),
                                  (16000, 400, 40, 0.0, 8000.0), (24000, 1920, 80, 0.0, 8000.0)])
def test_slaney_bank(api, args):
    got, want = api.melFilters(*args), R.mel_filters(*args)
    assert got.shape == want.shape
    assert np.array_equal(got != 0, want != 0), "support of every triangle must be identical"
    assert np.abs(got - want).max() <= 1e-5 * want.max()


def test_funasr_and_htk_banks(api):
    got, want = api.funASRMelFilters(), R.funasr_mel_filters()
    assert got.shape == (80, 200)
    assert np.abs(got - want).max() <= 1e-5 * want.max()
    got, want = api.melFiltersHTK(16000, 512, 80, 20.0, 8000.0), R.mel_filters_htk(16000, 512, 80, 20.0, 8000.0)
    assert np.array_equal(got, want), "integer-bin triangles are exact rationals"


def test_reflect_index_is_bit_exact(api):
    for n, p in [(5, 8), (3, 7), (1, 4), (2, 5), (300, 200), (1000, 200), (201, 200), (200, 200), (199, 200), (1441, 720)]:
        got = np.array([api.reflectPadIndex(i, n, p) for i in range(n + 2 * p)])
        assert np.array_equal(got, R.reflect_pad_index(n, p)), (n, p)


def test_frame_count_rules(built_lib):
    L = built_lib
    for n in (161, 200, 399, 400, 401, 16000, 479999, 480000, 480001, 960000):
        x = np.zeros(n, np.float32) + 1e-3
        assert L.b2a_whisper_num_frames(n, 0) == R.stft(x, R.whisper_hann_window(400), 400, 160).shape[0] - 1
        assert L.b2a_funasr_num_frames(n) == 1 + n // 160
        assert L.b2a_stft_num_frames(n, 400, 160, 1) == R.stft(x, R.whisper_hann_window(400), 400, 160).shape[0]
    assert L.b2a_whisper_num_frames(480000, 0) == 3000 and L.b2a_whisper_num_frames(480000, 480000) == 6000
    assert L.b2a_lfr_num_rows(2001, 6) == 334
    assert L.b2a_kaldi_num_frames(320000, 400, 160) == 1998 and L.b2a_kaldi_num_frames(399, 400, 160) < 0
    assert L.b2a_s3gen_num_frames(240000, 1920, 480) == 500 and L.b2a_s3gen_num_frames(600, 1920, 480) < 0
    assert L.b2a_s3gen_num_frames(700, 1920, 480) == R.s3gen_mel_spectrogram(np.zeros(700, np.float32)).shape[1]
    assert L.b2a_vocoder_stft_num_frames(720000, 16, 4) == 180001 and L.b2a_vocoder_stft_num_frames(720000, 20, 5) == 144001
    assert L.b2a_istft_out_length(180001, 4) == 720000
    assert L.b2a_stft_num_frames(100, 400, 160, 0) < 0
    assert L.b2a_next_power_of_2(400) == 512 and L.b2a_next_power_of_2(512) == 512 and L.b2a_next_power_of_2(1) == 1
    for a in (0, 159, 160, 16000, 320000):
        assert L.b2a_funasr_compute_feature_length(a, 160, 6) == R.compute_feature_length(a)


def test_mel_step_program_reproduces_dense_projection(built_lib, api):
    P = C.POINTER(C.c_float)
    rng = np.random.default_rng(0)
    banks = [(api.melFilters(16000, 400, 80, 0, 8000), False), (api.melFilters(16000, 400, 128, 0, 8000), False),
             (api.melFilters(16000, 400, 40, 0, 8000), False), (api.melFilters(24000, 1920, 80, 0, 8000), False),
             (api.funASRMelFilters(), False), (api.melFiltersHTK(16000, 512, 80, 20, 8000), True),
             (api.melFilters(16000, 400, 17, 300.0, 5000.0), False)]
    for bank, bin_major in banks:
        nm, nb = (bank.shape[1], bank.shape[0]) if bin_major else bank.shape
        p = rng.random(nb).astype(np.float32)
        out = np.full(nm, -1, np.float32)
        b = np.ascontiguousarray(bank)
        n = built_lib.b2a_debug_mel_program_apply(b.ctypes.data_as(P), nm, nb, int(bin_major), p.ctypes.data_as(P), out.ctypes.data_as(P))
        assert n > 0
        ref = (bank.T if bin_major else bank).astype(np.float64) @ p.astype(np.float64)
        assert np.abs(out - ref).max() <= 1e-6 * max(ref.max(), 1e-30)
    dense = rng.random((10, 50)).astype(np.float32)  # a bin feeding >2 filters: the generic kernel path is used instead
    out = np.zeros(10, np.float32)
    assert built_lib.b2a_debug_mel_program_apply(dense.ctypes.data_as(P), 10, 50, 0, rng.random(50).astype(np.float32).ctypes.data_as(P),
                                                 out.ctypes.data_as(P)) == -1


def test_bad_arguments_are_rejected_without_a_gpu(built_lib):
    L = built_lib
    out = np.zeros(4, np.float32)
    P = C.POINTER(C.c_float)
    assert L.b2a_window(99, 4, out.ctypes.data_as(P)) != 0
    assert L.b2a_window(0, 0, out.ctypes.data_as(P)) != 0
    assert L.b2a_reflect_pad_index(100, 5, 8) == -1
    assert L.b2a_whisper_log_mel_spectrogram(None, None, 1, 16000, 80, 0, None, 0) != 0  # null context
    # the ragged-batch and peer-memory entry points reject a null context / null buffers before any CUDA call
    assert L.b2a_whisper_log_mel_spectrogram_ragged(None, None, 1, 16000, None, 80, 0, None, None, 0) != 0
    assert L.b2a_funasr_preprocess_audio_ragged(None, None, 1, 16000, None, 80, 7, 6, 1, None, None, 0) != 0
    assert L.b2a_kaldi_fbank_campplus_ragged(None, None, 1, 16000, None, 16000, 80, 25.0, 10.0, 0, None, None, 0) != 0
    assert L.b2a_s3gen_mel_spectrogram_ragged(None, None, 1, 24000, None, 1920, 80, 24000, 480, 1920, 0, 8000, None, None, 0) != 0
    p = C.c_void_p()
    assert L.b2a_device_alloc(None, C.byref(p), 1024) != 0 and not p.value
    assert L.b2a_ipc_open(None, b"\0" * 64, C.byref(p)) != 0
    assert L.b2a_ipc_export(None, None, C.create_string_buffer(64)) != 0



def test_s3tokenizer_segment_plan_matches_the_oracle(built_lib):
    import ctypes as C
    from oracle import reference_dsp as R
    lens = np.array([7000, 3000, 3001, 1, 0, 5200, 5201], np.int64)
    I64, I32 = C.POINTER(C.c_int64), C.POINTER(C.c_int32)
    n = built_lib.b2a_s3tokenizer_plan_segments(lens.ctypes.data_as(I64), len(lens), 3000, 2600, None, None, None, 0)
    bi, st, ln = (np.empty(n, np.int32) for _ in range(3))
    assert built_lib.b2a_s3tokenizer_plan_segments(lens.ctypes.data_as(I64), len(lens), 3000, 2600, bi.ctypes.data_as(I32),
                                                   st.ctypes.data_as(I32), ln.ctypes.data_as(I32), n) == n
    mel = np.zeros((len(lens), 1, 7000), np.float32)
    _, want_len, info = R.s3tokenizer_segments(mel, lens)
    assert n == len(info) and np.array_equal(ln, want_len) and [b for b, _ in info] == bi.tolist()
    assert st.tolist() == [0, 2600, 5200, 0, 0, 2600, 0, 0, 0, 2600, 0, 2600, 5200]
    assert built_lib.b2a_s3tokenizer_plan_segments(lens.ctypes.data_as(I64), len(lens), 3000, 2600, bi.ctypes.data_as(I32),
                                                   st.ctypes.data_as(I32), ln.ctypes.data_as(I32), n - 1) == -1


@pytest.mark.parametrize("n_fft,n_mels", [(400, 128), (400, 80), (400, 40), (512, 80), (1920, 80)])
def test_exchange_buffer_layout_is_consistent(built_lib, n_fft, n_mels):
    """The frontend kernel uses the exchange buffer three times per tile.  Invariants of the layout: every spectrum bin has its
    own row inside the block of the stage-B item that produces it; the mel sums are staged in words that belong to no spectrum
    row and to no other filter, inside the buffer; writes by frame and reads by filter are bank-conflict free."""
    import ctypes as C
    nb = n_fft // 2 + 1
    slots = (C.c_int * nb)()
    words = (C.c_int * n_mels)()
    packed = built_lib.b2a_debug_plan_layout(n_fft, n_mels, slots, words)
    assert packed > 0
    ft, n1, n2 = packed >> 16, (packed >> 8) & 0xff, packed & 0xff
    assert n1 * n2 == n_fft
    slots, words = np.array(slots[:]), np.array(words[:])
    assert len(set(slots.tolist())) == nb and slots.min() >= 0 and slots.max() < n_fft
    for k in range(nb):
        item = k % n1 if k % n1 < n1 // 2 else n1 - k % n1          # the item that holds residue k mod n1 (mirrored above n1/2)
        if k % n1 in (0, n1 // 2):
            item = 0                                                    # k1 = 0 and k1 = n1/2 share item 0
        assert slots[k] // (2 * n2) == item
    spectrum_words = set()
    for r in slots:
        spectrum_words.update(range(r * ft, (r + 1) * ft))
    staged = set()
    for m in range(n_mels):
        w = set(range(words[m], words[m] + ft))
        assert not (w & spectrum_words) and not (w & staged) and words[m] >= 0 and words[m] + ft <= n_fft * ft
        staged |= w
    if ft == 32:
        for f in (0, 5, 31):   # 32 consecutive filters read at one frame index fall in 32 different banks
            for m0 in range(0, n_mels - 31, 32):
                assert len({(words[m] + f) % 32 for m in range(m0, m0 + 32)}) == 32


def test_lfr_slot_rule_of_the_interior_tile_store():
    """The main kernel's division-free LFR store (frontend.cu, OUT_LFR of the baked Fun-ASR bank) relies on this rule for the
    standard 7 / 6 stacking (FunASRAudio.swift:108-154): the output of a clip is a stream of slots, slot(i, j) = 7 i + j; an
    interior frame t (tile without frame 0 and frame T-1) fills slot u + u // 6 with u = t + 3, and also the slot before it
    when u is a multiple of 6; every other slot belongs to an edge tile.  Checked against the oracle's applyLFR."""
    for T in list(range(34, 300)) + [1998, 2001, 3000]:
        feat = np.arange(T, dtype=np.float32)[:, None] * np.ones((1, 2), np.float32)   # feature value = frame index
        src = R.apply_lfr(feat).reshape(-1, 7, 2)[:, :, 0].astype(np.int64).reshape(-1)   # slot -> source frame
        rows = (T + 5) // 6
        assert src.size == rows * 7
        got = np.full(rows * 7, -1, np.int64)
        for f0 in range(32, T, 32):
            if f0 + 32 >= T:
                continue   # the tile holding frame T-1 takes the generic path
            for t in range(f0, f0 + 32):
                u = t + 3
                i1 = u // 6
                if i1 <= rows - 1:
                    assert got[u + i1] == -1
                    got[u + i1] = t
                if u == 6 * i1:
                    assert got[u + i1 - 1] == -1
                    got[u + i1 - 1] = t
        interior = (src >= 32) & (src // 32 * 32 + 32 < T)
        assert np.array_equal(got[interior], src[interior]), f"T={T}: interior slots"
        assert np.all(got[~interior] == -1), f"T={T}: a slot of an edge tile was written by the interior rule"


@pytest.mark.parametrize("n_fft,n_mels,step", [(400, 128, 10), (400, 40, 10), (512, 80, 16), (1920, 80, 32)])
def test_incremental_walk_of_the_staging_offsets(built_lib, n_fft, n_mels, step):
    """The (M, T') store loops of the frontend kernel (OutBaseWalk in frontend.cu) advance quotient and remainder of m / CAP
    incrementally instead of dividing per row; the walk must reproduce the library's own staging map (output_words(), through
    b2a_debug_plan_layout) for every start slot and every filter.  step = items per pass of the plan (warps x sub-items)."""
    import ctypes as C
    slots = (C.c_int * (n_fft // 2 + 1))()
    words = (C.c_int * n_mels)()
    packed = built_lib.b2a_debug_plan_layout(n_fft, n_mels, slots, words)
    ft, n2 = packed >> 16, packed & 0xff
    cap = ((n2 - 1) * ft - (ft - 1)) // (ft + 1)
    for m0 in range(min(step, n_mels)):
        q, r = divmod(m0, cap)
        for m in range(m0, n_mels, step):
            assert (q * 2 * n2 + n2 + 1) * ft + ((cap * q) & (ft - 1)) + r * (ft + 1) == words[m], (m0, m)
            r += step % cap
            q += step // cap
            if r >= cap:
                r -= cap
                q += 1


def test_merge_tokenized_segments_matches_the_reference_rule():
    # mergeTokenizedSegments (S3TokenizerUtils.swift:71-88): bit-exact integer rule, incl. segments shorter than the dropped edges
    from mlx_swift_audio_b200 import api
    from oracle import reference_dsp as R
    rng = np.random.default_rng(11)
    cases = [
        [list(range(750)), list(range(1000, 1750)), list(range(2000, 2300))],      # S3Tokenizer's own shape: 30 s windows at 25 Hz
        [list(range(10))],                                                         # a single segment is kept whole
        [list(range(60)), list(range(100, 140)), list(range(200, 260))],           # middle segment shorter than 2 * 50: dropped
        [list(range(50)), list(range(100, 150))],                                  # exactly the dropped edge: nothing left of either
        [],
        [[], [1, 2, 3]],
    ]
    for _ in range(20):
        cases.append([list(rng.integers(0, 6561, int(rng.integers(0, 400)))) for _ in range(int(rng.integers(1, 6)))])
    for segs in cases:
        for overlap, rate in ((4, 25), (5, 25), (2, 50), (0, 25)):
            assert api.mergeTokenizedSegments(segs, overlap, rate) == [int(v) for v in R.merge_tokenized_segments(segs, overlap, rate)]


def test_resample_poly_filter_is_scipys_design():
    # non-parity extension: the polyphase resampler is DEFINED as scipy.signal.resample_poly (default Kaiser design); the host-side
    # filter + alignment must reproduce it (the GPU kernel is checked against scipy in tests/test_gpu_round2.py)
    from scipy import signal
    from mlx_swift_audio_b200 import _lib as L
    lib = L.load()
    rng = np.random.default_rng(5)
    for n, fr, to in [(2400, 24000, 16000), (1600, 16000, 24000), (4410, 44100, 16000), (1000, 48000, 16000), (777, 22050, 16000)]:
        up, dn, pre = C.c_int(), C.c_int(), C.c_int64()
        taps = lib.b2a_resample_poly_filter(n, fr, to, None, 0, C.byref(up), C.byref(dn), C.byref(pre))
        assert taps > 0
        h = np.zeros(taps, np.float32)
        assert lib.b2a_resample_poly_filter(n, fr, to, h.ctypes.data_as(C.POINTER(C.c_float)), taps, C.byref(up), C.byref(dn), C.byref(pre)) == taps
        x = rng.standard_normal(n)
        want = signal.resample_poly(x, up.value, dn.value)
        new_t = lib.b2a_resample_poly_length(n, fr, to)
        assert new_t == len(want)
        xu = np.zeros(n * up.value)
        xu[::up.value] = x
        full = np.convolve(xu, h.astype(np.float64))
        got = full[pre.value * dn.value::dn.value][:new_t]
        assert np.abs(got - want).max() <= 1e-6


@pytest.mark.parametrize("args", [(24000, 1920, 80, 0.0, 8000.0), (24000, 1920, 96, 0.0, 8000.0), (24000, 1920, 40, 0.0, 8000.0),
                                  (24000, 1920, 33, 0.0, 12000.0)])
def test_wpf_mel_schedule_equals_the_dense_bank(built_lib, api, args):
    """Mel schedule of the warp-per-frame n_fft 1920 kernel (csrc/wpf1920.cu, build_wpf_mel): the per-lane segment sums (four products per step from a
    16-byte aligned start bin), interpreted on the host exactly as the kernel runs them, equal the dense filterbank product (the segments only regroup the additions)."""
    import ctypes as C
    bank = np.ascontiguousarray(R.mel_filters(*args), np.float32)          # (M, 961)
    rng = np.random.default_rng(7)
    for p in (rng.random(961).astype(np.float32) * 100, np.ones(961, np.float32), np.eye(961, dtype=np.float32)[960],
              np.eye(961, dtype=np.float32)[0]):
        out = np.zeros(args[2], np.float32)
        fp = C.POINTER(C.c_float)
        words = built_lib.b2a_debug_wpf_mel_apply(bank.ctypes.data_as(fp), args[2], 961, 0, p.ctypes.data_as(fp), out.ctypes.data_as(fp))
        assert 0 < (words & 0xffff) <= 3584
        assert words >> 16 == 0, "every quarter-warp reads eight different 16-byte bank groups"
        want = bank.astype(np.float64) @ p.astype(np.float64)
        assert np.abs(out - want).max() <= 2e-6 * max(1.0, np.abs(want).max())


def test_wpf_mel_schedule_rejects_what_does_not_fit(built_lib):
    import ctypes as C
    fp = C.POINTER(C.c_float)
    bank = np.ascontiguousarray(R.mel_filters(24000, 1920, 128, 0.0, 8000.0), np.float32)   # more than 96 filters: the tiled kernel keeps it
    p = np.ones(961, np.float32)
    out = np.zeros(128, np.float32)
    assert built_lib.b2a_debug_wpf_mel_apply(bank.ctypes.data_as(fp), 128, 961, 0, p.ctypes.data_as(fp), out.ctypes.data_as(fp)) == -1


def test_process_wide_switches_are_host_only(built_lib):
    # A/B switches of the kernels' scheduling (dynamic tile walk, warp-per-frame n_fft 1920 kernel and its pruned stage B): plain host state,
    # settable without a GPU, and none of them may change results (tests/test_gpu_dyn_tiles.py, tests/test_gpu_wpf1920.py compare both sides)
    for fn, values in ((built_lib.b2a_debug_dyn_tiles, (0, 1)), (built_lib.b2a_debug_wpf1920, (0, 2, 1))):
        fn.restype = C.c_int
        fn.argtypes = [C.c_int]
        for v in values:
            assert fn(v) == 0


def test_pruned_band_of_the_1920_point_kernel():
    # wpf1920.cu PRUNE: with a bank that reads no bin above 640, bin k1 + 60 k2 (direct, k2 < 16) or 1920 - k1 - 60 k2 (mirror image, k2 > 16)
    # of every stage-B row k1 = 0..30 lies above 640 for k2 = 11..20 -- and for no other k2 in EVERY lane; S3Gen's bank ends at bin 640
    # (8000 Hz at 12.5 Hz per bin: the last filter's falling edge leaves a rounding-sized weight there)
    dead = []
    for k2 in range(32):
        bins = [(k1 + 60 * k2) if k2 < 16 else (960 - k1 if k2 == 16 else 1920 - k1 - 60 * k2) for k1 in range(31)]
        if all(b > 640 for b in bins):
            dead.append(k2)
    assert dead == list(range(11, 21))
    fb = R.mel_filters(24000, 1920, 80, 0.0, 8000.0)
    assert int(np.nonzero(fb.any(axis=0))[0].max()) <= 640
