"""Times the ragged entry points with all lengths equal (same work as the equal-length kernels beside them) and the run-time-configured
40-mel voice-encoder front end on full-size batches."""
import numpy as np
import torch
from mlx_swift_audio_b200 import api

ctx = api.Context(0, torch.cuda.current_stream().cuda_stream)
B, n = 512, 320000
g = torch.Generator(device="cuda").manual_seed(1)
x = 0.1 * torch.randn((B, n), generator=g, device="cuda")
lengths = [n] * B


def timeit(name, fn, reps=10):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"{name:34s} {e0.elapsed_time(e1) / reps:.3f} ms")


timeit("funasr preprocess (baked)", lambda: api.preprocessAudio(x, ctx=ctx))
timeit("funasr preprocess ragged", lambda: api.preprocessAudioRagged(x, lengths, ctx=ctx))
timeit("kaldi (baked)", lambda: api.kaldiFbankCAMPPlus(x, meanNorm=True, ctx=ctx))
timeit("kaldi ragged", lambda: api.kaldiFbankCAMPPlusRagged(x, lengths, meanNorm=True, ctx=ctx))
timeit("voice encoder 40-mel (generic)", lambda: api.voiceEncoderMelspectrogram(x, ctx=ctx))
timeit("chatterbox 128-mel MT (baked)", lambda: api.logMelSpectrogramChatterbox(x, ctx=ctx))
timeit("chatterbox ragged", lambda: api.logMelSpectrogramChatterboxRagged(x, lengths, ctx=ctx))
