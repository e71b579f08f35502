"""Cost of the ragged entry points when every clip has the same length (clip / tile tables, tail kernel), against the equal-length
entry points: python tools/gpu/ragged_equal.py  (CUDA events on torch's current stream, 20 calls each)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mlx_swift_audio_b200 import api  # noqa: E402

dev = torch.device("cuda", 0)
ctx = api.Context(0, torch.cuda.current_stream(dev).cuda_stream)


def timed(fn, n=20):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


g = torch.Generator(device=dev)
g.manual_seed(1)
for name, B, n, eq, rg in (
    ("funasr preprocessAudio", 512, 320000, lambda x: api.preprocessAudio(x, ctx=ctx), lambda x, l: api.preprocessAudioRagged(x, l, ctx=ctx)),
    ("kaldi fbank + mean-norm", 512, 320000, lambda x: api.kaldiFbankCAMPPlus(x, meanNorm=True, ctx=ctx),
     lambda x, l: api.kaldiFbankCAMPPlusRagged(x, l, meanNorm=True, ctx=ctx)),
    ("chatterbox 128-mel", 1024, 160000, lambda x: api.logMelSpectrogramChatterbox(x, nMels=128, ctx=ctx),
     lambda x, l: api.logMelSpectrogramChatterboxRagged(x, l, nMels=128, ctx=ctx)),
):
    x = 0.1 * torch.randn((B, n), generator=g, device=dev)
    lengths = [n] * B
    print(f"{name}: equal-length {timed(lambda: eq(x)):.4f} ms, ragged (all lengths equal) {timed(lambda: rg(x, lengths)):.4f} ms", flush=True)
