"""One equal-length and one ragged (all lengths equal) Fun-ASR preprocessAudio call, for an ncu launch list:
ncu --metrics gpu__time_duration.sum --clock-control none --csv python tools/gpu/ragged_launches.py"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from mlx_swift_audio_b200 import api  # noqa: E402

dev = torch.device("cuda", 0)
ctx = api.Context(0, torch.cuda.current_stream(dev).cuda_stream)
B, n = 512, 320000
x = 0.1 * torch.randn((B, n), device=dev)
for _ in range(3):
    api.preprocessAudio(x, ctx=ctx)
    api.preprocessAudioRagged(x, [n] * B, ctx=ctx)
torch.cuda.synchronize()
