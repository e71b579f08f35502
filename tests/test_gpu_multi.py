"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the N > 1 path of bench.py / production use -- every rank runs the CUDA
path on its contiguous shard of clips (no data-path collective) and the features are gathered onto rank 0 over NVLink, either
by NCCL (mlx_swift_audio_b200.shard.gather_features) or fused into the kernels' own stores through peer-mapped memory
(shard.FusedGather); both results must equal the single-GPU run bit for bit."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _worker(rank, world, port, n_clips, q):
    import torch
    import torch.distributed as dist
    from mlx_swift_audio_b200 import api
    from mlx_swift_audio_b200.shard import gather_features, shard_range
    from tests import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        x = synth.pcm(n_clips, 16000 * 2 + 11, seed=91)
        a, b = shard_range(n_clips, rank, world)
        local = api.whisperLogMelSpectrogram(torch.from_numpy(x[a:b]).cuda(), nMels=128)
        full = gather_features(local, n_clips, dst=0)
        torch.cuda.synchronize()
        if rank == 0:
            want = api.whisperLogMelSpectrogram(torch.from_numpy(x).cuda(), nMels=128)
            torch.cuda.synchronize()
            q.put(bool(torch.equal(full, want)))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 8])
def test_two_gpu_shard_and_nccl_gather(n_clips):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=10)


def _fused_worker(rank, world, port, n_clips, q):
    import torch
    import torch.distributed as dist
    from mlx_swift_audio_b200 import api
    from mlx_swift_audio_b200.shard import FusedGather, shard_range
    from tests import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        ctx = api.Context(rank, torch.cuda.current_stream(rank).cuda_stream)
        n = 16000 * 2 + 11
        x = synth.pcm(n_clips, n, seed=92)
        a, b = shard_range(n_clips, rank, world)
        xl = torch.from_numpy(x[a:b]).cuda()
        ok = True
        # Whisper (main kernel + clamp kernel, both writing through the peer mapping) and Fun-ASR (LFR store + CMVN pass)
        fg = FusedGather(ctx, n_clips, (n // 160, 128), dst=0)
        fg2 = FusedGather(ctx, n_clips, ((n // 160 + 1 + 5) // 6, 560), dst=0)
        for rep in range(2):   # the buffers are reused
            if b > a:
                api.whisperLogMelSpectrogram(xl * (rep + 1), nMels=128, ctx=ctx, out=fg.local_out())
                api.preprocessAudio(xl * (rep + 1), ctx=ctx, out=fg2.local_out())
            full, full2 = fg.finish(), fg2.finish()
            if rank == 0:
                xa = torch.from_numpy(x).cuda() * (rep + 1)
                ok = ok and bool(torch.equal(full, api.whisperLogMelSpectrogram(xa, nMels=128, ctx=ctx)))
                ok = ok and bool(torch.equal(full2, api.preprocessAudio(xa, ctx=ctx)))
                torch.cuda.synchronize()
            else:
                assert full is None and full2 is None
            fg.reuse()    # the consumer is done reading: the producers may overwrite the buffers
            fg2.reuse()
        fg.close()
        fg2.close()
        ctx.close()
        if rank == 0:
            q.put(ok)
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 8])
def test_two_gpu_fused_peer_store_gather(n_clips):
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    import torch.multiprocessing as mp
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_fused_worker, args=(r, 2, port, n_clips, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(300)
        assert p.exitcode == 0
    assert q.get(timeout=10)
