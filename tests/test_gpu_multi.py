"""GPU, 2 ranks over NCCL (skipped on a single-GPU box): the N > 1 path of bench.py / production use -- every rank runs the CUDA
path on its contiguous shard of clips (no data-path collective) and the features are gathered onto rank 0 over NVLink
(mlx_swift_audio_b200.shard.gather_features); the result must equal the single-GPU run bit for bit."""
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
