"""CPU, world_size 2 over gloo: the N>1 path of bench.py / production use -- contiguous clip shards per
rank with no data-path collective, plus the optional gather of features onto a consumer rank.  The
per-rank compute is stood in for by the oracle (the CUDA path needs a GPU)."""
import os
import socket

import numpy as np
import pytest

from mlx_swift_audio_b200.shard import shard_range


def test_shard_range_partitions_exactly():
    for n in (0, 1, 7, 8, 1024, 1025):
        for world in (1, 2, 3, 4, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1 and sizes == sorted(sizes, reverse=True)
    with pytest.raises(ValueError):
        shard_range(4, 2, 2)


def _worker(rank, world, port, n_clips, q):
    import torch
    import torch.distributed as dist
    from mlx_swift_audio_b200.shard import gather_features, shard_range as sr
    from oracle import reference_dsp as R
    from tests import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        x = synth.pcm(n_clips, 4000, seed=77)            # every rank can rebuild the batch; it only processes its shard
        a, b = sr(n_clips, rank, world)
        local = np.stack([R.whisper_log_mel_spectrogram(c, 80) for c in x[a:b]]) if b > a else np.zeros((0, 25, 80), np.float32)
        full = gather_features(torch.from_numpy(local), n_clips, dst=0)
        count = torch.tensor([b - a])
        dist.all_reduce(count)                            # bookkeeping only; not part of the data path
        if rank == 0:
            want = np.stack([R.whisper_log_mel_spectrogram(c, 80) for c in x])
            q.put((int(count.item()), bool(np.array_equal(full.numpy(), want))))
        else:
            assert full is None
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 4])
def test_two_rank_shard_and_gather(n_clips):
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
        p.join(120)
        assert p.exitcode == 0
    total, same = q.get(timeout=10)
    assert total == n_clips and same
