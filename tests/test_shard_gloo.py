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


# ---- FusedGather (gather fused into the producers' stores through peer-mapped memory): host logic on CPU -------------------
# CUDA IPC needs GPUs, so POSIX shared memory stands in for the peer mapping: b2a_device_alloc / b2a_ipc_export / b2a_ipc_open are
# stubbed with multiprocessing.shared_memory, the "kernel" is the oracle writing through the mapped address.  What is tested is
# everything FusedGather itself does: the handle broadcast, each rank's slice offset for uneven shards, reuse, and the barrier
# choreography of finish() / close().

class _ShmLib:
    def __init__(self):
        self.blocks = {}

    def _map(self, shm):
        import ctypes
        addr = ctypes.addressof(ctypes.c_char.from_buffer(shm.buf))
        self.blocks[addr] = shm
        return addr

    def b2a_device_alloc(self, h, pptr, nbytes):
        from multiprocessing import shared_memory
        pptr._obj.value = self._map(shared_memory.SharedMemory(create=True, size=int(nbytes)))
        return 0

    def b2a_ipc_export(self, h, base, buf):
        buf.value = self.blocks[base.value].name.encode()
        return 0

    def b2a_ipc_open(self, h, handle, pptr):
        from multiprocessing import shared_memory
        pptr._obj.value = self._map(shared_memory.SharedMemory(name=bytes(handle).split(b"\0")[0].decode()))
        return 0

    def b2a_ipc_close(self, h, ptr):
        self.blocks.pop(ptr.value)   # (the mapping itself goes away with the process: NumPy views may still reference it)
        return 0

    def b2a_device_free(self, h, ptr):
        self.blocks.pop(ptr.value).unlink()
        return 0


class _ShmCtx:
    device, h = 0, None

    def __init__(self):
        self.lib = _ShmLib()

    def check(self, rc):
        assert rc == 0

    def sync(self):
        pass


def _as_array(address, shape):
    import ctypes
    n = int(np.prod(shape))
    return np.ctypeslib.as_array((ctypes.c_float * n).from_address(address)).reshape(shape)


def _fused_worker(rank, world, port, n_clips, q):
    import torch.distributed as dist
    from mlx_swift_audio_b200.shard import FusedGather
    from oracle import reference_dsp as R
    from tests import synth
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)

    class Gather(FusedGather):
        def _view(self, base, shape):
            return _as_array(base, shape)

    try:
        x = synth.pcm(n_clips, 4000, seed=78)
        fg = Gather(_ShmCtx(), n_clips, (25, 80), dst=0)
        ok = True
        for rep in range(2):
            out = fg.local_out()
            assert out.shape == (fg.stop - fg.start, 25, 80)
            dst = _as_array(out.data_ptr(), out.shape) if fg.stop > fg.start else None
            for i, c in enumerate(x[fg.start:fg.stop]):          # the "kernel": stores straight into the consumer's buffer
                dst[i] = R.whisper_log_mel_spectrogram(c * (rep + 1), 80)
            full = fg.finish()
            if rank == 0:
                want = np.stack([R.whisper_log_mel_spectrogram(c * (rep + 1), 80) for c in x])
                ok = ok and np.array_equal(full, want)
            else:
                assert full is None
            fg.reuse()   # the consumer is done reading: the producers may overwrite
        fg.close()
        if rank == 0:
            q.put(bool(ok))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_clips", [5, 1])
def test_two_rank_fused_gather_host_logic(n_clips):
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
        p.join(120)
        assert p.exitcode == 0
    assert q.get(timeout=10)
