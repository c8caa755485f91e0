"""N>1 path on CPU: frame partitioning and the peaks all-gather with world_size 2 over gloo
(the GPU box runs the same code over NCCL)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pragma_dsp_b200 import sharding


def test_partition_covers_every_frame_once():
    for total in (0, 1, 7, 8, 9, 65536, 1_048_576, 28_122):
        for world in (1, 2, 4, 8):
            per = sharding.frames_per_rank(total, world)
            seen = []
            for r in range(world):
                a, b = sharding.shard_range(total, r, world)
                assert 0 <= b - a <= per
                seen.extend(range(a, b)) if total <= 65536 else None
                if r:
                    assert a == sharding.shard_range(total, r - 1, world)[1]
            assert sharding.shard_range(total, world - 1, world)[1] == total
            if total <= 65536:
                assert seen == list(range(total))
    assert sharding.stft_sample_span(10, 20, 1024, 4096) == (10240, 19 * 1024 + 4096)
    assert sharding.stft_sample_span(5, 5, 1024, 4096) == (0, 0)
    with pytest.raises(ValueError):
        sharding.shard_range(10, 2, 2)


def _worker(rank, world, port, total, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    per = sharding.frames_per_rank(total, world)
    a, b = sharding.shard_range(total, rank, world)
    # each rank fabricates the records of its own frames: index field = global frame number
    local = np.zeros(per, dtype=np.dtype([("index", "<i4"), ("_pad", "<i4"), ("frequency", "<f8"), ("amplitude", "<f8"), ("phase", "<f8")]))
    local["index"][: b - a] = np.arange(a, b)
    local["amplitude"][: b - a] = np.arange(a, b) * 0.5
    t = torch.from_numpy(local.view(np.uint8).reshape(per, 32).copy())
    out = sharding.gather_peaks(t, total)
    rec = sharding.peaks_from_bytes(out.numpy(), "f64")
    ok = bool((rec["index"] == np.arange(total)).all() and (rec["amplitude"] == np.arange(total) * 0.5).all())
    q.put((rank, ok, len(rec)))
    dist.barrier()
    dist.destroy_process_group()


def test_gather_peaks_world_size_2_gloo():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    total = 1001  # odd: rank 1 holds one padding record
    procs = [ctx.Process(target=_worker, args=(r, 2, port, total, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert res == [(0, True, total), (1, True, total)]
