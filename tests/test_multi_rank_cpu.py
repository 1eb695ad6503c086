"""Host-side logic of the N>1 path with world_size-2 gloo process groups on CPU: batch sharding, the byte-string
exchange that distributes the NCCL unique id, max-over-ranks timing, replica checksum agreement."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.multiprocessing as mp

from smelter_b200 import dist as sdist


def test_shard_range_partitions_exactly():
    for total in (1, 7, 32, 255, 256):
        for world in (1, 2, 3, 4, 8):
            spans = [sdist.shard_range(total, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == total
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1
    assert sdist.shard_range(256, 3, 8) == (96, 128)
    with pytest.raises(ValueError):
        sdist.shard_range(8, 2, 2)


def _free_port() -> int:
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank: int, world: int, port: int, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    d = sdist.init_process_group("gloo")
    try:
        assert sdist.env_rank_world() == (rank, rank, world)
        uid = sdist.share_bytes(bytes(range(128)) if rank == 0 else b"", 0)       # the NCCL unique id travels like this
        slow = sdist.max_over_ranks(10.0 + rank)
        total = sdist.sum_over_ranks(float(rank + 1))
        same = sdist.all_equal(0xDEADBEEFCAFEF00D)
        differ = sdist.all_equal(0xDEADBEEFCAFEF00D + rank)
        # sharded "inference": each rank owns a contiguous slice; gathered result equals the unsharded one
        x = torch.arange(10, dtype=torch.float32)
        lo, hi = sdist.shard_range(10, rank, world)
        part = x[lo:hi] * 2
        parts = [None] * world
        d.all_gather_object(parts, part)
        out.put((rank, uid == bytes(range(128)), slow, total, same, differ, torch.cat(parts).tolist()))
    finally:
        d.destroy_process_group()


def test_world_size_2_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    results = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, uid_ok, slow, total, same, differ, gathered in results:
        assert uid_ok and slow == 11.0 and total == 3.0 and same and not differ
        assert gathered == [float(2 * i) for i in range(10)]
