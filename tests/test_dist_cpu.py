"""N > 1 host logic on CPU: world_size-2 gloo processes exercise the sharding and the packed-model broadcast that
bench.py uses with NCCL on the GPU box."""
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "rife-ncnn-vulkan_b200"))
import dist_util  # noqa: E402


def test_shard_pairs_covers_stream_without_overlap():
    for n in (0, 1, 7, 8, 64, 65):
        for world in (1, 2, 3, 8):
            spans = [dist_util.shard_pairs(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for a, b in zip(spans, spans[1:]):
                assert a[1] == b[0]
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    blob = np.arange(100003, dtype=np.uint64).astype(np.uint8) if rank == 0 else None
    got = dist_util.broadcast_blob(blob, dist)
    mx = dist_util.max_over_ranks(10.0 + rank, dist)
    q.put((rank, int(got.sum()), got.size, mx))
    dist.destroy_process_group()


def test_broadcast_and_max_with_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    expect = int(np.arange(100003, dtype=np.uint64).astype(np.uint8).sum())
    assert [r[1] for r in res] == [expect, expect]
    assert [r[2] for r in res] == [100003, 100003]
    assert [r[3] for r in res] == [11.0, 11.0]
