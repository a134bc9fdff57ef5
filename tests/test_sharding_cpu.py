"""Multi-GPU host logic on CPU: stream partitioning and the max-over-ranks reduction, exercised with
a real world_size-2 gloo process group (the data path itself has no collective)."""
import os
import socket

import pytest

from audio_filtering_pipeline_b200.sharding import partition_streams


@pytest.mark.parametrize("total,world", [(16384, 8), (4096, 3), (7, 2), (2, 4), (65536, 8), (1, 1)])
def test_partition_covers_all_streams_once(total, world):
    seen = []
    for r in range(world):
        first, count = partition_streams(total, world, r)
        assert count >= 0
        seen += list(range(first, first + count))
        if first + count < total:
            assert count % 2 == 0            # pairs never straddle two GPUs
    assert seen == list(range(total))
    counts = [partition_streams(total, world, r)[1] for r in range(world)]
    assert max(counts) - min(counts) <= 2


def test_partition_rejects_bad_rank():
    with pytest.raises(ValueError):
        partition_streams(10, 2, 2)


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    import torch.distributed as dist
    from audio_filtering_pipeline_b200.sharding import max_over_ranks, sum_over_ranks, partition_streams
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        first, count = partition_streams(1000, world, rank)
        t = max_over_ranks(1.0 + rank)               # the slowest rank defines the step time
        n = sum_over_ranks(count)                    # whole-job units
        dist.barrier()
        q.put((rank, first, count, t, n))
    finally:
        dist.destroy_process_group()


def test_gloo_world_size_2():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert [r[1:3] for r in res] == [(0, 500), (500, 500)]
    assert all(r[3] == 2.0 for r in res)             # max over ranks
    assert all(r[4] == 1000.0 for r in res)          # sum over ranks
