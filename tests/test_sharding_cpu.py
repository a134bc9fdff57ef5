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


def test_reference_arm_prints_one_line_on_rank_0_only():
    """bench.py --impl reference under torchrun: rank 0 alone runs the CPU arm and prints the JSON line with
    the contract's keys; every other rank exits 0 without work or output."""
    import json
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2", "--steps", "1",
           "--warmup", "0"]
    env1 = dict(os.environ, RANK="1", WORLD_SIZE="2", LOCAL_RANK="1")
    r1 = subprocess.run(cmd, env=env1, capture_output=True, text=True, timeout=120)
    assert r1.returncode == 0 and r1.stdout.strip() == ""
    env0 = dict(os.environ, RANK="0", WORLD_SIZE="2", LOCAL_RANK="0")
    r0 = subprocess.run(cmd, env=env0, capture_output=True, text=True, timeout=300)
    assert r0.returncode == 0, r0.stderr[-500:]
    lines = [l for l in r0.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["n_gpus"] == 2 and d["unit"] == "Msamples/s" and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "Msamples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["gpu_launches"] == 0 and "workload" in d["config"]
