"""N > 1 host logic on CPU: world_size-2 gloo run of the stream sharding and the digest/timing
gather that bench.py performs over NCCL."""
import os
import socket

import numpy as np
import torch.multiprocessing as mp

from sdrreceiver_b200 import shard


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, per_rank, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ids = shard.stream_ids(rank, world, per_rank)
    # stand-in for the device work: each stream's "output" is a deterministic function of its id
    pcm = np.concatenate([np.full(16, g + 1, np.int16) for g in ids])
    stats, digests = shard.gather([10.0 + rank, 20.0 + 2 * rank, per_rank * 1000.0], shard.pcm_digest(pcm))
    if rank == 0:
        out.put((stats.tolist(), digests, ids))
    dist.barrier()
    dist.destroy_process_group()


def test_world2_sharding_and_gather():
    world, per_rank = 2, 3
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, per_rank, q)) for r in range(world)]
    for p in procs:
        p.start()
    stats, digests, ids0 = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    stats = np.array(stats)
    assert ids0 == [0, 2, 4]
    agg = shard.aggregate(stats, steps=5)
    assert agg["dev_ms"] == 11.0 and agg["e2e_ms"] == 22.0          # max over ranks
    assert agg["total_samples"] == 2 * 3 * 1000.0 * 5               # sum over ranks
    want0 = shard.pcm_digest(np.concatenate([np.full(16, g + 1, np.int16) for g in (0, 2, 4)]))
    want1 = shard.pcm_digest(np.concatenate([np.full(16, g + 1, np.int16) for g in (1, 3, 5)]))
    assert digests == [want0, want1]


def test_parity_verdict():
    ok = [[1, 2, 3, 7, 8, 9, 1], [4, 5, 6, 7, 8, 9, 0]]
    v = shard.parity_verdict(ok)
    assert v["ok"] and v["canary_digests_equal"] and v["canary_max_lsb_vs_reference"] == 1 and v["ranks"] == 2
    assert not shard.parity_verdict([[1, 2, 3, 7, 8, 9, 0], [4, 5, 6, 7, 8, 0, 0]])["ok"]          # one rank's canary differs
    assert not shard.parity_verdict([[1, 2, 3, 7, 8, 9, 2]])["ok"]                                # off by 2 LSB
    v = shard.parity_verdict([[1, 2, 3, 7, 8, 9, -1]])                                           # no golden file: not a pass
    assert not v["ok"] and not v["checked_against_golden"]


def test_partition_is_exact():
    for world in (1, 2, 4, 8):
        seen = sorted(g for r in range(world) for g in shard.stream_ids(r, world, 1024 // world))
        assert seen == list(range(1024))
        for g in (0, 5, 1023):
            r, slot = shard.owner(g, world)
            assert shard.stream_ids(r, world, 1024 // world)[slot] == g
    assert shard.rank_endpoint("tcp://*:6003", 3) == "tcp://*:6006"
    assert shard.rank_endpoint("ipc:///tmp/x", 3) == "ipc:///tmp/x"
