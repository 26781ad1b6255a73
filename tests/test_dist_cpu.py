"""World-size-2 gloo tests (CPU) of the multi-GPU host logic: round-robin sharding, the
count + padded-payload all-gather of SiftPoint arrays, and the all-pairs plan.  The oracle
plays the matcher so that the exchanged bytes are checked end to end."""
import os
import socket

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from cudasift_b200 import multi
from cudasift_b200.synth import synth_descriptors


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import oracle
        n = 40 + 25 * rank                                   # different counts per rank
        mine = synth_descriptors(n, seed=10 + rank)
        buf = torch.from_numpy(np.frombuffer(mine.tobytes(), np.uint8).copy())
        parts, counts = multi.allgather_records(dist, buf, n)
        assert counts == [40 + 25 * r for r in range(world)]
        for r in range(world):
            got = multi.records_from_bytes(parts[r])
            assert got.tobytes() == synth_descriptors(40 + 25 * r, seed=10 + r).tobytes()
        res = {}
        for j in multi.all_pairs_plan(world, rank):
            m = oracle.match(mine, multi.records_from_bytes(parts[j]))
            res[j] = m["match"].tolist()
        q.put((rank, res, multi.shard_indices(11, world, rank)))
    finally:
        dist.destroy_process_group()


def test_allgather_and_sharding_world2():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    [p.start() for p in procs]
    out = [q.get(timeout=120) for _ in range(world)]
    [p.join(timeout=60) for p in procs]
    assert all(p.exitcode == 0 for p in procs)
    out.sort()
    import oracle
    s0, s1 = synth_descriptors(40, 10), synth_descriptors(65, 11)
    assert out[0][1][1] == oracle.match(s0, s1)["match"].tolist()
    assert out[1][1][0] == oracle.match(s1, s0)["match"].tolist()
    assert out[0][2] == [0, 2, 4, 6, 8, 10] and out[1][2] == [1, 3, 5, 7, 9]


def test_shard_covers_all_items():
    for world in (1, 2, 4, 8):
        seen = sorted(i for r in range(world) for i in multi.shard_indices(512, world, r))
        assert seen == list(range(512))
        sizes = [len(multi.shard_indices(512, world, r)) for r in range(world)]
        assert max(sizes) - min(sizes) <= 1
    assert multi.all_pairs_plan(8, 3) == [0, 1, 2, 4, 5, 6, 7]
