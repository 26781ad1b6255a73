"""Multi-GPU plumbing: one process per GPU (torch.distributed), images sharded across ranks.

The hot path shards trivially (independent images): image i -> rank i mod world, no
collective on the extraction path (BASELINE config #4).  The cross-image
match step (config #5) needs exactly one exchange: an all-gather of the SiftPoint arrays.
Counts differ per image, so the counts are all-gathered first (a few bytes) and the
payload is padded to the largest count.

torch is used here for what it is good at -- device buffers, streams and the NCCL
process group; the compute goes through the C ABI with raw device pointers.
"""
import numpy as np

from . import SIFT_DTYPE

REC = SIFT_DTYPE.itemsize


def shard_indices(n_items, world, rank):
    """Round-robin shard: item i belongs to rank i % world."""
    return list(range(rank, n_items, world))


def padded_count(counts):
    return int(max(max(counts), 1))


def allgather_records(dist, local_bytes, count):
    """All-gather SiftPoint records.

    local_bytes: 1-D torch.uint8 tensor holding >= count*576 bytes (CPU tensor under gloo,
    CUDA tensor under nccl).  Returns (per-rank uint8 tensors trimmed to their counts, counts)."""
    import torch
    world = dist.get_world_size()
    dev = local_bytes.device
    cnt = torch.tensor([count], dtype=torch.int64, device=dev)
    all_cnt = [torch.zeros(1, dtype=torch.int64, device=dev) for _ in range(world)]
    dist.all_gather(all_cnt, cnt)
    counts = [int(c.item()) for c in all_cnt]
    pad = padded_count(counts) * REC
    send = torch.zeros(pad, dtype=torch.uint8, device=dev)
    send[:count * REC] = local_bytes[:count * REC]
    recv = torch.empty(world * pad, dtype=torch.uint8, device=dev)
    if dev.type == "cuda":
        dist.all_gather_into_tensor(recv, send)          # the one NCCL all-gather of config #5
    else:
        dist.all_gather([recv[r * pad:(r + 1) * pad] for r in range(world)], send)
    return [recv[r * pad: r * pad + counts[r] * REC] for r in range(world)], counts


def records_from_bytes(t):
    """uint8 tensor -> numpy SiftPoint records (host copy)."""
    return np.frombuffer(t.cpu().numpy().tobytes(), dtype=SIFT_DTYPE)


def all_pairs_plan(world, rank):
    """Rank g matches its own set against every other rank's set (row-sharded pair matrix)."""
    return [j for j in range(world) if j != rank]
