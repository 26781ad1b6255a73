#!/usr/bin/env python
"""BASELINE.json config #5: per-GPU ExtractSift + ONE NCCL all-gather of the SiftPoint arrays +
all-pairs MatchSiftData (rank g matches its image against every other rank's).

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
         --master-port 29511 scripts/allpairs.py [--check]

Device buffers are torch tensors (torch owns memory and the NCCL group); the compute goes through
the C ABI with raw pointers.  --check verifies every pair against the CPU oracle on rank 0's side.
"""
import argparse
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudasift_b200 as cs                     # noqa: E402
from cudasift_b200 import multi                # noqa: E402
from cudasift_b200.synth import synth_image    # noqa: E402

W, H, MAXPTS = 1920, 1080, 32768


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--check", action="store_true")
    ap.add_argument("--reps", type=int, default=5)
    args = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    cs.InitCuda(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    L = cs.lib()
    img = synth_image(W, H, seed=1000 + rank)
    d_img = torch.from_numpy(np.ascontiguousarray(img)).cuda()                       # pitch == width (1920 % 128 == 0)
    d_pts = torch.zeros(MAXPTS * cs.SIFT_DTYPE.itemsize, dtype=torch.uint8, device="cuda")
    times = []
    for rep in range(args.reps + 1):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        n = L.cs_extract(d_img.data_ptr(), W, H, W, 5, 1.0, 3.0, 0.0, 0, None, d_pts.data_ptr(), None, MAXPTS)
        assert n > 0, L.cs_last_error()
        t1 = time.perf_counter()
        if world > 1:
            parts, counts = multi.allgather_records(dist, d_pts, n)                  # the single exchange
        else:
            parts, counts = [d_pts[:n * 576]], [n]
        torch.cuda.synchronize()
        t2 = time.perf_counter()
        results = {}
        mine = d_pts
        for j in multi.all_pairs_plan(world, rank):
            other = parts[j].contiguous()
            ms = ctypes.c_double(0)
            r = L.cs_match(mine.data_ptr(), n, other.data_ptr(), counts[j], None, 0, ctypes.byref(ms))
            assert r == 0, L.cs_last_error()
            results[j] = np.frombuffer(mine[:n * 576].cpu().numpy().tobytes(), dtype=cs.SIFT_DTYPE)[["score", "ambiguity", "match"]].copy()
        torch.cuda.synchronize()
        t3 = time.perf_counter()
        if rep > 0:
            times.append((t1 - t0, t2 - t1, t3 - t2))
    t = np.median(np.array(times), axis=0) * 1e3
    tt = torch.tensor([t.sum()], device="cuda", dtype=torch.float64)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    ok = None
    if args.check:
        import oracle
        ok = True
        mine_h = multi.records_from_bytes(d_pts[:n * 576])
        for j in multi.all_pairs_plan(world, rank):
            want = oracle.match(mine_h, multi.records_from_bytes(parts[j]), threads=8)
            ok = ok and bool(np.array_equal(want["match"], results[j]["match"]) and np.array_equal(want["score"], results[j]["score"]))
    if rank == 0:
        print(json.dumps({"config": "all-pairs match, %d images 1920x1080" % world, "n_gpus": world, "features": counts,
                          "ms_extract": round(float(t[0]), 3), "ms_allgather": round(float(t[1]), 3),
                          "ms_match_all_pairs": round(float(t[2]), 3), "ms_total_max_over_ranks": round(float(tt.item()), 3),
                          "allgather_bytes_per_rank": int(max(counts)) * 576, "check_vs_oracle": ok}), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
