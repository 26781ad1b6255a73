#!/usr/bin/env python
"""bench.py -- headline benchmark of cudasift_b200 (BASELINE.json metric, config #2).

  python bench.py --gpus N --steps K --warmup W            # the product (default N=1)
  python bench.py --impl reference --gpus N ...            # the unmodified reference build
  (N>1: launched by `python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N`)

A "step" is one pass of ExtractSift over a batch of synthetic 1920x1080 float images that
are resident in HBM (32 distinct device buffers = 265 MB > the 126 MB L2, so every step
re-reads its inputs from DRAM).  `value` = images/s with device-resident inputs through the
pipelined extractor API; `e2e` = the same metric through the C-ABI call with HOST buffers
(pinned), H2D of every image and D2H of every result inside the timed region.
Extra objects: `roofline` (dominant kernel = the fused blur+DoG+extrema detector),
`cpu_baseline` (the oracle port timed on the host cores; rank 0, N=1 only), `match`
(MatchSiftData 10k x 10k, BASELINE.json config #3).  One JSON line on stdout (rank 0).
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

W, H, OCTAVES, INIT_BLUR, THRESH, MAX_PTS = 1920, 1080, 5, 1.0, 3.0, 32768
# dram__bytes_read.sum + dram__bytes_write.sum of detect_kernel per launch (ncu --set full, 1080p, profiles/)
DETECT_DRAM_BYTES = 11.10e6


def level_sizes(w=W, h=H, n=OCTAVES):
    out = []
    for _ in range(n):
        out.append((w, h))
        w, h = w // 2, h // 2
    return out


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json, burst copy)", d
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)", {}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.gpu), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1])); mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def dist_env():
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    return rank, world, local


def init_dist(world, local):
    if world <= 1:
        return None
    import torch
    import torch.distributed as dist
    torch.cuda.set_device(local)
    os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    return dist


def barrier_and_sync(dist):
    import cudasift_b200 as cs
    cs.lib().cs_device_sync()
    if dist is not None:
        import torch
        dist.barrier()
        torch.cuda.synchronize()


def reduce_max(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def reduce_sum(dist, x):
    if dist is None:
        return x
    import torch
    t = torch.tensor([x], dtype=torch.float64, device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def make_images(rank, distinct):
    from cudasift_b200.synth import synth_image
    return [synth_image(W, H, seed=1000 + 100 * rank + i) for i in range(distinct)]


# ------------------------------------------------------------------------------------------
def run_product(args):
    import cudasift_b200 as cs
    rank, world, local = dist_env()
    cs.InitCuda(local)
    dist = init_dist(world, local)
    L = cs.lib()
    B, S = args.batch, args.streams
    pitch = cs.iAlignUp(W, 128)
    imgs = make_images(rank, args.distinct)
    dbufs = []
    for i in range(B):                                # B distinct device buffers (> L2)
        img = cs.CudaImage().Allocate(W, H, pitch, False, None, imgs[i % len(imgs)])
        img.Download()
        dbufs.append(img)
    exs = [cs.Extractor(W, H, OCTAVES, MAX_PTS) for _ in range(S)]
    ev0 = [L.cs_event_create() for _ in range(S)]
    ev1 = [L.cs_event_create() for _ in range(S)]

    def step_device():
        for i in range(B):
            exs[i % S].submit_device(dbufs[i].d_data, pitch, INIT_BLUR, THRESH, 0.0)

    for _ in range(args.warmup):
        step_device()
    counts = [ex.wait() for ex in exs]
    # --- timed region 1: device-resident inputs ---
    sampler = ClockSampler(local)
    barrier_and_sync(dist)
    sampler.start()
    launches0 = L.cs_launch_count()
    for s in range(S):
        L.cs_event_record(ev0[s], exs[s].handle)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step_device()
    for s in range(S):
        L.cs_event_record(ev1[s], exs[s].handle)
    counts = [ex.wait() for ex in exs]
    barrier_and_sync(dist)
    wall = time.perf_counter() - t0
    dev_ms = max(L.cs_event_elapsed_ms(ev0[0], ev1[s]) for s in range(S))
    launches = L.cs_launch_count() - launches0
    clocks = sampler.stop()
    dev_ms = reduce_max(dist, dev_ms)
    wall = reduce_max(dist, wall)
    n_images = args.steps * B * world
    value = n_images / (dev_ms / 1e3)
    pts_per_image = float(np.mean(counts))

    # --- timed region 2: end to end through the C ABI with host (pinned) buffers ---
    himgs = []
    for i in range(min(B, 8)):
        p = L.cs_host_alloc_pinned(W * H * 4)
        ctypes.memmove(p, imgs[i % len(imgs)].ctypes.data, W * H * 4)
        himgs.append(p)
    e2e_images = max(B, 16) * max(1, min(args.steps, 100) // 2)

    def run_e2e(n):
        busy = [False] * S
        d2h = 0
        for i in range(n):
            s = i % S
            if busy[s]:
                d2h += exs[s].wait() * 576 + 8
            exs[s].submit_host(himgs[i % len(himgs)], INIT_BLUR, THRESH, 0.0)
            busy[s] = True
        for s in range(S):
            if busy[s]:
                d2h += exs[s].wait() * 576 + 8
        return d2h

    run_e2e(2 * S)
    barrier_and_sync(dist)
    t0 = time.perf_counter()
    d2h_bytes = run_e2e(e2e_images)
    barrier_and_sync(dist)
    e2e_s = reduce_max(dist, time.perf_counter() - t0)
    e2e_value = e2e_images * world / e2e_s

    # --- extension: 8-bit uploads (4x less PCIe traffic, exact conversion on the device) ---
    h8 = []
    for i in range(min(B, 8)):
        p = L.cs_host_alloc_pinned(W * H)
        a8 = np.clip(np.rint(imgs[i % len(imgs)]), 0, 255).astype(np.uint8)
        ctypes.memmove(p, a8.ctypes.data, W * H)
        h8.append(p)

    def run_e2e_u8(n):
        busy = [False] * S
        for i in range(n):
            s = i % S
            if busy[s]:
                exs[s].wait()
            exs[s].submit_host_u8(h8[i % len(h8)], INIT_BLUR, THRESH, 0.0)
            busy[s] = True
        for s in range(S):
            if busy[s]:
                exs[s].wait()
    run_e2e_u8(2 * S)
    barrier_and_sync(dist)
    t0 = time.perf_counter()
    run_e2e_u8(e2e_images)
    barrier_and_sync(dist)
    e2e_u8_value = e2e_images * world / reduce_max(dist, time.perf_counter() - t0)

    # --- synchronous classic call, one image at a time (reference-shaped usage) ---
    hp = np.zeros(MAX_PTS, cs.SIFT_DTYPE)
    t0 = time.perf_counter()
    nsync = 16
    for i in range(2):       # warm the internal pipeline of the synchronous path
        L.cs_extract_host(himgs[i % len(himgs)], W, H, OCTAVES, INIT_BLUR, THRESH, 0.0, 0, hp.ctypes.data, MAX_PTS)
    t0 = time.perf_counter()
    for i in range(nsync):
        L.cs_extract_host(himgs[i % len(himgs)], W, H, OCTAVES, INIT_BLUR, THRESH, 0.0, 0, hp.ctypes.data, MAX_PTS)
    sync_ms = (time.perf_counter() - t0) / nsync * 1e3

    out = {
        "metric": "1920x1080 images/sec ExtractSift", "value": round(value, 1), "unit": "images/s",
        "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(dev_ms / args.steps, 4), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": "ExtractSift 1920x1080 float, 5 octaves, initBlur 1.0, thresh 3.0 (BASELINE config #2)",
                   "images_per_step_per_gpu": B, "streams": S, "distinct_images": len(imgs),
                   "l2_policy": "inputs larger than L2 (%d device images = %.0f MB per GPU)" % (B, B * pitch * H * 4 / 1e6),
                   "features_per_image": round(pts_per_image, 1), "parallelism": "images sharded one process per GPU, no collective"},
        "clocks": clocks,
        "e2e": {"value": round(e2e_value, 1), "unit": "images/s", "h2d_bytes_per_step": B * W * H * 4,
                "d2h_bytes_per_step": int(d2h_bytes / e2e_images * B), "images": e2e_images * world,
                "api": "cs_extractor_submit_host/cs_extractor_wait (pinned host buffers, %d in flight)" % S,
                "sync_call_ms": round(sync_ms, 3),
                "u8_upload_images_per_s": round(e2e_u8_value, 1)},
        "gpu_launches": int(launches * world),
        "wall_s": round(wall, 4),
    }
    if rank == 0:
        # ---- roofline of the dominant kernel (single stream, CUDA events at stage boundaries) ----
        prof = []
        for i in range(min(B, 16)):
            n, ms = exs[0].profile(dbufs[i].d_data, pitch, INIT_BLUR, THRESH, 0.0)
            prof.append(ms)
        prof = np.array(prof[2:])
        lowpass_ms, sd_ms, detect_ms, describe_ms, total_ms = prof.mean(axis=0)
        lv = level_sizes()
        detect_bytes = sum(4 * w * h for w, h in lv)                       # each octave base image read once
        pipeline_bytes = 4 * W * H + sum(2 * 4 * w * h for w, h in lv) + 576 * pts_per_image
        peak, how, _ = peaks()
        ach = detect_bytes / (detect_ms * 1e-3) / 1e9
        out["roofline"] = {"bound": "hbm", "kernel": "detect_kernel (8-scale blur + DoG + 3x3x3 extrema, all octaves)",
                           "achieved": round(ach, 1), "peak": peak, "unit": "GB/s", "frac": round(ach / peak, 4),
                           "traffic": DETECT_DRAM_BYTES, "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one "
                           "ncu --set full capture of detect_kernel (profiles/r01_prof_detect.txt)", "peak_source": how, "algorithmic_bytes_per_launch": detect_bytes,
                           "avg_launch_ms": round(float(detect_ms), 4),
                           "stage_ms": {"lowpass": round(float(lowpass_ms), 4), "scaledown_x4": round(float(sd_ms), 4),
                                        "detect": round(float(detect_ms), 4), "describe": round(float(describe_ms), 4),
                                        "pipeline_total": round(float(total_ms), 4)},
                           "lowpass_gbs": round(2 * 4 * W * H / (lowpass_ms * 1e-3) / 1e9, 1),
                           "pipeline_algorithmic_bytes": int(pipeline_bytes),
                           "pipeline_frac_at_value": round(pipeline_bytes * value / world / 1e9 / peak, 4)}
        out["match"] = bench_match(cs, args)
        if world == 1 and not args.no_cpu:
            out["cpu_baseline"] = cpu_baseline(imgs)
        print(json.dumps(out), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def bench_match(cs, args):
    """BASELINE.json config #3: MatchSiftData 10000 x 10000 synthetic 128-D descriptors."""
    from cudasift_b200.synth import synth_descriptors
    res = {}
    n = 10000
    s1, s2 = synth_descriptors(n, 1), synth_descriptors(n, 2)
    d1 = cs.InitSiftData(cs.SiftData(), n, False, True)
    d2 = cs.InitSiftData(cs.SiftData(), n, False, True)
    d1._buf.upload(s1); d2._buf.upload(s2)
    d1.numPts = d2.numPts = n
    for mode, name in ((1, "exact_fp32"), (2, "tensor")):
        for _ in range(3):
            cs.MatchSiftData(d1, d2, mode=mode)
        ts = [cs.MatchSiftData(d1, d2, mode=mode) for _ in range(20)]
        ms = float(np.median(ts))
        res[name] = {"ms": round(ms, 4), "gpair_per_s": round(n * n / (ms * 1e-3) / 1e9, 2),
                     "tflops_algorithmic": round(2 * n * n * 128 / (ms * 1e-3) / 1e12, 2)}
    res["stats_tensor"] = cs.match_stats()
    res["n"] = n
    # tensor roofline of the single-pass matcher: three FP16 MMAs (hi*hi + hi*lo + lo*hi) per algorithmic product
    _, _, pk = peaks()
    tpeak = float(pk.get("bf16_tflops", 0.0)) or 1693.0
    t = res["tensor"]
    res["tensor_roofline"] = {
        "bound": "tensor", "kernel": "t3_gemm_kernel (tcgen05.mma kind::f16, M128 N128 K16, A in TMEM, K_eff = 384)",
        "achieved_algorithmic": t["tflops_algorithmic"], "achieved_executed": round(3 * t["tflops_algorithmic"], 2),
        "peak": tpeak, "unit": "TFLOP/s", "frac_algorithmic": round(t["tflops_algorithmic"] / tpeak, 4),
        "frac_executed": round(3 * t["tflops_algorithmic"] / tpeak, 4),
        "peak_source": "MEASURED_PEAKS.json bf16_tflops (cuBLAS burst)" if pk else "fallback",
        "note": "whole MatchSiftData call (prep + GEMM + resolve + D2H of 5 fields); the GEMM kernel alone keeps the tensor "
                "pipe active 73 % of its cycles (profiles/r01_prof_match.txt, sm__pipe_tensor_cycles_active)"}
    return res


def cpu_baseline(imgs):
    """The oracle port (oracle/sift_oracle.c) on the host cores: one image per thread."""
    import oracle
    cores = os.cpu_count() or 1
    threads = min(cores, 32)
    per_thread = 8
    oracle.extract(imgs[0], OCTAVES, INIT_BLUR, THRESH)          # warm (loads the .so)

    def work(i):
        for j in range(per_thread):
            oracle.extract(imgs[(i + j) % len(imgs)], OCTAVES, INIT_BLUR, THRESH)
    t0 = time.perf_counter()
    ths = [threading.Thread(target=work, args=(i,)) for i in range(threads)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    dt = time.perf_counter() - t0
    n = threads * per_thread
    out = {"value": round(n / dt, 2), "unit": "images/s", "cores": threads, "kind": "port",
           "sample": "%d synthetic 1920x1080 images through oracle_extract (scalar C restatement), %d threads, %.1f s"
                     % (n, threads, dt)}
    out["opencv_sift"] = opencv_sift_baseline(imgs, cores)
    return out


def opencv_sift_baseline(imgs, cores, budget_s=6.0):
    """BASELINE.json's "scalar CPU OpenCV-SIFT baseline on the host cores, core count stated": a different
    algorithm variant (OpenCV's SIFT, 8-bit input), reported for orientation only; 1 thread, then all cores."""
    try:
        import cv2
    except Exception as e:                                    # not part of the contract: absent -> say so
        return {"unavailable": str(e)[:80]}
    im8 = [np.clip(np.rint(im), 0, 255).astype(np.uint8) for im in imgs[:4]]
    sift = cv2.SIFT_create(nOctaveLayers=5)
    res = {"host_cores": cores, "nOctaveLayers": 5}
    for name, nthreads in (("scalar", 1), ("all_cores", cores)):
        cv2.setNumThreads(nthreads)
        sift.detectAndCompute(im8[0], None)                   # warm
        t0, k, nk = time.perf_counter(), 0, 0
        while time.perf_counter() - t0 < budget_s / 2 and k < 64:
            kp, _ = sift.detectAndCompute(im8[k % len(im8)], None)
            nk += len(kp); k += 1
        dt = time.perf_counter() - t0
        res[name] = {"images_per_s": round(k / dt, 2), "threads": nthreads, "images": k, "features_per_image": round(nk / max(k, 1))}
    return res


# ------------------------------------------------------------------------------------------
def run_reference(args):
    """The unmodified reference (oracle/_ref/libcudasift_ref.so, sm_100 build of
    Celebrandil/CudaSift) through its own C++ API, same metric and config.  The reference has
    no CPU implementation of this path; if the library did not travel, the oracle port is
    timed on the host cores instead."""
    rank, world, local = dist_env()
    import cudasift_b200 as cs
    import reflib
    imgs = make_images(rank, args.distinct)
    path = reflib.REF_LIB
    if not os.path.exists(path):
        if rank == 0:
            cb = cpu_baseline(imgs)
            print(json.dumps({"impl": "reference", "metric": "1920x1080 images/sec ExtractSift", "value": cb["value"],
                              "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                              "higher_is_better": True, "scaling": "weak", "dtype": "f32", "data": "synthetic",
                              "config": {"workload": "ExtractSift 1920x1080 float, 5 octaves (oracle port on CPU)"},
                              "cpu_baseline": cb, "e2e": {"value": cb["value"], "unit": "images/s",
                                                          "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}), flush=True)
        return
    cs.InitCuda(local)                       # selects the device for this process
    dist = init_dist(world, local)
    ref = reflib.CxxSiftLib(path, device=local)
    B = args.batch
    c = ctypes
    images = [ref.image(imgs[i % len(imgs)]) for i in range(B)]
    sd = reflib.CSiftData()
    ref._init(c.byref(sd), MAX_PTS, True, True)
    tmp = ref._alloc(W, H, OCTAVES, False)

    def step(download):
        for i in range(B):
            if download:
                ref._imgdown(c.byref(images[i]))
            ref._extract(c.byref(sd), c.byref(images[i]), OCTAVES, INIT_BLUR, THRESH, 0.0, False, tmp)
    with reflib.quiet_stdout():
        for _ in range(max(1, args.warmup)):
            step(False)
        barrier_and_sync(dist)
        t0 = time.perf_counter()
        for _ in range(args.steps):
            step(False)
        barrier_and_sync(dist)
        dt = time.perf_counter() - t0
        npts = sd.numPts
        t0 = time.perf_counter()
        esteps = max(1, args.steps // 2)
        for _ in range(esteps):
            step(True)
        barrier_and_sync(dist)
        dte = time.perf_counter() - t0
    dt, dte = reduce_max(dist, dt), reduce_max(dist, dte)
    match = None
    if rank == 0:
        from cudasift_b200.synth import synth_descriptors
        n = 10000
        s1, s2 = synth_descriptors(n, 1), synth_descriptors(n, 2)
        ts = [ref.match(s1, s2)[1] for _ in range(6)][1:]
        ms = float(np.median(ts))
        match = {"n": n, "ms": round(ms, 4), "gpair_per_s": round(n * n / (ms * 1e-3) / 1e9, 2),
                 "how": "reference MatchSiftData (FindMaxCorr10), its own TimerGPU incl. the 5-field D2H copy"}
    value = args.steps * B * world / dt
    e2e = esteps * B * world / dte
    if rank == 0:
        print(json.dumps({
            "impl": "reference", "metric": "1920x1080 images/sec ExtractSift", "value": round(value, 1),
            "unit": "images/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": round(dt / args.steps * 1e3, 4), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "ExtractSift 1920x1080 float, 5 octaves, initBlur 1.0, thresh 3.0 (BASELINE config #2)",
                       "images_per_step_per_gpu": B, "features_per_image": npts,
                       "how": "unmodified Celebrandil/CudaSift built for sm_100 (oracle/_ref), its own ExtractSift "
                              "loop as in mainSift.cpp:65-69, pre-allocated temp memory, one process per GPU"},
            "cpu_baseline": {"value": round(value, 1), "unit": "images/s", "cores": 1, "kind": "reference",
                             "sample": "the reference is a CUDA library: timed on the GPU (1 host thread drives it), "
                                       "%d images" % (args.steps * B)},
            "e2e": {"value": round(e2e, 1), "unit": "images/s", "h2d_bytes_per_step": B * W * H * 4,
                    "d2h_bytes_per_step": B * npts * 576,
                    "api": "CudaImage::Download + ExtractSift (host copy of the points included, cudaSiftH.cu:139-140)"},
            "match": match,
        }), flush=True)
    if dist is not None:
        dist.barrier()
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=32, help="images per step per GPU")
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--distinct", type=int, default=8, help="distinct synthetic images per rank")
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "b200" else args.warmup
    if args.impl == "reference":
        run_reference(args)
    else:
        run_product(args)


if __name__ == "__main__":
    main()
