#!/usr/bin/env python
"""Time the compiled detector variants (cs_set_tuning "detect_variant"): isolated detect_kernel time
(single stream, CUDA events at the stage boundaries) and whole-pipeline throughput (8 pipelined
extractors, device-resident 1080p inputs).  Usage: python scripts/detect_variants.py [ids...]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudasift_b200 as cs
from cudasift_b200.synth import synth_image

W, H, S, B, STEPS = 1920, 1080, 8, 32, 10
ids = [int(a) for a in sys.argv[1:]] or list(range(8))   # id = variant + 100 * skip mask
cs.InitCuda(0)
L = cs.lib()
pitch = cs.iAlignUp(W, 128)
imgs = [synth_image(W, H, seed=1000 + i) for i in range(8)]
dbufs = []
for i in range(B):
    im = cs.CudaImage().Allocate(W, H, pitch, False, None, imgs[i % 8]); im.Download(); dbufs.append(im)
ref_counts = None
for v in ids:
    assert L.cs_set_tuning(b"detect_variant", v % 100) == 0 and L.cs_set_tuning(b"detect_skip", v // 100) == 0
    exs = [cs.Extractor(W, H, 5, 32768) for _ in range(S)]
    ev0, ev1 = L.cs_event_create(), [L.cs_event_create() for _ in range(S)]
    def step():
        for i in range(B):
            exs[i % S].submit_device(dbufs[i].d_data, pitch, 1.0, 3.0, 0.0)
    for _ in range(3):
        step()
    counts = [ex.wait() for ex in exs]
    L.cs_device_sync()
    L.cs_event_record(ev0, exs[0].handle)
    for _ in range(STEPS):
        step()
    for s in range(S):
        L.cs_event_record(ev1[s], exs[s].handle)
    counts = [ex.wait() for ex in exs]
    L.cs_device_sync()
    ms = max(L.cs_event_elapsed_ms(ev0, ev1[s]) for s in range(S))
    prof = np.array([exs[0].profile(dbufs[i].d_data, pitch, 1.0, 3.0, 0.0)[1] for i in range(12)][2:]).mean(axis=0)
    one = [exs[0].profile(dbufs[i].d_data, pitch, 1.0, 3.0, 0.0)[0] for i in range(8)]
    if ref_counts is None:
        ref_counts = one
    print("variant %d: detect %.1f us  pipeline(single stream) %.1f us  throughput %.0f img/s  counts_equal %s" % (
        v, prof[2] * 1e3, prof[4] * 1e3, STEPS * B / (ms / 1e3), one == ref_counts), flush=True)
    del exs
