"""Small workload for ncu: a few 1080p extractions (single stream) and 10k x 10k matches."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudasift_b200 as cs
from cudasift_b200.synth import synth_descriptors, synth_image
what = sys.argv[1] if len(sys.argv) > 1 else "both"
cs.InitCuda(0)
if what in ("extract", "both"):
    W, H = 1920, 1080
    imgs = [synth_image(W, H, seed=1000 + i) for i in range(2)]
    d = [cs.CudaImage().Allocate(W, H, 1920, False, None, im) for im in imgs]
    [x.Download() for x in d]
    ex = cs.Extractor(W, H, 5, 32768)
    for i in range(6):
        ex.submit_device(d[i % 2].d_data, 1920, 1.0, 3.0, 0.0)
        print("pts", ex.wait())
if what in ("match", "both"):
    n = 10000
    s1, s2 = synth_descriptors(n, 1), synth_descriptors(n, 2)
    d1 = cs.InitSiftData(cs.SiftData(), n, False, True); d2 = cs.InitSiftData(cs.SiftData(), n, False, True)
    d1._buf.upload(s1); d2._buf.upload(s2); d1.numPts = d2.numPts = n
    for mode in (2, 2, 2, 1):
        print("match mode", mode, cs.MatchSiftData(d1, d2, mode=mode), cs.match_stats())
