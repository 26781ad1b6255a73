"""Profiling driver (run under ncu): a few batched extractions of synthetic 1080p images."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudasift_b200 as cs
from cudasift_b200.synth import synth_image
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 3
cs.InitCuda(0)
w, h = 1920, 1080
pitch = cs.iAlignUp(w, 128)
imgs = [synth_image(w, h, seed=1000 + i) for i in range(min(B, 8))]
cis = []
for i in range(B):
    ci = cs.CudaImage().Allocate(w, h, pitch, False, None, imgs[i % len(imgs)])
    ci.Download()
    cis.append(ci)
ex = cs.Extractor(w, h, 5, 32768, False, batch=B)
for r in range(reps):
    n, ms = ex.profile_batch([c.d_data for c in cis], pitch, 1.0, 3.0, 0.0)
    print(r, n, [round(x * 1e3 / B, 1) for x in ms])
