"""Steady-state timing of the tensor matcher (10k x 10k) under the CS_TC_DBG experiments."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudasift_b200 as cs
from cudasift_b200.synth import synth_descriptors
cs.InitCuda(0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 10000
s1, s2 = synth_descriptors(n, 1), synth_descriptors(n, 2)
d1 = cs.InitSiftData(cs.SiftData(), n, False, True); d2 = cs.InitSiftData(cs.SiftData(), n, False, True)
d1._buf.upload(s1); d2._buf.upload(s2); d1.numPts = d2.numPts = n
for dbg in ("0", "1", "2"):
    os.environ["CS_TC_DBG"] = dbg
    for _ in range(3): cs.MatchSiftData(d1, d2, mode=2)
    ts = [cs.MatchSiftData(d1, d2, mode=2) for _ in range(20)]
    print("dbg", dbg, "median ms %.4f  min %.4f" % (np.median(ts), min(ts)), cs.match_stats(), flush=True)
os.environ["CS_TC_DBG"] = "0"
