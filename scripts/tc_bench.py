"""Steady-state timing of the tensor matcher (default 10k x 10k), single-pass vs exact path, with a
bit-exact comparison of the five output fields.  Usage: python scripts/tc_bench.py [n]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudasift_b200 as cs
from cudasift_b200.synth import synth_descriptors
cs.InitCuda(0)
for n in ([int(a) for a in sys.argv[1:]] or [2000, 10000]):
    s1, s2 = synth_descriptors(n, 1), synth_descriptors(n, 2)
    a, _ = cs.match_host(s1, s2, mode=1)
    b, _ = cs.match_host(s1, s2, mode=2)
    st = cs.match_stats()
    bad = {f: int((a[f] != b[f]).sum()) for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos")}
    d1 = cs.InitSiftData(cs.SiftData(), n, False, True); d2 = cs.InitSiftData(cs.SiftData(), n, False, True)
    d1._buf.upload(s1); d2._buf.upload(s2); d1.numPts = d2.numPts = n
    for mode in (2, 1):
        for _ in range(3): cs.MatchSiftData(d1, d2, mode=mode)
        ts = [cs.MatchSiftData(d1, d2, mode=mode) for _ in range(20)]
        print("n %d mode %d median ms %.4f  min %.4f  Gpair/s %.0f" % (n, mode, np.median(ts), min(ts), n * n / np.median(ts) / 1e6), flush=True)
    print("n %d mismatches vs exact path %s  stats(groups, chains, fallback rows, mode) %s" % (n, bad, st), flush=True)
