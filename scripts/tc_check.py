"""GPU-box check of the tcgen05 matcher against the oracle (sizes given on the command line)."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudasift_b200 as cs, oracle
from cudasift_b200.synth import synth_descriptors
cs.InitCuda(0)
sizes = [tuple(int(v) for v in a.split("x")) for a in sys.argv[1:]] or [(256, 256)]
for n1, n2 in sizes:
    s1, s2 = synth_descriptors(n1, 1), synth_descriptors(n2, 2)
    want = oracle.match(s1, s2, threads=16)
    got, ms = cs.match_host(s1, s2, mode=2)
    st = cs.match_stats()
    bad = {f: int((got[f] != want[f]).sum()) for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos")}
    print("%dx%d ms=%.3f stats(emitted,chains,fallback,path)=%s mismatches=%s" % (n1, n2, ms, st, bad), flush=True)
    if bad["match"]:
        idx = np.nonzero(got["match"] != want["match"])[0][:5]
        for i in idx:
            print("  row", i, "got", got["match"][i], got["score"][i], "want", want["match"][i], want["score"][i])
