#!/usr/bin/env python
"""Generate the golden fixtures from the UNMODIFIED reference library.

Run on the GPU box (the reference is CUDA code):   python tests/golden/make_golden.py
Reads oracle/_ref/libcudasift_ref.so (built from /root/reference by cudasift_b200/build.py),
feeds it seeded synthetic inputs (regenerable from cudasift_b200/synth.py) plus a crop of
the reference's demo image data/left.pgm, and writes small .npz files to
gpurun_out/golden/, which are then committed under tests/golden/.

Fixtures pin the oracle (tests/test_golden_cpu.py) and the CUDA path (test_golden_gpu.py):
  stages.npz   sha256 + 32x32 crops of LowPass / ScaleDown / ScaleUp / 7 DoG planes (bit-exact)
  extract_*.npz  canonical-sorted SiftPoint records of ExtractSift (two reference runs, so the
                 reference's own run-to-run noise is on file)
  match.npz    the five MatchSiftData output fields on seeded descriptor sets (bit-exact)
"""
import hashlib
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import cudasift_b200 as cs            # noqa: E402
import reflib                         # noqa: E402
from cudasift_b200 import build       # noqa: E402
from cudasift_b200.synth import synth_descriptors, synth_image   # noqa: E402

OUT = os.path.join(ROOT, "gpurun_out", "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def canon(p):
    return p[np.lexsort((p["orientation"], p["scale"], p["xpos"], p["ypos"], p["subsampling"]))]


def slim(p):
    """Keep the extraction outputs only (match fields are uninitialised device memory)."""
    q = np.zeros(len(p), cs.SIFT_DTYPE)
    for f in ("xpos", "ypos", "scale", "sharpness", "edgeness", "orientation", "subsampling", "data"):
        q[f] = p[f]
    return q


def main():
    os.makedirs(OUT, exist_ok=True)
    cs.InitCuda(0)
    ref = reflib.load_reference()
    assert ref is not None, "reference library missing: run cudasift_b200/build.py where /root/reference exists"
    # ---- image stages ----
    img = synth_image(320, 240, seed=7)
    st = {"image_seed": 7, "image_size": (320, 240)}
    lp = ref.lowpass(img, 1.0)
    sd = ref.scaledown(img)
    su = ref.scaleup(np.ascontiguousarray(img[:100, :128]))
    st.update(lowpass_sha=sha(lp), lowpass_crop=lp[100:132, 100:132], scaledown_sha=sha(sd),
              scaledown_crop=sd[40:72, 40:72], scaleup_sha=sha(su), scaleup_crop=su[:32, :32])
    for octave in (5, 3):
        d = ref.dog(img, 5, octave)
        st["dog%d_sha" % octave] = sha(d)
        st["dog%d_crop" % octave] = d[:, 100:132, 100:132]
    np.savez_compressed(os.path.join(OUT, "stages.npz"), **st)
    # ---- extraction ----
    cases = {"synth320": (img, 3.0), "synth640": (synth_image(640, 480, seed=1000), 3.0)}
    left = os.path.join(build.REF_DIR, "data", "left.pgm")
    if os.path.exists(left):
        import cv2
        cases["left_crop"] = (np.ascontiguousarray(cv2.imread(left, 0).astype(np.float32)[200:680, 300:940]), 4.5)
    for name, (im, thresh) in cases.items():
        a = slim(canon(ref.extract(im, thresh=thresh)))
        b = slim(canon(ref.extract(im, thresh=thresh)))
        extra = {"image": im.astype(np.uint8)} if name == "left_crop" else {}
        np.savez_compressed(os.path.join(OUT, "extract_%s.npz" % name), run1=a, run2=b, thresh=thresh,
                            shape=im.shape, **extra)
        print(name, len(a), len(b))
    # ---- matching ----
    m = {}
    for n1, n2, sa, sb in ((300, 352, 3, 4), (1000, 1031, 5, 6)):
        s1, s2 = synth_descriptors(n1, sa), synth_descriptors(n2, sb)
        out, _ = ref.match(s1, s2)
        for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
            m["%dx%d_%s" % (n1, n2, f)] = out[f]
        m["%dx%d_seeds" % (n1, n2)] = (sa, sb)
    np.savez_compressed(os.path.join(OUT, "match.npz"), **m)
    print("golden fixtures written to", OUT)


if __name__ == "__main__":
    main()
