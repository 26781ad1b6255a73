#!/usr/bin/env python
"""Golden vectors for ImproveHomography.  Generated in the build container with OpenCV's own
cv::solve(DECOMP_CHOLESKY) through cv2 -- the one third-party call of geomFuncs.cpp:55 -- inside a statement-by-
statement port of the routine (oracle/geom.py: improve_homography_cv2).

  python tests/golden/make_geom_golden.py      -> tests/golden/improve_homography.npz
"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

CASES = ((5, 0.0, 0.80, 3.0), (1, 0.85, 0.95, 5.0), (8, 0.0, 1.0, 2.0), (2, 2.0, 0.0, 3.0))


def start_homography(H_TRUE):
    H0 = H_TRUE.copy(); H0[0, 2] += 1.5; H0[1, 2] -= 1.0; H0[0, 0] *= 1.001
    return (H0 * 1.7).astype(np.float32)


def main():
    import cv2
    from oracle.geom import improve_homography_cv2
    from test_homography import H_TRUE, planted
    p, _ = planted(n=900, seed=11, noise=0.25)
    H0 = start_homography(H_TRUE)
    out = {"points": p, "H0": H0, "cases": np.array(CASES, np.float64), "opencv": np.array(cv2.__version__)}
    for k, (loops, mins, maxa, thr) in enumerate(CASES):
        H, nfit, err = improve_homography_cv2(p, H0, int(loops), mins, maxa, thr)
        out["H_%d" % k], out["numfit_%d" % k], out["err_%d" % k] = H, np.int32(nfit), err
        print(k, nfit, H)
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", "improve_homography.npz"), **out)


if __name__ == "__main__":
    main()
