"""oracle/geom.py -- numpy restatement of ImproveHomography (TEST INFRASTRUCTURE ONLY).

Follows geomFuncs.cpp:6-72 of the reference line by line (float32 for den/dx/dy/err, float64 for the
8x8 normal equations), with numpy.linalg.solve standing in for cv::solve(DECOMP_CHOLESKY).

PARITY UNPINNED: the reference routine needs OpenCV, which is not in this image, so the
reference itself cannot be run here and it has no golden vectors; the restatement is checked only
against geometry with a known answer (tests/test_homography.py).
"""
import numpy as np


def improve_homography(pts, homography, numLoops=5, minScore=0.0, maxAmbiguity=0.80, thresh=3.0):
    """Returns (9-float homography, numfit, match_error array); `pts` is a SIFT_DTYPE record array."""
    f32 = np.float32
    h = np.asarray(homography, f32).reshape(9)
    A = (h[:8] / h[8]).astype(np.float64)                      # geomFuncs.cpp:20-21
    limit = f32(thresh) * f32(thresh)
    x, y = pts["xpos"].astype(f32), pts["ypos"].astype(f32)
    u, v = pts["match_xpos"].astype(f32), pts["match_ypos"].astype(f32)
    xd, yd = x.astype(np.float64), y.astype(np.float64)

    def err2(A):
        den = (A[6] * xd + A[7] * yd + 1.0).astype(f32)        # :29
        dx = ((A[0] * xd + A[1] * yd + A[2]) / den.astype(np.float64) - u).astype(f32)
        dy = ((A[3] * xd + A[4] * yd + A[5]) / den.astype(np.float64) - v).astype(f32)
        return dx * dx + dy * dy                                # :32, float

    gate = ~((pts["score"] < f32(minScore)) | (pts["ambiguity"] > f32(maxAmbiguity)))   # :26
    for _ in range(numLoops):
        w = gate & (err2(A) < limit)                            # :33
        n = int(w.sum())
        one, zero = np.ones(n), np.zeros(n)
        xs, ys = xd[w], yd[w]
        Yu = np.stack([xs, ys, one, zero, zero, zero, -(x[w] * u[w]).astype(np.float64), -(y[w] * u[w]).astype(np.float64)], 1)
        Yv = np.stack([zero, zero, zero, xs, ys, one, -(x[w] * v[w]).astype(np.float64), -(y[w] * v[w]).astype(np.float64)], 1)
        M = Yu.T @ Yu + Yv.T @ Yv                               # :40-43, :52-55
        X = Yu.T @ u[w].astype(np.float64) + Yv.T @ v[w].astype(np.float64)
        try:
            np.linalg.cholesky(M)
            A = np.linalg.solve(M, X)                           # :57
        except np.linalg.LinAlgError:
            A = np.zeros(8)                                     # cv::solve leaves zeros when the factorisation fails
    e = err2(A)
    numfit = int((e < limit).sum())                             # :66-67
    out = np.concatenate([A, [1.0]]).astype(f32)
    return out, numfit, np.sqrt(e).astype(f32)
