"""oracle/geom.py -- numpy restatement of ImproveHomography (TEST INFRASTRUCTURE ONLY).

Follows geomFuncs.cpp:6-72 of the reference line by line (float32 for den/dx/dy/err, float64 for the
8x8 normal equations), with numpy.linalg.solve standing in for cv::solve(DECOMP_CHOLESKY).

Pinned (round 2): the reference routine itself needs OpenCV's C++ headers, which this image lacks, but cv2
(OpenCV 4.13) is importable, so improve_homography_cv2() below restates geomFuncs.cpp statement by statement
around the very cv::solve(DECOMP_CHOLESKY) the reference calls; tests/golden/improve_homography.npz holds its
outputs (tests/golden/make_geom_golden.py) and tests/test_homography.py checks this vectorised restatement and
the product's ImproveHomography against them.
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


def improve_homography_cv2(pts, homography, numLoops=5, minScore=0.0, maxAmbiguity=0.80, thresh=3.0):
    """The reference routine statement by statement (geomFuncs.cpp:6-72), with the reference's own third-party
    solver: cv2.solve(M, X, flags=cv2.DECOMP_CHOLESKY) is the Python binding of the cv::solve call at :55
    (OpenCV 4.13 in this image; the reference pins no version).  Point-by-point accumulation in the reference's
    order and types (float products for Y[6], Y[7]; double for M, X; float for den/dx/dy/err).  This is what
    pins improve_homography() above and the product's ImproveHomography: tests/golden/improve_homography.npz is
    generated from it (tests/golden/make_geom_golden.py) and tests/test_homography.py re-runs it when cv2 imports."""
    import cv2
    f32, f64 = np.float32, np.float64
    h = np.asarray(homography, f32).reshape(9)
    A = np.array([f64(h[i] / h[8]) for i in range(8)], f64).reshape(8, 1)      # :20-21 (float division)
    limit = f32(thresh) * f32(thresh)
    minScore, maxAmbiguity = f32(minScore), f32(maxAmbiguity)
    xs, ys = pts["xpos"].astype(f32), pts["ypos"].astype(f32)
    us, vs = pts["match_xpos"].astype(f32), pts["match_ypos"].astype(f32)
    sc, am = pts["score"].astype(f32), pts["ambiguity"].astype(f32)

    def residual(A, x, y, u, v, one):
        a = A[:, 0]
        den = f32(a[6] * f64(x) + a[7] * f64(y) + one)                         # :29 / :61
        dx = f32((a[0] * f64(x) + a[1] * f64(y) + a[2]) / f64(den) - f64(u))
        dy = f32((a[3] * f64(x) + a[4] * f64(y) + a[5]) / f64(den) - f64(v))
        return f32(f32(dx * dx) + f32(dy * dy))

    for _ in range(numLoops):
        M = np.zeros((8, 8), f64)
        X = np.zeros((8, 1), f64)
        for i in range(len(pts)):
            x, y, u, v = xs[i], ys[i], us[i], vs[i]
            if sc[i] < minScore or am[i] > maxAmbiguity:                       # :26
                continue
            err = residual(A, x, y, u, v, f64(f32(1.0)))
            wei = f64(1.0) if err < limit else f64(0.0)                        # :33
            Y = np.array([x, y, 1.0, 0.0, 0.0, 0.0, -f32(x * u), -f32(y * u)], f64)   # :34-39
            M += np.outer(Y, Y) * wei                                          # :40-42
            X += (Y * f64(u) * wei).reshape(8, 1)                              # :43
            Y = np.array([0.0, 0.0, 0.0, x, y, 1.0, -f32(x * v), -f32(y * v)], f64)   # :44-49
            M += np.outer(Y, Y) * wei
            X += (Y * f64(v) * wei).reshape(8, 1)
        ok, sol = cv2.solve(M, X, flags=cv2.DECOMP_CHOLESKY)                   # :55
        A = sol if ok else np.zeros((8, 1), f64)
    numfit = 0
    errs = np.zeros(len(pts), f32)
    for i in range(len(pts)):
        err = residual(A, xs[i], ys[i], us[i], vs[i], f64(1.0))
        numfit += int(err < limit)                                             # :65-66
        errs[i] = np.sqrt(err)                                                 # :67
    out = np.concatenate([A[:, 0], [1.0]]).astype(f32)                         # :69-71
    return out, numfit, errs
