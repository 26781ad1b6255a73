"""Keypoint-set comparison: canonical one-to-one pairing with tolerance."""
import numpy as np


def pair_points(a, b, pos_tol=0.05, scale_rel=0.02, ori_tol=2.0):
    """Greedy one-to-one pairing of records a -> b on (subsampling, x, y, scale, orientation).
    Returns (index pairs, unmatched_a, unmatched_b)."""
    used = np.zeros(len(b), bool)
    pairs = []
    order = np.lexsort((b["xpos"], b["ypos"]))
    by = b["ypos"][order]
    for i in range(len(a)):
        p = a[i]
        tol = pos_tol * max(1.0, p["subsampling"])
        lo, hi = np.searchsorted(by, p["ypos"] - tol), np.searchsorted(by, p["ypos"] + tol)
        best, bestd = -1, 1e9
        for j in order[lo:hi]:
            if used[j]:
                continue
            q = b[j]
            if q["subsampling"] != p["subsampling"]:
                continue
            if abs(q["xpos"] - p["xpos"]) > tol:
                continue
            if abs(q["scale"] - p["scale"]) > scale_rel * p["scale"]:
                continue
            do = abs(q["orientation"] - p["orientation"]) % 360.0
            do = min(do, 360.0 - do)
            if not (do <= ori_tol):
                continue
            d = abs(q["xpos"] - p["xpos"]) + abs(q["ypos"] - p["ypos"]) + do * 0.01
            if d < bestd:
                best, bestd = j, d
        if best >= 0:
            used[best] = True
            pairs.append((i, best))
    ua = sorted(set(range(len(a))) - {i for i, _ in pairs})
    ub = list(np.nonzero(~used)[0])
    return pairs, ua, ub


def compare_sets(a, b, **kw):
    """Summary dict: matched fraction (both ways) and worst relative errors over the pairs."""
    pairs, ua, ub = pair_points(a, b, **kw)
    out = {"na": len(a), "nb": len(b), "pairs": len(pairs), "unmatched_a": len(ua), "unmatched_b": len(ub)}
    if pairs:
        ia = np.array([i for i, _ in pairs]); ib = np.array([j for _, j in pairs])
        pa, pb = a[ia], b[ib]
        out["pos_err"] = float(np.max(np.maximum(np.abs(pa["xpos"] - pb["xpos"]), np.abs(pa["ypos"] - pb["ypos"]))
                                      / np.maximum(1.0, pa["subsampling"])))
        out["scale_rel"] = float(np.max(np.abs(pa["scale"] - pb["scale"]) / pa["scale"]))
        do = np.abs(pa["orientation"] - pb["orientation"]) % 360.0
        out["ori_err"] = float(np.max(np.minimum(do, 360.0 - do)))
        dd = np.abs(pa["data"] - pb["data"])
        finite = np.isfinite(dd).all(axis=1)
        out["desc_max"] = float(dd[finite].max()) if finite.any() else 0.0
        out["desc_med"] = float(np.median(dd[finite].max(axis=1))) if finite.any() else 0.0
        out["desc_bad"] = int((dd[finite].max(axis=1) > 1e-3).sum())
        out["sharp_rel"] = float(np.max(np.abs(pa["sharpness"] - pb["sharpness"]) / np.maximum(1e-3, np.abs(pa["sharpness"]))))
    return out
