"""Deterministic synthetic inputs for tests and benchmarks.

Images: random filled discs/rectangles + uniform noise, Gaussian-blurred (sigma 1 px),
rescaled to mean 128 / std 60 and clipped to [1, 254] -- no exactly flat regions (the
reference's descriptor kernel degenerates on exactly-zero gradients, quirks Q21/Q22).
Descriptor sets: uniform[0,1) 128-vectors, L2-normalised (the pattern of match.cu:945-957).
"""
import numpy as np

try:
    from .records import SIFT_DTYPE
except ImportError:      # loaded as a plain module (bench.py --impl reference never imports the package)
    from records import SIFT_DTYPE


def _blur(img, sigma):
    try:
        from scipy.ndimage import gaussian_filter
        return gaussian_filter(img, sigma, mode="nearest")
    except Exception:  # separable numpy fallback
        r = int(4 * sigma + 0.5)
        x = np.arange(-r, r + 1)
        k = np.exp(-x * x / (2.0 * sigma * sigma))
        k /= k.sum()
        pad = np.pad(img, r, mode="edge")
        t = sum(k[i] * pad[:, i:i + img.shape[1]] for i in range(2 * r + 1))
        return sum(k[i] * t[i:i + img.shape[0], :] for i in range(2 * r + 1))


def synth_image(width=1920, height=1080, seed=1000, shapes=600, small=536, noise=40.0, sigma=1.0):
    """Feature density close to the reference's demo photos (about 1.9k points at 1080p,
    thresh 3.0): `small` small shapes + (shapes-small) large ones, +-noise, blur sigma."""
    rng = np.random.default_rng(seed)
    img = np.full((height, width), 128.0, np.float64)
    for i in range(shapes):
        val = rng.uniform(0, 255)
        cx, cy = rng.uniform(0, width), rng.uniform(0, height)
        lo, hi = (0.004, 0.03) if i < small else (0.02, 0.15)
        if rng.random() < 0.5:
            r = rng.uniform(lo, hi) * min(width, height)
            x0, x1 = max(0, int(cx - r) - 1), min(width, int(cx + r) + 2)
            y0, y1 = max(0, int(cy - r) - 1), min(height, int(cy + r) + 2)
            yy, xx = np.mgrid[y0:y1, x0:x1]
            img[y0:y1, x0:x1][(xx - cx) ** 2 + (yy - cy) ** 2 < r * r] = val
        else:
            hw, hh = rng.uniform(lo, hi * 1.3) * width, rng.uniform(lo, hi * 1.3) * height
            x0, x1 = max(0, int(cx - hw) - 1), min(width, int(cx + hw) + 2)
            y0, y1 = max(0, int(cy - hh) - 1), min(height, int(cy + hh) + 2)
            yy, xx = np.mgrid[y0:y1, x0:x1]
            img[y0:y1, x0:x1][(np.abs(xx - cx) < hw) & (np.abs(yy - cy) < hh)] = val
    img += rng.uniform(-noise, noise, size=img.shape)
    img = _blur(img, sigma)
    img = (img - img.mean()) / (img.std() + 1e-9) * 60.0 + 128.0
    return np.clip(img, 1.0, 254.0).astype(np.float32)


def synth_descriptors(n, seed, width=1920, height=1080, sift_like=False):
    rng = np.random.default_rng(seed)
    pts = np.zeros(n, SIFT_DTYPE)
    d = rng.random((n, 128), dtype=np.float32)
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    if sift_like:   # clamp at 0.2 and renormalise as the descriptor stage does
        d = np.minimum(d, 0.2)
        d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts["data"] = d.astype(np.float32)
    pts["xpos"] = rng.uniform(0, width, n).astype(np.float32)
    pts["ypos"] = rng.uniform(0, height, n).astype(np.float32)
    pts["scale"] = rng.uniform(1, 8, n).astype(np.float32)
    pts["subsampling"] = 1.0
    pts["match"] = -1
    return pts
