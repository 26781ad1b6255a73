"""Round-2 GPU check: batched TMA pipeline vs the round-1 (legacy) kernels on the same inputs.
usage: python scripts/r2_check.py [pyramid|extract|batch|time|all]   (run on the GPU box)"""
import ctypes
import os
import sys
import time
import traceback

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudasift_b200 as cs
from cudasift_b200.synth import synth_image

KEY = ("subsampling", "ypos", "xpos", "scale", "orientation")


def canon(p):
    return p[np.lexsort(tuple(p[k] for k in reversed(KEY)))]


def dev_image(img):
    h, w = img.shape
    pitch = cs.iAlignUp(w, 128)
    ci = cs.CudaImage().Allocate(w, h, pitch, False, None, img)
    ci.Download()
    return ci, pitch


def run_extractor(img, legacy, octaves=5, thresh=3.0, scaleUp=False, levels=True):
    cs.set_tuning("legacy", 1 if legacy else 0)
    h, w = img.shape
    ex = cs.Extractor(w, h, octaves, 32768, scaleUp)
    ci, pitch = dev_image(img)
    ex.submit_device(ci.d_data, pitch, 1.0, thresh, 0.0)
    n = ex.wait()
    pts = ex.device_points_at(0, n)
    lv = []
    if levels:
        nl = octaves
        for l in range(nl):
            try:
                lv.append(ex.read_level(0, l))
            except cs.CudaSiftError:
                break
    ex.close()
    cs.set_tuning("legacy", 0)
    return pts, lv


def cmp_sets(a, b, tag):
    a, b = canon(a), canon(b)
    if len(a) != len(b):
        print("  %s: COUNT MISMATCH %d vs %d" % (tag, len(a), len(b)))
        # how many of a are in b by exact position/scale
        sa = set(zip(a["xpos"].tolist(), a["ypos"].tolist(), a["scale"].tolist(), a["orientation"].tolist()))
        sb = set(zip(b["xpos"].tolist(), b["ypos"].tolist(), b["scale"].tolist(), b["orientation"].tolist()))
        print("    common %d, only new %d, only legacy %d" % (len(sa & sb), len(sa - sb), len(sb - sa)))
        only = sorted(sa - sb)[:5], sorted(sb - sa)[:5]
        print("    e.g. only new", only[0], "only legacy", only[1])
        return False
    ok = True
    for f in ("xpos", "ypos", "scale", "sharpness", "edgeness", "orientation", "subsampling", "data"):
        eq = np.array_equal(a[f], b[f])
        if not eq:
            d = np.abs(a[f].astype(np.float64) - b[f].astype(np.float64))
            print("  %s: field %s differs: max %.3g, rows %d" % (tag, f, d.max(), int((d.reshape(len(a), -1).max(axis=1) > 0).sum())))
            ok = False
    print("  %s: %d points, %s" % (tag, len(a), "IDENTICAL" if ok else "DIFFERENT"))
    return ok


def stage_pyramid():
    for (w, h, oct_) in ((1920, 1080, 5), (1280, 960, 5), (641, 479, 4), (150, 100, 3), (1000, 700, 7), (300, 200, 1), (257, 131, 2)):
        img = synth_image(w, h, seed=7)
        _, lnew = run_extractor(img, False, oct_)
        _, lold = run_extractor(img, True, oct_)
        res = []
        for l, (a, b) in enumerate(zip(lnew, lold)):
            res.append("L%d %s %s" % (l, a.shape, "ok" if a.shape == b.shape and np.array_equal(a, b) else
                                      "DIFF(%d px, max %.3g)" % (int((a != b).sum()), float(np.abs(a - b).max())) if a.shape == b.shape else "SHAPE"))
        print("pyramid %dx%d oct %d: %s" % (w, h, oct_, "; ".join(res)), flush=True)


def stage_extract():
    for (w, h, oct_, th, up) in ((1920, 1080, 5, 3.0, False), (1280, 960, 5, 3.0, False), (641, 479, 4, 2.0, False),
                                 (150, 100, 3, 1.0, False), (640, 480, 5, 3.0, True), (300, 200, 1, 2.0, False),
                                 (1000, 700, 7, 3.0, False)):
        img = synth_image(w, h, seed=11)
        pn, _ = run_extractor(img, False, oct_, th, up, levels=False)
        po, _ = run_extractor(img, True, oct_, th, up, levels=False)
        cmp_sets(pn, po, "extract %dx%d oct %d thresh %.1f up %d" % (w, h, oct_, th, up))
        sys.stdout.flush()
    # the drop-in synchronous call
    img = synth_image(1280, 960, seed=3)
    cs.set_tuning("legacy", 0)
    a = cs.extract_host(img)
    cs.set_tuning("legacy", 1)
    b = cs.extract_host(img)
    cs.set_tuning("legacy", 0)
    cmp_sets(a, b, "drop-in cs_extract_host 1280x960")
    # dense noise, cap on/off
    rng = np.random.default_rng(5)
    noise = np.clip(128 + 60 * rng.standard_normal((480, 640)), 1, 254).astype(np.float32)
    for cap in (1, 0):
        cs.set_tuning("cap32", cap)
        a = cs.extract_host(noise, thresh=0.5)
        print("  dense noise thresh 0.5 cap32=%d: %d points" % (cap, len(a)))
    cs.set_tuning("cap32", 1)
    cs.set_tuning("legacy", 1)
    b = cs.extract_host(noise, thresh=0.5)
    cs.set_tuning("legacy", 0)
    print("  dense noise legacy (no cap): %d points" % len(b), flush=True)


def stage_batch():
    w, h = 1920, 1080
    imgs = [synth_image(w, h, seed=100 + i) for i in range(6)]
    singles = [run_extractor(im, False, levels=False)[0] for im in imgs]
    cs.set_tuning("legacy", 0)
    ex = cs.Extractor(w, h, 5, 32768, False, batch=6)
    cis = [dev_image(im) for im in imgs]
    for rep in range(3):
        ex.submit_device_batch([c[0].d_data for c in cis], cis[0][1], 1.0, 3.0, 0.0)
        counts = ex.wait_batch(6)
    ok = True
    for i in range(6):
        ok &= cmp_sets(ex.device_points_at(i, counts[i]), singles[i], "batch slot %d" % i)
    # host path
    ptrs = []
    for i in range(6):
        hp = cs.lib().cs_extractor_host_image_at(ex.handle, i)
        ctypes.memmove(hp, imgs[i].ctypes.data, w * h * 4)
        ptrs.append(hp)
    ex.submit_host_batch(ptrs, 1.0, 3.0, 0.0)
    counts = ex.wait_batch(6)
    for i in range(6):
        ok &= cmp_sets(ex.host_points_at(i, counts[i]), singles[i], "host batch slot %d" % i)
    # partial batch
    ex.submit_device_batch([c[0].d_data for c in cis[:3]], cis[0][1], 1.0, 3.0, 0.0)
    counts = ex.wait_batch(3)
    for i in range(3):
        ok &= cmp_sets(ex.device_points_at(i, counts[i]), singles[i], "partial batch slot %d" % i)
    print("batch:", "ALL IDENTICAL" if ok else "DIFFERENCES", flush=True)
    ex.close()


def stage_time():
    w, h = 1920, 1080
    imgs = [synth_image(w, h, seed=1000 + i) for i in range(8)]
    cis = [dev_image(imgs[i % 8]) for i in range(32)]
    ptrs = [c[0].d_data for c in cis]
    pitch = cis[0][1]
    for legacy in (1, 0):
        cs.set_tuning("legacy", legacy)
        ex = cs.Extractor(w, h, 5, 32768, False, batch=1)
        t = []
        for i in range(12):
            n, ms = ex.profile(ptrs[i], pitch, 1.0, 3.0, 0.0)
            t.append(ms)
        t = np.array(t[2:]).mean(axis=0) * 1e3
        print("single image %s: pyrA %.1f chain %.1f detect %.1f describe %.1f total %.1f us (%d pts)" %
              ("legacy" if legacy else "new", t[0], t[1], t[2], t[3], t[4], n), flush=True)
        ex.close()
    cs.set_tuning("legacy", 0)
    for B in (4, 16, 32):
        for hs, pa in ((0, 0), (16, 0), (32, 0), (64, 0), (0, 16), (0, 128)):
            if B != 32 and (hs, pa) != (0, 0):
                continue
            cs.set_tuning("d2_hs", hs); cs.set_tuning("pa_rows", pa)
            ex = cs.Extractor(w, h, 5, 32768, False, batch=B)
            t = []
            for i in range(8):
                n, ms = ex.profile_batch(ptrs[:B], pitch, 1.0, 3.0, 0.0)
                t.append(ms)
            t = np.array(t[2:]).mean(axis=0) * 1e3 / B
            print("batch %2d hs %2d pa %3d per image: pyrA %.1f chain %.1f detect %.1f describe %.1f total %.1f us (%.0f img/s, %d pts/img)" %
                  (B, hs, pa, t[0], t[1], t[2], t[3], t[4], 1e6 / t[4], n // B), flush=True)
            # graph submit throughput
            for _ in range(3):
                ex.submit_device_batch(ptrs[:B], pitch, 1.0, 3.0, 0.0)
            ex.wait_batch(B)
            t0 = time.perf_counter()
            reps = 20
            for _ in range(reps):
                ex.submit_device_batch(ptrs[:B], pitch, 1.0, 3.0, 0.0)
            ex.wait_batch(B)
            dt = time.perf_counter() - t0
            print("     graph submits: %.0f img/s" % (reps * B / dt), flush=True)
            ex.close()
    cs.set_tuning("d2_hs", 0); cs.set_tuning("pa_rows", 0)
    # two batch extractors in flight
    B = 16
    exs = [cs.Extractor(w, h, 5, 32768, False, batch=B) for _ in range(2)]
    for rep in range(3):
        for k, ex in enumerate(exs):
            ex.submit_device_batch(ptrs[k * B:(k + 1) * B], pitch, 1.0, 3.0, 0.0)
    [ex.wait_batch(B) for ex in exs]
    t0 = time.perf_counter()
    reps = 20
    for rep in range(reps):
        for k, ex in enumerate(exs):
            ex.submit_device_batch(ptrs[k * B:(k + 1) * B], pitch, 1.0, 3.0, 0.0)
    [ex.wait_batch(B) for ex in exs]
    dt = time.perf_counter() - t0
    print("2 x batch16 in flight: %.0f img/s" % (reps * 2 * B / dt), flush=True)


if __name__ == "__main__":
    what = sys.argv[1] if len(sys.argv) > 1 else "all"
    cs.InitCuda(0)
    for name, fn in (("pyramid", stage_pyramid), ("extract", stage_extract), ("batch", stage_batch), ("time", stage_time)):
        if what in (name, "all"):
            print("==== %s" % name, flush=True)
            try:
                fn()
            except Exception:
                traceback.print_exc()
                sys.stdout.flush()
