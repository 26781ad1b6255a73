"""The oracle against the golden fixtures produced by the UNMODIFIED reference library on a
B200 (tests/golden/make_golden.py).  This is what pins the oracle: the deterministic image
stages and the matcher must agree bit for bit, extraction as a set within tolerance (the
CPU cannot reproduce the GPU's MUFU approximations and texture blend exactly)."""
import hashlib
import os

import numpy as np
import pytest

import oracle
from compare import compare_sets
from cudasift_b200.synth import synth_descriptors, synth_image

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def canon(p):
    return p[np.lexsort((p["orientation"], p["scale"], p["xpos"], p["ypos"], p["subsampling"]))]


def _load(name):
    path = os.path.join(G, name)
    if not os.path.exists(path):
        pytest.skip("golden fixture %s not generated yet" % name)
    return np.load(path)


def test_image_stages_bit_exact_vs_reference_fixture():
    st = _load("stages.npz")
    img = synth_image(320, 240, seed=int(st["image_seed"]))
    lp = oracle.lowpass(img, 1.0)
    assert np.array_equal(lp[100:132, 100:132], st["lowpass_crop"]) and sha(lp) == str(st["lowpass_sha"])
    sd = oracle.scaledown(img)
    assert np.array_equal(sd[40:72, 40:72], st["scaledown_crop"]) and sha(sd) == str(st["scaledown_sha"])
    su = oracle.scaleup(np.ascontiguousarray(img[:100, :128]))
    assert sha(su) == str(st["scaleup_sha"])
    for octave in (5, 3):
        d = oracle.dog(img, 5, octave)
        assert np.array_equal(d[:, 100:132, 100:132], st["dog%d_crop" % octave])
        assert sha(d) == str(st["dog%d_sha" % octave]), "DoG planes (octave %d) differ from the reference" % octave


@pytest.mark.parametrize("name", ["synth320", "synth640", "left_crop"])
def test_extract_vs_reference_fixture(name):
    f = _load("extract_%s.npz" % name)
    if name == "left_crop":
        img = f["image"].astype(np.float32)
    else:
        w = int(name[5:])
        img = synth_image(w, w * 3 // 4, seed=7 if w == 320 else 1000)
    ref = f["run1"]
    pts, _ = oracle.extract(img, 5, 1.0, float(f["thresh"]))
    assert len(pts) == len(ref), (len(pts), len(ref))          # detection is exact arithmetic
    rep = compare_sets(canon(pts), canon(ref), pos_tol=0.01, ori_tol=3.0)
    assert rep["pairs"] >= 0.99 * len(ref), rep
    assert rep["pos_err"] < 1e-3 and rep["scale_rel"] < 1e-4 and rep["sharp_rel"] < 1e-3, rep
    assert rep["desc_med"] < 2e-3, rep


def test_match_bit_exact_vs_reference_fixture():
    m = _load("match.npz")
    for n1, n2 in ((300, 352), (1000, 1031)):
        sa, sb = [int(v) for v in m["%dx%d_seeds" % (n1, n2)]]
        got = oracle.match(synth_descriptors(n1, sa), synth_descriptors(n2, sb), threads=4)
        for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
            assert np.array_equal(got[f], m["%dx%d_%s" % (n1, n2, f)]), (n1, n2, f)
