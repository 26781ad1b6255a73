"""The CUDA path against the committed golden fixtures of the reference (no reference
library needed at run time)."""
import hashlib
import os

import numpy as np
import pytest

from compare import compare_sets
from cudasift_b200.synth import synth_descriptors, synth_image

pytestmark = pytest.mark.gpu
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


def test_stages_vs_fixture(cs, selflib):
    st = _load("stages.npz")
    img = synth_image(320, 240, seed=int(st["image_seed"]))
    assert sha(selflib.lowpass(img, 1.0)) == str(st["lowpass_sha"])
    assert sha(selflib.scaledown(img)) == str(st["scaledown_sha"])
    assert sha(selflib.scaleup(np.ascontiguousarray(img[:100, :128]))) == str(st["scaleup_sha"])
    src = cs.CudaImage().Allocate(320, 240, None, False, None, img); src.Download()
    for octave in (5, 3):
        buf = cs.DeviceBuffer(7 * 240 * src.pitch * 4); buf.zero()
        assert cs.lib().cs_dog_planes(src.d_data, buf.ptr, 320, 240, src.pitch, 5, octave) == 0
        d = buf.download(np.float32, 7 * 240 * src.pitch).reshape(7, 240, src.pitch)[:, :, :320]
        assert sha(d) == str(st["dog%d_sha" % octave])


@pytest.mark.parametrize("name", ["synth320", "synth640", "left_crop"])
def test_extract_vs_fixture(cs, name):
    f = _load("extract_%s.npz" % name)
    if name == "left_crop":
        img = f["image"].astype(np.float32)
    else:
        w = int(name[5:])
        img = synth_image(w, w * 3 // 4, seed=7 if w == 320 else 1000)
    r1, r2 = f["run1"], f["run2"]
    got = canon(cs.extract_host(img, thresh=float(f["thresh"])))
    assert len(got) == len(r1)
    noise = compare_sets(canon(r1), canon(r2))
    rep = compare_sets(got, canon(r1))
    assert rep["pairs"] >= noise["pairs"] - 1, (rep, noise)
    assert rep["pos_err"] < 1e-3 and rep["scale_rel"] < 1e-3 and rep["ori_err"] < 0.36, rep   # 1e-3 of 360 deg
    assert rep["desc_bad"] <= noise["desc_bad"] + max(1, 0.002 * rep["pairs"]), (rep, noise)


@pytest.mark.parametrize("mode", [1, 2])
def test_match_vs_fixture(cs, mode):
    m = _load("match.npz")
    for n1, n2 in ((300, 352), (1000, 1031)):
        sa, sb = [int(v) for v in m["%dx%d_seeds" % (n1, n2)]]
        got, _ = cs.match_host(synth_descriptors(n1, sa), synth_descriptors(n2, sb), mode=mode)
        for f in ("score", "ambiguity", "match", "match_xpos", "match_ypos"):
            assert np.array_equal(got[f], m["%dx%d_%s" % (n1, n2, f)]), (mode, n1, n2, f)
