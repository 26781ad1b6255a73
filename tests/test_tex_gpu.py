"""The oracle's software texture fetch against the B200 texture unit (same set-up as the
gather stages, cudaSiftH.cu:186-205)."""
import numpy as np
import pytest

import oracle
from cudasift_b200.synth import synth_image

pytestmark = pytest.mark.gpu


def test_texture_emulation(cs):
    w, h = 256, 128
    arr = synth_image(w, h, seed=21)
    img = cs.CudaImage().Allocate(w, h, None, False, None, arr)
    img.Download()
    rng = np.random.default_rng(1)
    n = 20000
    xs = rng.uniform(-3, w + 3, n).astype(np.float32)
    ys = rng.uniform(-3, h + 3, n).astype(np.float32)
    xs[:100] = np.arange(100) + 0.5          # texel centres
    ys[:100] = 10.5
    dx, dy, do = cs.DeviceBuffer(n * 4), cs.DeviceBuffer(n * 4), cs.DeviceBuffer(n * 4)
    dx.upload(xs); dy.upload(ys)
    assert cs.lib().cs_tex_probe(img.d_data, w, h, img.pitch, dx.ptr, dy.ptr, n, do.ptr) == 0
    hw = do.download(np.float32, n)
    sw = oracle.tex2d(arr, xs, ys)
    assert np.array_equal(hw[:100], arr[10, :100])
    err = np.abs(hw - sw)
    # same 1.8 fixed-point weights and weight products -> only the final rounding can differ
    assert (hw == sw).mean() > 0.99, "texture emulation: only %.4f of the fetches are bit-identical" % (hw == sw).mean()
    assert err.max() < 1e-4, "texture emulation off: max=%g" % err.max()
