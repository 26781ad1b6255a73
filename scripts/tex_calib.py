"""Probe the texture unit's linear-filter weight quantisation on a ramp image.
Writes gpurun_out/tex_calib.npz (xs, hw) for offline analysis."""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cudasift_b200 as cs

cs.InitCuda(0)
w, h = 256, 8
ramp = np.tile(np.arange(w, dtype=np.float32) * 256.0, (h, 1))      # T[i] = 256*i  -> tex = 256*(i + alpha_q)
img = cs.CudaImage().Allocate(w, h, None, False, None, ramp); img.Download()
n = 1 << 16
xs = (10.5 + np.arange(n, dtype=np.float64) / 8192.0).astype(np.float32)      # 8 texels, 1/8192 steps
ys = np.full(n, 3.5, np.float32)
dx, dy, do = cs.DeviceBuffer(n * 4), cs.DeviceBuffer(n * 4), cs.DeviceBuffer(n * 4)
dx.upload(xs); dy.upload(ys)
cs.lib().cs_tex_probe(img.d_data, w, h, img.pitch, dx.ptr, dy.ptr, n, do.ptr)
hw = do.download(np.float32, n)
# 2-D: random image, random coords, to learn the blend arithmetic
rng = np.random.default_rng(0)
im2 = rng.uniform(0, 255, (64, 64)).astype(np.float32)
img2 = cs.CudaImage().Allocate(64, 64, None, False, None, im2); img2.Download()
m = 4096
x2 = rng.uniform(1, 62, m).astype(np.float32); y2 = rng.uniform(1, 62, m).astype(np.float32)
dx2, dy2, do2 = cs.DeviceBuffer(m * 4), cs.DeviceBuffer(m * 4), cs.DeviceBuffer(m * 4)
dx2.upload(x2); dy2.upload(y2)
cs.lib().cs_tex_probe(img2.d_data, 64, 64, img2.pitch, dx2.ptr, dy2.ptr, m, do2.ptr)
hw2 = do2.download(np.float32, m)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "tex_calib.npz"), xs=xs, hw=hw, im2=im2, x2=x2, y2=y2, hw2=hw2)
print("tex calib written")

# ---- hot-pixel probe: the four bilinear weights as functions of the fractional position ----
hot = np.zeros((16, 16), np.float32); hot[8, 8] = 65536.0
img3 = cs.CudaImage().Allocate(16, 16, None, False, None, hot); img3.Download()
g = np.arange(0, 2 * 256 + 1, dtype=np.float64) / 256.0 + 7.5        # x in [7.5, 9.5]: both sides of texel 8
sub = (np.arange(0, 8) / 2048.0)                                      # sub-steps of 1/2048
gx = (g[::4, None] + sub[None, :]).ravel()
X, Y = np.meshgrid(gx, gx)
x3, y3 = X.ravel().astype(np.float32), Y.ravel().astype(np.float32)
m3 = len(x3)
dx3, dy3, do3 = cs.DeviceBuffer(m3 * 4), cs.DeviceBuffer(m3 * 4), cs.DeviceBuffer(m3 * 4)
dx3.upload(x3); dy3.upload(y3)
cs.lib().cs_tex_probe(img3.d_data, 16, 16, img3.pitch, dx3.ptr, dy3.ptr, m3, do3.ptr)
hw3 = do3.download(np.float32, m3)
# column ramp: y quantisation alone
rampy = np.tile((np.arange(64, dtype=np.float32) * 256.0)[:, None], (1, 16))
img4 = cs.CudaImage().Allocate(16, 64, None, False, None, rampy); img4.Download()
ys4 = (10.5 + np.arange(n, dtype=np.float64) / 8192.0).astype(np.float32)
xs4 = np.full(n, 3.5, np.float32)
dx.upload(xs4); dy.upload(ys4)
cs.lib().cs_tex_probe(img4.d_data, 16, 64, img4.pitch, dx.ptr, dy.ptr, n, do.ptr)
hw4 = do.download(np.float32, n)
np.savez_compressed(os.path.join(ROOT, "gpurun_out", "tex_calib2.npz"), x3=x3, y3=y3, hw3=hw3, ys4=ys4, hw4=hw4)
print("tex calib2 written")
